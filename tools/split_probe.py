#!/usr/bin/env python
"""Split decode against one lane per stream on the cells of the K x block grid that have few
streams, and on single large buffers (device-resident, CUDA events).  Prints a Markdown table."""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from sweep import gen_biased, time_ms  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1 << 30)
    ap.add_argument("--cells", default="4:128,4:256,4:512,4:1024,8:256,8:512,8:1024,16:512,16:1024,32:1024,32:128")
    ap.add_argument("--singles", default="32:64,4:64,32:1,8:1", help="K:MiB single buffers")
    args = ap.parse_args()
    huf = importlib.import_module("huffman-avx512_b200")
    huf.load(build_if_missing=False)
    dev = torch.device("cuda", 0)
    raw = gen_biased(args.size, dev, 7)
    print("| K | block | streams | one lane per stream GB/s | split GB/s | ok |\n|---|---|---|---|---|---|")

    def cell(k, bs, n):
        codec = huf.BlockCodec(k, bs, device=dev)
        r = raw[:n]
        slots, sizes = codec.compress(r)
        offs = codec.slot_offsets(n)
        out = torch.empty(n, dtype=torch.uint8, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        work = torch.empty(codec.split_work_bytes(n), dtype=torch.uint8, device=dev)
        t0 = time_ms(lambda: codec.decompress(slots, offs, sizes, n, out=out, status=status, split=False), 3)
        out.zero_()
        t1 = time_ms(lambda: codec.decompress(slots, offs, sizes, n, out=out, status=status, split=True, work=work), 3)
        ok = bool(torch.equal(out, r)) and int(status.item()) == 0
        nb = codec.n_blocks(n)
        print(f"| {k} | {bs >> 10} KiB | {nb * k} | {n / t0 / 1e6:.1f} | {n / t1 / 1e6:.1f} | {'ok' if ok else 'MISMATCH'} |", flush=True)

    for c in args.cells.split(","):
        if not c:
            continue
        k, kib = (int(x) for x in c.split(":"))
        cell(k, kib << 10, args.size)
    for c in args.singles.split(","):
        if not c:
            continue
        k, mib = (int(x) for x in c.split(":"))
        cell(k, mib << 20, mib << 20)


if __name__ == "__main__":
    main()
