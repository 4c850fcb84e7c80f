#!/usr/bin/env python
"""Tuning probe: compress / decompress rates when the working set stays in L2 (small input,
no flush) against the usual DRAM-resident size -- tells how much of a kernel's time is memory
latency.  python tools/l2probe.py [--k 32] [--block 131072]"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--block", type=int, default=128 << 10)
    ap.add_argument("--sizes", default="16,32,64,1024")
    args = ap.parse_args()
    huf = importlib.import_module("huffman-avx512_b200")
    huf.load(build_if_missing=False)
    dev = torch.device("cuda", 0)
    codec = huf.BlockCodec(args.k, args.block, device=dev)
    g = torch.Generator(device=dev).manual_seed(7)
    for mib in [int(x) for x in args.sizes.split(",")]:
        n = mib << 20
        u = torch.rand(n, device=dev, generator=g).clamp_(min=1e-30)
        raw = (torch.floor(torch.log(u) / float(np.log(0.8))).to(torch.int64) % 256).to(torch.uint8)
        del u
        slots, sizes = codec.alloc_slots(n)
        offsets = codec.slot_offsets(n)
        out = torch.empty(n, dtype=torch.uint8, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        for _ in range(3):
            codec.compress(raw, slots=slots, sizes=sizes, status=status)
            codec.decompress(slots, offsets, sizes, n, out=out, status=status)
        assert torch.equal(out, raw)
        iters = max(10, 2048 // mib)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(iters):
            codec.compress(raw, slots=slots, sizes=sizes, status=status)
        ev[1].record()
        for _ in range(iters):
            codec.decompress(slots, offsets, sizes, n, out=out, status=status)
        ev[2].record()
        torch.cuda.synchronize()
        c = n * iters / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9
        d = n * iters / (ev[1].elapsed_time(ev[2]) * 1e-3) / 1e9
        print(f"{mib:5d} MiB ({n // args.block} blocks): compress {c:7.0f} GB/s  decompress {d:7.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
