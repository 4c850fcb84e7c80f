#!/usr/bin/env python
"""BASELINE configs 3-5 on one GPU (or one rank per GPU under torchrun):

  * config 5: stream count K in {4,8,16,32,48} x block size 16 KiB..1 MiB, compress and decompress
    GB/s of raw bytes and the fraction of the measured HBM roofline on N(1+rho) algorithmic bytes;
  * config 4: English-letter-frequency text, 128 KiB x 32, per-block and shared tables;
  * config 3: byte histogram on uniform and skewed inputs (8 GiB with --hist-gib 8).

Writes a Markdown report (default profiles/sweep.md) and prints one JSON object.  Device-resident
timing with CUDA events, 3 warm-up + `--iters` timed launches per cell; every cell is round-trip
checked on the device.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gen_biased(n, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    step = 1 << 27
    for lo in range(0, n, step):
        m = min(step, n - lo)
        u = torch.rand(m, device=dev, generator=g).clamp_(min=1e-30)
        out[lo:lo + m] = (torch.floor(torch.log(u) / float(np.log(0.8))).to(torch.int64) % 256).to(torch.uint8)
    return out


def gen_english(n, dev, seed):
    freq = torch.tensor([8.167, 1.492, 2.782, 4.253, 12.702, 2.228, 2.015, 6.094, 6.966, 0.153, 0.772, 4.025,
                         2.406, 6.749, 7.507, 1.929, 0.095, 5.987, 6.327, 9.056, 2.758, 0.978, 2.360, 0.150,
                         1.974, 0.074, 21.0], device=dev)
    syms = torch.tensor(list(range(ord("a"), ord("z") + 1)) + [ord(" ")], dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(seed)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    step = 1 << 26
    for lo in range(0, n, step):
        m = min(step, n - lo)
        out[lo:lo + m] = syms[torch.multinomial(freq, m, replacement=True, generator=g)]
    return out


def time_ms(fn, iters):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def cell(huf, raw, k, bs, iters, peak, shared=False):
    n = raw.numel()
    codec = huf.BlockCodec(k, bs, device=raw.device)
    slots, sizes = codec.alloc_slots(n)
    offs = codec.slot_offsets(n)
    out = torch.empty(n, dtype=torch.uint8, device=raw.device)
    status = torch.zeros(1, dtype=torch.int32, device=raw.device)
    table = None
    if shared:
        table = codec.build_table(codec.histogram(raw))
    tc = time_ms(lambda: codec.compress(raw, slots=slots, sizes=sizes, table=table, status=status), iters)
    td = time_ms(lambda: codec.decompress(slots, offs, sizes, n, out=out, status=status), iters)
    ok = bool(torch.equal(out, raw)) and int(status.item()) == 0
    rho = float(sizes[:codec.n_blocks(n)].to(torch.int64).sum().item()) / n
    alg = n * (1 + rho)
    return {"k": k, "block": bs, "ratio": rho, "ok": ok, "comp_GBps": n / tc / 1e6, "dec_GBps": n / td / 1e6,
            "comp_frac": alg / tc / 1e6 / peak, "dec_frac": alg / td / 1e6 / peak}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1 << 30)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--hist-gib", type=float, default=2.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "sweep.md"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    huf = importlib.import_module("huffman-avx512_b200")
    huf.load(build_if_missing=False)
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    peak = 6528.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    res = {"peak_GBps": peak, "size": args.size}
    lines = ["# Sweep (tools/sweep.py)", "", f"One B200, {args.size / (1 << 30):g} GiB per cell, device-resident, CUDA events, "
             f"{args.iters} timed launches after 3 warm-ups; fractions are of the measured HBM peak "
             f"({peak:.0f} GB/s) on N(1+ratio) algorithmic bytes. Every cell round-trips on the device.", ""]

    raw = gen_biased(args.size, dev, 7)
    ks = (4, 8, 16, 32, 48)
    blocks = [16 << 10, 64 << 10, 128 << 10, 256 << 10, 1 << 20] if args.quick else [16 << 10, 32 << 10, 64 << 10, 128 << 10, 256 << 10, 512 << 10, 1 << 20]
    res["config5"] = []
    lines += ["## Config 5: biased input, K x block size (compress / decompress GB/s of raw bytes; roofline fraction)", "",
              "| block | " + " | ".join(f"K={k}" for k in ks) + " |", "|---|" + "---|" * len(ks)]
    for bs in blocks:
        row = []
        for k in ks:
            c = cell(huf, raw, k, bs, args.iters, peak)
            res["config5"].append(c)
            row.append(f"{c['comp_GBps']:.0f} / {c['dec_GBps']:.0f} ({c['comp_frac']:.2f} / {c['dec_frac']:.2f})"
                       + ("" if c["ok"] else " **MISMATCH**"))
        lines.append(f"| {bs >> 10} KiB | " + " | ".join(row) + " |")
    del raw

    eng = gen_english(args.size, dev, 11)
    res["config4"] = {}
    lines += ["", "## Config 4: English-letter-frequency text, 128 KiB x 32", "",
              "| table | ratio | compress GB/s | decompress GB/s | roofline frac (c / d) |", "|---|---|---|---|---|"]
    for name, shared in (("per-block", False), ("shared", True)):
        c = cell(huf, eng, 32, 128 << 10, args.iters, peak, shared=shared)
        res["config4"][name] = c
        lines.append(f"| {name} | {c['ratio']:.4f} | {c['comp_GBps']:.0f} | {c['dec_GBps']:.0f} | "
                     f"{c['comp_frac']:.2f} / {c['dec_frac']:.2f}" + ("" if c["ok"] else " **MISMATCH**") + " |")
    del eng

    nh = int(args.hist_gib * (1 << 30))
    codec = huf.BlockCodec(32, 128 << 10, device=dev)
    hist = torch.empty(256, dtype=torch.int64, device=dev)
    res["config3"] = {}
    lines += ["", f"## Config 3: byte histogram, {args.hist_gib:g} GiB per GPU", "",
              "| input | GB/s | frac of measured HBM peak | total check |", "|---|---|---|---|"]
    g = torch.Generator(device=dev).manual_seed(3)
    uni = torch.empty(nh, dtype=torch.uint8, device=dev)
    step = 1 << 28
    for lo in range(0, nh, step):
        m = min(step, nh - lo)
        uni[lo:lo + m] = torch.randint(0, 256, (m,), dtype=torch.uint8, device=dev, generator=g)
    for name in ("uniform", "skewed"):
        if name == "skewed":  # 2^i copies of 'A'+i, i < 18, shuffled (codec/histogram_benchmark.cpp:30-40): ~50% one symbol
            u = torch.empty(nh, dtype=torch.uint8, device=dev)
            for lo in range(0, nh, step):
                m = min(step, nh - lo)
                r = torch.rand(m, device=dev, generator=g)
                u[lo:lo + m] = (ord("A") + 17 - torch.clamp(torch.floor(-torch.log2(r)), max=17)).to(torch.uint8)
            uni = u
        t = time_ms(lambda: codec.histogram(uni, out=hist), args.iters)
        okh = int(hist.sum().item()) == nh
        res["config3"][name] = {"GBps": nh / t / 1e6, "frac": nh / t / 1e6 / peak, "ok": okh}
        lines.append(f"| {name} | {nh / t / 1e6:.0f} | {nh / t / 1e6 / peak:.2f} | {'ok' if okh else 'MISMATCH'} |")
    open(args.out, "w").write("\n".join(lines) + "\n")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
