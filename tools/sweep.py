#!/usr/bin/env python
"""BASELINE configs 3-5 on one GPU, or on N GPUs of one box under torchrun (one rank per GPU over
NCCL; every cell is timed as the max over ranks and reported as the sum over ranks):

  * config 5: stream count K in {4,8,16,32,48} x block size 16 KiB..1 MiB, compress and decompress
    GB/s of raw bytes and the fraction of the measured HBM roofline on N(1+rho) algorithmic bytes;
  * config 4: English-letter-frequency text, 128 KiB x 32, per-block and shared tables;
  * config 3: byte histogram on uniform and skewed inputs (8 GiB with --hist-gib 8).

Writes a Markdown report (default profiles/sweep.md) and prints one JSON object (rank 0).
Device-resident timing with CUDA events, 3 warm-up + `--iters` timed launches per cell; every cell
is round-trip checked on the device.  Config 5 is weak scaling (`--size` bytes per GPU); configs 3
and 4 take TOTAL sizes (`--hist-total-gib`, `--english-total-gib`) cut into one contiguous shard
per GPU, with the 256-bin histogram all-reduce inside the timed region where a table is shared.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 tools/sweep.py --hist-total-gib 8 --english-total-gib 16 --out profiles/r2_sweep_8gpu.md
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))

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
    cdf = torch.cumsum(freq / freq.sum(), 0)
    cdf[-1] = 1.0
    step = 1 << 27
    for lo in range(0, n, step):  # inverse-CDF draw
        m = min(step, n - lo)
        idx = torch.searchsorted(cdf, torch.rand(m, device=dev, generator=g)).clamp_(max=26)
        out[lo:lo + m] = syms[idx]
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
    ms = a.elapsed_time(b) / iters
    if WORLD > 1:  # a cell takes as long as its slowest rank
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def all_ok(ok):
    if WORLD > 1:
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
    return ok


def cell(huf, raw, k, bs, iters, peak, shared=False):
    """One (K, block size) cell on this rank's `raw`; rates are sums over the ranks, fractions are
    of WORLD x the measured HBM peak.  shared: one table for all blocks of all ranks -- the timed
    compress then includes the histogram kernel, the all-reduce and the table build."""
    n = raw.numel()
    codec = huf.BlockCodec(k, bs, device=raw.device)
    slots, sizes = codec.alloc_slots(n)
    offs = codec.slot_offsets(n)
    out = torch.empty(n, dtype=torch.uint8, device=raw.device)
    status = torch.zeros(1, dtype=torch.int32, device=raw.device)
    hist = torch.empty(256, dtype=torch.int64, device=raw.device)
    table = torch.empty(codec.table_bytes, dtype=torch.uint8, device=raw.device)

    def comp():
        t = None
        if shared:
            codec.histogram(raw, out=hist)
            huf.sharded.allreduce_histogram(hist)
            t = codec.build_table(hist, out=table)
        codec.compress(raw, slots=slots, sizes=sizes, table=t, status=status)

    tc = time_ms(comp, iters)
    # few streams in all (n_blocks * k): the library takes its split decode, which wants a workspace
    split = bool(codec.L.hufb200_decompress_prefers_split(k, codec.n_blocks(n), n))
    work = torch.empty(codec.split_work_bytes(n), dtype=torch.uint8, device=raw.device) if split else None
    td = time_ms(lambda: codec.decompress(slots, offs, sizes, n, out=out, status=status, split=split, work=work), iters)
    ok = all_ok(bool(torch.equal(out, raw)) and int(status.item()) == 0)
    csum = sizes[:codec.n_blocks(n)].to(torch.int64).sum()
    if WORLD > 1:
        dist.all_reduce(csum)
    tot = n * WORLD
    rho = float(csum.item()) / tot
    alg = tot * (1 + rho)
    return {"k": k, "block": bs, "ratio": rho, "ok": ok, "split_decode": split, "comp_GBps": tot / tc / 1e6, "dec_GBps": tot / td / 1e6,
            "comp_frac": alg / tc / 1e6 / (peak * WORLD), "dec_frac": alg / td / 1e6 / (peak * WORLD)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1 << 30, help="config 5: raw bytes per GPU per cell")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--hist-total-gib", type=float, default=8.0, help="config 3: total bytes over all GPUs")
    ap.add_argument("--english-total-gib", type=float, default=None,
                    help="config 4: total bytes over all GPUs (default: 2 GiB per GPU, i.e. 16 GiB on 8)")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "sweep.md"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--ks", default="", help="config 5: comma list of stream counts (default 4,8,16,32,48)")
    ap.add_argument("--blocks-kib", default="", help="config 5: comma list of block sizes in KiB")
    ap.add_argument("--skip", default="", help="comma list of config3,config4,config5")
    args = ap.parse_args()
    huf = importlib.import_module("huffman-avx512_b200")
    huf.load(build_if_missing=False)
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if WORLD > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    skip = set(args.skip.split(","))
    peak = 6528.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    res = {"peak_GBps": peak, "size_per_gpu": args.size, "n_gpus": WORLD}
    lines = [f"# Sweep (tools/sweep.py), {WORLD} x B200", "",
             f"{WORLD} GPU(s) of one box, one process per GPU (NCCL); device-resident, CUDA events, {args.iters} timed "
             f"launches after 3 warm-ups, a cell's time is the slowest rank's; rates are sums over the GPUs, fractions "
             f"are of {WORLD} x the measured HBM peak ({peak:.0f} GB/s per GPU) on N(1+ratio) algorithmic bytes. "
             "Every cell round-trips on the device on every rank.", ""]

    if "config5" not in skip:
        raw = gen_biased(args.size, dev, 7 + 100 * RANK)
        ks = tuple(int(x) for x in args.ks.split(",")) if args.ks else (4, 8, 16, 32, 48)
        blocks = [16 << 10, 64 << 10, 128 << 10, 256 << 10, 1 << 20] if args.quick else [16 << 10, 32 << 10, 64 << 10, 128 << 10, 256 << 10, 512 << 10, 1 << 20]
        if args.blocks_kib:
            blocks = [int(x) << 10 for x in args.blocks_kib.split(",")]
        res["config5"] = []
        lines += [f"## Config 5: biased input, {args.size / (1 << 30):g} GiB per GPU, K x block size "
                  "(compress / decompress GB/s of raw bytes; roofline fraction)", "",
                  "| block | " + " | ".join(f"K={k}" for k in ks) + " |", "|---|" + "---|" * len(ks)]
        for bs in blocks:
            row = []
            for k in ks:
                c = cell(huf, raw, k, bs, args.iters, peak)
                res["config5"].append(c)
                row.append(f"{c['comp_GBps']:.0f} / {c['dec_GBps']:.0f}{'*' if c['split_decode'] else ''} "
                           f"({c['comp_frac']:.2f} / {c['dec_frac']:.2f})" + ("" if c["ok"] else " **MISMATCH**"))
            lines.append(f"| {bs >> 10} KiB | " + " | ".join(row) + " |")
        lines += ["", "\\* split decode (n_blocks x K <= 16384 streams: every stream cut into items, one lane per item)"]
        del raw

    if "config4" not in skip:
        tot = int((args.english_total_gib if args.english_total_gib else 2.0 * WORLD) * (1 << 30))
        per = (tot // WORLD) // (128 << 10) * (128 << 10)
        eng = gen_english(per, dev, 11 + 100 * RANK)
        res["config4"] = {"total_bytes": per * WORLD}
        lines += ["", f"## Config 4: English-letter-frequency text, {per * WORLD / (1 << 30):g} GiB block-sharded over "
                  f"{WORLD} GPU(s), 128 KiB x 32", "",
                  "| table | ratio | compress GB/s | decompress GB/s | roofline frac (c / d) |", "|---|---|---|---|---|"]
        for name, shared in (("per-block", False), ("shared (histogram + all-reduce + table build inside the timed compress)", True)):
            c = cell(huf, eng, 32, 128 << 10, args.iters, peak, shared=shared)
            res["config4"]["shared" if shared else "per-block"] = c
            lines.append(f"| {name} | {c['ratio']:.4f} | {c['comp_GBps']:.0f} | {c['dec_GBps']:.0f} | "
                         f"{c['comp_frac']:.2f} / {c['dec_frac']:.2f}" + ("" if c["ok"] else " **MISMATCH**") + " |")
        del eng

    if "config3" not in skip:
        nh_tot = int(args.hist_total_gib * (1 << 30))
        nh = nh_tot // WORLD // 16 * 16
        codec = huf.BlockCodec(32, 128 << 10, device=dev)
        hist = torch.empty(256, dtype=torch.int64, device=dev)
        res["config3"] = {"total_bytes": nh * WORLD}
        lines += ["", f"## Config 3: byte histogram, {nh * WORLD / (1 << 30):g} GiB pre-sharded over {WORLD} GPU(s), "
                  "256 x i64 all-reduce inside the timed region", "",
                  "| input | GB/s (all GPUs) | frac of measured HBM peak | of which all-reduce | total check |", "|---|---|---|---|---|"]
        g = torch.Generator(device=dev).manual_seed(3 + 100 * RANK)
        step = 1 << 28
        for name in ("uniform", "skewed"):
            u = torch.empty(nh, dtype=torch.uint8, device=dev)
            for lo in range(0, nh, step):
                m = min(step, nh - lo)
                if name == "uniform":
                    u[lo:lo + m] = torch.randint(0, 256, (m,), dtype=torch.uint8, device=dev, generator=g)
                else:  # 2^i copies of 'A'+i, i < 18, shuffled (codec/histogram_benchmark.cpp:30-40): ~50% one symbol
                    r = torch.rand(m, device=dev, generator=g)
                    u[lo:lo + m] = (ord("A") + 17 - torch.clamp(torch.floor(-torch.log2(r)), max=17)).to(torch.uint8)

            def hist_step():
                codec.histogram(u, out=hist)
                huf.sharded.allreduce_histogram(hist)

            t = time_ms(hist_step, args.iters)
            t_k = time_ms(lambda: codec.histogram(u, out=hist), args.iters)
            hist_step()
            okh = int(hist.sum().item()) == nh * WORLD
            tot = nh * WORLD
            res["config3"][name] = {"GBps": tot / t / 1e6, "frac": tot / t / 1e6 / (peak * WORLD), "ok": okh,
                                    "kernel_only_GBps": tot / t_k / 1e6, "allreduce_ms": max(0.0, t - t_k)}
            lines.append(f"| {name} | {tot / t / 1e6:.0f} | {tot / t / 1e6 / (peak * WORLD):.2f} | "
                         f"{max(0.0, t - t_k) * 1e3:.0f} us of {t * 1e3:.0f} us | {'ok' if okh else 'MISMATCH'} |")
            del u
    if RANK == 0:
        open(args.out, "w").write("\n".join(lines) + "\n")
        print(json.dumps(res))
    if WORLD > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
