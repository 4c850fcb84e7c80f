#!/usr/bin/env python
"""bench.py -- headline benchmark of the multi-stream Huffman hot path (BASELINE.json config 2).

A step = one pass of the hot path over one batch of synthetic input: compress (per-block
histogram + table build + 32-stream encode) and then decompress (header parse + multi-symbol
table + decode) of `--size` bytes per GPU (default 1 GiB) in 128 KiB blocks x 32 streams.
`value` = raw bytes that went through the whole round trip per second, summed over ranks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torch.distributed.run, one rank per GPU; every rank owns its own shard
(weak scaling, no data-path collective).  `--impl reference` times the reference's own CPU
implementation (oracle/_ref, else the oracle port) on the host cores, rank 0 only.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "huffman_roundtrip_raw_GBps"
UNIT = "GB/s"
GB = 1e9


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1 << 30, help="raw bytes per GPU per step")
    ap.add_argument("--block", type=int, default=128 << 10)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=2.0, help="per direction, cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--shared-table", action="store_true", help="one table for all blocks (histogram all-reduce)")
    return ap.parse_args()


def workload_name(args):
    return (f"biased p_i=0.8^i*0.2, {args.size / (1 << 30):g} GiB per GPU, {args.block >> 10} KiB blocks x "
            f"{args.k} streams, compress+decompress")


def config(args, n_gpus):
    return {"workload": workload_name(args), "block_bytes": args.block, "streams": args.k,
            "raw_bytes_per_gpu": args.size, "table": "shared" if args.shared_table else "per-block",
            "l2": "inputs larger than L2 (no flush needed)" if args.size > (256 << 20) else "L2 flushed between steps",
            "sharding": f"contiguous block ranges over {n_gpus} GPU(s), no data-path collective"}


# ----------------------------------------------------------------------------- CPU legs

def cpu_sample(args, n_blocks=64):
    import numpy as np
    from _cases import biased
    n_blocks = max(1, min(n_blocks, args.size // args.block))
    data = np.frombuffer(biased(n_blocks * args.block, seed=12345), dtype=np.uint8)
    return data, n_blocks


def cpu_reference_run(args, seconds, threads):
    """Times the reference's CPU codec on a bounded sample; returns (dict, kind)."""
    import numpy as np
    from _libs import Oracle, Ref, have_ref
    data, nb = cpu_sample(args)
    sample = f"{nb} blocks x {args.block >> 10} KiB of the same biased input, >= {seconds:g} s per direction"
    if have_ref():
        r = Ref()
        res = {}
        # scalar is what the GPU output is compared with; the AVX-512 paths are the reference's fastest
        variants = [("scalar", r.SCALAR)]
        flags = open("/proc/cpuinfo").read()
        avx = all(f in flags for f in ("avx512f", "avx512bw", "avx512vbmi"))
        if avx and args.k % 8 == 0:
            variants += [("avx512_gather", r.GATHER), ("avx512_permute", r.PERMUTE)]
        for name, v in variants:
            c, ratio = r.bench(args.k, v, 0, data, args.block, args.block, nb, threads, seconds)
            d, _ = r.bench(args.k, v, 1, data, args.block, args.block, nb, threads, seconds)
            res[name] = {"compress_GBps": c / GB, "decompress_GBps": d / GB,
                         "roundtrip_GBps": 1.0 / (GB / c + GB / d), "ratio": ratio}
        best = max(res.values(), key=lambda x: x["roundtrip_GBps"])
        return {"value": best["roundtrip_GBps"], "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": sample, "avx512": avx, "paths": res,
                "huff0": "unavailable (FiniteStateEntropy source not vendored; README: 1946/3636 MiB/s on a 9950X)"}
    # oracle port, single thread
    o = Oracle()
    blk = data[: args.block].tobytes()
    t0 = time.perf_counter()
    it = 0
    while time.perf_counter() - t0 < seconds:
        comp = o.compress(args.k, blk)
        it += 1
    tc = (time.perf_counter() - t0) / it
    t0 = time.perf_counter()
    it = 0
    while time.perf_counter() - t0 < seconds:
        o.decompress(args.k, comp)
        it += 1
    td = (time.perf_counter() - t0) / it
    return {"value": args.block / (tc + td) / GB, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "1 block, oracle port (reference build absent)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, args.steps)
    # each step is a bounded sample; the whole run stays within a few minutes
    per_step = min(args.cpu_seconds, 60.0 / (2 * (steps + args.warmup)))
    t0 = time.perf_counter()
    vals = []
    base = None
    for s in range(args.warmup + steps):
        base = cpu_reference_run(args, per_step, threads)
        if s >= args.warmup:
            vals.append(base["value"])
    wall = time.perf_counter() - t0
    v = sum(vals) / len(vals)
    base["value"] = v
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * wall / (steps + args.warmup), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": config(args, args.gpus), "cpu_baseline": base,
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ----------------------------------------------------------------------------- clocks

class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag.set()
        if self.is_alive():
            self.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------- GPU arm

def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line (NCCL_DEBUG=VERSION would print a banner there)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    huf = importlib.import_module("huffman-avx512_b200")
    huf.load(build_if_missing=False)  # fail loudly if the CUDA library is missing
    codec = huf.BlockCodec(args.k, args.block, device=dev)
    sharded = huf.sharded.ShardedCodec(codec)
    n = args.size
    nb = codec.n_blocks(n)

    # synthetic shard, resident in HBM before the timed region (seed differs per rank)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    raw = torch.empty(n, dtype=torch.uint8, device=dev)
    chunk = 1 << 27
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        u = torch.rand(m, device=dev, generator=g).clamp_(min=1e-30)
        raw[lo: lo + m] = (torch.floor(torch.log(u) / float(np.log(0.8))).to(torch.int64) % 256).to(torch.uint8)
        del u
    slots, sizes = codec.alloc_slots(n)
    offsets = codec.slot_offsets(n)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    flush = None
    if n <= (256 << 20):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(evs=None):
        if flush is not None:
            flush.zero_()
        if evs:
            evs[0].record()
        table = None
        if args.shared_table:
            table, _ = sharded.shared_table(raw)
        codec.compress(raw, slots=slots, sizes=sizes, table=table, status=status)
        if evs:
            evs[1].record()
        codec.decompress(slots, offsets, sizes, n, out=out, status=status)
        if evs:
            evs[2].record()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    if not os.environ.get("HUF_EXPERIMENT"):  # tuning builds may produce garbage on purpose
        assert int(status.item()) == 0, "kernel reported a malformed block"
        assert torch.equal(out, raw), "round trip mismatch"
    comp_bytes = int(sizes[:nb].to(torch.int64).sum().item())
    rho = comp_bytes / n

    sampler = ClockSampler(local)
    launches0 = huf.launch_count()
    per_step_events = [[ev(), ev(), ev()] for _ in range(args.steps)]
    barrier()
    sampler.start()
    t_start, t_end = ev(), ev()
    t_start.record()
    for s in range(args.steps):
        step(per_step_events[s])
    t_end.record()
    barrier()
    clocks = sampler.result()
    launches = huf.launch_count() - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    comp_ms = sum(e[0].elapsed_time(e[1]) for e in per_step_events) / args.steps
    dec_ms = sum(e[1].elapsed_time(e[2]) for e in per_step_events) / args.steps

    # histogram-only kernel (BASELINE config 3 shape, on this shard), same hygiene
    hist_out = torch.empty(256, dtype=torch.int64, device=dev)
    for _ in range(3):
        codec.histogram(raw, out=hist_out)
    h0, h1 = ev(), ev()
    torch.cuda.synchronize()
    h0.record()
    for _ in range(10):
        codec.histogram(raw, out=hist_out)
    h1.record()
    torch.cuda.synchronize()
    hist_ms = h0.elapsed_time(h1) / 10

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        host_raw = torch.empty(n, dtype=torch.uint8).pin_memory()
        host_raw.copy_(raw)
        bound = huf.lib.hufb200_container_bound(n, args.block, args.k)
        host_comp = torch.empty(bound, dtype=torch.uint8).pin_memory()
        host_out = torch.empty(n, dtype=torch.uint8).pin_memory()
        clen = C.c_size_t(0)
        olen = C.c_size_t(0)
        L = huf.load()

        def e2e_step():
            huf.binding.check(L.hufb200_compress_blocks(args.k, args.block, C.c_void_p(host_raw.data_ptr()), n,
                                                        C.c_void_p(host_comp.data_ptr()), bound, C.byref(clen)))
            huf.binding.check(L.hufb200_decompress_blocks(C.c_void_p(host_comp.data_ptr()), clen.value,
                                                          C.c_void_p(host_out.data_ptr()), n, C.byref(olen)))

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        assert olen.value == n and torch.equal(host_out, host_raw), "e2e round trip mismatch"
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * n / dt / GB, "unit": UNIT, "h2d_bytes_per_step": n + clen.value,
               "d2h_bytes_per_step": clen.value + n, "ms_per_step": 1e3 * dt,
               "api": "hufb200_compress_blocks + hufb200_decompress_blocks (host pointers, pinned)"}
        del host_raw, host_comp, host_out

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s"
        alg_bytes = n * (1.0 + rho)
        dom = "k_compress_blocks" if comp_ms >= dec_ms else "k_decompress_blocks"
        dom_ms = max(comp_ms, dec_ms)
        achieved = alg_bytes / (dom_ms * 1e-3) / GB
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        except Exception:
            pass
        ms_per_step = elapsed_ms / args.steps
        res = {
            "metric": METRIC, "value": world * n / (ms_per_step * 1e-3) / GB, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config(args, world),
            "compress_GBps_per_gpu": n / (comp_ms * 1e-3) / GB, "decompress_GBps_per_gpu": n / (dec_ms * 1e-3) / GB,
            "histogram_GBps_per_gpu": n / (hist_ms * 1e-3) / GB, "compression_ratio": rho,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": dom_ms,
                         "frac_of_8TBps_nominal": achieved / 8000.0,
                         "other": {"kernel": "k_decompress_blocks" if dom == "k_compress_blocks" else "k_compress_blocks",
                                   "achieved": alg_bytes / (min(comp_ms, dec_ms) * 1e-3) / GB,
                                   "frac": alg_bytes / (min(comp_ms, dec_ms) * 1e-3) / GB / peak},
                         "histogram": {"achieved": n / (hist_ms * 1e-3) / GB, "frac": n / (hist_ms * 1e-3) / GB / peak}},
            "clocks": clocks, "gpu_launches": launches,
        }
        if e2e:
            res["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            res["cpu_baseline"] = cpu_reference_run(args, args.cpu_seconds, os.cpu_count() or 1)
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
