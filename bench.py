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
    ap.add_argument("--sustained-seconds", type=float, default=3.0, help="length of the sustained loop (0 = skip)")
    ap.add_argument("--no-shared-leg", action="store_true", help="skip the shared-table sub-record")
    ap.add_argument("--parity-blocks", type=int, default=256, help="blocks compared with the CPU checker in the cpu_baseline leg")
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


def cpu_reference_run(args, seconds, threads, breadth=False):
    """Times the reference's CPU codec on a bounded sample; returns the cpu_baseline record.
    breadth: additionally the reference's published method (one thread, README.md:38) for
    K in {4,8,16,32,48} and its histogram functions."""
    import numpy as np
    from _libs import Oracle, Ref, have_ref
    data, nb = cpu_sample(args)
    sample = f"{nb} blocks x {args.block >> 10} KiB of the same biased input, >= {seconds:g} s per direction"
    if have_ref():
        r = Ref()
        res = {}
        # scalar is what the GPU output is compared with; the AVX-512 paths are the reference's fastest
        flags = open("/proc/cpuinfo").read()
        avx = all(f in flags for f in ("avx512f", "avx512bw", "avx512vbmi"))

        def variants(k):
            v = [("scalar", r.SCALAR)] if k in (1, 2, 4, 8, 16, 32, 48) else []
            if avx and k % 8 == 0:
                v += [("avx512_gather", r.GATHER), ("avx512_permute", r.PERMUTE)]
            return v

        def run(k, v, thr, secs):
            c, ratio = r.bench(k, v, 0, data, args.block, args.block, nb, thr, secs)
            d, _ = r.bench(k, v, 1, data, args.block, args.block, nb, thr, secs)
            return {"compress_GBps": c / GB, "decompress_GBps": d / GB, "roundtrip_GBps": 1.0 / (GB / c + GB / d),
                    "ratio": ratio}

        for name, v in variants(args.k):
            res[name] = run(args.k, v, threads, seconds)
        best = max(res.values(), key=lambda x: x["roundtrip_GBps"])
        out = {"value": best["roundtrip_GBps"], "unit": UNIT, "cores": threads, "kind": "reference",
               "sample": sample, "avx512": avx, "paths": res,
               "huff0": "unavailable (FiniteStateEntropy source not vendored; README: 1946/3636 MiB/s on a 9950X)"}
        if breadth:
            short = min(0.25, seconds)
            one = {}
            for k in (4, 8, 16, 32, 48):
                for name, v in variants(k):
                    one[f"{name}/K{k}"] = run(k, v, 1, short)
            out["single_thread"] = {"method": "one thread, the reference's published method (README.md:38)",
                                    "seconds_per_case": short, "paths": one}
            hsample = data[: min(data.size, 8 << 20)]
            hist = {}
            for which, name in ((0, "MakeHistogram"), (1, "Simple"), (2, "Multi"), (3, "Vectorized"), (4, "GatherScatter")):
                try:
                    hist[name] = {"one_thread_GBps": r.lib.ref_bench_histogram(which, hsample.ctypes.data_as(
                        C.POINTER(C.c_uint8)), hsample.size, 1, short) / GB,
                        "all_cores_GBps": r.lib.ref_bench_histogram(which, hsample.ctypes.data_as(
                            C.POINTER(C.c_uint8)), hsample.size, threads, short) / GB}
                except Exception as e:  # the shim knows fewer variants
                    hist[name] = str(e)
            out["histogram"] = hist
        return out
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
        self.first = threading.Event()  # set once the thread is up and has its first sample
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
            self.first.set()
            time.sleep(0.01)

    def start_sampling(self):
        """Starts the thread and returns once it is running: its start-up (thread creation, the first
        NVML call) must not take the interpreter away from the launch loop inside the timed region --
        with an empty stream that shows up as an idle GPU in the first step."""
        self.start()
        if self.ok:
            self.first.wait(0.5)

    def result(self):
        self.stop_flag.set()
        if self.is_alive():
            self.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------- config 1

def config1_table(huf, seconds=0.25, fixture="proba02_100k.bin", label="GenerateProbaData(0.2, 102400), 100 KiB",
                  large=True):
    """BASELINE config 1, the reference's own benchmark unit (codec/huffman_benchmark.cpp:61-81): ONE
    100 KiB biased buffer (GenerateProbaData(0.2, 102400), tests/golden/proba02_100k.bin) through
    the single-buffer drop-in calls that HuffmanCompressorB200<K>::Compress / Decompress make
    (hufb200_compress / hufb200_decompress, host pointers, copies and synchronisation inside), for
    K in {4,8,16,32,48}: microseconds per call and MiB/s -- beside the reference's scalar / AVX-512
    paths on the same buffer, one thread, its published method.  With another fixture the same
    table is the reference's file benchmark (BM_CompressFile / BM_DecompressFile, :218-248)."""
    import numpy as np
    from _cases import golden
    from _libs import Ref, have_ref
    buf = np.frombuffer(golden(fixture), dtype=np.uint8)
    n = buf.size
    L = huf.load()
    out = {"buffer": label, "unit": "MiB/s", "rows": {}}
    cap = L.hufb200_compress_bound(n, 64) + 64
    comp = np.empty(cap, dtype=np.uint8)
    back = np.empty(n, dtype=np.uint8)
    clen, olen = C.c_size_t(0), C.c_size_t(0)
    for k in (4, 8, 16, 32, 48):
        def cstep():
            huf.binding.check(L.hufb200_compress(k, C.c_void_p(buf.ctypes.data), n, C.c_void_p(comp.ctypes.data), cap,
                                                 C.byref(clen)))

        def dstep():
            huf.binding.check(L.hufb200_decompress(k, C.c_void_p(comp.ctypes.data), clen.value,
                                                   C.c_void_p(back.ctypes.data), n, C.byref(olen)))
        res = {}
        for name, fn in (("compress", cstep), ("decompress", dstep)):
            for _ in range(5):
                fn()
            t0 = time.perf_counter()
            it = 0
            while time.perf_counter() - t0 < seconds:
                fn()
                it += 1
            us = (time.perf_counter() - t0) / it * 1e6
            res[name + "_us_per_call"] = us
            res[name + "_MiBps"] = n / (us * 1e-6) / 2 ** 20
        assert olen.value == n and np.array_equal(back, buf), "config 1 round trip mismatch"
        res["compressed_bytes"] = clen.value
        out["rows"][f"HuffmanCompressorB200<{k}>"] = res
    # one LARGE buffer through the same two calls (pinned host memory): compress spreads pieces of
    # the K streams over the device; decompress cuts the K streams into items (split decode)
    if large:
        _config1_large(huf, L, out, clen, olen)
    if have_ref():
        r = Ref()
        flags = open("/proc/cpuinfo").read()
        avx = all(f in flags for f in ("avx512f", "avx512bw", "avx512vbmi"))
        for k in (4, 8, 16, 32, 48):
            vs = [("HuffmanCompressorMulti", r.SCALAR)]
            if avx and k % 8 == 0:
                vs += [("HuffmanCompressorAvxGather", r.GATHER), ("HuffmanCompressorAvxPermute", r.PERMUTE)]
            for cls, v in vs:
                c, _ = r.bench(k, v, 0, buf, n, n, 1, 1, seconds)
                d, _ = r.bench(k, v, 1, buf, n, n, 1, 1, seconds)
                out["rows"][f"{cls}<{k}>"] = {"compress_MiBps": c / 2 ** 20, "decompress_MiBps": d / 2 ** 20,
                                              "compress_us_per_call": n / c * 1e6, "decompress_us_per_call": n / d * 1e6,
                                              "threads": 1}
    return out


def _config1_large(huf, L, out, clen, olen):
    import torch
    from _cases import biased
    nl = 64 << 20
    big = torch.frombuffer(bytearray(biased(nl, seed=64)), dtype=torch.uint8).pin_memory()
    capl = L.hufb200_compress_bound(nl, 32) + 64
    compl = torch.empty(capl, dtype=torch.uint8).pin_memory()
    backl = torch.empty(nl, dtype=torch.uint8).pin_memory()

    def lc():
        huf.binding.check(L.hufb200_compress(32, C.c_void_p(big.data_ptr()), nl, C.c_void_p(compl.data_ptr()), capl,
                                             C.byref(clen)))

    def ld():
        huf.binding.check(L.hufb200_decompress(32, C.c_void_p(compl.data_ptr()), clen.value,
                                               C.c_void_p(backl.data_ptr()), nl, C.byref(olen)))
    tms = {}
    for name, fn, reps in (("compress", lc, 5), ("decompress", ld, 5)):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        tms[name] = (time.perf_counter() - t0) / reps
    assert olen.value == nl and torch.equal(backl, big), "large single buffer round trip mismatch"
    out["single_buffer_64MiB_K32"] = {"host_memory": "pinned", "compress_ms": tms["compress"] * 1e3,
                                      "compress_GBps": nl / tms["compress"] / GB, "decompress_ms": tms["decompress"] * 1e3,
                                      "decompress_GBps": nl / tms["decompress"] / GB,
                                      "note": "both through the host-pointer calls (PCIe inside); decompress = split decode, "
                                              "32 streams cut into items of 4 Kbit, one lane per item"}
    del big, compl, backl


def split_decode_record(huf, raw, dev):
    """Device-resident decode of inputs with FEW streams (the split decode's ground): the first
    256 MiB of the bench input as 1 MiB x K = 4 blocks, and its first 64 MiB as ONE buffer of 32
    streams -- one lane per stream against the split path, CUDA events, output compared."""
    import torch
    out = {"unit": "GB/s of raw bytes", "cases": {}}
    for name, k, bs, n in (("blocks_1MiB_K4_256MiB", 4, 1 << 20, 256 << 20), ("one_buffer_64MiB_K32", 32, 64 << 20, 64 << 20)):
        n = min(n, raw.numel() // bs * bs)
        if n == 0:
            continue
        codec = huf.BlockCodec(k, bs, device=dev)
        r = raw[:n]
        slots, sizes = codec.compress(r)
        offs = codec.slot_offsets(n)
        back = torch.empty(n, dtype=torch.uint8, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        work = torch.empty(codec.split_work_bytes(n), dtype=torch.uint8, device=dev)
        res = {"streams": codec.n_blocks(n) * k}
        for label, split, reps in (("one_lane_per_stream", False, 1 if n // (codec.n_blocks(n) * k) > (1 << 20) else 3), ("split", True, 5)):
            codec.decompress(slots, offs, sizes, n, out=back, status=status, split=split, work=work)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                codec.decompress(slots, offs, sizes, n, out=back, status=status, split=split, work=work)
            b.record()
            torch.cuda.synchronize()
            res[label + "_GBps"] = n / (a.elapsed_time(b) / reps * 1e-3) / GB
            assert torch.equal(back, r) and int(status.item()) == 0, f"{name}/{label}: decode mismatch"
            back.zero_()
        out["cases"][name] = res
        del slots, sizes, back, work
    return out


# ----------------------------------------------------------------------------- parity sample

def parity_sample(args, huf, codec, raw, slots, sizes, status, sh_table):
    """Compares `--parity-blocks` seeded blocks of the GPU output with the CPU checker."""
    import numpy as np
    import torch
    from _parity import compare_blocks
    n = raw.numel()
    nb = codec.n_blocks(n)
    rng = np.random.default_rng(20261018)
    pick = sorted(set([0, nb - 1] + [int(b) for b in rng.integers(0, nb, max(0, args.parity_blocks - 2))]))
    raw_h = raw.cpu().numpy()
    codec.compress(raw, slots=slots, sizes=sizes, status=status)
    torch.cuda.synchronize()
    sz = sizes[:nb].cpu().numpy().astype(np.int64)
    offs = np.arange(nb, dtype=np.int64) * codec.slot_stride
    sl = slots.cpu().numpy()
    name, checked, bad = compare_blocks(raw_h, sl, offs, sz, args.k, args.block, blocks=pick)
    out = {"checker": name, "per_block_tables": {"blocks": checked, "mismatches": len(bad)}}
    assert not bad, f"parity: {len(bad)} of {checked} blocks differ from {name}: {bad[:8]}"
    if sh_table is not None:
        from _libs import Oracle
        o = Oracle()
        codec.compress(raw, slots=slots, sizes=sizes, table=sh_table, status=status)
        torch.cuda.synchronize()
        sz = sizes[:nb].cpu().numpy().astype(np.int64)
        sl = slots.cpu().numpy()
        # the table the device built from the (all-reduced) histogram, as the header of block 0 carries it
        b0 = sl[: int(sz[0])].tobytes()
        mask = int.from_bytes(b0[4:8], "little")
        lens = [l for l in range(13) if (mask >> l) & 1]
        counts = list(b0[8: 8 + len(lens)])
        if len(lens) == 1 and counts[0] == 0:
            counts[0] = 256
        lc = np.zeros(13, dtype=np.uint16)
        for l, c in zip(lens, counts):
            lc[l] = c
        nsyms = int(lc.sum())
        syms = b0[8 + len(lens): 8 + len(lens) + nsyms]
        # ... which must be the table of the global histogram
        want = o.make_coding(np.bincount(raw_h, minlength=256).astype(np.uint32))
        ok_table = bool(np.array_equal(want["len_count"], lc)) and want["sorted_syms"] == syms
        bad_s = 0
        sub = pick[: max(8, len(pick) // 8)]
        for b in sub:
            lo = b * args.block
            w = o.compress_with_table(args.k, raw_h[lo: lo + args.block].tobytes(), lc, syms)
            if w != sl[offs[b]: offs[b] + int(sz[b])].tobytes():
                bad_s += 1
        out["shared_table"] = {"blocks": len(sub), "mismatches": bad_s, "table_equals_global_histogram_table": ok_table}
        assert bad_s == 0 and ok_table, "parity: shared-table blocks differ from the oracle"
    return out


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
    sampler.start_sampling()
    barrier()
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

    # ---- sustained: the same step in a seconds-long loop (the timed region above is a burst of
    # args.steps steps), clocks sampled throughout
    sustained = None
    if args.sustained_seconds > 0:
        s_sampler = ClockSampler(local)
        per = max(1e-3, elapsed_ms / args.steps * 1e-3)
        n_sus = max(args.steps, int(args.sustained_seconds / per) + 1)
        s_sampler.start_sampling()
        barrier()
        s0, s1 = ev(), ev()
        s0.record()
        for _ in range(n_sus):
            step()
        s1.record()
        barrier()
        s_clocks = s_sampler.result()
        sus_ms = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([sus_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sus_ms = float(t.item())
        sustained = {"value": world * n * n_sus / (sus_ms * 1e-3) / GB, "unit": UNIT, "steps": n_sus,
                     "seconds": sus_ms * 1e-3, "clocks": s_clocks}

    # ---- shared-table mode (the one collective of the path): histogram kernel -> all-reduce of
    # 256 x i64 over NCCL -> table build -> compress with that table, every rank on its shard
    shared = None
    if not args.no_shared_leg:
        sh_hist = torch.empty(256, dtype=torch.int64, device=dev)
        sh_table = torch.empty(codec.table_bytes, dtype=torch.uint8, device=dev)
        sh_steps = max(3, min(args.steps, 10))

        def shared_step(evs=None):
            if evs:
                evs[0].record()
            codec.histogram(raw, out=sh_hist)
            if evs:
                evs[1].record()
            huf.sharded.allreduce_histogram(sh_hist)
            if evs:
                evs[2].record()
            codec.build_table(sh_hist, out=sh_table)
            if evs:
                evs[3].record()
            codec.compress(raw, slots=slots, sizes=sizes, table=sh_table, status=status)
            if evs:
                evs[4].record()

        for _ in range(3):
            shared_step()
        sh_events = [[ev() for _ in range(5)] for _ in range(sh_steps)]
        barrier()
        for i in range(sh_steps):
            shared_step(sh_events[i])
        barrier()
        avg = lambda a, b: sum(e[a].elapsed_time(e[b]) for e in sh_events) / sh_steps
        sh_total_ms = avg(0, 4)
        if world > 1:
            t = torch.tensor([sh_total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sh_total_ms = float(t.item())
        # the shared table must decode: round trip on the device
        codec.decompress(slots, offsets, sizes, n, out=out, status=status)
        torch.cuda.synchronize()
        sh_ok = int(status.item()) == 0 and bool(torch.equal(out, raw))
        sh_rho = int(sizes[:nb].to(torch.int64).sum().item()) / n
        shared = {"GBps": world * n / (sh_total_ms * 1e-3) / GB, "ms_per_step": sh_total_ms,
                  "histogram_ms": avg(0, 1), "allreduce_us": 1e3 * avg(1, 2), "table_build_us": 1e3 * avg(2, 3),
                  "compress_ms": avg(3, 4), "compression_ratio": sh_rho, "roundtrip_ok": sh_ok,
                  "collective": f"all_reduce(SUM) of 256 x int64 over {'NCCL, ' + str(world) + ' ranks' if world > 1 else 'no process group (1 rank: no-op)'}"}
        assert sh_ok, "shared-table round trip mismatch"
        shared["_table"] = sh_table  # kept for the parity sample of the cpu_baseline leg
        # back to per-block tables for what follows
        codec.compress(raw, slots=slots, sizes=sizes, status=status)
        torch.cuda.synchronize()

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

        # pure-copy ceiling of the same step: the same byte counts over PCIe with nothing else --
        # first n bytes up while clen bytes come down (the compress call), then clen up while n
        # come down (the decompress call); pinned memory, two streams, full duplex
        csz = clen.value
        dev_a = torch.empty(n, dtype=torch.uint8, device=dev)
        dev_b = torch.empty(n, dtype=torch.uint8, device=dev)
        st_up, st_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def copy_step():
            for up, dn in ((n, csz), (csz, n)):
                with torch.cuda.stream(st_up):
                    dev_a[:up].copy_(host_raw[:up], non_blocking=True)
                with torch.cuda.stream(st_dn):
                    host_out[:dn].copy_(dev_b[:dn], non_blocking=True)
                st_up.synchronize()
                st_dn.synchronize()

        copy_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            copy_step()
        torch.cuda.synchronize()
        dtc = (time.perf_counter() - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([dtc], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtc = float(t.item())
        e2e["pcie_ceiling_GBps"] = world * n / dtc / GB
        e2e["pcie_ceiling_ms_per_step"] = 1e3 * dtc
        e2e["fraction_of_pcie_ceiling"] = dtc / dt
        del dev_a, dev_b

        # the same calls with PAGEABLE host memory (what a std::string caller of the policy class
        # has): the driver stages such copies through its own pinned buffers
        import numpy as np_
        pg_raw = np_.empty(n, dtype=np_.uint8)
        pg_raw[:] = host_raw.numpy()
        pg_comp = np_.empty(bound, dtype=np_.uint8)
        pg_out = np_.empty(n, dtype=np_.uint8)

        def pageable_step():
            huf.binding.check(L.hufb200_compress_blocks(args.k, args.block, C.c_void_p(pg_raw.ctypes.data), n,
                                                        C.c_void_p(pg_comp.ctypes.data), bound, C.byref(clen)))
            huf.binding.check(L.hufb200_decompress_blocks(C.c_void_p(pg_comp.ctypes.data), clen.value,
                                                          C.c_void_p(pg_out.ctypes.data), n, C.byref(olen)))

        pageable_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, args.e2e_steps - 1)):
            pageable_step()
        torch.cuda.synchronize()
        dtp = (time.perf_counter() - t0) / max(1, args.e2e_steps - 1)
        assert olen.value == n and np_.array_equal(pg_out, pg_raw), "pageable e2e round trip mismatch"
        if world > 1:
            t = torch.tensor([dtp], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtp = float(t.item())
        e2e["pageable"] = {"value": world * n / dtp / GB, "unit": UNIT, "ms_per_step": 1e3 * dtp,
                           "note": "same calls, pageable host buffers (driver-staged copies)"}
        del host_raw, host_comp, host_out, pg_raw, pg_comp, pg_out

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
        if sustained:
            res["sustained"] = sustained
        sh_table = shared.pop("_table") if shared else None
        if shared:
            res["shared_table"] = shared
        if not args.no_shared_leg and world == 1:
            res["split_decode"] = split_decode_record(huf, raw, dev)
        if not args.no_cpu_baseline and world == 1:
            res["cpu_baseline"] = cpu_reference_run(args, args.cpu_seconds, os.cpu_count() or 1, breadth=True)
            # the checker's other job in this leg: a seeded sample of the blocks the timed steps
            # produced, byte for byte against the CPU implementation (per-block tables), and the
            # same for shared-table mode against compress-with-that-table
            res["cpu_baseline"]["parity"] = parity_sample(args, huf, codec, raw, slots, sizes, status, sh_table)
            res["cpu_baseline"]["config1"] = config1_table(huf)
            try:  # the reference's file benchmark on real text (an extra table: never costs the bench line)
                res["cpu_baseline"]["config1_file"] = config1_table(
                    huf, seconds=0.15, fixture="real_text_100k.bin", large=False,
                    label="real English prose, first 100 KiB (tests/golden/real_text_100k.bin; stands in for enwik8, "
                          "codec/huffman_benchmark.cpp:218-248)")
            except Exception as e:  # noqa: BLE001
                res["cpu_baseline"]["config1_file"] = {"error": repr(e)}
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
