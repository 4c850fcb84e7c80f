#!/usr/bin/env python
"""Re-hosted report tooling (the reference's make_table.py:17-67).

Turns our measurement files into (1) a google-benchmark-schema JSON
(`{"benchmarks": [{"name": "BM_CompressBiased<::huffman::HuffmanCompressorB200<32>>",
"bytes_per_second": ...}, ...]}`), which the reference's own make_table.py can read for the rows it
knows (Scalar / AVX-512 Gather / AVX-512 Permute), and (2) the README-style Markdown table with the
B200 rows next to the CPU rows measured in the same bench run.

    python tools/report.py profiles/r1_bench.json profiles/r1_sweep.json [--json out.json]
"""
import argparse
import json


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("bench_json")
    ap.add_argument("sweep_json", nargs="?")
    ap.add_argument("--json", default=None, help="write the google-benchmark-schema file here")
    args = ap.parse_args()
    bench = json.loads(open(args.bench_json).read().strip().splitlines()[-1])
    k = bench["config"]["streams"]
    rows = []  # (method, streams, compress B/s, decompress B/s, where)
    gb = []
    names = {"scalar": ("Scalar", "::huffman::HuffmanCompressorMulti"),
             "avx512_gather": ("AVX-512 Gather", "::huffman::HuffmanCompressorAvxGather"),
             "avx512_permute": ("AVX-512 Permute", "::huffman::HuffmanCompressorAvxPermute")}
    cpu = bench.get("cpu_baseline", {})
    for key, (label, cls) in names.items():
        p = cpu.get("paths", {}).get(key)
        if not p:
            continue
        c, d = p["compress_GBps"] * 1e9, p["decompress_GBps"] * 1e9
        rows.append((label, k, c, d, f"host, {cpu.get('cores')} threads"))
        gb.append({"name": f"BM_CompressBiased<{cls}<{k}>>", "bytes_per_second": c, "threads": cpu.get("cores")})
        gb.append({"name": f"BM_DecompressBiased<{cls}<{k}>>", "bytes_per_second": d, "threads": cpu.get("cores")})
    cells = [(k, bench["config"]["block_bytes"], bench["compress_GBps_per_gpu"] * 1e9, bench["decompress_GBps_per_gpu"] * 1e9)]
    if args.sweep_json:
        sw = json.load(open(args.sweep_json))
        cells += [(c["k"], c["block"], c["comp_GBps"] * 1e9, c["dec_GBps"] * 1e9) for c in sw.get("config5", [])
                  if c["block"] == bench["config"]["block_bytes"] and c["k"] != k]
    for kk, blk, c, d in sorted(cells):
        rows.append(("B200 (1 GPU, device-resident block container)", kk, c, d, f"{blk >> 10} KiB blocks"))
        gb.append({"name": f"BM_CompressBiasedDeviceBlocks<::hufb200::BlockCodec<{kk}>>", "bytes_per_second": c})
        gb.append({"name": f"BM_DecompressBiasedDeviceBlocks<::hufb200::BlockCodec<{kk}>>", "bytes_per_second": d})
    # BASELINE config 1: the reference's own benchmark unit, one 100 KiB buffer per call
    # ... and the same through the reference's file benchmark (BM_CompressFile / BM_DecompressFile,
    # codec/huffman_benchmark.cpp:218-248) on real text
    c1_rows, file_rows = [], []
    for key, bm, dest in (("config1", "Biased", c1_rows), ("config1_file", "File", file_rows)):
        for name, r in cpu.get(key, {}).get("rows", {}).items():
            ns = "::hufb200::" if name.startswith("HuffmanCompressorB200") else "::huffman::"
            gb.append({"name": f"BM_Compress{bm}<{ns}{name}>", "bytes_per_second": r["compress_MiBps"] * 2 ** 20,
                       "real_time": r["compress_us_per_call"], "time_unit": "us"})
            gb.append({"name": f"BM_Decompress{bm}<{ns}{name}>", "bytes_per_second": r["decompress_MiBps"] * 2 ** 20,
                       "real_time": r["decompress_us_per_call"], "time_unit": "us"})
            dest.append((name, r))
    rows.append(("Huff0", 4, None, None, cpu.get("huff0", "unavailable")))
    if args.json:
        json.dump({"context": {"source": args.bench_json, "metric": bench["metric"]}, "benchmarks": gb},
                  open(args.json, "w"), indent=1)
    mib = lambda x: "n/a" if x is None else f"{int(x / 2 ** 20)} MiB/s"
    print("Method | Streams | Compress | Decompress | Where")
    print("-------|---|---|---|---")
    for m, s, c, d, w in rows:
        print(f"{m} | {s} | {mib(c)} | {mib(d)} | {w}")
    for title, table in (
            ("Config 1: ONE 100 KiB biased buffer per call (codec/huffman_benchmark.cpp:61-81), host pointers, one thread",
             c1_rows),
            ("File benchmark: ONE 100 KiB buffer of real English text per call (codec/huffman_benchmark.cpp:218-248; "
             "tests/golden/real_text_100k.bin stands in for enwik8), host pointers, one thread", file_rows)):
        if not table:
            continue
        print()
        print(title)
        print()
        print("Compressor | Compress | us/call | Decompress | us/call | compressed bytes")
        print("-----------|---|---|---|---|---")
        for name, r in table:
            print(f"{name} | {int(r['compress_MiBps'])} MiB/s | {r['compress_us_per_call']:.1f} | "
                  f"{int(r['decompress_MiBps'])} MiB/s | {r['decompress_us_per_call']:.1f} | "
                  f"{r.get('compressed_bytes', '')}")


if __name__ == "__main__":
    main()
