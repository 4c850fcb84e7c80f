#!/usr/bin/env python
"""Per-source-line summary of an ncu report: instructions executed and stall samples.

    python profiles/ncu_lines.py gpurun_out/prof.ncu-rep k_compress [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", f"regex:{kern}", "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file = ""
    lines = []
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit() and len(r) > 9:
            def g(name):
                try:
                    return int(float(r[hdr.index(name)]))
                except Exception:
                    return 0
            lines.append((g("Instructions Executed"), g("Warp Stall Sampling (All Samples)"),
                          g("Thread Instructions Executed"), cur_file, int(r[0]), r[1].strip()))
    tot_i = sum(l[0] for l in lines)
    tot_s = sum(l[1] for l in lines)
    print(f"total warp-instructions {tot_i}, stall samples {tot_s}")
    print("by instructions:")
    for l in sorted(lines, reverse=True)[:top]:
        print(f"{100 * l[0] / tot_i:5.1f}% inst {100 * l[1] / max(tot_s, 1):5.1f}% smp  {l[3]}:{l[4]}  {l[5][:100]}")
    print("by stall samples:")
    for l in sorted(lines, key=lambda x: -x[1])[:top // 2]:
        print(f"{100 * l[0] / tot_i:5.1f}% inst {100 * l[1] / max(tot_s, 1):5.1f}% smp  {l[3]}:{l[4]}  {l[5][:100]}")


if __name__ == "__main__":
    main()
