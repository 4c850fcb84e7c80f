"""Generates the committed fixtures under tests/golden/ from the UNMODIFIED reference
(oracle/_ref/libhufref.so, built by oracle/Makefile from /root/reference).

    python tests/golden/make_golden.py

Inputs: byte strings whose generators depend on libstdc++/glibc (mt19937, rand()).
Outputs: reference CompressMulti<K> images (small ones in full, large ones as SHA-256),
MakeCanonicalCoding results and Decoder2x tables for a few histograms.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _libs import Ref  # noqa: E402


def main():
    r = Ref()
    w = lambda name, data: open(os.path.join(HERE, name), "wb").write(data)
    w("proba02_100k.bin", r.gen_proba(0.2, 100 << 10))
    w("uniform_100k.bin", r.gen_rand(1, 100 << 10))
    w("long_random.bin", r.gen_rand(0, 100_000))
    w("equal_counts.bin", r.gen_equal_counts())
    many = r.gen_many_random()
    w("many_random.bin", b"".join(many))
    w("many_random_lens.bin", np.array([len(m) for m in many], dtype="<i4").tobytes())
    w("hist_biased_3125.bin", r.gen_hist_biased(3125))
    # Real text: the reference's second benchmark table takes the first LEN = 100 KiB of a file
    # named on the command line (enwik8; codec/huffman_benchmark.cpp:38-58, :218-248, README.md:66-68).
    # There is no network for enwik8; the stand-in is real English prose this image ships: the
    # licence texts under /usr/share/common-licenses (verbatim copies are permitted by each of
    # them), concatenated in name order, first 100 KiB.
    lic = "/usr/share/common-licenses"
    names = sorted(n for n in os.listdir(lic) if not os.path.islink(os.path.join(lic, n)))
    text = b"".join(open(os.path.join(lic, n), "rb").read() for n in names)[: 100 << 10]
    assert len(text) == 100 << 10
    w("real_text_100k.bin", text)

    from _cases import reference_test_cases, extra_cases, KS  # after the inputs exist
    vectors = {}
    for name, data in reference_test_cases() + extra_cases():
        if name.startswith("ManyRandom") and int(name[10:]) >= 8:
            continue
        for k in KS:
            comp = r.compress(k, data)
            ent = {"raw_len": len(data), "comp_len": len(comp), "sha256": hashlib.sha256(comp).hexdigest()}
            if len(comp) <= 600:
                ent["hex"] = comp.hex()
            vectors[f"{name}/K{k}"] = ent
    # tables
    tables = {}
    for name, data in reference_test_cases()[:8] + extra_cases():
        hist = r.histogram(data, which=1)
        cd = r.make_coding(hist)
        ent = {"len_count": [int(x) for x in cd["len_count"]], "sorted_syms": cd["sorted_syms"].hex(),
               "len_mask": cd["len_mask"]}
        if cd["num_syms"]:
            ent["dtable2x_sha256"] = hashlib.sha256(
                r.dtable(2, cd["len_count"], cd["sorted_syms"]).tobytes()).hexdigest()
        tables[name] = ent
    json.dump({"compress": vectors, "tables": tables}, open(os.path.join(HERE, "vectors.json"), "w"), indent=1,
              sort_keys=True)
    print("wrote", len(vectors), "compress vectors,", len(tables), "tables")


if __name__ == "__main__":
    main()
