"""CPU: the plain-C oracle against the committed golden vectors (generated from the unmodified
reference by tests/golden/make_golden.py) and the three known answers of SURVEY.md section 3.1."""
import hashlib
import json
import os

import numpy as np

from _cases import GOLDEN, KS, extra_cases, reference_test_cases

ALL = dict(reference_test_cases() + extra_cases())
VEC = json.load(open(os.path.join(GOLDEN, "vectors.json")))


def test_known_answers(oracle):
    hello2 = ("0b0000001c000000010502" "6c6f20485764657" "20b000000" "0000000000000000" "80099c"
              "0000000000000000" "ccab")
    assert oracle.compress(2, b"Hello World").hex() == hello2
    assert oracle.compress(1, b"AAA").hex() == "0300000001000000" "0141" + "00" * 8
    empty4 = oracle.compress(4, b"")
    assert len(empty4) == 52 and empty4[:8] == bytes(8)
    assert np.frombuffer(empty4[8:20], dtype="<u4").tolist() == [8, 16, 24]
    assert empty4[20:] == bytes(32)


def test_compress_vectors(oracle):
    n = 0
    for key, ent in VEC["compress"].items():
        name, ks = key.split("/K")
        got = oracle.compress(int(ks), ALL[name])
        assert len(got) == ent["comp_len"], key
        assert hashlib.sha256(got).hexdigest() == ent["sha256"], key
        if "hex" in ent:
            assert got.hex() == ent["hex"], key
        assert oracle.decompress(int(ks), got) == ALL[name], key
        n += 1
    assert n == len(VEC["compress"]) and n > 150


def test_table_vectors(oracle):
    for name, ent in VEC["tables"].items():
        cd = oracle.make_coding(oracle.histogram(ALL[name]))
        assert [int(x) for x in cd["len_count"]] == ent["len_count"], name
        assert cd["sorted_syms"].hex() == ent["sorted_syms"], name
        assert cd["len_mask"] == ent["len_mask"], name
        if "dtable2x_sha256" in ent:
            t = oracle.dtable(2, cd["len_count"], cd["sorted_syms"])
            assert hashlib.sha256(t.tobytes()).hexdigest() == ent["dtable2x_sha256"], name


def test_config1_sizes(oracle):
    # biased 100 KiB, K=32 -> 47142 bytes, K=4 -> 46792; uniform -> 103045
    from _cases import golden
    b = golden("proba02_100k.bin")
    assert len(oracle.compress(32, b)) == 47142
    assert len(oracle.compress(4, b)) == 46792
    assert len(oracle.compress(32, golden("uniform_100k.bin"))) == 103045


def test_ragged_and_edge_roundtrips(oracle):
    rng = np.random.default_rng(5)
    for n in list(range(0, 70)) + [255, 256, 257, 1000, 4095, 4096, 4097]:
        data = bytes(rng.integers(0, 1 + n % 7 * 40, n, dtype=np.uint8)) if n else b""
        for k in KS + (3, 64):
            assert oracle.decompress(k, oracle.compress(k, data)) == data, (n, k)
