"""Input cases shared by the CPU and GPU tests.  The first group restates the inputs of the
reference's own tests (codec/huffman_test.cpp:56-184, codec/histogram_test.cpp:18-42); the
byte strings that depend on libstdc++/glibc generators are stored under tests/golden/ (made by
tests/golden/make_golden.py from the reference build)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

LOREM = b"""
Lorem ipsum dolor sit amet, consectetur adipiscing elit, sed do eiusmod
tempor incididunt ut labore et dolore magna aliqua. Ut enim ad minim
veniam, quis nostrud exercitation ullamco laboris nisi ut aliquip ex ea
commodo consequat. Duis aute irure dolor in reprehenderit in voluptate
velit esse cillum dolore eu fugiat nulla pariatur. Excepteur sint
occaecat cupidatat non proident, sunt in culpa qui officia deserunt
mollit anim id est laborum.
    """


def long_codes(log_size=16):
    """LongCodes, codec/huffman_test.cpp:144-156: 2^i copies of 'A'+i."""
    return b"".join(bytes([ord("A") + i]) * (1 << i) for i in range(log_size))


def golden(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def biased(n, seed=0, p=0.2):
    """Same distribution as GenerateProbaData (codec/huffman_benchmark.cpp:27-36), numpy RNG."""
    rng = np.random.default_rng(seed)
    u = rng.random(n)
    u[u == 0] = 0.5
    return (np.floor(np.log(u) / np.log(1 - p)).astype(np.int64) % 256).astype(np.uint8).tobytes()


def english(n, seed=0):
    """iid letters a-z + space with English frequencies (BASELINE config 4)."""
    freq = np.array([8.167, 1.492, 2.782, 4.253, 12.702, 2.228, 2.015, 6.094, 6.966, 0.153, 0.772, 4.025,
                     2.406, 6.749, 7.507, 1.929, 0.095, 5.987, 6.327, 9.056, 2.758, 0.978, 2.360, 0.150,
                     1.974, 0.074, 21.0])
    syms = np.array(list(range(ord("a"), ord("z") + 1)) + [ord(" ")], dtype=np.uint8)
    rng = np.random.default_rng(seed)
    return syms[rng.choice(27, size=n, p=freq / freq.sum())].tobytes()


def reference_test_cases():
    """(name, bytes) for every input of the reference's CompressorTest suite."""
    cases = [
        ("Hello", b"Hello World"),
        ("LongerText", LOREM),
        ("EqualCounts", golden("equal_counts.bin")),
        ("LongRandom", golden("long_random.bin")),
        ("SingleSymbolAAA", b"AAA"),
        ("SingleSymbol1000a", b"a" * 1000),
        ("LongCodes", long_codes()),
        ("EmptyString", b""),
    ]
    many = golden("many_random.bin")
    lens = np.frombuffer(golden("many_random_lens.bin"), dtype="<i4")
    pos = 0
    for i, l in enumerate(lens):
        cases.append((f"ManyRandom{i}", many[pos: pos + l]))
        pos += int(l)
    return cases


def extra_cases():
    rng = np.random.default_rng(1234)
    return [
        ("Biased100K", golden("proba02_100k.bin")),
        ("Uniform100K", golden("uniform_100k.bin")),
        # real English prose, 100 KiB: stand-in for the reference's enwik8 file benchmark
        # (codec/huffman_benchmark.cpp:218-248); see tests/golden/make_golden.py
        ("RealText100K", golden("real_text_100k.bin")),
        ("Biased70001", biased(70001, seed=3, p=0.05)),
        ("English50000", english(50000, seed=5)),
        ("OneByte", b"x"),
        ("TwoSyms", b"ab" * 777 + b"a"),
        ("Ragged31", bytes(rng.integers(0, 256, 31, dtype=np.uint8))),
        ("Ragged4097", bytes(rng.integers(0, 7, 4097, dtype=np.uint8))),
        ("All256x1", bytes(range(256))),
        ("Fib", b"".join(bytes([i]) * f for i, f in enumerate(
            [1, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2584, 4181, 6765]))),
    ]


KS = (1, 2, 4, 8, 16, 32, 48)
