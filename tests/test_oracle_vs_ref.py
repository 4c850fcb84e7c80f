"""CPU: pins the oracle restatement against the UNMODIFIED reference build (oracle/_ref).
Skipped where the reference sources are absent and no prebuilt _ref travelled."""
import numpy as np
import pytest

from _cases import KS, biased, english, extra_cases, reference_test_cases

ALL = reference_test_cases() + extra_cases()


def test_generators_match_fixtures(ref):
    from _cases import golden
    assert ref.gen_proba(0.2, 100 << 10) == golden("proba02_100k.bin")
    assert ref.gen_rand(0, 100_000) == golden("long_random.bin")
    assert ref.gen_equal_counts() == golden("equal_counts.bin")


@pytest.mark.parametrize("k", KS)
def test_compress_bytes_equal(oracle, ref, k):
    for name, data in ALL:
        c = ref.compress(k, data)
        assert oracle.compress(k, data) == c, (name, k)
        assert oracle.decompress(k, c) == data, (name, k)
        assert ref.decompress(k, c) == data, (name, k)


def test_avx_paths_equal_scalar(ref):
    # AvxCheckCompressor, codec/huffman_test.cpp:15-32, extended to K in {8,16,32,48}
    for name, data in ALL[:10] + extra_cases()[:4]:
        for k in (8, 16, 32, 48):
            s = ref.compress(k, data)
            assert ref.compress(k, data, ref.GATHER) == s, (name, k)
            assert ref.compress(k, data, ref.PERMUTE) == s, (name, k)
            assert ref.decompress(k, s, ref.GATHER) == data
            assert ref.decompress(k, s, ref.PERMUTE) == data


def test_histogram_variants(oracle, ref):
    from _cases import golden
    for data in (b"foobar", golden("long_random.bin")[:2007], golden("hist_biased_3125.bin"), biased(50_000), b""):
        want = oracle.histogram(data)
        for which in range(5):
            assert np.array_equal(ref.histogram(data, which), want), which


def test_sort_clone_matches_std_sort(oracle, ref):
    """The libstdc++ introsort restatement on tie-heavy inputs."""
    rng = np.random.default_rng(11)
    for trial in range(400):
        n = int(rng.integers(1, 257))
        hi = int(rng.choice([1, 2, 3, 5, 20, 1000]))
        hist = np.zeros(256, dtype=np.uint32)
        syms = np.sort(rng.permutation(256)[:n]).astype(np.uint8)
        hist[syms] = rng.integers(1, hi + 1, n)
        assert oracle.sort_syms(hist, syms.tobytes()) == ref.sort_syms(hist, syms.tobytes()), trial
    # adversarial orders (force deep recursion / the heapsort fallback)
    for n in (64, 128, 200, 256):
        for pattern in range(4):
            hist = np.zeros(256, dtype=np.uint32)
            idx = np.arange(n)
            if pattern == 0:
                vals = np.where(idx % 2 == 0, idx + 1, n + idx)            # organ-pipe like
            elif pattern == 1:
                vals = np.concatenate([np.arange(1, n // 2 + 1), np.arange(n // 2, 0, -1)])[:n]
            elif pattern == 2:
                vals = (idx * 7919) % 13 + 1
            else:  # median-of-3 killer
                vals = np.zeros(n, dtype=np.int64)
                half = n // 2
                for i in range(half):
                    vals[i] = i + 1 if i % 2 == 0 else half + i + (1 if i % 2 else 0)
                    vals[half + i] = 2 * (i + 1)
            hist[:n] = vals[:n]
            syms = bytes(range(n))
            assert oracle.sort_syms(hist, syms) == ref.sort_syms(hist, syms), (n, pattern)


def test_make_coding_and_limit(oracle, ref):
    rng = np.random.default_rng(3)
    hists = [oracle.histogram(d) for _, d in ALL[:10] + extra_cases()]
    for _ in range(200):
        n = int(rng.integers(1, 257))
        h = np.zeros(256, dtype=np.uint32)
        kind = rng.integers(0, 3)
        idx = rng.permutation(256)[:n]
        if kind == 0:
            h[idx] = rng.integers(1, 4, n)
        elif kind == 1:
            # optimal depths stay <= 32: beyond that the reference's CollectCodeLen writes past
            # len_count[33] (codec/huffman.cpp:329-337, undefined behaviour), see DESIGN.md
            h[idx] = (1.5 ** rng.integers(0, 36, n)).astype(np.int64)
        else:
            h[idx] = rng.integers(1, 100000, n)
        hists.append(h)
    for h in hists:
        a, b = oracle.make_coding(h), ref.make_coding(h)
        for key in ("num_syms", "len_mask", "sorted_syms"):
            assert a[key] == b[key], key
        for key in ("len_count", "code_bits", "code_len"):
            assert np.array_equal(a[key], b[key]), key
        if a["num_syms"]:
            for which in (1, 2):
                assert np.array_equal(oracle.dtable(which, a["len_count"], a["sorted_syms"]),
                                      ref.dtable(which, b["len_count"], b["sorted_syms"]))
    for _ in range(100):
        lc = np.zeros(33, dtype=np.uint16)
        # a valid (complete) prefix-code length profile that is deeper than 12
        depth = int(rng.integers(13, 30))
        lc[1:depth] = 1
        lc[depth] = 2
        assert np.array_equal(oracle.limit_code_lengths(lc), ref.limit_code_lengths(lc))


def test_random_sizes_and_k(oracle, ref):
    rng = np.random.default_rng(17)
    for _ in range(60):
        n = int(rng.integers(0, 20000))
        data = [biased(n, seed=int(rng.integers(1 << 30))), english(n, seed=1),
                bytes(rng.integers(0, 256, n, dtype=np.uint8))][int(rng.integers(0, 3))]
        for k in (1, 4, 32, 48):
            assert oracle.compress(k, data) == ref.compress(k, data)
