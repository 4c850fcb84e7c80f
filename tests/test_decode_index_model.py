"""CPU: a numpy model of how k_decompress_blocks addresses its decode table (DESIGN.md section 4.4;
huffman-avx512_b200/csrc/huf_kernels.cu, build_dtable and the lookup loop), checked exhaustively
over every stream window for the length tables of all test inputs.

The table has a BITS-bit first level and behind it one entry per longer code, length by length.
The kernel reads entry  max over l = BITS..12 of (p_l + c_l)  with p_l = the window's first l bits,
and forms the first three levels from ONE shifted window W = window >> (30 - BITS): byte addresses
W + b0, 2W + b1, 4W + b2, whose two lowest bits are junk and are cleared after the max.  What must
hold, for every window that starts with a code of the table (every window, since the code is
complete):
  * a code of at most BITS bits resolves to first-level entry p_BITS;
  * a longer code of length l resolves to the entry of exactly that code;
  * the junk bits never change the winner.
The lengths come from the CPU checker's MakeCanonicalCoding restatement (codec/huffman.cpp:339-437)."""
import numpy as np

from _cases import biased, english, extra_cases, reference_test_cases

MAXLEN = 12


def _tables(oracle):
    rng = np.random.default_rng(5)
    hists = [oracle.histogram(d) for _, d in reference_test_cases()[:12] + extra_cases() if len(d) > 1]
    hists.append(oracle.histogram(biased(200000, seed=1, p=0.02)))      # many 12-bit codes
    hists.append(oracle.histogram(english(100000, seed=2)))
    hists.append(np.array([1 << min(i, 30) for i in range(256)], dtype=np.uint32))  # length-limited
    fib = [1, 1]
    while len(fib) < 40:
        fib.append(fib[-1] + fib[-2])
    hists.append(np.array(fib + [0] * 216, dtype=np.uint32))
    for _ in range(40):                                                   # random shapes
        n = int(rng.integers(2, 257))
        h = np.zeros(256, dtype=np.uint32)
        h[rng.permutation(256)[:n]] = np.maximum(1, (rng.pareto(0.7, n) * 10).astype(np.int64)).astype(np.uint32)
        hists.append(h)
    out = []
    for h in hists:
        cd = oracle.make_coding(h)
        if cd["num_syms"] > 1:
            out.append(np.array(cd["len_count"][:MAXLEN + 1], dtype=np.int64))
    return out


def _model(len_count, bits):
    """(exact entry index, entry index as the kernel forms it, expected entry) for every 13-bit window."""
    # left-aligned 12-bit end of the code range of each length (parse_header: code_end)
    code_end = np.cumsum(len_count << (MAXLEN - np.arange(MAXLEN + 1)))
    assert code_end[MAXLEN] == 1 << MAXLEN, "complete code (Kraft sum 1)"
    w13 = np.arange(1 << 13, dtype=np.int64)       # the window's first 13 bits
    w12 = w13 >> 1
    # c_l: level l reads entry p_l + c_l; level BITS is the first-level table itself (c = 0)
    c = {bits: 0}
    base = int(code_end[bits] >> (MAXLEN - bits))  # P: first prefix no code of at most BITS bits owns
    for l in range(bits + 1, MAXLEN + 1):
        first = int(code_end[l - 1] >> (MAXLEN - l))
        c[l] = base - first
        base += int(len_count[l])
    exact = np.full(w13.shape, -(1 << 30), dtype=np.int64)
    for l in range(bits, MAXLEN + 1):
        exact = np.maximum(exact, (w12 >> (MAXLEN - l)) + c[l])
    # the kernel's form on byte addresses (table at address 0)
    W = w13 >> (13 - (bits + 2))                   # window >> (30 - BITS): BITS + 2 bits
    ea = W + 4 * c[bits]
    if bits + 1 <= MAXLEN:
        ea = np.maximum(ea, 2 * W + 4 * c[bits + 1])
    if bits + 2 <= MAXLEN:
        ea = np.maximum(ea, 4 * W + 4 * c[bits + 2])
    for l in range(bits + 3, MAXLEN + 1):
        ea = np.maximum(ea, 4 * ((w12 >> (MAXLEN - l)) + c[l]))
    kernel = (ea & ~3) >> 2
    # what the entry must be: the code the window starts with
    length = np.searchsorted(code_end, w12, side="right")  # first l with w12 < code_end[l]
    expect = np.where(length <= bits, w12 >> (MAXLEN - bits), 0)
    base = int(code_end[bits] >> (MAXLEN - bits))
    for l in range(bits + 1, MAXLEN + 1):
        first = int(code_end[l - 1] >> (MAXLEN - l))
        expect = np.where(length == l, base + (w12 >> (MAXLEN - l)) - first, expect)
        base += int(len_count[l])
    return exact, kernel, expect, base


def test_one_window_index_equals_the_exact_index_and_hits_the_code(oracle):
    n = 0
    for lc in _tables(oracle):
        for bits in (9, 10, 11):
            exact, kernel, expect, used = _model(lc, bits)
            assert np.array_equal(exact, expect), (lc.tolist(), bits)
            assert np.array_equal(kernel, expect), (lc.tolist(), bits)
            # the table stays inside what dec_entries() reserves
            assert used <= (1 << bits) + (128 if bits == 11 else 256), (lc.tolist(), bits)
            assert int(kernel.min()) >= 0
            n += 1
    assert n >= 150
