"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the
same inputs -- bit-exact, both directions.  Re-hosts the reference's CompressorTest /
HistogramTest suites (codec/huffman_test.cpp:47-184, codec/histogram_test.cpp:13-52) with
round-trip checks upgraded to byte equality of the compressed image and cross-decoding."""
import hashlib
import json
import os

import numpy as np
import pytest

from _cases import GOLDEN, KS, extra_cases, golden, reference_test_cases, biased, english

pytestmark = pytest.mark.gpu

ALL_CASES = reference_test_cases() + extra_cases()


@pytest.mark.parametrize("k", KS)
def test_compress_matches_oracle_and_cross_decodes(huf, oracle, k):
    for name, data in ALL_CASES:
        want = oracle.compress(k, data)
        got = huf.compress(k, data)
        assert got == want, f"{name} K={k}: compressed image differs ({len(got)} vs {len(want)} bytes)"
        assert huf.decompress(k, want) == data, f"{name} K={k}: GPU decode of oracle stream"
        assert oracle.decompress(k, got) == data, f"{name} K={k}: oracle decode of GPU stream"


def test_compress_matches_golden_vectors(huf):
    vec = json.load(open(os.path.join(GOLDEN, "vectors.json")))["compress"]
    cases = dict(ALL_CASES)
    n = 0
    for key, ent in vec.items():
        name, ks = key.split("/K")
        got = huf.compress(int(ks), cases[name])
        assert len(got) == ent["comp_len"], key
        assert hashlib.sha256(got).hexdigest() == ent["sha256"], key
        if "hex" in ent:
            assert got.hex() == ent["hex"], key
        n += 1
    assert n > 150


def test_compress_matches_reference_build(huf, ref):
    """Against the unmodified reference where it was built (skipped if oracle/_ref is absent)."""
    for name, data in ALL_CASES[:12] + extra_cases():
        for k in (4, 32, 48):
            assert huf.compress(k, data) == ref.compress(k, data), (name, k)
            for variant in (ref.GATHER, ref.PERMUTE) if k % 8 == 0 else ():
                assert huf.decompress(k, ref.compress(k, data, variant)) == data
                assert ref.decompress(k, huf.compress(k, data), variant) == data


def test_odd_stream_counts(huf, oracle):
    for k in (3, 5, 7, 24, 40, 64):
        for name, data in extra_cases():
            got = huf.compress(k, data)
            assert got == oracle.compress(k, data), (name, k)
            assert huf.decompress(k, got) == data, (name, k)


def test_histogram(huf, oracle):
    # HistogramTest.ShortSanity / Long (codec/histogram_test.cpp:18-42)
    h = huf.MakeHistogram(b"foobar")
    assert (h[ord("f")], h[ord("o")], h[ord("b")], h[ord("a")], h[ord("r")], h[ord("q")]) == (1, 2, 1, 1, 1, 0)
    for data in (golden("long_random.bin")[:2007], golden("hist_biased_3125.bin"), golden("uniform_100k.bin"),
                 b"", b"z", biased(1 << 20, seed=9), bytes(1 << 16), english(333_333, seed=2)):
        assert np.array_equal(huf.MakeHistogram(data), oracle.histogram(data))
    # unaligned views
    buf = np.frombuffer(biased(300_001, seed=4), dtype=np.uint8)
    for off in (1, 3, 7, 15):
        assert np.array_equal(huf.MakeHistogram(buf[off:]), oracle.histogram(buf[off:].tobytes()))


def test_table_build_matches_oracle(huf, oracle):
    rng = np.random.default_rng(7)
    hists = [oracle.histogram(d) for _, d in ALL_CASES[:10] + extra_cases()]
    # tie-heavy histograms: the order of equal counts must follow libstdc++'s introsort
    for n in (1, 2, 3, 15, 16, 17, 33, 64, 100, 200, 256):
        for hi in (1, 2, 3, 10):
            h = np.zeros(256, dtype=np.uint32)
            idx = rng.permutation(256)[:n]
            h[idx] = rng.integers(1, hi + 1, n)
            hists.append(h)
    # median-of-three killer style patterns and long-code histograms
    h = np.zeros(256, dtype=np.uint32)
    h[:] = np.arange(256, 0, -1)
    hists.append(h)
    h = np.zeros(256, dtype=np.uint32)
    h[:40] = [min(2 ** i, 2 ** 31) for i in range(40)]
    hists.append(h)
    for h in hists:
        want = oracle.make_coding(h)
        got = huf.make_table(h)
        assert got["num_syms"] == want["num_syms"]
        assert got["len_mask"] == want["len_mask"]
        assert np.array_equal(got["len_count"], want["len_count"])
        assert got["sorted_syms"] == want["sorted_syms"]
        assert np.array_equal(got["code_len"], want["code_len"])
        assert np.array_equal(got["code_bits"], want["code_bits"])


def test_decode_table_matches_oracle(huf, oracle):
    for _, data in ALL_CASES[:10] + extra_cases():
        if not data:
            continue
        cd = oracle.make_coding(oracle.histogram(data))
        want = oracle.dtable(2, cd["len_count"], cd["sorted_syms"])
        got = huf.decode_table(cd["len_count"], cd["sorted_syms"])
        assert np.array_equal(got, want)


def test_decode_table1x_matches_oracle(huf, oracle):
    """Decoder1x (codec/huffman.cpp:594-632): {code_len, sym} per 12-bit prefix, from the kernel's builder."""
    for _, data in ALL_CASES[:10] + extra_cases():
        if not data:
            continue
        cd = oracle.make_coding(oracle.histogram(data))
        want = oracle.dtable(1, cd["len_count"], cd["sorted_syms"])
        got = huf.decode_table1x(cd["len_count"], cd["sorted_syms"])
        assert np.array_equal(got, want)


def test_compress_with_table(huf, oracle):
    data = golden("proba02_100k.bin")
    other = biased(50_000, seed=11)
    cd = oracle.make_coding(oracle.histogram(data))  # table from a different buffer
    for k in (4, 32):
        want = oracle.compress_with_table(k, other, cd["len_count"], cd["sorted_syms"])
        got = huf.compress_with_table(k, other, cd["len_count"], cd["sorted_syms"])
        assert got == want
        assert huf.decompress(k, got) == other
    # a symbol without a code must be reported, not encoded
    with pytest.raises(huf.HufError) as ei:
        huf.compress_with_table(4, b"\xff" * 100, cd["len_count"], cd["sorted_syms"])
    assert ei.value.code == -4


def test_error_paths(huf):
    with pytest.raises(huf.HufError) as ei:
        huf.compress(0, b"abc")
    assert ei.value.code == -1
    with pytest.raises(huf.HufError) as ei:
        huf.compress(65, b"abc")
    assert ei.value.code == -1
    good = huf.compress(4, b"hello hello hello")
    with pytest.raises(huf.HufError):
        huf.decompress(4, good[:6])
    bad = bytearray(good)
    bad[4] = 0xff
    bad[5] = 0xff  # length mask with bits above 12
    with pytest.raises(huf.HufError) as ei:
        huf.decompress(4, bytes(bad))
    assert ei.value.code == -4
    # truncated payload must not crash the device
    try:
        huf.decompress(4, good[:-9])
    except huf.HufError:
        pass
    assert huf.decompress(4, good) == b"hello hello hello"


def test_block_container_roundtrip(huf, oracle):
    data = biased(1_000_000, seed=21)
    for k, bs in ((32, 131072), (4, 16384), (48, 65536), (8, 1 << 20)):
        cont = huf.compress_blocks(k, bs, data)
        assert huf.decompress_blocks(cont) == data
        # every block is an independent reference-format buffer
        nb = (len(data) + bs - 1) // bs
        sizes = np.frombuffer(cont[32: 32 + 4 * nb], dtype="<u4")
        pos = 32 + 4 * nb
        for b in range(nb):
            blk = cont[pos: pos + int(sizes[b])]
            pos += int(sizes[b])
            raw_blk = data[b * bs: (b + 1) * bs]
            if b in (0, nb - 1) or b % 3 == 0:
                assert blk == oracle.compress(k, raw_blk), (k, bs, b)
        assert pos == len(cont)
    assert huf.decompress_blocks(huf.compress_blocks(32, 131072, b"")) == b""


def test_small_block_batches_equal_the_reference(huf, oracle):
    """Blocks of at most 32 KiB go through k_compress_small_blocks (four blocks per CTA and
    iteration, their tables built side by side): every block, incl. a ragged last one in a
    batch that is not full, must equal the reference's bytes."""
    for k, bs, nblk in ((8, 16384, 37), (32, 32768, 22), (16, 4096, 131), (1, 8192, 9)):
        data = biased(nblk * bs - 777, seed=100 + k) if k != 16 else english(nblk * bs - 5, seed=3)
        cont = huf.compress_blocks(k, bs, data)
        assert huf.decompress_blocks(cont) == data
        nb = (len(data) + bs - 1) // bs
        assert nb == nblk
        sizes = np.frombuffer(cont[32: 32 + 4 * nb], dtype="<u4")
        pos = 32 + 4 * nb
        for b in range(nb):
            blk = cont[pos: pos + int(sizes[b])]
            pos += int(sizes[b])
            assert blk == oracle.compress(k, data[b * bs: (b + 1) * bs]), (k, bs, b)
        assert pos == len(cont)


def test_block_sizes_off_the_16_byte_grid(huf, oracle):
    """The container takes any block size (the reference's buffers have no alignment either): blocks
    then start off a 16-byte boundary in the raw input and in the decoder's output."""
    for k, bs, n in ((32, 1000, 17_003), (8, 4099, 50_000), (48, 70_001, 300_000), (4, 33, 1_000), (16, 131_073, 400_000)):
        data = biased(n, seed=bs)
        cont = huf.compress_blocks(k, bs, data)
        assert huf.decompress_blocks(cont) == data, (k, bs)
        nb = (len(data) + bs - 1) // bs
        sizes = np.frombuffer(cont[32: 32 + 4 * nb], dtype="<u4")
        pos = 32 + 4 * nb
        for b in range(nb):
            blk = cont[pos: pos + int(sizes[b])]
            pos += int(sizes[b])
            assert blk == oracle.compress(k, data[b * bs: (b + 1) * bs]), (k, bs, b)
        assert pos == len(cont)


def test_staged_encoder_overflow_fallback(huf, oracle):
    """One slice made only of rare symbols (12-bit codes, > 10 bits/symbol) overflows the per-warp
    staging buffer of the staged encoder and must take the ring path; bytes still equal the oracle."""
    rng = np.random.default_rng(99)
    k, bs = 32, 131072
    sl = bs // k
    blk = np.full(bs, ord("a"), dtype=np.uint8)
    blk[::7] = ord("b")
    blk[::11] = ord("c")
    for s in (5, 17, 31):  # three slices of rare symbols
        blk[s * sl:(s + 1) * sl] = rng.integers(40, 256, sl, dtype=np.uint8)
    data = blk.tobytes() * 2 + bytes(rng.integers(0, 256, 1000, dtype=np.uint8))
    want0 = oracle.compress(k, data[:bs])
    sizes = np.frombuffer(want0[8 + 13:], dtype=np.uint8)  # not used, just make sure it parses
    assert huf.compress(k, data[:bs]) == want0
    cont = huf.compress_blocks(k, bs, data)
    assert huf.decompress_blocks(cont) == data
    nb = 3
    idx = np.frombuffer(cont[32:32 + 4 * nb], dtype="<u4")
    assert cont[32 + 4 * nb: 32 + 4 * nb + int(idx[0])] == want0
    # the rare slices really are longer than 10 bits/symbol
    cd = oracle.make_coding(oracle.histogram(data[:bs]))
    bits = int(cd["code_len"][blk[5 * sl:6 * sl]].sum())
    assert bits > 10 * sl


def test_decoder_survives_corrupt_input(huf):
    """Hardened decoder: corrupted buffers give an error code or bounded
    garbage -- never a fault, a hang or an out-of-bounds write -- and the device stays usable."""
    rng = np.random.default_rng(2024)
    good_data = biased(200_000, seed=5)
    for k in (4, 32, 48):
        good = bytearray(huf.compress(k, good_data))
        for trial in range(40):
            bad = bytearray(good)
            kind = trial % 4
            if kind == 0:    # flip bytes in the header / table / end offsets
                for _ in range(3):
                    bad[int(rng.integers(0, 8 + 13 + 60 + 4 * k))] ^= int(rng.integers(1, 256))
            elif kind == 1:  # flip payload bytes
                for _ in range(20):
                    bad[int(rng.integers(0, len(bad)))] ^= int(rng.integers(1, 256))
            elif kind == 2:  # truncate
                bad = bad[: int(rng.integers(1, len(bad)))]
            else:            # end offsets beyond the payload
                pos = 8 + 13 + 40
                bad[pos: pos + 4] = (0xfffffff0).to_bytes(4, "little")
            try:
                out = huf.decompress(k, bytes(bad))
                assert len(out) == int.from_bytes(bytes(bad[:4]), "little")
            except huf.HufError as e:
                assert e.code in (-1, -2, -4)
        assert huf.decompress(k, bytes(good)) == good_data
    # same through the block container
    cont = bytearray(huf.compress_blocks(32, 65536, good_data))
    for trial in range(20):
        bad = bytearray(cont)
        for _ in range(10):
            bad[int(rng.integers(32, len(bad)))] ^= int(rng.integers(1, 256))
        try:
            huf.decompress_blocks(bytes(bad))
        except huf.HufError as e:
            assert e.code in (-1, -2, -4)
    assert huf.decompress_blocks(bytes(cont)) == good_data


def test_stream_longer_than_the_position_window(huf, oracle):
    """One stream of more than 2^26 symbols: the decoder keeps positions in 26 bits and has to
    move its window along (found by a 70 MiB single-stream round trip)."""
    from _cases import biased
    n = (65 << 20) + 4321
    blk = biased(1 << 20, seed=5)
    data = (blk * ((n >> 20) + 1))[:n]
    comp = huf.compress(1, data)
    assert comp == oracle.compress(1, data)
    assert huf.decompress(1, comp) == data


def test_table_build_sort_order_stress(huf, oracle):
    """The warp-parallel restatement of libstdc++'s introsort partition phase against the oracle:
    tie-heavy random histograms and orders that force deep recursion / the heapsort fallback."""
    rng = np.random.default_rng(11)
    hists = []
    for trial in range(400):
        n = int(rng.integers(1, 257))
        hi = int(rng.choice([1, 2, 3, 5, 20, 1000]))
        h = np.zeros(256, dtype=np.uint32)
        h[rng.permutation(256)[:n]] = rng.integers(1, hi + 1, n)
        hists.append(h)
    for n in (64, 128, 200, 256):
        for pattern in range(4):
            h = np.zeros(256, dtype=np.uint32)
            idx = np.arange(n)
            if pattern == 0:
                vals = np.where(idx % 2 == 0, idx + 1, n + idx)
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
            h[:n] = vals[:n]
            hists.append(h)
    # similar weights (what incompressible input looks like): the batched form of the two-queue
    # merge; with counts near 2^31 the u32 node weights wrap like the reference's
    for n in (32, 33, 64, 100, 255, 256):
        for base, spread in ((500, 40), (3, 2), (1 << 20, 1 << 18), ((1 << 31) - 100, 90)):
            h = np.zeros(256, dtype=np.uint32)
            h[rng.permutation(256)[:n]] = (base + rng.integers(-spread, spread + 1, n)).astype(np.uint32)
            hists.append(h)
    for t, h in enumerate(hists):
        want = oracle.make_coding(h)
        got = huf.make_table(h)
        assert got["sorted_syms"] == want["sorted_syms"], t
        assert np.array_equal(got["len_count"], want["len_count"]), t
        assert np.array_equal(got["code_bits"], want["code_bits"]), t


def test_staged_mode_for_long_slices_and_its_fallback(huf, oracle):
    """Slices of 8192 symbols are staged when the table's mean code length says the streams fit
    their staging buffers; a stream that does not fit after all takes the ring path, and a block
    whose mean says 'no' is not staged at all.  Bytes equal the oracle either way."""
    rng = np.random.default_rng(5)
    k, bs = 16, 131072
    sl = bs // k
    assert sl == 8192
    mostly = np.full(bs, ord("a"), dtype=np.uint8)
    mostly[::5] = ord("b")
    one_hot = mostly.copy()
    one_hot[5 * sl:6 * sl] = rng.integers(0, 256, sl, dtype=np.uint8)  # one incompressible slice
    noise = rng.integers(0, 256, bs, dtype=np.uint8)                    # mean code length 8: ring path
    for blk in (mostly, one_hot, noise):
        data = blk.tobytes()
        want = oracle.compress(k, data)
        assert huf.compress(k, data) == want
    data = mostly.tobytes() + one_hot.tobytes() + noise.tobytes() + one_hot.tobytes()[:70001]
    assert huf.decompress_blocks(huf.compress_blocks(k, bs, data)) == data
    # the incompressible slice really overflows a staging buffer (1276 words) while the block's mean is low
    cd = oracle.make_coding(oracle.histogram(one_hot.tobytes()))
    assert int(cd["code_len"][one_hot[5 * sl:6 * sl]].sum()) > 1276 * 32
    assert int(cd["code_len"][one_hot].sum()) / bs * sl < 1276 * 28


def test_corruption_is_reported_not_decoded(huf):
    """Detectable corruption must come back as HUFB200_E_CORRUPT, not as garbage with status OK:
    an over- or under-subscribed length table (Kraft sum != 1), a payload whose codes do not end
    where the stream ends, end offsets that are not cumulative."""
    data = biased(300_000, seed=9)
    for k in (4, 32):
        good = bytearray(huf.compress(k, data))
        mask = int.from_bytes(good[4:8], "little")
        npop = bin(mask).count("1")
        # (a) length table: one code more / one code less of some length
        for delta in (+1, -1):
            bad = bytearray(good)
            bad[8 + npop - 1] = (bad[8 + npop - 1] + delta) & 0xff
            with pytest.raises(huf.HufError) as ei:
                huf.decompress(k, bytes(bad))
            assert ei.value.code == -4
        # (b) payload: a flipped byte garbles a few symbols, then a Huffman decoder falls back into
        #     step; it is caught whenever the garbled part decodes to a different NUMBER of symbols
        #     (the stream then does not end on its last byte).  The format has no checksum, so
        #     flips that keep the count are undetectable -- here, by the reference, by anyone.
        hits = 0
        rng = np.random.default_rng(k)
        for trial in range(30):
            bad = bytearray(good)
            pos = int(rng.integers(len(good) // 2, len(good) - 64))
            bad[pos] ^= 0x55
            try:
                huf.decompress(k, bytes(bad))  # (a flip in slop or padding bits changes nothing)
            except huf.HufError as e:
                assert e.code == -4
                hits += 1
        assert hits >= 1, hits
        # (c) end offsets: swap two neighbours (no longer cumulative)
        if k > 2:
            nsyms = sum(good[8: 8 + npop]) or 256
            ends = 8 + npop + nsyms
            bad = bytearray(good)
            bad[ends: ends + 4], bad[ends + 4: ends + 8] = good[ends + 4: ends + 8], good[ends: ends + 4]
            with pytest.raises(huf.HufError) as ei:
                huf.decompress(k, bytes(bad))
            assert ei.value.code == -4
        assert huf.decompress(k, bytes(good)) == data


def test_corrupt_chunk_with_pinned_buffers_leaves_nothing_in_flight(huf):
    """A container of several pipeline chunks whose FIRST chunk is corrupt, decoded into pinned
    memory: the call must return E_CORRUPT only after every copy it queued against the caller's
    buffers has finished (the output buffer is overwritten right after the call; a later decode
    must be unaffected)."""
    import ctypes as C
    import torch
    data = biased(100 << 20, seed=21)  # 100 MiB -> four 32 MiB chunks
    k, bs = 32, 131072
    cont = huf.compress_blocks(k, bs, data)
    L = huf.load()
    host_c = torch.frombuffer(bytearray(cont), dtype=torch.uint8).pin_memory()
    host_o = torch.empty(len(data), dtype=torch.uint8).pin_memory()
    olen = C.c_size_t(0)
    info = huf.binding.container_info(np.frombuffer(cont, dtype=np.uint8))
    first_payload = 32 + 4 * info["n_blocks"]
    saved = host_c[first_payload + 4: first_payload + 8].clone()
    host_c[first_payload + 4: first_payload + 8] = torch.tensor([0xff, 0xff, 0x00, 0x00], dtype=torch.uint8)  # len_mask
    rc = L.hufb200_decompress_blocks(C.c_void_p(host_c.data_ptr()), len(cont), C.c_void_p(host_o.data_ptr()),
                                     len(data), C.byref(olen))
    assert rc == -4
    host_o.fill_(0xAB)  # would race with a copy still in flight
    torch.cuda.synchronize()
    assert bool((host_o == 0xAB).all())
    host_c[first_payload + 4: first_payload + 8] = saved
    rc = L.hufb200_decompress_blocks(C.c_void_p(host_c.data_ptr()), len(cont), C.c_void_p(host_o.data_ptr()),
                                     len(data), C.byref(olen))
    assert rc == 0 and olen.value == len(data)
    assert host_o.numpy().tobytes() == data


def test_cpp_policy_class_runs_on_the_device(huf, oracle):
    """include/hufb200.hpp's HuffmanCompressorB200<K> -- the template argument the reference's
    TYPED_TEST_SUITE / DEFINE_BENCHMARKS take (codec/huffman_test.cpp:47-54,
    codec/huffman_benchmark.cpp:252-281) -- built with g++ and RUN on the device over the
    reference's CompressorTest inputs: Decompress(Compress(x)) == x inside the program, and the
    bytes it returns equal the oracle's."""
    import struct
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = r"""
#include "hufb200.hpp"
#include <cstdio>
#include <fstream>
#include <iterator>
#include <vector>
template <int K> int run(const std::vector<std::string>& in, std::vector<std::string>* out) {
  using T = hufb200::HuffmanCompressorB200<K>;
  for (const std::string& raw : in) {
    std::string c = T::Compress(raw);
    if (T::Decompress(c) != raw) { std::printf("%s: round trip differs\n", T::name().c_str()); return 1; }
    out->push_back(c);
  }
  return 0;
}
int main(int argc, char** argv) {
  const int k = std::atoi(argv[1]);
  std::ifstream f(argv[2], std::ios::binary);
  std::string all((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::vector<std::string> in, out;
  for (size_t p = 0; p + 4 <= all.size();) {
    uint32_t n; std::memcpy(&n, all.data() + p, 4); p += 4;
    in.emplace_back(all.substr(p, n)); p += n;
  }
  int rc = 3;
  try {
    switch (k) {
      case 1: rc = run<1>(in, &out); break;
      case 4: rc = run<4>(in, &out); break;
      case 8: rc = run<8>(in, &out); break;
      case 16: rc = run<16>(in, &out); break;
      case 32: rc = run<32>(in, &out); break;
      case 48: rc = run<48>(in, &out); break;
    }
    hufb200::ByteHistogram h = hufb200::MakeHistogram(in.empty() ? std::string() : in[1]);
    unsigned long long tot = 0; for (uint32_t c : h) tot += c;
    if (!in.empty() && tot != in[1].size()) rc |= 4;
  } catch (const std::exception& e) { std::printf("exception: %s\n", e.what()); return 2; }
  std::ofstream o(argv[3], std::ios::binary);
  for (const std::string& c : out) { uint32_t n = (uint32_t)c.size(); o.write((const char*)&n, 4); o.write(c.data(), n); }
  return rc;
}
"""
    libdir = os.path.dirname(huf.lib_path())
    exe = "/tmp/hufb200_policy_gpu"
    with open(exe + ".cpp", "w") as f:
        f.write("#include <cstring>\n#include <cstdlib>\n" + src)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"), exe + ".cpp", "-o", exe,
                        "-L", libdir, "-lhufb200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cases = [d for _, d in reference_test_cases()[:12] + extra_cases()]
    with open(exe + ".in", "wb") as f:
        for d in cases:
            f.write(struct.pack("<I", len(d)) + d)
    for k in (1, 4, 8, 16, 32, 48):
        rc = subprocess.run([exe, str(k), exe + ".in", exe + ".out"], capture_output=True, text=True)
        assert rc.returncode == 0, (k, rc.stdout, rc.stderr)
        blob = open(exe + ".out", "rb").read()
        p = 0
        for d in cases:
            (n,) = struct.unpack_from("<I", blob, p)
            assert blob[p + 4: p + 4 + n] == oracle.compress(k, d), (k, len(d))
            p += 4 + n
        assert p == len(blob)


def test_large_single_buffer_spread_over_the_device(huf, oracle):
    """hufb200_compress of ONE buffer of 256 KiB or more runs per-stream histograms, one plan and
    pieces of 3072 symbols on many CTAs (launch_compress_single): the bytes must still be exactly
    CompressMulti<K>'s -- ragged sizes, every K class, text and incompressible input, a supplied
    table, and a supplied table that lacks a symbol."""
    rng = np.random.default_rng(99)
    cases = [
        ("biased256k", biased(256 << 10, seed=1)),
        ("biased1m+", biased((1 << 20) + 12345, seed=2)),
        ("english3m", english(3_000_001, seed=3)),
        ("uniform700k", bytes(rng.integers(0, 256, 700_003, dtype=np.uint8))),
        ("single_symbol", b"z" * 400_000),
        ("two_symbols", (b"ab" * 200_000) + b"a"),
        ("long_codes", b"".join(bytes([65 + i]) * (1 << i) for i in range(19))),
    ]
    for name, data in cases:
        for k in (1, 4, 8, 32, 48, 5):
            want = oracle.compress(k, data)
            got = huf.compress(k, data)
            assert got == want, f"{name} K={k}: {len(got)} vs {len(want)} bytes"
            assert huf.decompress(k, got) == data, (name, k)
    data = biased(2_000_003, seed=8)
    # a table that is not the buffer's own but has a code for each of its symbols
    cd = oracle.make_coding(oracle.histogram(data) + 3 * oracle.histogram(english(500_000, seed=9)))
    for k in (4, 32):
        want = oracle.compress_with_table(k, data, cd["len_count"], cd["sorted_syms"])
        assert huf.compress_with_table(k, data, cd["len_count"], cd["sorted_syms"]) == want
    with pytest.raises(huf.HufError) as ei:
        huf.compress_with_table(4, b"\xff" * 300_000, cd["len_count"], cd["sorted_syms"])
    assert ei.value.code == -4
    assert huf.decompress(32, huf.compress_with_table(32, data, cd["len_count"], cd["sorted_syms"])) == data
