"""Split decode (hufb200_decompress_split_dev): every stream cut into items, one lane per item.
It must give exactly what DecompressMulti<K> gives (codec/huffman.cpp:892-955) -- for codes that
fall into step quickly, for codes that never do (all codes of one length: the serial tail), for a
lone symbol, for empty and ragged slices -- on streams written by the checker's encoder (the
reference's format) as well as on our own."""
import os
import subprocess
import sys

import numpy as np
import pytest

from _cases import LOREM, biased, english, long_codes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _t(a):
    import torch
    return torch.frombuffer(bytearray(a), dtype=torch.uint8).cuda()


def _inputs():
    rng = np.random.default_rng(5)
    return {
        "biased": lambda n: biased(n, seed=n % 97),
        "english": lambda n: english(n, seed=n % 89),
        # 128 equally likely symbols: every code has 7 bits, a decoder started off a code boundary
        # never falls into step -- everything behind the first item takes the serial tail
        "seven_bit": lambda n: rng.integers(0, 128, n, dtype=np.uint8).tobytes(),
        "uniform": lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes(),
        "two_symbols": lambda n: rng.integers(0, 2, n, dtype=np.uint8).tobytes(),
        "lone_symbol": lambda n: b"a" * n,
        "long_codes": lambda n: (long_codes(16) * (n // 65535 + 1))[:n],
        "text": lambda n: (LOREM * (n // len(LOREM) + 1))[:n],
    }


def _pack_host(blocks):
    """Blocks back to back on the device: (comp tensor, offsets, sizes)."""
    import torch
    sizes = np.array([len(b) for b in blocks], dtype=np.int32)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    comp = _t(b"".join(blocks) + b"\0" * 64)
    return comp, torch.from_numpy(offs).cuda(), torch.from_numpy(sizes).cuda()


@pytest.mark.parametrize("name", ["biased", "english", "seven_bit", "uniform", "two_symbols", "lone_symbol",
                                  "long_codes", "text"])
@pytest.mark.parametrize("k,bs,n", [(4, 1 << 18, (1 << 18) * 2 + 12345), (1, 100000, 100000), (32, 1 << 17, 3 * (1 << 17) + 1),
                                    (8, 1 << 20, (1 << 20) + 77), (48, 65536, 65536 * 2 + 5), (3, 4096, 4096 + 17)])
def test_split_decode_of_reference_format_streams(huf, oracle, name, k, bs, n):
    """Streams written by the checker's encoder, decoded by the split path and by one lane per stream."""
    import torch
    data = _inputs()[name](n)
    blocks = [oracle.compress(k, data[o:o + bs]) for o in range(0, n, bs)]
    comp, offs, sizes = _pack_host(blocks)
    codec = huf.BlockCodec(k, bs)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = codec.decompress(comp, offs, sizes, n, status=status, split=True)
    torch.cuda.synchronize()
    assert int(status.item()) == 0, (name, k, bs)
    assert out[:n].cpu().numpy().tobytes() == data, (name, k, bs)
    out2 = codec.decompress(comp, offs, sizes, n, status=status, split=False)
    assert torch.equal(out[:n], out2[:n])


def test_split_decode_small_items_in_a_fresh_process():
    """Items of 64 and 96 bits (HUFB200_SPLIT_BITS, read once per process): item boundaries every few
    codes, so every start/exit/last-item rule is hit thousands of times per stream."""
    code = r"""
import sys, numpy as np, torch, importlib
sys.path.insert(0, %r); sys.path.insert(0, %r)
from _libs import Oracle
from _cases import biased, english
huf = importlib.import_module("huffman-avx512_b200")
o = Oracle()
rng = np.random.default_rng(1)
cases = [("biased", biased(50001, seed=3)), ("english", english(70000, seed=4)),
         ("seven", rng.integers(0, 128, 30011, dtype=np.uint8).tobytes()), ("lone", b"z" * 5000), ("empty1", b"q")]
for name, data in cases:
    for k in (1, 4, 32):
        bs = 1 << 15
        n = len(data)
        blocks = [o.compress(k, data[i:i + bs]) for i in range(0, n, bs)]
        sizes = np.array([len(b) for b in blocks], dtype=np.int32)
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        comp = torch.frombuffer(bytearray(b"".join(blocks) + b"\0" * 64), dtype=torch.uint8).cuda()
        st = torch.zeros(1, dtype=torch.int32, device="cuda")
        out = huf.BlockCodec(k, bs).decompress(comp, torch.from_numpy(offs).cuda(), torch.from_numpy(sizes).cuda(), n,
                                               status=st, split=True)
        torch.cuda.synchronize()
        assert int(st.item()) == 0, (name, k)
        assert out[:n].cpu().numpy().tobytes() == data, (name, k)
print("ok")
""" % (ROOT, os.path.join(ROOT, "tests"))
    for bits in ("64", "96"):
        env = dict(os.environ, HUFB200_SPLIT_BITS=bits)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0 and "ok" in r.stdout, (bits, r.stdout[-2000:], r.stderr[-2000:])


def test_split_decode_of_our_own_slots_at_size(huf):
    """256 MiB in 1 MiB x 4 blocks (the shape the split path is for): compress on the device, decode
    both ways, compare with the input."""
    import torch
    n = 1 << 28
    g = torch.Generator(device="cuda").manual_seed(11)
    u = torch.rand(n, device="cuda", generator=g).clamp_(min=1e-30)
    raw = (torch.floor(torch.log(u) / float(np.log(0.8))).to(torch.int64) % 256).to(torch.uint8)
    del u
    codec = huf.BlockCodec(4, 1 << 20)
    slots, sizes = codec.compress(raw)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = codec.decompress(slots, codec.slot_offsets(n), sizes, n, status=status, split=True)
    assert int(status.item()) == 0
    assert torch.equal(out, raw)


def test_single_buffer_calls_take_the_split_path(huf, oracle):
    """hufb200_decompress of one large buffer (K streams only) goes through the split decode."""
    L = huf.load()
    for k, n in ((32, 8 << 20), (4, (1 << 20) + 3), (1, 300000), (48, 3 << 20), (8, 100 << 10), (32, 200000)):
        data = english(n, seed=k)
        comp = oracle.compress(k, data)
        before = huf.launch_count()
        assert huf.decompress(k, comp) == data
        launches = huf.launch_count() - before
        if len(comp) <= (1 << 20) and n <= (2 << 20):
            assert launches == 1  # the whole split decode as one CTA (k_split_small)
        else:
            assert launches > 1  # plan, scan, passes, item scan, write: not the single launch


@pytest.mark.parametrize("name", ["biased", "english", "seven_bit", "uniform", "two_symbols", "lone_symbol",
                                  "long_codes", "text"])
def test_small_single_buffers_one_cta_split(huf, oracle, name):
    """hufb200_decompress of small buffers (k_split_small: one launch, up to eight CTAs), every K class, ragged sizes."""
    for k in (1, 3, 4, 8, 16, 32, 48, 64):
        for n in (k * 1024, 100 << 10, 33333 + 1024 * k, (1 << 20) + 17):
            data = _inputs()[name](n)
            comp = oracle.compress(k, data)
            assert huf.decompress(k, comp) == data, (name, k, n)


def test_small_single_buffer_corruption(huf, oracle):
    """The one-CTA form on damaged buffers: an error or output of the right length, never a fault."""
    k = 8
    data = biased(100 << 10, seed=4)
    good = bytearray(oracle.compress(k, data))
    rng = np.random.default_rng(3)
    for trial in range(80):
        bad = bytearray(good)
        for _ in range(int(rng.integers(1, 6))):
            bad[int(rng.integers(0, len(bad)))] = int(rng.integers(0, 256))
        bad[0:4] = good[0:4]  # keep raw_size: the caller sizes its buffer from it
        try:
            out = huf.decompress(k, bytes(bad))
            assert len(out) == len(data)
        except huf.HufError as e:
            assert e.code == -4  # HUFB200_E_CORRUPT
    assert huf.decompress(k, bytes(good)) == data


def test_split_decode_survives_corrupt_input(huf, oracle):
    """Flipped bytes anywhere in the block: no fault, nothing written outside the output, and either
    an error status or output of the right length."""
    import torch
    k, bs = 4, 1 << 16
    data = biased(bs, seed=9)
    good = bytearray(oracle.compress(k, data))
    rng = np.random.default_rng(2)
    codec = huf.BlockCodec(k, bs)
    guard = 4096
    for trial in range(60):
        bad = bytearray(good)
        for _ in range(int(rng.integers(1, 6))):
            bad[int(rng.integers(0, len(bad)))] = int(rng.integers(0, 256))
        if trial % 10 == 0:  # garbage end offsets / header
            lo = int(rng.integers(0, 64))
            bad[lo:lo + 8] = bytes(rng.integers(0, 256, 8, dtype=np.uint8))
        comp, offs, sizes = _pack_host([bytes(bad)])
        buf = torch.full((bs + 2 * guard,), 0xA5, dtype=torch.uint8, device="cuda")
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        codec.decompress(comp, offs, sizes, bs, out=buf[guard:guard + bs], status=status, split=True)
        torch.cuda.synchronize()
        assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + bs:] == 0xA5).all()), trial
    # and the untouched buffer still decodes
    comp, offs, sizes = _pack_host([bytes(good)])
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = codec.decompress(comp, offs, sizes, bs, status=status, split=True)
    assert int(status.item()) == 0 and out[:bs].cpu().numpy().tobytes() == data


def test_small_single_buffer_with_unequal_streams_retries(huf, oracle):
    """Streams of very unequal bit length: the first eighth of the buffer is incompressible, the rest
    one symbol -- the CTA that gets the first streams has far more than 1.5 times its share of the
    bits, reports it, and the call falls back to the spread form."""
    rng = np.random.default_rng(8)
    n = 1 << 20
    data = rng.integers(0, 256, n // 8, dtype=np.uint8).tobytes() + b"\x07" * (n - n // 8)
    for k in (8, 32):
        comp = oracle.compress(k, data)
        before = huf.launch_count()
        assert huf.decompress(k, comp) == data
        assert huf.launch_count() - before > 1
