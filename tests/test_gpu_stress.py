"""Seeded randomized GPU-vs-oracle sweep over sizes, stream counts and symbol distributions:
exercises every region alignment (E mod 4), slices that end on / off the 512-symbol iteration and
32-bit word boundaries, empty slices, staged and ring encoder modes, aligned and unaligned decoder
rounds.  Bit-exact both ways."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(rng, n, kind):
    if n == 0:
        return b""
    if kind == 0:    # geometric, like GenerateProbaData
        u = rng.random(n)
        u[u == 0] = 0.5
        p = rng.choice([0.05, 0.2, 0.5, 0.9])
        return (np.floor(np.log(u) / np.log(1 - p)).astype(np.int64) % 256).astype(np.uint8).tobytes()
    if kind == 1:    # uniform over a random alphabet size
        return rng.integers(0, int(rng.integers(1, 257)), n, dtype=np.uint8).tobytes()
    if kind == 2:    # few symbols, long runs
        return np.repeat(rng.integers(0, 256, max(1, n // 50 + 1), dtype=np.uint8), 50)[:n].tobytes()
    if kind == 3:    # power-of-two counts: long codes, length limiting
        syms = np.concatenate([np.full(min(1 << i, n), 65 + i, dtype=np.uint8) for i in range(18)])
        rng.shuffle(syms)
        return np.resize(syms, n).tobytes()
    return bytes([int(rng.integers(0, 256))]) * n   # one symbol


def test_random_small_buffers(huf, oracle):
    rng = np.random.default_rng(20260101)
    sizes = list(range(0, 40)) + [511, 512, 513, 1023, 1024, 1025, 4095, 4096, 4097, 16383, 16384, 16385]
    n_cases = 0
    for trial in range(260):
        n = int(rng.choice(sizes)) if trial % 2 == 0 else int(rng.integers(0, 70000))
        k = int(rng.choice([1, 2, 3, 4, 7, 8, 16, 24, 32, 40, 48, 64]))
        data = _data(rng, n, int(rng.integers(0, 5)))
        want = oracle.compress(k, data)
        got = huf.compress(k, data)
        assert got == want, (trial, n, k)
        assert huf.decompress(k, want) == data, (trial, n, k)
        n_cases += 1
    assert n_cases == 260


def test_random_block_shapes(huf, oracle):
    """Device block API across shapes that hit both encoder modes (slice <= 4096 staged, longer: ring)."""
    import torch
    rng = np.random.default_rng(7)
    for trial in range(14):
        k = int(rng.choice([4, 8, 16, 32, 48, 64]))
        bs = int(rng.choice([4096, 16384, 65536, 131072, 262144, 524288]))
        nb = int(rng.integers(1, 5))
        n = nb * bs - int(rng.integers(0, bs)) if trial % 3 else nb * bs
        data = _data(rng, n, int(rng.integers(0, 4)))
        codec = huf.BlockCodec(k, bs)
        raw = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        slots, sizes = codec.compress(raw)
        out = codec.decompress(slots, codec.slot_offsets(n), sizes, n)
        torch.cuda.synchronize()
        assert out[:n].cpu().numpy().tobytes() == data, (trial, k, bs, n)
        sz = sizes.cpu().numpy()
        sl = slots.cpu().numpy()
        for b in range(codec.n_blocks(n)):
            blk = sl[b * codec.slot_stride: b * codec.slot_stride + int(sz[b])].tobytes()
            assert blk == oracle.compress(k, data[b * bs: (b + 1) * bs]), (trial, k, bs, b)
