"""Out-of-bounds WRITE checks of our own (compute-sanitizer is closed on this pool of GPUs):
every output buffer the kernels get is surrounded by / filled with a canary pattern, and the
canaries must survive -- for well-formed input on the encoder's rare paths (staging overflow, ring
fallback, long slices, unaligned slices) and for systematically corrupted input on the decoder."""
import numpy as np
import pytest

from _cases import biased, english

pytestmark = pytest.mark.gpu

CANARY = 0xCD


def _t(data):
    import torch
    return torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()


@pytest.mark.parametrize("k,bs,kind", [(32, 131072, "biased"), (32, 131072, "uniform"), (48, 131072, "biased"),
                                       (4, 16384, "uniform"), (8, 1 << 20, "biased"), (16, 262144, "english"),
                                       (32, 65536, "mixed"), (5, 40000 // 16 * 16, "uniform")])
def test_encoder_writes_stay_inside_each_block(huf, k, bs, kind):
    import torch
    n = 5 * bs + bs // 3
    rng = np.random.default_rng(k * 1000 + bs)
    if kind == "biased":
        data = biased(n, seed=k)
    elif kind == "english":
        data = english(n, seed=k)
    elif kind == "uniform":  # incompressible: 8+ bits per symbol -> staging overflow / ring fallback paths
        data = bytes(rng.integers(0, 256, n, dtype=np.uint8))
    else:  # runs of incompressible and of highly compressible bytes: streams of very different sizes
        parts = []
        while sum(map(len, parts)) < n:
            m = int(rng.integers(100, 9000))
            parts.append(bytes(rng.integers(0, 256, m, dtype=np.uint8)) if rng.random() < 0.5 else bytes([int(rng.integers(0, 3))]) * m)
        data = b"".join(parts)[:n]
    codec = huf.BlockCodec(k, bs)
    raw = _t(data)
    nb = codec.n_blocks(n)
    guard = 1 << 16
    buf = torch.full((nb * codec.slot_stride + 2 * guard,), CANARY, dtype=torch.uint8, device="cuda")
    slots = buf[guard: guard + nb * codec.slot_stride]
    assert slots.data_ptr() % 16 == 0
    sizes = torch.zeros(nb, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    codec.compress(raw, slots=slots, sizes=sizes, status=status)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    h = buf.cpu().numpy()
    sz = sizes.cpu().numpy()
    assert (h[:guard] == CANARY).all() and (h[guard + nb * codec.slot_stride:] == CANARY).all()
    for b in range(nb):
        lo = guard + b * codec.slot_stride
        tail = h[lo + int(sz[b]) + 3: lo + codec.slot_stride]  # up to 3 zero bytes past the end are documented
        assert (tail == CANARY).all(), (b, int(np.argmax(tail != CANARY)))
    out = codec.decompress(slots, codec.slot_offsets(n), sizes, n, status=status)
    assert int(status.item()) == 0 and out[:n].cpu().numpy().tobytes() == data


def test_decoder_writes_stay_inside_the_output(huf):
    import torch
    rng = np.random.default_rng(31337)
    for k, bs in ((32, 131072), (4, 16384), (48, 65536), (8, 262144)):
        n = 6 * bs + 777
        data = biased(n, seed=k + 1)
        codec = huf.BlockCodec(k, bs)
        raw = _t(data)
        nb = codec.n_blocks(n)
        slots, sizes = codec.compress(raw)
        packed, offsets, total = codec.pack(slots, sizes, nb)
        tot = int(total.item())
        good = packed[:tot].clone()
        good_sizes = sizes.clone()
        guard = 1 << 16
        obuf = torch.full((n + 2 * guard,), CANARY, dtype=torch.uint8, device="cuda")
        out = obuf[guard: guard + n]
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        flagged = 0
        for trial in range(60):
            bad = good.clone()
            bsz = good_sizes.clone()
            kind = trial % 5
            if kind == 0:    # random bytes anywhere (headers, tables, end offsets, payload)
                idx = torch.from_numpy(rng.integers(0, tot, 64)).cuda()
                bad[idx] = torch.from_numpy(rng.integers(0, 256, 64, dtype=np.uint8)).cuda()
            elif kind == 1:  # garbage over the first 64 bytes of every block
                for b in range(nb):
                    o = int(offsets[b].item())
                    bad[o: o + 64] = torch.from_numpy(rng.integers(0, 256, 64, dtype=np.uint8)).cuda()
            elif kind == 2:  # sizes that lie (shorter / longer than the block really is)
                bsz = (bsz.to(torch.int64) + torch.from_numpy(rng.integers(-2000, 2000, bsz.numel())).cuda()).clamp_(min=0).to(torch.int32)
            elif kind == 3:  # end offsets all ones
                for b in range(nb):
                    o = int(offsets[b].item()) + 8 + 13 + 30
                    bad[o: o + 4 * (k - 1)] = 0xFF
            else:            # zeroed payload tail
                bad[tot // 2:] = 0
            status.zero_()
            obuf.fill_(CANARY)
            # the packed buffer is handed over with slack behind it: sizes may claim more than a block has
            slack = torch.zeros(tot + 4096, dtype=torch.uint8, device="cuda")
            slack[:tot] = bad
            codec.decompress(slack, offsets, bsz, n, out=out, status=status)
            torch.cuda.synchronize()
            h = obuf.cpu().numpy()
            assert (h[:guard] == CANARY).all() and (h[guard + n:] == CANARY).all(), (k, bs, trial)
            flagged += int(status.item() != 0)
        assert flagged >= 30, flagged
        status.zero_()
        codec.decompress(good, offsets, good_sizes, n, out=out, status=status)
        assert int(status.item()) == 0 and out.cpu().numpy().tobytes() == data
