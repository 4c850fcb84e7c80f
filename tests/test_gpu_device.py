"""Device-resident API (what bench.py times): torch tensors in HBM, kernels on torch's stream."""
import numpy as np
import pytest

from _cases import biased, english

pytestmark = pytest.mark.gpu


def _t(data):
    import torch
    return torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()


@pytest.mark.parametrize("k,bs", [(32, 131072), (4, 16384), (8, 65536), (16, 32768), (48, 131072), (32, 1 << 20)])
def test_blocks_dev_parity_and_roundtrip(huf, oracle, k, bs):
    import torch
    n = 3 * bs + bs // 2 + 16  # ragged last block
    data = biased(n, seed=k + bs)
    codec = huf.BlockCodec(k, bs)
    raw = _t(data)
    slots, sizes = codec.compress(raw)
    nb = codec.n_blocks(n)
    torch.cuda.synchronize()
    sz = sizes.cpu().numpy()
    sl = slots.cpu().numpy()
    for b in range(nb):
        blk = sl[b * codec.slot_stride: b * codec.slot_stride + sz[b]].tobytes()
        assert blk == oracle.compress(k, data[b * bs: (b + 1) * bs]), (k, bs, b)
    # decode straight from the slots
    out = codec.decompress(slots, codec.slot_offsets(n), sizes, n)
    assert out[:n].cpu().numpy().tobytes() == data
    # and from the packed layout
    packed, offsets, total = codec.pack(slots, sizes, nb)
    assert int(total.item()) == int(sz[:nb].sum())
    assert np.array_equal(offsets[:nb].cpu().numpy(), np.concatenate([[0], np.cumsum(sz[:nb])[:-1]]))
    out2 = codec.decompress(packed, offsets, sizes, n)
    assert out2[:n].cpu().numpy().tobytes() == data


def test_histogram_dev_large(huf):
    import torch
    n = (1 << 28) + 12345
    g = torch.Generator(device="cuda").manual_seed(1)
    raw = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda", generator=g)
    raw[: n // 2] = 65  # skewed half
    codec = huf.BlockCodec(32, 131072)
    h = codec.histogram(raw)
    want = torch.bincount(raw.to(torch.int32), minlength=256)  # checker only
    assert torch.equal(h, want.to(torch.int64))
    assert int(h.sum().item()) == n


def test_shared_table_mode(huf, oracle):
    import torch
    k, bs = 32, 131072
    data = english(4 * bs, seed=3)
    codec = huf.BlockCodec(k, bs)
    raw = _t(data)
    hist = codec.histogram(raw)
    table = codec.build_table(hist)
    slots, sizes = codec.compress(raw, table=table)
    torch.cuda.synchronize()
    cd = oracle.make_coding(oracle.histogram(data))
    sz = sizes.cpu().numpy()
    sl = slots.cpu().numpy()
    for b in range(4):
        blk = sl[b * codec.slot_stride: b * codec.slot_stride + sz[b]].tobytes()
        want = oracle.compress_with_table(k, data[b * bs: (b + 1) * bs], cd["len_count"], cd["sorted_syms"])
        assert blk == want, b
    out = codec.decompress(slots, codec.slot_offsets(len(data)), sizes, len(data))
    assert out.cpu().numpy().tobytes() == data


def test_full_size_every_block_equals_reference(huf):
    """BASELINE config 2 at full size (1 GiB biased, 128 KiB x 32): EVERY one of the 8192 blocks
    byte-equal (SHA-256) to the reference's scalar CompressMulti<32> of the same bytes (the
    oracle restatement where the reference build is absent), plus the size-independent
    properties: decode(encode(x)) == x on the device, sum of block sizes == packed size."""
    import torch
    from _parity import compare_blocks
    k, bs, n = 32, 131072, 1 << 30
    g = torch.Generator(device="cuda").manual_seed(2)
    u = torch.rand(n, device="cuda", generator=g).clamp_(min=1e-30)
    raw = (torch.floor(torch.log(u) / np.log(0.8)).to(torch.int64) % 256).to(torch.uint8)
    del u
    codec = huf.BlockCodec(k, bs)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    slots, sizes = codec.compress(raw, status=status)
    out = codec.decompress(slots, codec.slot_offsets(n), sizes, n, status=status)
    assert int(status.item()) == 0
    assert torch.equal(out, raw)
    del out
    nb = codec.n_blocks(n)
    sz = sizes.cpu().numpy().astype(np.int64)
    ratio = sz.sum() / n
    assert 0.44 < ratio < 0.48, ratio
    packed, offsets, total = codec.pack(slots, sizes, nb)
    assert int(total.item()) == int(sz.sum())
    del slots
    name, checked, bad = compare_blocks(raw.cpu().numpy(), packed.cpu().numpy(), offsets.cpu().numpy(), sz, k, bs)
    assert checked == nb == 8192
    assert not bad, f"{len(bad)} blocks differ from {name}: {bad[:8]}"


def test_concurrent_launches_on_two_streams(huf, oracle):
    """Compress launches hand their blocks out through per-launch work counters: launches that
    overlap on different streams must not disturb each other."""
    import torch
    k, bs = 32, 16384
    nblk = 96
    datas = [biased(nblk * bs, seed=100 + i) for i in range(2)]
    codec = huf.BlockCodec(k, bs)
    raws = [_t(d) for d in datas]
    streams = [torch.cuda.Stream() for _ in range(2)]
    outs = [codec.alloc_slots(nblk * bs) for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(8):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                codec.compress(raws[i], slots=outs[i][0], sizes=outs[i][1])
    torch.cuda.synchronize()
    for i in range(2):
        sz = outs[i][1].cpu().numpy()
        sl = outs[i][0].cpu().numpy()
        for b in (0, 1, nblk // 2, nblk - 1):
            blk = sl[b * codec.slot_stride: b * codec.slot_stride + sz[b]].tobytes()
            assert blk == oracle.compress(k, datas[i][b * bs: (b + 1) * bs]), (i, b)
        out = codec.decompress(outs[i][0], codec.slot_offsets(nblk * bs), outs[i][1], nblk * bs)
        assert out[: nblk * bs].cpu().numpy().tobytes() == datas[i]


def test_many_launches_wrap_the_counter_ring(huf, oracle):
    """More launches than the ring of work counters has entries (4096): every pair must come back
    armed."""
    import torch
    k, bs = 8, 4096
    data = biased(3 * bs + 100, seed=7)
    codec = huf.BlockCodec(k, bs)
    raw = _t(data)
    slots, sizes = codec.alloc_slots(len(data))
    for _ in range(4500):
        codec.compress(raw, slots=slots, sizes=sizes)
    torch.cuda.synchronize()
    sz = sizes.cpu().numpy()
    sl = slots.cpu().numpy()
    for b in range(codec.n_blocks(len(data))):
        blk = sl[b * codec.slot_stride: b * codec.slot_stride + sz[b]].tobytes()
        assert blk == oracle.compress(k, data[b * bs: (b + 1) * bs]), b
