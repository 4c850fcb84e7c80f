"""Multi-GPU path on real devices (skipped with fewer than two): one process per GPU over NCCL,
every rank compresses ITS byte range of ONE common input (sharded.shard_bytes); the ranks' blocks
put together must equal what one GPU produces for the whole input -- per-block tables, and
shared-table mode, where the only collective of the path (the 256-bin histogram all-reduce) runs
on the product's own k_histogram output."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, k, block, outdir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    huf = importlib.import_module("huffman-avx512_b200")
    from _cases import biased
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    data = np.frombuffer(biased(n, seed=4242), dtype=np.uint8)  # the same input on every rank
    lo, hi = huf.sharded.shard_bytes(n, block, rank, world)
    shard = torch.from_numpy(data[lo:hi].copy()).to(dev)
    codec = huf.BlockCodec(k, block, device=dev)
    sh = huf.sharded.ShardedCodec(codec)
    res = {}
    for mode in ("per_block", "shared"):
        table = None
        if mode == "shared":
            table, hist = sh.shared_table(shard)  # k_histogram -> NCCL all-reduce -> k_build_table
            res["hist"] = hist.cpu().numpy()
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        slots, sizes = codec.compress(shard, table=table, status=status)
        nb = codec.n_blocks(hi - lo)
        packed, offsets, total = codec.pack(slots, sizes, nb)
        out = codec.decompress(packed, offsets, sizes, hi - lo, status=status)
        assert int(status.item()) == 0
        assert bool(torch.equal(out[: hi - lo], shard))
        res[mode] = (packed[: int(total.item())].cpu().numpy(), sizes[:nb].cpu().numpy())
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), lo=lo, hi=hi, hist=res["hist"],
             pb=res["per_block"][0], pb_sizes=res["per_block"][1], sh=res["shared"][0], sh_sizes=res["shared"][1])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks_put_together_equal_one_gpu(huf, oracle, tmp_path, world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from _cases import biased
    n, k, block = 37 * 65536 + 12345, 32, 65536  # 38 blocks, the last one short; uneven over the ranks
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, k, block, str(tmp_path)), nprocs=world, join=True)
    data = biased(n, seed=4242)
    # one GPU, whole input
    raw = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    codec = huf.BlockCodec(k, block)
    nb = codec.n_blocks(n)
    hist1 = codec.histogram(raw)
    table1 = codec.build_table(hist1)
    want = {}
    for mode, table in (("pb", None), ("sh", table1)):
        slots, sizes = codec.compress(raw, table=table)
        packed, offsets, total = codec.pack(slots, sizes, nb)
        want[mode] = (packed[: int(total.item())].cpu().numpy().tobytes(), sizes[:nb].cpu().numpy())
    parts = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    assert parts[0]["lo"] == 0 and parts[-1]["hi"] == n
    for a, b in zip(parts, parts[1:]):
        assert a["hi"] == b["lo"]
    want_hist = np.bincount(np.frombuffer(data, dtype=np.uint8), minlength=256)
    for p in parts:
        assert np.array_equal(p["hist"], want_hist)  # every rank holds the global histogram
    assert np.array_equal(hist1.cpu().numpy(), want_hist)
    for mode in ("pb", "sh"):
        got = b"".join(p[mode].tobytes() for p in parts)
        got_sizes = np.concatenate([p[mode + "_sizes"] for p in parts])
        assert np.array_equal(got_sizes, want[mode][1]), mode
        assert got == want[mode][0], mode
    # and the one-GPU per-block output is the oracle's
    off = 0
    for b in range(nb):
        s = int(want["pb"][1][b])
        assert want["pb"][0][off: off + s] == oracle.compress(k, data[b * block: (b + 1) * block]), b
        off += s
