"""Multi-GPU sharding of the block codec: one process per GPU, contiguous block ranges per
rank, no payload exchange.  The only collective is the optional 256-bin histogram all-reduce
when one shared table is requested.  Works under torch.distributed
with the nccl backend (CUDA tensors) and, for the host-side logic, gloo (CPU tensors)."""
import torch
import torch.distributed as dist


def shard_blocks(n_blocks, rank, world):
    """Contiguous block range [b0, b1) of `rank`: ceil(n_blocks / world) blocks per rank."""
    per = (n_blocks + world - 1) // world
    b0 = min(rank * per, n_blocks)
    b1 = min(b0 + per, n_blocks)
    return b0, b1


def shard_bytes(n, block_size, rank, world):
    """Byte range [lo, hi) of the raw input owned by `rank` (whole blocks; the last block of the
    input may be short)."""
    n_blocks = (n + block_size - 1) // block_size
    b0, b1 = shard_blocks(n_blocks, rank, world)
    return min(b0 * block_size, n), min(b1 * block_size, n)


def allreduce_histogram(hist, group=None):
    """Sum of the ranks' 256-bin int64 histograms, in place.  No-op without a process group."""
    assert hist.dtype == torch.int64 and hist.numel() == 256
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


class ShardedCodec:
    """Each rank compresses / decompresses its own contiguous block range on its own GPU."""

    def __init__(self, codec, group=None):
        self.codec = codec
        self.group = group

    def shared_table(self, raw_shard):
        """Histogram of the local shard -> all-reduce (2 KiB) -> identical table on every rank."""
        hist = self.codec.histogram(raw_shard)
        allreduce_histogram(hist, self.group)
        return self.codec.build_table(hist), hist

    def compress(self, raw_shard, shared_table=False, slots=None, sizes=None):
        table = None
        if shared_table:
            table, _ = self.shared_table(raw_shard)
        return self.codec.compress(raw_shard, slots=slots, sizes=sizes, table=table)

    def decompress(self, comp, offsets, sizes, raw_n, out=None):
        return self.codec.decompress(comp, offsets, sizes, raw_n, out=out)
