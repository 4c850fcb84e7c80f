"""Device-resident block codec over torch tensors (torch is only the allocator / stream provider).

Every method launches the library's own kernels through the *_dev entry points of
include/hufb200.h on torch's current CUDA stream; nothing here computes on the host.
"""
import ctypes as C

import torch

from .binding import check, load


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class BlockCodec:
    """N-stream Huffman codec over independent fixed-size blocks, each block a complete
    reference-format buffer (CompressMulti<k>, codec/huffman.cpp:738-846)."""

    def __init__(self, k, block_size, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("huffman-avx512_b200 needs a CUDA device; there is no CPU fallback")
        self.k = int(k)
        self.block_size = int(block_size)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.L = load()
        self.slot_stride = self.L.hufb200_slot_stride(self.block_size, self.k)
        self.table_bytes = self.L.hufb200_table_bytes()

    # ---- geometry
    def n_blocks(self, n):
        return self.L.hufb200_blocks_count(n, self.block_size)

    def alloc_slots(self, n):
        nb = self.n_blocks(n)
        slots = torch.empty(max(nb, 1) * self.slot_stride, dtype=torch.uint8, device=self.device)
        sizes = torch.zeros(max(nb, 1), dtype=torch.int32, device=self.device)
        return slots, sizes

    # ---- kernels
    def histogram(self, raw, out=None):
        """256 x int64 byte histogram of a uint8 CUDA tensor (k_histogram)."""
        assert raw.is_cuda and raw.dtype == torch.uint8 and raw.is_contiguous()
        if out is None:
            out = torch.empty(256, dtype=torch.int64, device=raw.device)
        with torch.cuda.device(raw.device):
            check(self.L.hufb200_histogram_dev(_ptr(raw), raw.numel(), _ptr(out), _stream()))
        return out

    def build_table(self, hist, out=None):
        """Shared table from a 256 x int64 histogram (k_build_table)."""
        assert hist.is_cuda and hist.dtype == torch.int64 and hist.numel() == 256
        if out is None:
            out = torch.empty(self.table_bytes, dtype=torch.uint8, device=hist.device)
        with torch.cuda.device(hist.device):
            check(self.L.hufb200_build_table_dev(_ptr(hist), _ptr(out), _stream()))
        return out

    def compress(self, raw, slots=None, sizes=None, table=None, status=None):
        """Compresses every block of `raw`; returns (slots, sizes).  Block b sits at
        slots[b*slot_stride : b*slot_stride + sizes[b]]."""
        assert raw.is_cuda and raw.dtype == torch.uint8 and raw.is_contiguous()
        if slots is None or sizes is None:
            slots, sizes = self.alloc_slots(raw.numel())
        with torch.cuda.device(raw.device):
            check(self.L.hufb200_compress_blocks_dev(self.k, self.block_size, _ptr(raw), raw.numel(), _ptr(slots),
                                                     self.slot_stride, _ptr(sizes), _ptr(table), _ptr(status),
                                                     _stream()))
        return slots, sizes

    def slot_offsets(self, n):
        nb = self.n_blocks(n)
        return torch.arange(nb, dtype=torch.int64, device=self.device) * self.slot_stride

    def pack(self, slots, sizes, n_blocks, packed=None):
        """Slot layout -> packed layout; returns (packed, offsets[int64], total[int64 scalar tensor])."""
        offsets = torch.empty(max(n_blocks, 1), dtype=torch.int64, device=slots.device)
        total = torch.zeros(1, dtype=torch.int64, device=slots.device)
        with torch.cuda.device(slots.device):
            if packed is None:
                check(self.L.hufb200_pack_blocks_dev(_ptr(slots), self.slot_stride, _ptr(sizes), n_blocks,
                                                     C.c_void_p(0), _ptr(offsets), _ptr(total), _stream()))
                tot = int(total.item())
                packed = torch.empty(tot + 32, dtype=torch.uint8, device=slots.device)
            check(self.L.hufb200_pack_blocks_dev(_ptr(slots), self.slot_stride, _ptr(sizes), n_blocks,
                                                 _ptr(packed), _ptr(offsets), _ptr(total), _stream()))
        return packed, offsets, total

    def decompress(self, comp, offsets, sizes, raw_n, out=None, status=None, split=None, work=None):
        """Decodes the blocks at comp[offsets[b] : offsets[b]+sizes[b]] into `raw_n` bytes.
        split: None = the library's choice (hufb200_decompress_prefers_split), True / False = the
        split decode (streams cut into items, one lane per item) / one lane per stream.
        work: workspace tensor for the split decode (split_work_bytes(raw_n) bytes), kept by the
        caller across calls; allocated here when missing."""
        assert comp.is_cuda and comp.dtype == torch.uint8
        nb = self.n_blocks(raw_n)
        if out is None:
            out = torch.empty(max(raw_n, 1), dtype=torch.uint8, device=comp.device)
        if split is None:
            split = bool(self.L.hufb200_decompress_prefers_split(self.k, nb, raw_n))
        with torch.cuda.device(comp.device):
            if split and nb:
                need = self.split_work_bytes(raw_n)
                if work is None or work.numel() < need:
                    work = torch.empty(need, dtype=torch.uint8, device=comp.device)
                check(self.L.hufb200_decompress_split_dev(self.k, self.block_size, _ptr(comp), _ptr(offsets),
                                                          _ptr(sizes), nb, _ptr(out), raw_n, _ptr(work), work.numel(),
                                                          _ptr(status), _stream()))
            else:
                check(self.L.hufb200_decompress_blocks_dev(self.k, self.block_size, _ptr(comp), _ptr(offsets),
                                                           _ptr(sizes), nb, _ptr(out), raw_n, _ptr(status), _stream()))
        return out

    def split_work_bytes(self, raw_n):
        with torch.cuda.device(self.device):
            return self.L.hufb200_decompress_split_work_bytes(self.k, self.block_size, self.n_blocks(raw_n), raw_n)
