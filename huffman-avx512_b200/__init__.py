"""huffman-avx512_b200 -- B200-native multi-stream Huffman codec (hot path of ahartik/huffman-avx512).

The product is `libhufb200.so` (hand-written sm_100a kernels behind the C ABI of
include/hufb200.h).  This package is the thin Python host layer used by the tests and
bench.py: a ctypes binding of that ABI (`lib`), a mirror of the reference's compressor
policy classes (`HuffmanCompressorB200`, codec/huffman.h:42-52), a device-resident block
codec over torch tensors (`BlockCodec`) and the multi-GPU sharding helper (`sharded`).

There is no CPU fallback: importing works without a GPU (so the ABI surface can be checked),
but every compute call raises `HufError` when no CUDA device is usable.
"""
from .binding import (HufError, lib, lib_path, load, HuffmanCompressorB200, MakeHistogram, compress,
                      decompress, compress_with_table, compress_blocks, decompress_blocks, make_table,
                      decode_table, decode_table1x, compress_bound, slot_stride, blocks_count, launch_count, ABI_SYMBOLS)
from .device import BlockCodec
from . import sharded

__all__ = ["HufError", "lib", "lib_path", "load", "HuffmanCompressorB200", "MakeHistogram", "compress",
           "decompress", "compress_with_table", "compress_blocks", "decompress_blocks", "make_table",
           "decode_table", "decode_table1x", "compress_bound", "slot_stride", "blocks_count", "launch_count", "BlockCodec",
           "sharded", "ABI_SYMBOLS"]
