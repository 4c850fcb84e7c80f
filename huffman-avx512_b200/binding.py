"""ctypes binding of include/hufb200.h plus the host-pointer convenience calls."""
import ctypes as C
import os

import numpy as np

from . import build as _build

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
szp = C.POINTER(C.c_size_t)

OK, E_INVALID, E_NOSPACE, E_CUDA, E_CORRUPT, E_NODEVICE = 0, -1, -2, -3, -4, -5

# every symbol include/hufb200.h declares: (name, restype, argtypes)
ABI_SYMBOLS = [
    ("hufb200_version", C.c_int, []),
    ("hufb200_last_error", C.c_char_p, []),
    ("hufb200_device_count", C.c_int, []),
    ("hufb200_launch_count", C.c_uint64, []),
    ("hufb200_release_workspace", None, []),
    ("hufb200_histogram", C.c_int, [C.c_void_p, C.c_size_t, u32p]),
    ("hufb200_histogram64", C.c_int, [C.c_void_p, C.c_size_t, u64p]),
    ("hufb200_histogram_dev", C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    ("hufb200_make_table", C.c_int, [u32p, u16p, u8p, C.POINTER(C.c_int), u32p, u16p, u16p]),
    ("hufb200_decode_table", C.c_int, [u16p, C.c_void_p, C.c_int, u8p]),
    ("hufb200_decode_table1x", C.c_int, [u16p, C.c_void_p, C.c_int, u8p]),
    ("hufb200_compress_bound", C.c_size_t, [C.c_size_t, C.c_int]),
    ("hufb200_compress", C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, szp]),
    ("hufb200_decompress", C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, szp]),
    ("hufb200_raw_size", C.c_int, [C.c_void_p, C.c_size_t, szp]),
    ("hufb200_compress_with_table", C.c_int,
     [C.c_int, C.c_void_p, C.c_size_t, u16p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, szp]),
    ("hufb200_blocks_count", C.c_size_t, [C.c_size_t, C.c_size_t]),
    ("hufb200_container_bound", C.c_size_t, [C.c_size_t, C.c_size_t, C.c_int]),
    ("hufb200_compress_blocks", C.c_int,
     [C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, szp]),
    ("hufb200_decompress_blocks", C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, szp]),
    ("hufb200_container_info", C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), szp, szp, szp]),
    ("hufb200_slot_stride", C.c_size_t, [C.c_size_t, C.c_int]),
    ("hufb200_compress_blocks_dev", C.c_int,
     [C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
      C.c_void_p, C.c_void_p]),
    ("hufb200_decompress_blocks_dev", C.c_int,
     [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
      C.c_void_p, C.c_void_p]),
    ("hufb200_decompress_split_work_bytes", C.c_size_t, [C.c_int, C.c_size_t, C.c_size_t, C.c_size_t]),
    ("hufb200_decompress_prefers_split", C.c_int, [C.c_int, C.c_size_t, C.c_size_t]),
    ("hufb200_decompress_split_dev", C.c_int,
     [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
      C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    ("hufb200_pack_blocks_dev", C.c_int,
     [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("hufb200_table_bytes", C.c_size_t, []),
    ("hufb200_build_table_dev", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
]


class HufError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hufb200 error {code}: {msg}")
        self.code = code


_lib = None


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Loads libhufb200.so (building it in-tree first if it is missing).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise HufError(E_CUDA, f"{path} is missing; run `python __graft_entry__.py build`")
        _build.build()
    L = C.CDLL(path)
    for name, res, args in ABI_SYMBOLS:
        f = getattr(L, name)  # AttributeError if the library does not export it
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


class _LazyLib:
    def __getattr__(self, name):
        return getattr(load(), name)


lib = _LazyLib()


def check(rc):
    if rc != OK:
        raise HufError(rc, (load().hufb200_last_error() or b"").decode("utf-8", "replace"))


def _np_u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data.view(np.uint8).reshape(-1))
    return np.frombuffer(bytes(data), dtype=np.uint8)


def _vp(a):
    return C.c_void_p(a.ctypes.data) if a.size else C.c_void_p(0)


def compress_bound(n, k):
    return load().hufb200_compress_bound(n, k)


def slot_stride(block_size, k):
    return load().hufb200_slot_stride(block_size, k)


def blocks_count(n, block_size):
    return load().hufb200_blocks_count(n, block_size)


def launch_count():
    return load().hufb200_launch_count()


def MakeHistogram(data):
    """huffman::MakeHistogram (codec/histogram.h:12): 256 x u32 counts."""
    a = _np_u8(data)
    h = np.zeros(256, dtype=np.uint32)
    check(load().hufb200_histogram(_vp(a), a.size, h.ctypes.data_as(u32p)))
    return h


def histogram64(data):
    a = _np_u8(data)
    h = np.zeros(256, dtype=np.uint64)
    check(load().hufb200_histogram64(_vp(a), a.size, h.ctypes.data_as(u64p)))
    return h


def compress(k, raw):
    """huffman::CompressMulti<k> (codec/huffman.h:9-10)."""
    a = _np_u8(raw)
    out = np.empty(compress_bound(a.size, k), dtype=np.uint8)
    n = C.c_size_t(0)
    check(load().hufb200_compress(k, _vp(a), a.size, _vp(out), out.size, C.byref(n)))
    return out[: n.value].tobytes()


MAX_SINGLE_BUFFER = 1 << 30  # the library's single-buffer limit (32-bit offsets, codec/huffman.cpp:772)


def decompress(k, comp, max_raw=MAX_SINGLE_BUFFER):
    """huffman::DecompressMulti<k> (codec/huffman.h:11-12).  The output is sized from the buffer's
    untrusted raw_size field, so that is checked against `max_raw` before anything is allocated."""
    a = _np_u8(comp)
    rs = C.c_size_t(0)
    check(load().hufb200_raw_size(_vp(a), a.size, C.byref(rs)))
    if rs.value > max_raw:
        raise HufError(E_CORRUPT, f"header claims {rs.value} raw bytes, more than max_raw={max_raw}")
    out = np.empty(max(rs.value, 1), dtype=np.uint8)
    n = C.c_size_t(0)
    check(load().hufb200_decompress(k, _vp(a), a.size, _vp(out), rs.value, C.byref(n)))
    return out[: n.value].tobytes()


def compress_with_table(k, raw, len_count, sorted_syms):
    a = _np_u8(raw)
    lc = np.ascontiguousarray(np.asarray(len_count, dtype=np.uint16))
    assert lc.size == 13
    sy = _np_u8(sorted_syms)
    out = np.empty(compress_bound(a.size, k), dtype=np.uint8)
    n = C.c_size_t(0)
    check(load().hufb200_compress_with_table(k, _vp(a), a.size, lc.ctypes.data_as(u16p), _vp(sy), sy.size,
                                             _vp(out), out.size, C.byref(n)))
    return out[: n.value].tobytes()


def make_table(hist):
    """MakeCanonicalCoding (codec/huffman.cpp:339-437) on the device."""
    h = np.ascontiguousarray(np.asarray(hist, dtype=np.uint32))
    assert h.size == 256
    lc = np.zeros(13, dtype=np.uint16)
    sy = np.zeros(256, dtype=np.uint8)
    ns = C.c_int(0)
    mask = np.zeros(1, dtype=np.uint32)
    cb = np.zeros(256, dtype=np.uint16)
    cl = np.zeros(256, dtype=np.uint16)
    check(load().hufb200_make_table(h.ctypes.data_as(u32p), lc.ctypes.data_as(u16p), sy.ctypes.data_as(u8p),
                                    C.byref(ns), mask.ctypes.data_as(u32p), cb.ctypes.data_as(u16p),
                                    cl.ctypes.data_as(u16p)))
    return dict(len_count=lc, sorted_syms=sy[: ns.value].tobytes(), num_syms=ns.value, len_mask=int(mask[0]),
                code_bits=cb, code_len=cl)


def decode_table(len_count, sorted_syms):
    """Decoder2x table (codec/huffman.cpp:642-681) as built by the decode kernel: (4096, 4) u8."""
    lc = np.ascontiguousarray(np.asarray(len_count, dtype=np.uint16))
    sy = _np_u8(sorted_syms)
    out = np.zeros(4096 * 4, dtype=np.uint8)
    check(load().hufb200_decode_table(lc.ctypes.data_as(u16p), _vp(sy), sy.size, out.ctypes.data_as(u8p)))
    return out.reshape(4096, 4)


def decode_table1x(len_count, sorted_syms):
    """Decoder1x table (codec/huffman.cpp:594-632) from the same builder: (4096, 2) u8 {code_len, sym}."""
    lc = np.ascontiguousarray(np.asarray(len_count, dtype=np.uint16))
    sy = _np_u8(sorted_syms)
    out = np.zeros(4096 * 2, dtype=np.uint8)
    check(load().hufb200_decode_table1x(lc.ctypes.data_as(u16p), _vp(sy), sy.size, out.ctypes.data_as(u8p)))
    return out.reshape(4096, 2)


def compress_blocks(k, block_size, raw):
    a = _np_u8(raw)
    out = np.empty(load().hufb200_container_bound(a.size, block_size, k), dtype=np.uint8)
    n = C.c_size_t(0)
    check(load().hufb200_compress_blocks(k, block_size, _vp(a), a.size, _vp(out), out.size, C.byref(n)))
    return out[: n.value].tobytes()


def container_info(container):
    a = _np_u8(container)
    k = C.c_int(0)
    bs, rs, nb = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    check(load().hufb200_container_info(_vp(a), a.size, C.byref(k), C.byref(bs), C.byref(rs), C.byref(nb)))
    return dict(k=k.value, block_size=bs.value, raw_size=rs.value, n_blocks=nb.value)


def decompress_blocks(container, max_raw=None):
    """Decodes a block container.  A few kilobytes of header can claim terabytes of raw data (a
    block of one repeated symbol has no payload bits), so pass `max_raw` when the container is
    untrusted: the claimed size is checked against it before the output is allocated."""
    a = _np_u8(container)
    info = container_info(a)
    if max_raw is not None and info["raw_size"] > max_raw:
        raise HufError(E_CORRUPT, f"container claims {info['raw_size']} raw bytes, more than max_raw={max_raw}")
    out = np.empty(max(info["raw_size"], 1), dtype=np.uint8)
    n = C.c_size_t(0)
    check(load().hufb200_decompress_blocks(_vp(a), a.size, _vp(out), info["raw_size"], C.byref(n)))
    return out[: n.value].tobytes()


class HuffmanCompressorB200:
    """Mirror of huffman::HuffmanCompressorMulti<K> (codec/huffman.h:42-52): Compress / Decompress / name."""

    def __init__(self, k):
        self.k = k

    def Compress(self, raw):
        return compress(self.k, raw)

    def Decompress(self, compressed):
        return decompress(self.k, compressed)

    def name(self):
        return f"HuffmanB200<{self.k}>"
