"""ctypes loaders for the two CPU checkers (oracle/libhuforacle.so, oracle/_ref/libhufref.so).

Test infrastructure only: nothing under huffman-avx512_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


def _ptr(a, t):
    return a.ctypes.data_as(t)


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


class Oracle:
    """oracle/huf_oracle.c (the plain-C restatement)."""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "libhuforacle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = L = C.CDLL(path)
        L.hufo_compress_bound.restype = C.c_size_t
        L.hufo_compress_bound.argtypes = [C.c_size_t, C.c_int]
        for f in (L.hufo_compress, L.hufo_decompress):
            f.restype = C.c_int
            f.argtypes = [C.c_int, u8p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.hufo_compress_with_table.restype = C.c_int
        L.hufo_compress_with_table.argtypes = [C.c_int, u8p, C.c_size_t, u16p, u8p, C.c_int, u8p,
                                               C.c_size_t, C.POINTER(C.c_size_t)]
        L.hufo_histogram.argtypes = [u8p, C.c_size_t, u32p]
        L.hufo_histogram64.argtypes = [u8p, C.c_size_t, u64p]
        L.hufo_sort_syms.argtypes = [u32p, u8p, C.c_int]
        L.hufo_limit_code_lengths.argtypes = [u16p]
        L.hufo_dtable1x.argtypes = [u16p, u8p, C.c_int, u8p]
        L.hufo_dtable2x.argtypes = [u16p, u8p, C.c_int, u8p]
        L.hufo_assign_codes.argtypes = [u16p, u8p, C.c_int, u16p, u16p]

    def compress_bound(self, n, k):
        return self.lib.hufo_compress_bound(n, k)

    def compress(self, k, raw):
        raw = np.ascontiguousarray(np.frombuffer(bytes(raw), dtype=np.uint8))
        out = np.empty(self.compress_bound(raw.size, k), dtype=np.uint8)
        n = C.c_size_t(0)
        rc = self.lib.hufo_compress(k, _ptr(raw, u8p), raw.size, _ptr(out, u8p), out.size, C.byref(n))
        if rc:
            raise RuntimeError(f"hufo_compress rc={rc}")
        return out[: n.value].tobytes()

    def compress_with_table(self, k, raw, len_count, syms):
        raw = np.ascontiguousarray(np.frombuffer(bytes(raw), dtype=np.uint8))
        lc = np.ascontiguousarray(np.asarray(len_count, dtype=np.uint16))
        sy = np.ascontiguousarray(np.frombuffer(bytes(syms), dtype=np.uint8))
        out = np.empty(self.compress_bound(raw.size, k), dtype=np.uint8)
        n = C.c_size_t(0)
        rc = self.lib.hufo_compress_with_table(k, _ptr(raw, u8p), raw.size, _ptr(lc, u16p),
                                               _ptr(sy, u8p), sy.size, _ptr(out, u8p), out.size,
                                               C.byref(n))
        if rc:
            raise RuntimeError(f"hufo_compress_with_table rc={rc}")
        return out[: n.value].tobytes()

    def decompress(self, k, comp):
        comp = np.ascontiguousarray(np.frombuffer(bytes(comp), dtype=np.uint8))
        raw_size = int(np.frombuffer(comp[:4].tobytes(), dtype="<u4")[0]) if comp.size >= 4 else 0
        out = np.empty(max(raw_size, 1), dtype=np.uint8)
        n = C.c_size_t(0)
        rc = self.lib.hufo_decompress(k, _ptr(comp, u8p), comp.size, _ptr(out, u8p), out.size, C.byref(n))
        if rc:
            raise RuntimeError(f"hufo_decompress rc={rc}")
        return out[: n.value].tobytes()

    def histogram(self, data):
        a = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        h = np.zeros(256, dtype=np.uint32)
        self.lib.hufo_histogram(_ptr(a, u8p), a.size, _ptr(h, u32p))
        return h

    def make_coding(self, hist):
        class Coding(C.Structure):
            _fields_ = [("code_bits", C.c_uint16 * 256), ("code_len", C.c_uint16 * 256),
                        ("sorted_syms", C.c_uint8 * 256), ("num_syms", C.c_int),
                        ("len_count", C.c_uint16 * 33), ("len_mask", C.c_uint32)]
        h = np.ascontiguousarray(np.asarray(hist, dtype=np.uint32))
        cd = Coding()
        self.lib.hufo_make_coding.argtypes = [u32p, C.POINTER(Coding)]
        self.lib.hufo_make_coding(_ptr(h, u32p), C.byref(cd))
        return dict(len_count=np.array(cd.len_count[:13], dtype=np.uint16),
                    sorted_syms=bytes(cd.sorted_syms[: cd.num_syms]), num_syms=cd.num_syms,
                    len_mask=cd.len_mask, code_bits=np.array(cd.code_bits[:], dtype=np.uint16),
                    code_len=np.array(cd.code_len[:], dtype=np.uint16))

    def sort_syms(self, hist, syms):
        h = np.ascontiguousarray(np.asarray(hist, dtype=np.uint32))
        s = np.array(np.frombuffer(bytes(syms), dtype=np.uint8))
        self.lib.hufo_sort_syms(_ptr(h, u32p), _ptr(s, u8p), s.size)
        return s.tobytes()

    def limit_code_lengths(self, len_count33):
        a = np.array(len_count33, dtype=np.uint16)
        assert a.size == 33
        self.lib.hufo_limit_code_lengths(_ptr(a, u16p))
        return a

    def dtable(self, which, len_count, syms):
        lc = np.ascontiguousarray(np.asarray(len_count, dtype=np.uint16))
        sy = np.ascontiguousarray(np.frombuffer(bytes(syms), dtype=np.uint8))
        w = 2 if which == 1 else 4
        out = np.zeros(4096 * w, dtype=np.uint8)
        f = self.lib.hufo_dtable1x if which == 1 else self.lib.hufo_dtable2x
        f(_ptr(lc, u16p), _ptr(sy, u8p), sy.size, _ptr(out, u8p))
        return out.reshape(4096, w)


def ref_path():
    return os.path.join(ORACLE_DIR, "_ref", "libhufref.so")


def have_ref():
    return os.path.exists(ref_path())


class Ref:
    """oracle/_ref/libhufref.so: the unmodified reference behind oracle/ref_shim.cpp."""

    SCALAR, GATHER, PERMUTE = 0, 1, 2

    def __init__(self):
        self.lib = L = C.CDLL(ref_path())
        for f in (L.ref_compress, L.ref_decompress):
            f.restype = C.c_int
            f.argtypes = [C.c_int, C.c_int, u8p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ref_histogram.argtypes = [C.c_int, u8p, C.c_size_t, u32p]
        L.ref_make_coding.argtypes = [u32p, u16p, u8p, C.POINTER(C.c_int), u32p, u16p, u16p]
        L.ref_limit_code_lengths.argtypes = [u16p]
        L.ref_dtable2x.argtypes = [u16p, u8p, C.c_int, u8p]
        L.ref_dtable1x.argtypes = [u16p, u8p, C.c_int, u8p]
        L.ref_sort_syms.argtypes = [u32p, u8p, C.c_int]
        L.ref_gen_proba.argtypes = [C.c_double, u8p, C.c_size_t]
        L.ref_gen_rand.argtypes = [C.c_int, u8p, C.c_size_t]
        L.ref_gen_equal_counts.argtypes = [u8p]
        L.ref_gen_many_random.restype = C.c_size_t
        L.ref_gen_many_random.argtypes = [u8p, C.POINTER(C.c_int)]
        L.ref_gen_hist_biased.argtypes = [u8p, C.c_size_t]
        L.ref_bench.restype = C.c_double
        L.ref_bench.argtypes = [C.c_int, C.c_int, C.c_int, u8p, C.c_size_t, C.c_size_t, C.c_int,
                                C.c_int, C.c_double, C.POINTER(C.c_double)]
        L.ref_bench_histogram.restype = C.c_double
        L.ref_bench_histogram.argtypes = [C.c_int, u8p, C.c_size_t, C.c_int, C.c_double]

    def compress(self, k, raw, variant=0):
        raw = np.ascontiguousarray(np.frombuffer(bytes(raw), dtype=np.uint8))
        out = np.empty(8 + 13 + 256 + 4 * k + (raw.size * 12 + 7) // 8 + 9 * k + 64, dtype=np.uint8)
        n = C.c_size_t(0)
        rc = self.lib.ref_compress(k, variant, _ptr(raw, u8p), raw.size, _ptr(out, u8p), out.size, C.byref(n))
        if rc:
            raise RuntimeError(f"ref_compress rc={rc}")
        return out[: n.value].tobytes()

    def decompress(self, k, comp, variant=0):
        comp = np.ascontiguousarray(np.frombuffer(bytes(comp), dtype=np.uint8))
        raw_size = int(np.frombuffer(comp[:4].tobytes(), dtype="<u4")[0])
        out = np.empty(max(raw_size, 1), dtype=np.uint8)
        n = C.c_size_t(0)
        rc = self.lib.ref_decompress(k, variant, _ptr(comp, u8p), comp.size, _ptr(out, u8p), out.size, C.byref(n))
        if rc:
            raise RuntimeError(f"ref_decompress rc={rc}")
        return out[: n.value].tobytes()

    def histogram(self, data, which=0):
        a = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        h = np.zeros(256, dtype=np.uint32)
        rc = self.lib.ref_histogram(which, _ptr(a, u8p), a.size, _ptr(h, u32p))
        assert rc == 0
        return h

    def make_coding(self, hist):
        h = np.ascontiguousarray(np.asarray(hist, dtype=np.uint32))
        lc = np.zeros(13, dtype=np.uint16)
        sy = np.zeros(256, dtype=np.uint8)
        ns = C.c_int(0)
        mask = np.zeros(1, dtype=np.uint32)
        cb = np.zeros(256, dtype=np.uint16)
        cl = np.zeros(256, dtype=np.uint16)
        self.lib.ref_make_coding(_ptr(h, u32p), _ptr(lc, u16p), _ptr(sy, u8p), C.byref(ns),
                                 _ptr(mask, u32p), _ptr(cb, u16p), _ptr(cl, u16p))
        return dict(len_count=lc, sorted_syms=sy[: ns.value].tobytes(), num_syms=ns.value,
                    len_mask=int(mask[0]), code_bits=cb, code_len=cl)

    def sort_syms(self, hist, syms):
        h = np.ascontiguousarray(np.asarray(hist, dtype=np.uint32))
        s = np.array(np.frombuffer(bytes(syms), dtype=np.uint8))
        self.lib.ref_sort_syms(_ptr(h, u32p), _ptr(s, u8p), s.size)
        return s.tobytes()

    def limit_code_lengths(self, len_count33):
        a = np.array(len_count33, dtype=np.uint16)
        assert a.size == 33
        self.lib.ref_limit_code_lengths(_ptr(a, u16p))
        return a

    def dtable(self, which, len_count, syms):
        lc = np.ascontiguousarray(np.asarray(len_count, dtype=np.uint16))
        sy = np.ascontiguousarray(np.frombuffer(bytes(syms), dtype=np.uint8))
        w = 2 if which == 1 else 4
        out = np.zeros(4096 * w, dtype=np.uint8)
        f = self.lib.ref_dtable1x if which == 1 else self.lib.ref_dtable2x
        f(_ptr(lc, u16p), _ptr(sy, u8p), sy.size, _ptr(out, u8p))
        return out.reshape(4096, w)

    # generators
    def gen_proba(self, p, n):
        a = np.empty(n, dtype=np.uint8)
        self.lib.ref_gen_proba(p, _ptr(a, u8p), n)
        return a.tobytes()

    def gen_rand(self, which, n):
        a = np.empty(n, dtype=np.uint8)
        self.lib.ref_gen_rand(which, _ptr(a, u8p), n)
        return a.tobytes()

    def gen_equal_counts(self):
        a = np.empty(1024, dtype=np.uint8)
        self.lib.ref_gen_equal_counts(_ptr(a, u8p))
        return a.tobytes()

    def gen_many_random(self):
        a = np.empty(100_000, dtype=np.uint8)
        lens = (C.c_int * 100)()
        tot = self.lib.ref_gen_many_random(_ptr(a, u8p), lens)
        out, pos = [], 0
        for l in lens:
            out.append(a[pos: pos + l].tobytes())
            pos += l
        assert pos == tot
        return out

    def gen_hist_biased(self, n):
        a = np.empty(n, dtype=np.uint8)
        self.lib.ref_gen_hist_biased(_ptr(a, u8p), n)
        return a.tobytes()

    def bench(self, k, variant, direction, bufs, length, stride, n_bufs, threads, min_seconds):
        a = np.ascontiguousarray(bufs)
        ratio = C.c_double(0)
        r = self.lib.ref_bench(k, variant, direction, _ptr(a, u8p), length, stride, n_bufs, threads,
                               min_seconds, C.byref(ratio))
        return r, ratio.value
