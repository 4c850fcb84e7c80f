"""CPU: the C-ABI library loads and exports every symbol include/hufb200.h declares, pure host
logic (bounds, container parsing, argument validation) and the failure mode without a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(huf):
    hdr = open(os.path.join(ROOT, "include", "hufb200.h")).read()
    declared = set(re.findall(r"\b(hufb200_[a-z0-9_]+)\s*\(", hdr))
    bound = {name for name, _, _ in huf.ABI_SYMBOLS}
    assert declared == bound, declared ^ bound
    L = huf.load()
    for name in declared:
        assert getattr(L, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", huf.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (hufb200_[a-z0-9_]+)", out))
    assert declared <= exported


def test_pure_host_functions(huf, oracle):
    L = huf.load()
    assert L.hufb200_version() >= 100
    for n in (0, 1, 1000, 131072, 1 << 20):
        for k in (1, 4, 32, 48, 64):
            assert L.hufb200_compress_bound(n, k) >= oracle.compress_bound(n, k) - 64
            assert L.hufb200_slot_stride(n, k) % 256 == 0
            assert L.hufb200_slot_stride(n, k) >= L.hufb200_compress_bound(n, k) + 3
    assert L.hufb200_blocks_count(0, 131072) == 0
    assert L.hufb200_blocks_count(131073, 131072) == 2
    assert L.hufb200_table_bytes() == 1024 + 1024 + 256 + 32 + 16  # enc, enc2, sorted_syms, len_count, scalars


def test_split_decode_choice_is_host_logic(huf):
    """hufb200_decompress_prefers_split: few streams in all and long enough slices -> split decode."""
    L = huf.load()
    MiB = 1 << 20
    assert L.hufb200_decompress_prefers_split(32, 1, 64 * MiB) == 1          # one large buffer: 32 streams
    assert L.hufb200_decompress_prefers_split(4, 1024, 1024 * MiB) == 1      # 1 MiB x 4 blocks: 4096 streams
    assert L.hufb200_decompress_prefers_split(32, 8192, 1024 * MiB) == 0     # BASELINE config 2: 262144 streams
    assert L.hufb200_decompress_prefers_split(32, 1, 100 << 10) == 0         # 3200 symbols per stream: too short
    assert L.hufb200_decompress_prefers_split(1, 1, 1024 * MiB) == 0         # above the split path's per-stream limit
    assert L.hufb200_decompress_prefers_split(4, 1, 1024 * MiB) == 1
    assert L.hufb200_decompress_prefers_split(4, 0, 0) == 0


def test_bound_covers_worst_case(huf, oracle):
    rng = np.random.default_rng(0)
    for n in (1, 100, 4097):
        data = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        for k in (1, 32, 64):
            assert len(oracle.compress(k, data)) <= huf.compress_bound(n, k)


def test_argument_validation_needs_no_device(huf):
    L = huf.load()
    n = C.c_size_t(0)
    buf = (C.c_uint8 * 64)()
    assert L.hufb200_compress(0, buf, 10, buf, 64, C.byref(n)) == -1
    assert L.hufb200_compress(65, buf, 10, buf, 64, C.byref(n)) == -1
    assert L.hufb200_decompress(4, buf, 4, buf, 64, C.byref(n)) == -4
    assert b"header" in L.hufb200_last_error()
    assert L.hufb200_container_info(buf, 64, None, None, None, None) == -4
    assert L.hufb200_compress_blocks(32, 0, buf, 10, buf, 64, C.byref(n)) == -1  # block size 0


def test_no_cpu_fallback_without_device(huf):
    """On a box without a GPU every compute call must fail loudly (never route to the oracle)."""
    if huf.lib.hufb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(huf.HufError) as ei:
        huf.compress(32, b"hello world")
    assert ei.value.code in (-3, -5)
    with pytest.raises(huf.HufError):
        huf.MakeHistogram(b"abc")
    with pytest.raises(RuntimeError):
        huf.BlockCodec(32, 131072)


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "huffman-avx512_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read().lower()
                assert "oracle" not in text and "_libs" not in text, f"{f} mentions the oracle"


def test_cpp_policy_header_compiles(huf):
    """include/hufb200.hpp (mirror of codec/huffman.h:42-52) compiles and links against the ABI."""
    src = r'''
#include "hufb200.hpp"
#include <cstdio>
template <typename T> int run() {   // the shape of the reference's TYPED_TEST bodies
  std::string raw = "Hello World";
  try {
    std::string c = T::Compress(raw);
    return T::Decompress(c) == raw ? 0 : 1;
  } catch (const std::exception& e) { std::printf("%s: %s\n", T::name().c_str(), e.what()); return 2; }
}
int main() { return run<hufb200::HuffmanCompressorB200<32>>() | run<hufb200::HuffmanCompressorB200<4>>(); }
'''
    libdir = os.path.dirname(huf.lib_path())
    exe = os.path.join("/tmp", "hufb200_policy_test")
    with open(exe + ".cpp", "w") as f:
        f.write(src)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), exe + ".cpp", "-o", exe,
                        "-L", libdir, "-lhufb200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rc = subprocess.run([exe], capture_output=True, text=True)
    if huf.lib.hufb200_device_count() > 0:
        assert rc.returncode == 0, rc.stdout
    else:
        assert rc.returncode == 2 and "failed" in rc.stdout  # loud failure, no fallback
