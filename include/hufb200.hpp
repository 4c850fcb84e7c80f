// hufb200.hpp -- header-only C++ mirror of the reference's compressor policy classes
// (codec/huffman.h:42-97) and histogram entry point (codec/histogram.h:10-12) on top of
// the C ABI in hufb200.h.  A test or benchmark templated on "a compressor" (the reference's
// TYPED_TEST_SUITE, codec/huffman_test.cpp:47-54, and DEFINE_BENCHMARKS,
// codec/huffman_benchmark.cpp:252-281) accepts hufb200::HuffmanCompressorB200<K> unchanged.
//
// Error behaviour: the reference has no error channel (assert/abort); these wrappers throw
// std::runtime_error carrying hufb200_last_error() instead of aborting.
#pragma once

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <string_view>

#include "hufb200.h"

namespace hufb200 {

using ByteHistogram = std::array<uint32_t, 256>;  // codec/histogram.h:10

inline void check(int rc, const char* what) {
  if (rc != HUFB200_OK)
    throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + hufb200_last_error());
}

// huffman::MakeHistogram, codec/histogram.h:12
inline ByteHistogram MakeHistogram(std::string_view str) {
  ByteHistogram h{};
  check(hufb200_histogram(reinterpret_cast<const uint8_t*>(str.data()), str.size(), h.data()), "hufb200_histogram");
  return h;
}

// huffman::CompressMulti<K>, codec/huffman.h:9-10
template <int K>
std::string CompressMulti(std::string_view raw) {
  std::string out(hufb200_compress_bound(raw.size(), K), '\0');
  size_t n = 0;
  check(hufb200_compress(K, reinterpret_cast<const uint8_t*>(raw.data()), raw.size(),
                         reinterpret_cast<uint8_t*>(out.data()), out.size(), &n),
        "hufb200_compress");
  out.resize(n);
  return out;
}

// huffman::DecompressMulti<K>, codec/huffman.h:11-12
template <int K>
std::string DecompressMulti(std::string_view compressed) {
  size_t raw_size = 0;
  check(hufb200_raw_size(reinterpret_cast<const uint8_t*>(compressed.data()), compressed.size(), &raw_size),
        "hufb200_raw_size");
  // raw_size is untrusted; the library's single-buffer limit (2^30, codec/huffman.cpp:772) bounds the allocation
  if (raw_size > (size_t(1) << 30)) throw std::runtime_error("hufb200: header claims more than 2^30 raw bytes");
  std::string out(raw_size, '\0');
  size_t n = 0;
  check(hufb200_decompress(K, reinterpret_cast<const uint8_t*>(compressed.data()), compressed.size(),
                           reinterpret_cast<uint8_t*>(out.data()), out.size(), &n),
        "hufb200_decompress");
  out.resize(n);
  return out;
}

// Same shape as huffman::HuffmanCompressorMulti<K>, codec/huffman.h:42-52.
template <int K>
class HuffmanCompressorB200 {
 public:
  static_assert(K >= 1 && K <= HUFB200_MAX_K, "K outside 1..64");
  static std::string Compress(std::string_view raw) { return CompressMulti<K>(raw); }
  static std::string Decompress(std::string_view compressed) { return DecompressMulti<K>(compressed); }
  static std::string name() { return "HuffmanB200<" + std::to_string(K) + ">"; }
};

}  // namespace hufb200
