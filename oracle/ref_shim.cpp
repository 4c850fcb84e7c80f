// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" shim around the UNMODIFIED reference sources under
// /root/reference (compiled where they lie; nothing is copied).  It textually
// includes codec/huffman.cpp so the anonymous-namespace internals
// (MakeCanonicalCoding, Decoder2x, ...) are reachable, and instantiates the
// scalar K=48 codec the reference does not instantiate itself
// (codec/huffman.cpp:1977-2004).  Built by oracle/Makefile into
// oracle/_ref/libhufref.so.  Consumers: tests/, bench.py's reference arm and
// cpu_baseline leg, __graft_entry__.smoke().
#include "codec/huffman.cpp"  // NOLINT: deliberate, see above

#include <chrono>
#include <cmath>
#include <cstdio>
#include <random>
#include <thread>

namespace {

using huffman::ByteHistogram;

template <int K>
int CompressK(int variant, std::string_view raw, std::string* out) {
  switch (variant) {
    case 0: *out = huffman::CompressMulti<K>(raw); return 0;
    case 1:
      if constexpr (K % 8 == 0) { *out = huffman::CompressMultiAvx512Gather<K>(raw); return 0; }
      return -2;
    case 2:
      if constexpr (K % 8 == 0) { *out = huffman::CompressMultiAvx512Permute<K>(raw); return 0; }
      return -2;
  }
  return -2;
}

template <int K>
int DecompressK(int variant, std::string_view comp, std::string* out) {
  switch (variant) {
    case 0: *out = huffman::DecompressMulti<K>(comp); return 0;
    case 1:
      if constexpr (K % 8 == 0) { *out = huffman::DecompressMultiAvx512Gather<K>(comp); return 0; }
      return -2;
    case 2:
      if constexpr (K % 8 == 0) { *out = huffman::DecompressMultiAvx512Permute<K>(comp); return 0; }
      return -2;
  }
  return -2;
}

int Compress(int k, int variant, std::string_view raw, std::string* out) {
  switch (k) {
    case 1: return CompressK<1>(variant, raw, out);
    case 2: return CompressK<2>(variant, raw, out);
    case 4: return CompressK<4>(variant, raw, out);
    case 8: return CompressK<8>(variant, raw, out);
    case 16: return CompressK<16>(variant, raw, out);
    case 24: return variant == 0 ? -2 : CompressK<24>(variant, raw, out);
    case 32: return CompressK<32>(variant, raw, out);
    case 40: return variant == 0 ? -2 : CompressK<40>(variant, raw, out);
    case 48: return CompressK<48>(variant, raw, out);
  }
  return -2;
}

int Decompress(int k, int variant, std::string_view comp, std::string* out) {
  switch (k) {
    case 1: return DecompressK<1>(variant, comp, out);
    case 2: return DecompressK<2>(variant, comp, out);
    case 4: return DecompressK<4>(variant, comp, out);
    case 8: return DecompressK<8>(variant, comp, out);
    case 16: return DecompressK<16>(variant, comp, out);
    case 24: return variant == 0 ? -2 : DecompressK<24>(variant, comp, out);
    case 32: return DecompressK<32>(variant, comp, out);
    case 40: return variant == 0 ? -2 : DecompressK<40>(variant, comp, out);
    case 48: return DecompressK<48>(variant, comp, out);
  }
  return -2;
}

}  // namespace

// Scalar K=48 is not instantiated by the reference; do it here (the template
// definitions are visible because huffman.cpp is included above).
template std::string huffman::CompressMulti<48>(std::string_view);
template std::string huffman::DecompressMulti<48>(std::string_view);

extern "C" {

// variant: 0 scalar (CompressMulti), 1 AVX-512 gather, 2 AVX-512 permute.
// Returns 0, -1 if cap is too small (out_len still set), -2 unsupported (k, variant).
int ref_compress(int k, int variant, const uint8_t* raw, size_t n, uint8_t* out, size_t cap,
                 size_t* out_len) {
  std::string s;
  int rc = Compress(k, variant, std::string_view(reinterpret_cast<const char*>(raw), n), &s);
  if (rc) return rc;
  *out_len = s.size();
  if (s.size() > cap) return -1;
  memcpy(out, s.data(), s.size());
  return 0;
}

int ref_decompress(int k, int variant, const uint8_t* comp, size_t n, uint8_t* out, size_t cap,
                   size_t* out_len) {
  std::string s;
  int rc = Decompress(k, variant, std::string_view(reinterpret_cast<const char*>(comp), n), &s);
  if (rc) return rc;
  *out_len = s.size();
  if (s.size() > cap) return -1;
  memcpy(out, s.data(), s.size());
  return 0;
}

// which: 0 MakeHistogram, 1 Simple, 2 Multi, 3 Vectorized, 4 GatherScatter.
int ref_histogram(int which, const uint8_t* in, size_t n, uint32_t out[256]) {
  std::string_view v(reinterpret_cast<const char*>(in), n);
  ByteHistogram h;
  switch (which) {
    case 0: h = huffman::MakeHistogram(v); break;
    case 1: h = huffman::MakeHistogramSimple(v); break;
    case 2: h = huffman::MakeHistogramMulti(v); break;
    case 3: h = huffman::MakeHistogramVectorized(v); break;
    case 4: h = huffman::MakeHistogramGatherScatter(v); break;
    default: return -2;
  }
  memcpy(out, h.data(), sizeof(uint32_t) * 256);
  return 0;
}

// MakeCanonicalCoding (codec/huffman.cpp:339-437) on a caller-supplied histogram.
// code_bits/code_len are indexed by symbol; len_count has 13 entries.
int ref_make_coding(const uint32_t hist[256], uint16_t len_count[13], uint8_t sorted_syms[256],
                    int* num_syms, uint32_t* len_mask, uint16_t code_bits[256],
                    uint16_t code_len[256]) {
  ByteHistogram h;
  memcpy(h.data(), hist, sizeof(uint32_t) * 256);
  huffman::CanonicalCoding c = huffman::MakeCanonicalCoding(h);
  for (int i = 0; i <= huffman::kMaxCodeLength; ++i) len_count[i] = c.len_count[i];
  memcpy(sorted_syms, c.sorted_syms, 256);
  *num_syms = c.num_syms;
  *len_mask = c.len_mask;
  for (int i = 0; i < 256; ++i) {
    code_bits[i] = c.codes[i].bits;
    code_len[i] = c.codes[i].len;
  }
  return 0;
}

// LimitCodeLengths (codec/huffman.cpp:297-327) on a 33-entry length histogram, in place.
void ref_limit_code_lengths(uint16_t len_count[33]) { huffman::LimitCodeLengths(len_count); }

// Decoder2x table (codec/huffman.cpp:642-704): 4096 entries of 4 bytes
// {num_bits, sym0, sym1, num_syms}.  Entries the reference never fills are
// reported as they are in a zero-initialised table.
int ref_dtable2x(const uint16_t len_count[13], const uint8_t* syms, int num_syms,
                 uint8_t out[4096 * 4]) {
  huffman::Decoder2x dec(len_count, syms, num_syms);
  memcpy(out, dec.dtable(), 4096 * 4);
  return 0;
}

// Decoder1x table (codec/huffman.cpp:594-632): 4096 entries {code_len, sym}.
int ref_dtable1x(const uint16_t len_count[13], const uint8_t* syms, int num_syms,
                 uint8_t out[4096 * 2]) {
  huffman::Decoder1x dec(len_count, syms, num_syms);
  memcpy(out, dec.dtable(), 4096 * 2);
  return 0;
}

// The comparator-only std::sort of codec/huffman.cpp:353-354, exposed so the C
// restatement of libstdc++'s introsort can be property-tested against it.
void ref_sort_syms(const uint32_t hist[256], uint8_t* syms, int n) {
  std::sort(syms, syms + n, [&](uint8_t a, uint8_t b) { return hist[a] > hist[b]; });
}

// ---- input generators restated from the reference's tests/benchmarks (they
// live in files that need gtest / google-benchmark, which are not installed) ----

// GenerateProbaData, codec/huffman_benchmark.cpp:27-36.
void ref_gen_proba(double p, uint8_t* out, size_t len) {
  std::mt19937_64 mt;
  std::uniform_real_distribution<> dist(0.0, 1.0);
  double logp = log(1 - p);
  for (size_t i = 0; i < len; ++i) out[i] = (uint8_t)(char)(int(log(dist(mt)) / logp) % 256);
}

// which: 0 LongRandom / BM_*Short bytes `rand()&rand()&rand()` after srand(0)
//          (codec/huffman_test.cpp:115-131, codec/huffman_benchmark.cpp:143-150)
//        1 uniform `uint8_t(rand())` after srand(0) (codec/huffman_benchmark.cpp:110-116)
void ref_gen_rand(int which, uint8_t* out, size_t len) {
  srand(0);
  for (size_t i = 0; i < len; ++i)
    out[i] = which == 0 ? uint8_t((rand() & rand() & rand()) & 0xff) : uint8_t(rand());
}

// EqualCounts, codec/huffman_test.cpp:100-113: 4x each byte, shuffled with default mt19937.
void ref_gen_equal_counts(uint8_t out[1024]) {
  std::string raw;
  for (int i = 0; i < 4; ++i)
    for (int c = 0; c < 256; ++c) raw.push_back(c);
  std::shuffle(raw.begin(), raw.end(), std::mt19937());
  memcpy(out, raw.data(), 1024);
}

// ManyRandom, codec/huffman_test.cpp:164-184: 100 strings; writes them back to back
// into out (cap >= 100000) and their lengths into lens[100]. Returns total bytes.
size_t ref_gen_many_random(uint8_t* out, int lens[100]) {
  std::mt19937 mt;
  size_t pos = 0;
  for (int k = 0; k < 100; ++k) {
    int len = 1 + mt() % 1000;
    lens[k] = len;
    for (int i = 0; i < len; ++i) {
      uint8_t ch = 0;
      do {
        ch = (mt() & mt()) & 0xff;
        ch ^= 'A';
      } while (!std::isprint(ch));
      out[pos++] = ch;
    }
  }
  return pos;
}

// Histogram benchmark's skewed input (codec/histogram_benchmark.cpp:27-41): 2^i copies
// of 'A'+i, i<18, shuffled with default mt19937; first `len` bytes (len <= 262143).
void ref_gen_hist_biased(uint8_t* out, size_t len) {
  std::string text;
  for (int i = 0; i < 18; ++i)
    for (int j = 0; j < (1 << i); ++j) text.push_back('A' + i);
  std::shuffle(text.begin(), text.end(), std::mt19937());
  memcpy(out, text.data(), len < text.size() ? len : text.size());
}

// ---- timing harness: chrono re-creation of BM_CompressBiased / BM_DecompressBiased
// (codec/huffman_benchmark.cpp:61-81; google-benchmark is not installed) ----

// Runs `threads` host threads; each loops over the `n_bufs` buffers (stride
// `stride`, length `len` each) calling the chosen codec until `min_seconds`
// elapsed; returns total raw bytes processed per second over all threads.
// dir: 0 compress, 1 decompress (buffers are compressed once up front).
double ref_bench(int k, int variant, int dir, const uint8_t* bufs, size_t len, size_t stride,
                 int n_bufs, int threads, double min_seconds, double* out_ratio) {
  std::vector<std::string> comp(n_bufs);
  size_t csum = 0;
  for (int i = 0; i < n_bufs; ++i) {
    if (Compress(k, 0, std::string_view(reinterpret_cast<const char*>(bufs + i * stride), len),
                 &comp[i]))
      return -1.0;
    csum += comp[i].size();
  }
  if (out_ratio) *out_ratio = double(csum) / (double(len) * n_bufs);
  std::vector<double> bytes(threads, 0.0), secs(threads, 0.0);
  auto work = [&](int t) {
    std::string tmp;
    // one untimed warm-up call
    if (dir == 0) Compress(k, variant, std::string_view(reinterpret_cast<const char*>(bufs), len), &tmp);
    else Decompress(k, variant, comp[0], &tmp);
    auto t0 = std::chrono::steady_clock::now();
    double done = 0;
    int i = t % n_bufs;
    for (;;) {
      if (dir == 0)
        Compress(k, variant, std::string_view(reinterpret_cast<const char*>(bufs + i * stride), len), &tmp);
      else
        Decompress(k, variant, comp[i], &tmp);
      done += len;
      i = (i + 1) % n_bufs;
      double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (el >= min_seconds) { secs[t] = el; break; }
    }
    bytes[t] = done;
  };
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
  for (auto& x : th) x.join();
  double rate = 0;
  for (int t = 0; t < threads; ++t) rate += bytes[t] / secs[t];
  return rate;
}

// Histogram timing: which as in ref_histogram.
double ref_bench_histogram(int which, const uint8_t* in, size_t n, int threads, double min_seconds) {
  std::vector<double> rate(threads, 0.0);
  auto work = [&](int t) {
    uint32_t h[256];
    ref_histogram(which, in, n, h);
    auto t0 = std::chrono::steady_clock::now();
    double done = 0;
    volatile uint32_t sink = 0;
    for (;;) {
      ref_histogram(which, in, n, h);
      sink = sink + h[7];
      done += n;
      double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (el >= min_seconds) { rate[t] = done / el; break; }
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
  for (auto& x : th) x.join();
  double r = 0;
  for (double x : rate) r += x;
  return r;
}

}  // extern "C"
