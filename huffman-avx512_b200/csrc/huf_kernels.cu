// huf_kernels.cu -- sm_100a kernels of the multi-stream Huffman hot path.
//
//   k_histogram          byte histogram, lane-privatised shared-memory bins, 128-bit loads
//   k_build_table        canonical table from a 256 x u64 histogram (shared-table mode)
//   k_make_table         same from a u32 histogram, dumping every field (ABI/table parity)
//   k_compress_blocks    fused per-block histogram -> table -> encode: warp-specialised CTA, blocks
//                        handed out dynamically, streams drawn by ticket, staged in shared memory
//                        and placed through an mbarrier chain (long streams piecewise after a length pass)
//   k_decompress_blocks  header parse (validated) -> multi-symbol table with a 9/10/11-bit first
//                        level and one entry per longer code -> one lane per stream decode
//   k_dump_dtable        the decode kernel's table builder in the reference's formats
//                        (Decoder1x, Decoder2x), for parity tests
//   k_histogram_streams, k_single_plan, k_piece_lengths, k_encode_pieces
//                        one large buffer spread over the device
//   k_scan_sizes/k_pack  slot layout -> packed layout
//
// Reference behaviour: ahartik/huffman-avx512 codec/huffman.cpp, codec/histogram.cpp.
#include <atomic>
#include <cstdlib>
#include <type_traits>

#include "huf_device.cuh"
#include "huf_kernels.h"

namespace hufb200 {

// ===========================================================================
// Histogram (MakeHistogram, codec/histogram.cpp:193-201; 8 private tables there,
// :18-20 -- here 32 lane-private columns so that no two lanes of a warp ever
// touch the same bank or address, whatever the symbol distribution).
// bins[sym * 32 + lane]
// ===========================================================================
// One `red.shared.add` per byte; the bin address is formed by a byte extract (PRMT) and one
// shift-add (LEA) on the lane's 32-bit shared-space base.
// Cache modifiers of the kernels' global accesses.  Every input byte is used once per pass and
// every output byte is written once, so nothing should take a line of the small L1 that is left
// next to the shared-memory carve-out (measured: profiles/r2_cache_modifiers.md).
#ifndef HUF_HIST_LDMOD
#define HUF_HIST_LDMOD ".L1::no_allocate"
#endif
#ifndef HUF_ENC_LDMOD
#define HUF_ENC_LDMOD ".L1::no_allocate"
#endif
#ifndef HUF_COMP_STMOD
#define HUF_COMP_STMOD ".L1::no_allocate"
#endif
__device__ __forceinline__ uint4 ldg16_hist(const uint4* p) {
  uint4 v;
  asm volatile("ld.global" HUF_HIST_LDMOD ".v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ldg16_enc(const void* p) {
  uint4 v;
  asm volatile("ld.global" HUF_ENC_LDMOD ".v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_out(uint32_t* p, uint32_t v) {
  asm volatile("st.global" HUF_COMP_STMOD ".u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void bins_add_word(uint32_t* bins_lane, uint32_t w) {
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(bins_lane);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t a = base + (__byte_perm(w, 0, 0x4440 + i) << 7);
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
  }
}
__device__ __forceinline__ void bins_add_vec(uint32_t* bins_lane, const uint4& v) {
  bins_add_word(bins_lane, v.x);
  bins_add_word(bins_lane, v.y);
  bins_add_word(bins_lane, v.z);
  bins_add_word(bins_lane, v.w);
}
// column sum of bin `t`, rotated start so the 32 threads of a warp hit distinct banks
__device__ __forceinline__ uint32_t bins_reduce(const uint32_t* bins, int t) {
  uint32_t s = 0;
#pragma unroll 8
  for (int i = 0; i < 32; ++i) s += bins[(t << 5) + ((i + t) & 31)];
  return s;
}
// same, leaving the bins zeroed (the compress kernel keeps its bins / staging union all-zero
// between phases, so neither needs a clearing pass)
__device__ __forceinline__ uint32_t bins_reduce_clear(uint32_t* bins, int t) {
  uint32_t s = 0;
#pragma unroll 8
  for (int i = 0; i < 32; ++i) {
    uint32_t* p = bins + (t << 5) + ((i + t) & 31);
    s += *p;
    *p = 0;
  }
  return s;
}

// Accumulates the bytes [p, p+n) into the CTA's lane-private bins (all threads).
// (t, nt): index of the calling thread among the nt threads that take part (whole warps).
__device__ __forceinline__ void bins_accumulate(uint32_t* bins, const uint8_t* p, uint64_t n, uint32_t t, uint32_t nt) {
  uint32_t* bl = bins + lane_id();
  const uint64_t mis = (16 - ((uintptr_t)p & 15)) & 15;
  const uint64_t head = mis < n ? mis : n;
  for (uint64_t i = t; i < head; i += nt) atomicAdd(bl + ((uint32_t)p[i] << 5), 1u);
  const uint4* v = reinterpret_cast<const uint4*>(p + head);
  const uint64_t nvec = (n - head) >> 4;
  uint64_t i = t;
  const uint64_t step = nt;
  if (i + 3 * step < nvec) {  // batches of 4 loads per thread, the next batch in flight while one is counted
    uint4 a = ldg16_hist(v + i), b = ldg16_hist(v + i + step), c = ldg16_hist(v + i + 2 * step), d = ldg16_hist(v + i + 3 * step);
    i += 4 * step;
    for (; i + 3 * step < nvec; i += 4 * step) {
      const uint4 a2 = ldg16_hist(v + i), b2 = ldg16_hist(v + i + step), c2 = ldg16_hist(v + i + 2 * step), d2 = ldg16_hist(v + i + 3 * step);
      bins_add_vec(bl, a);
      bins_add_vec(bl, b);
      bins_add_vec(bl, c);
      bins_add_vec(bl, d);
      a = a2;
      b = b2;
      c = c2;
      d = d2;
    }
    bins_add_vec(bl, a);
    bins_add_vec(bl, b);
    bins_add_vec(bl, c);
    bins_add_vec(bl, d);
  }
  for (; i < nvec; i += step) bins_add_vec(bl, v[i]);
  const uint64_t done = head + (nvec << 4);
  for (uint64_t j = done + t; j < n; j += nt) atomicAdd(bl + ((uint32_t)p[j] << 5), 1u);
}

__global__ void __launch_bounds__(kHistThreads)
k_histogram(const uint8_t* __restrict__ in, uint64_t n, unsigned long long* __restrict__ out) {
  __shared__ uint32_t bins[256 * 32];
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) bins[i] = 0;
  __syncthreads();
  // contiguous 16-byte-granular share per CTA
  const uint64_t nvec_total = (n + 15) >> 4;
  const uint64_t per = (nvec_total + gridDim.x - 1) / gridDim.x;
  const uint64_t beg = (uint64_t)blockIdx.x * per * 16;
  if (beg < n) {
    uint64_t len = per * 16;
    if (beg + len > n) len = n - beg;
    // 32-bit lane counters: the launcher sizes the grid so that a CTA's share stays below 2^32 bytes
    bins_accumulate(bins, in + beg, len, threadIdx.x, blockDim.x);
  }
  __syncthreads();
  if (threadIdx.x < 256) {
    const uint32_t s = bins_reduce(bins, threadIdx.x);
    if (s) atomicAdd(out + threadIdx.x, (unsigned long long)s);
  }
}

// ===========================================================================
// Table kernels
// ===========================================================================
__global__ void __launch_bounds__(32) k_build_table(const unsigned long long* __restrict__ hist,
                                                    HufTable* __restrict__ out) {
  __shared__ HufTable tab;
  __shared__ TableScratch sc;
  __shared__ unsigned long long h[256];
  for (int i = threadIdx.x; i < 256; i += 32) h[i] = hist[i];
  __syncwarp();
  build_table_warp<unsigned long long, unsigned long long>(h, &tab, &sc);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&tab);
  uint32_t* dst = reinterpret_cast<uint32_t*>(out);
  for (int i = threadIdx.x; i < (int)(sizeof(HufTable) / 4); i += 32) dst[i] = src[i];
}

// mode 0: build from hist (u32). mode 1: build from (len_count, syms).
__global__ void __launch_bounds__(32) k_make_table(const uint32_t* __restrict__ hist,
                                                   const uint16_t* __restrict__ in_len_count,
                                                   const uint8_t* __restrict__ in_syms, int in_n,
                                                   int mode, HufTable* __restrict__ out) {
  __shared__ HufTable tab;
  __shared__ TableScratch sc;
  __shared__ uint32_t h[256];
  __shared__ uint16_t lc[16];
  __shared__ uint8_t sy[256];
  if (mode == 0) {
    for (int i = threadIdx.x; i < 256; i += 32) h[i] = hist[i];
    __syncwarp();
    build_table_warp<uint32_t, unsigned long long>(h, &tab, &sc);
  } else {
    for (int i = threadIdx.x; i < 13; i += 32) lc[i] = in_len_count[i];
    for (int i = threadIdx.x; i < in_n; i += 32) sy[i] = in_syms[i];
    __syncwarp();
    table_from_lengths_warp(lc, sy, in_n, &tab, &sc);
  }
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&tab);
  uint32_t* dst = reinterpret_cast<uint32_t*>(out);
  for (int i = threadIdx.x; i < (int)(sizeof(HufTable) / 4); i += 32) dst[i] = src[i];
}

// ===========================================================================
// Fused compress kernel: one CTA per block (CompressMulti<K>, codec/huffman.cpp:738-846)
// ===========================================================================
// Staged mode (slices of at most kStageSlice symbols): a warp keeps the whole bitstream of its
// stream in shared memory, so the stream's size is known without a separate length pass.
constexpr int kStageSlice = 4096;
constexpr int kStageWords = 1280;  // 5 KiB per warp = 10 bits/symbol of a 4096-symbol slice; the first 256 words (1 KiB-aligned)
                                   // double as the ring of the per-stream fallback encoder
constexpr int kStageFront = 2;     // spare words before stream word 0 (see stage_put64_end)

struct CompSmem {
  union {
    uint32_t bins[256 * 32];                  // histogram phase
    uint32_t stage[kCompWarps][kStageWords];  // encode phase, staged mode (each 1 KiB-aligned)
  } u;
  uint32_t hist[2][256];  // [cur] block being encoded, [cur^1] the CTA's next block
  HufTable tab[2];        // same double buffering: the builder warp fills [cur^1] during the encode of [cur]
  TableScratch sc;
  unsigned long long stream_bits[kMaxK];
  uint32_t region_end[kMaxK];  // cumulative end offsets relative to the payload start (:772-786)
  uint32_t bad;
  uint32_t blk_new, blk_new2;  // block indices just drawn from the launch's counter
  uint32_t ticket;             // next stream of the current block to be encoded (staged mode)
  unsigned long long placed[kMaxK];  // mbarriers: [s] completes once per staged block, when region_end[s] is known
};

// Small blocks (k_compress_small_blocks): a batch of kBatch blocks per CTA and iteration.  Same
// layout and size as CompSmem, except that the table builds' scratch areas alias the bins /
// staging union (idle while tables are built) and the room of CompSmem::sc holds two more
// histograms and tables.
constexpr int kBatch = 4;
struct CompSmemBatch {
  union {
    uint32_t bins[256 * 32];
    uint32_t stage[kCompWarps][kStageWords];
    TableScratch sc[kBatch];  // dirty after a build: zeroed again before the union's next use
  } u;
  uint32_t hist[kBatch][256];
  HufTable tab[kBatch];
  unsigned long long stream_bits[kMaxK];
  uint32_t region_end[kMaxK];
  uint32_t bad;
  uint32_t blk_new, blk_new2;
  uint32_t ticket;
  unsigned long long placed[kMaxK];
};
static_assert(sizeof(TableScratch) * kBatch <= sizeof(uint32_t) * 256 * 32, "the scratch areas fit the union");
static_assert(sizeof(CompSmemBatch) <= sizeof(CompSmem), "the batch kernel keeps the residency of the block kernel");

__device__ __forceinline__ uint32_t shl_c(uint32_t x, uint32_t s) {  // shift >= 32 gives 0
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
  return r;
}
__device__ __forceinline__ uint32_t shr_c(uint32_t x, uint32_t s) {
  uint32_t r;
  asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
  return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void red_or_shared(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// table base + 4 * index as one multiply-add: it runs on the FMA pipe, next to the lookup loop's
// many ALU-pipe shifts and logic ops
__device__ __forceinline__ uint32_t entry_addr(uint32_t base, uint32_t idx) {
  uint32_t a;
  asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(a) : "r"(idx), "r"(base));
  return a;
}

template <int M>
__device__ __forceinline__ uint32_t mad_u32(uint32_t x, uint32_t c) {  // x * M + c, kept a multiply-add (FMA pipe)
  uint32_t a;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(x), "n"(M), "r"(c));
  return a;
}
__device__ __forceinline__ uint32_t mad_u32_rr(uint32_t x, uint32_t m, uint32_t c) {
  uint32_t a;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(x), "r"(m), "r"(c));
  return a;
}
// read-only data (the decode table inside the lookup loop): not volatile, no memory clobber, so
// the compiler may schedule it across the loop's other shared-memory traffic
__device__ __forceinline__ uint32_t lds_u32_ro(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// byte i of w, zero-extended (one PRMT)
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int i) { return __byte_perm(w, 0, 0x4440 + i); }

// Loads the 16 symbols [off, off+16) of a slice as four little-endian words; symbols
// beyond `valid` read as 0 and are masked by the caller.  Slices that do not start on a
// 16-byte boundary (K does not divide the block nicely, e.g. K = 48) take two aligned 128-bit
// loads and a byte funnel shift; the misalignment is the same for every lane and iteration of
// a stream, so the switch is warp-uniform.  `lim` = end of the input buffer (no reads past it).
__device__ __forceinline__ uint4 load16(const uint8_t* sp, uint32_t off, uint32_t valid, bool aligned,
                                        const uint8_t* lim) {
  const uint8_t* a = sp + off;
  if (valid >= 16) {
    if (aligned) return ldg16_enc(a);
    const uintptr_t base = (uintptr_t)a & ~(uintptr_t)15;
    if (base + 32 <= (uintptr_t)lim) {
      const uint4 A = *reinterpret_cast<const uint4*>(base);
      const uint4 B = *reinterpret_cast<const uint4*>(base + 16);
      const uint32_t sh = 8u * (uint32_t)((uintptr_t)a & 3);
      switch (((uintptr_t)a >> 2) & 3) {
        case 0:
          return make_uint4(__funnelshift_r(A.x, A.y, sh), __funnelshift_r(A.y, A.z, sh),
                            __funnelshift_r(A.z, A.w, sh), __funnelshift_r(A.w, B.x, sh));
        case 1:
          return make_uint4(__funnelshift_r(A.y, A.z, sh), __funnelshift_r(A.z, A.w, sh),
                            __funnelshift_r(A.w, B.x, sh), __funnelshift_r(B.x, B.y, sh));
        case 2:
          return make_uint4(__funnelshift_r(A.z, A.w, sh), __funnelshift_r(A.w, B.x, sh),
                            __funnelshift_r(B.x, B.y, sh), __funnelshift_r(B.y, B.z, sh));
        default:
          return make_uint4(__funnelshift_r(A.w, B.x, sh), __funnelshift_r(B.x, B.y, sh),
                            __funnelshift_r(B.y, B.z, sh), __funnelshift_r(B.z, B.w, sh));
      }
    }
  }
  uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if ((uint32_t)i < valid) w[i >> 2] |= (uint32_t)a[i] << (8 * (i & 3));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Sum of the four entries of the symbols in w: code lengths add up in bits 16.., the low
// halves (codes < 2^12) cannot carry into them.
__device__ __forceinline__ uint32_t word_entries(const uint32_t* enc, uint32_t w) {
  return enc[byte_of(w, 0)] + enc[byte_of(w, 1)] + enc[byte_of(w, 2)] + enc[byte_of(w, 3)];
}

// Per-stream bit total (the reference derives it from per-stream histograms, :776-782).
__device__ inline unsigned long long stream_length_warp(const uint32_t* enc, const uint8_t* sp,
                                                        uint32_t sz, uint32_t* bad, const uint8_t* lim) {
  const int lane = lane_id();
  const bool aligned = (((uintptr_t)sp) & 15) == 0;
  unsigned long long acc = 0;
  uint32_t flag = 0;
  const uint32_t full = sz & ~511u;
  for (uint32_t base = 0; base < full; base += 512) {
    const uint4 v = load16(sp, base + lane * 16, 16, aligned, lim);
    // 16 entries: lengths <= 16*12 in bits 16..23, kEncInvalid entries pile up in bits 30+
    const uint32_t s = ((word_entries(enc, v.x) >> 16) + (word_entries(enc, v.y) >> 16)) +
                       ((word_entries(enc, v.z) >> 16) + (word_entries(enc, v.w) >> 16));
    flag |= s;
    acc += s;
  }
  if (full < sz) {
    const uint32_t off = full + lane * 16;
    const uint32_t valid = off < sz ? (sz - off < 16 ? sz - off : 16) : 0;
    for (uint32_t i = 0; i < valid; ++i) {
      const uint32_t l = enc[sp[off + i]] >> 16;
      flag |= l;
      acc += l;
    }
  }
  if (flag & 0xfffff000u) atomicOr(bad, 1u);  // a symbol without a code (kEncInvalid)
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  return acc;
}

// ORs the `len` (<= 32) low bits of `code` into the warp's ring at stream bit position `pos`.
// Bits of `code` above `len` are ignored.  ring_base: shared-space address, 1 KiB aligned.
__device__ __forceinline__ void ring_put(uint32_t ring_base, uint32_t pos, uint32_t code, uint32_t len) {
  const uint32_t t = shl_c(code, 32u - len);  // left-aligned; len == 0 gives 0
  const uint32_t o = pos & 31u;
  const uint32_t a0 = ring_base | ((pos >> 3) & 0x3fcu);
  const uint32_t a1 = ring_base | ((a0 + 4u) & 0x3fcu);
  red_or_shared(a0, t >> o);
  red_or_shared(a1, shl_c(t, 32u - o));
}

// Same into a linear (non-wrapping) staging buffer.
__device__ __forceinline__ void stage_put(uint32_t stage_base, uint32_t pos, uint32_t code, uint32_t len) {
  const uint32_t t = shl_c(code, 32u - len);
  const uint32_t o = pos & 31u;
  const uint32_t a0 = stage_base + ((pos >> 3) & ~3u);
  red_or_shared(a0, t >> o);
  red_or_shared(a0 + 4u, shl_c(t, 32u - o));
}

// Up to 64 bits (eight symbols) at once: three words.  Bits of `code` above `len` are ignored.
__device__ __forceinline__ void stage_put64(uint32_t stage_base, uint32_t pos, unsigned long long code, uint32_t len) {
  unsigned long long t;
  asm("shl.b64 %0, %1, %2;" : "=l"(t) : "l"(code), "r"(64u - len));  // left-aligned; len == 0 gives 0
  const uint32_t hi = (uint32_t)(t >> 32), lo = (uint32_t)t;
  const uint32_t o = pos & 31u;
  const uint32_t a0 = stage_base + ((pos >> 3) & ~3u);
  red_or_shared(a0, hi >> o);
  red_or_shared(a0 + 4u, __funnelshift_r(lo, hi, o));
  red_or_shared(a0 + 8u, shl_c(lo, 32u - o));
}

// Up to 64 bits whose LAST bit sits just before stream bit position `end`: the value is taken
// right-aligned (bits above its length must be zero), so nothing has to be left-aligned first:
// with r = end % 32 the three words ending at word end/32 are hi >> r, (hi:lo) >> r and
// lo << (32 - r).  Words that get nothing receive an OR with 0 (hence two spare words in front
// of the stream, kStageFront).
__device__ __forceinline__ void stage_put64_end(uint32_t stream_base, uint32_t end, uint32_t hi, uint32_t lo) {
  const uint32_t a = entry_addr(stream_base, end >> 5);
  red_or_shared(a - 8u, __funnelshift_r(hi, 0u, end));  // hi >> r
  red_or_shared(a - 4u, __funnelshift_r(lo, hi, end));  // shift amount = end & 31
  red_or_shared(a, __funnelshift_r(0u, lo, end));       // lo << (32 - r), 0 for r == 0
}

// Four table entries (first symbol first) -> the two pair codes and lengths.  c01 may carry
// garbage above l01 bits (it always ends up left-aligned by a shift); c23 is clean.
template <bool CLEAN = false>
__device__ __forceinline__ void quad_code(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3, uint32_t& c01,
                                          uint32_t& l01, uint32_t& c23, uint32_t& l23) {
  const uint32_t l1 = e1 >> 16, l3 = e3 >> 16;
  c01 = ((CLEAN ? (e0 & 0xffffu) : e0) << l1) | (e1 & 0xffffu);
  c23 = ((e2 & 0xffffu) << l3) | (e3 & 0xffffu);
  l01 = (e0 + e1) >> 16;
  l23 = (e2 + e3) >> 16;
}

// Encodes one stream (slice sp[0..sz)) whose region ends at byte offset e_off of dst and is
// `region` bytes long (CodeWriter semantics, codec/huffman.cpp:439-500: MSB-first bitstream
// whose first byte is the region's LAST byte; the first 8 bytes of the region stay zero).
//
// Stream word q (32 stream bits, MSB first) occupies bytes e_off-4q-4 .. e_off-4q-1 as a
// little-endian u32.  With r = ((e_off-1) & 3) + 1 (1..4) the aligned output word number m at
// dst + (e_off - r) - 4m equals the top half of (W[m-1]:W[m]) << 8r (W[-1] = 0), so the warp
// emits aligned 128-byte rows; complete stream words leave the ring 32 at a time.
__device__ inline void encode_stream_warp(const uint32_t* enc, uint32_t ring_base, const uint8_t* sp,
                                          uint32_t sz, unsigned long long bits, uint8_t* dst,
                                          uint32_t e_off, uint32_t region, const uint8_t* lim) {
  const int lane = lane_id();
  const bool aligned = (((uintptr_t)sp) & 15) == 0;
  const uint32_t r = ((e_off - 1u) & 3u) + 1u;
  const uint32_t sh = 8u * r;
  const uint32_t e_al = e_off - r;
  const uint32_t s_al = (e_off - region + 3u) & ~3u;
  const uint32_t m_last = (e_al - s_al) >> 2;  // inclusive
  uint32_t* out_lane = reinterpret_cast<uint32_t*>(dst + e_al) - lane;  // word m at out_lane[lane - m]
  const uint32_t ring_lane = ring_base + 4u * (uint32_t)lane;

  unsigned long long bitpos = 0;
  uint32_t m_next = 0;  // next output word to write
  uint32_t carry = 0;   // W[m_next - 1]
  for (uint32_t base = 0; base < sz; base += 512) {
    const uint32_t off = base + lane * 16;
    const uint32_t valid = off < sz ? (sz - off < 16 ? sz - off : 16) : 0;
    const uint4 v = load16(sp, off, valid, aligned, lim);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t c01[4], l01[4], c23[4], l23[4];
    if (valid == 16) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        quad_code(enc[byte_of(w[j], 0)], enc[byte_of(w[j], 1)], enc[byte_of(w[j], 2)], enc[byte_of(w[j], 3)],
                  c01[j], l01[j], c23[j], l23[j]);
    } else {  // last iteration of the slice: symbols past its end contribute no bits
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = (uint32_t)(4 * j + i) < valid ? enc[byte_of(w[j], i)] : 0u;
        quad_code(e[0], e[1], e[2], e[3], c01[j], l01[j], c23[j], l23[j]);
      }
    }
    uint32_t lq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) lq[j] = l01[j] + l23[j];
    const uint32_t lane_len = (lq[0] + lq[1]) + (lq[2] + lq[3]);
    const uint32_t incl = warp_incl_scan(lane_len);
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t pos = (uint32_t)bitpos + (incl - lane_len);  // only the low bits matter
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (lq[j] <= 32) {  // almost always: one put for four symbols
        ring_put(ring_base, pos, (c01[j] << l23[j]) | c23[j], lq[j]);
      } else {
        ring_put(ring_base, pos, c01[j], l01[j]);
        ring_put(ring_base, pos + l01[j], c23[j], l23[j]);
      }
      pos += lq[j];
    }
    bitpos += total;
    __syncwarp();
    // move complete stream words out of the ring, 32 at a time
    const uint32_t wc = (uint32_t)(bitpos >> 5);
    while (wc - m_next >= 32) {
      const uint32_t a = ring_lane + ((m_next << 2) & 0x3fcu);  // lanes wrap together (m_next % 32 == 0)
      const uint32_t lo = lds_u32(a);
      sts_u32(a, 0);
      uint32_t hi = __shfl_up_sync(0xffffffffu, lo, 1);
      if (lane == 0) hi = carry;
      carry = __shfl_sync(0xffffffffu, lo, 31);
      *(out_lane - m_next) = __funnelshift_lc(lo, hi, sh);
      m_next += 32;
    }
    __syncwarp();
  }
  // tail: remaining complete words, the partial word, padding and the zero slop below the stream
  const uint32_t wtot = (uint32_t)((bits + 31) >> 5);
  for (; m_next <= m_last; m_next += 32) {
    const uint32_t m = m_next + lane;
    uint32_t lo = 0;
    if (m < wtot) {
      const uint32_t a = ring_base + ((m << 2) & 0x3fcu);
      lo = lds_u32(a);
      sts_u32(a, 0);
    }
    uint32_t hi = __shfl_up_sync(0xffffffffu, lo, 1);
    if (lane == 0) hi = carry;
    carry = __shfl_sync(0xffffffffu, lo, 31);
    if (m <= m_last) *(out_lane - m_next) = __funnelshift_lc(lo, hi, sh);
  }
  __syncwarp();
}

// ---- staged encoder -------------------------------------------------------------------------
// What the staging puts may touch: every put stays inside [stream word -2, stream word
// kStageLimitBits/32 + 2).
constexpr uint32_t kStageLimitBits = (uint32_t)(kStageWords - kStageFront - 2) * 32;

__device__ __forceinline__ uint32_t lds_u32_ro_enc2(uint32_t addr) {  // entry of HufTable::enc2 (1 KiB behind enc)
  uint32_t v;
  asm("ld.shared.u32 %0, [%1+1024];" : "=r"(v) : "r"(addr));
  return v;
}
template <int OFF>
__device__ __forceinline__ void red_or_shared_off(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.or.b32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(OFF) : "memory");
}
// Up to 64 bits, TOP-aligned in hi:lo (first bit = bit 31 of hi, everything behind the last bit
// zero), whose first bit goes to stream bit position `pos`: with r = pos % 32 the three words
// from word pos/32 on receive hi >> r, (hi:lo) >> r and lo << (32 - r) -- three funnel shifts
// that take pos as it is (the hardware uses pos % 32).
__device__ __forceinline__ void stage_put64_start(uint32_t stream_base, uint32_t pos, uint32_t hi, uint32_t lo) {
  const uint32_t a = entry_addr(stream_base, pos >> 5);
  red_or_shared_off<0>(a, __funnelshift_r(hi, 0u, pos));
  red_or_shared_off<4>(a, __funnelshift_r(lo, hi, pos));
  red_or_shared_off<8>(a, __funnelshift_r(0u, lo, pos));  // 0 for r == 0
}

// Staged mode, step 1: encode the whole stream into the warp's linear staging buffer (zeroed by
// the previous copy-out).  Returns the stream's bit total (all ones if a symbol has no code);
// *overflow is set when it does not fit (then only the count is valid and the caller falls back
// to the ring path).  enc_addr: shared-space address of the HufTable (enc, then enc2).
//
// A trip handles 512 symbols, 16 per lane (one 128-bit load), with the entries of enc2:
// code << (32 - len) | len, i.e. the code TOP-aligned and its length in the low bits, bits
// 4..19 zero.  Then
//   * a funnel shift takes its amount from the low 5 bits of a register, so "append code b behind
//     code a" is a | (b >> a) with the ENTRY a as the shift amount -- no length is ever
//     extracted -- and the pair's length is the low bits of a + b (the sums of up to 16 entries
//     keep the length total exact in bits 0..7: nothing carries into them);
//   * pair = (e0 | e1 >> e0) & ~15 (the AND, free inside the same LOP3, drops the length bits),
//     quad = pair01 | pair23 >> (e0 + e1): 9 instructions for four symbols, exact whenever the
//     quad has at most 32 bits;
//   * two quads make a top-aligned 64-bit value that goes to the lane's scanned bit position with
//     three red.shared.or (stage_put64_start).
// A quad can have up to 48 bits: its low word (pair23 << (32 - len01), one more funnel shift) is
// kept as well, and a lane with a quad longer than 32 bits puts its four quads one by one
// instead of as two octets -- lengths and positions come from the entry sums either way.  A
// symbol without a code has an enc2 entry of no bits and a marker bit that survives the sums:
// such a stream is reported.
template <bool kPiece = false>
__device__ inline unsigned long long encode_stream_staged_warp(uint32_t enc_addr, uint32_t stage_base,
                                                               const uint8_t* sp, uint32_t sz, bool* overflow,
                                                               const uint8_t* lim, uint32_t bitpos0 = 0) {
  const int lane = lane_id();
  // A slice that does not start on a 16-byte boundary (K does not divide the block nicely, e.g.
  // K = 48) is read from the boundary in front of it: every load is an aligned 128-bit load, and
  // the `mis` symbols in front of the slice -- all in lane 0 of the first group -- contribute no
  // bits.  (The input base and the block size are multiples of 16, so that boundary is never in
  // front of the buffer.)  mis is re-derived from sp where it is needed instead of being kept in
  // a register across the loop.
  constexpr bool aligned = true;
  uint32_t bitpos = kPiece ? bitpos0 : 0u;  // where the first code goes (long streams are staged piecewise, see below); the end position is returned
  bool over = false;
  uint32_t flag = 0;  // OR of the entry sums: bits 12..16 set <=> some symbol had no code
  uint32_t sb = stage_base + 4u * kStageFront;  // stream word 0
  auto entry2 = [&](uint32_t w, int i) { return lds_u32_ro_enc2(entry_addr(enc_addr, byte_of(w, i))); };
  // A trip in two halves, so that the registers holding its 16 symbols are free -- and the load
  // of the trip after next can be issued into them -- before the long half starts.
  // kFull: every lane has 16 symbols; else `valid` of them (symbols past the slice's end contribute no bits)
  // kFull: every lane has 16 symbols; else the symbols [first, valid) of the 16 (symbols outside the slice contribute no bits)
  auto gather = [&](const uint4& v, auto full_tag, uint32_t valid, uint32_t (&E)[16], uint32_t first = 0) {
    constexpr bool kFull = decltype(full_tag)::value;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        E[4 * j + i] = (kFull || ((uint32_t)(4 * j + i) < valid && (uint32_t)(4 * j + i) >= first)) ? entry2(w[j], i) : 0u;
  };
  auto finish = [&](const uint32_t (&E)[16]) {
    uint32_t Q[4], Qlo[4], S[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t* e = &E[4 * j];
      const uint32_t s01 = e[0] + e[1];
      S[j] = s01 + (e[2] + e[3]);
      const uint32_t p01 = (e[0] | __funnelshift_r(e[1], 0u, e[0])) & ~15u;
      const uint32_t p23 = (e[2] | __funnelshift_r(e[3], 0u, e[2])) & ~15u;
      Q[j] = p01 | __funnelshift_r(p23, 0u, s01);
      Qlo[j] = __funnelshift_r(0u, p23, s01);  // bits 33.. of the quad; 0 unless it is longer than 32 bits
    }
    const uint32_t sA = S[0] + S[1], sB = S[2] + S[3];
    const uint32_t T = sA + sB;
    flag |= T;
    // quad longer than 32 bits <=> bit 6 of (its length + 31); lengths sit in bits 0..5 of S
    const uint32_t chk = ((S[0] + 31u) | (S[1] + 31u)) | ((S[2] + 31u) | (S[3] + 31u));
    const uint32_t lane_len = T & 0xffu;
    const uint32_t incl = warp_incl_scan(lane_len);
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (bitpos + total > kStageLimitBits) over = true;  // the same in every lane
    if (!over) {
      uint32_t pos = bitpos + (incl - lane_len);
      if (!(chk & 0x40u)) {
#pragma unroll
        for (int h = 0; h < 4; h += 2) {
          const uint32_t a = S[h] & 0x3fu;  // 0..32: clamping shifts
          stage_put64_start(sb, pos, Q[h] | __funnelshift_rc(Q[h + 1], 0u, a), __funnelshift_rc(0u, Q[h + 1], a));
          if (h == 0) pos += sA & 0x7fu;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          stage_put64_start(sb, pos, Q[j], Qlo[j]);
          pos += S[j] & 0x3fu;
        }
      }
    }
    bitpos += total;
  };

  // full groups of 512 symbols: every lane has 16, nothing to mask; the symbols of the trip after
  // next are requested as soon as a trip's lookups are issued (two trips per iteration, so the
  // two register sets just swap roles).  The loop exists twice: slices that start on a 16-byte
  // boundary (all shapes where K divides the block nicely) take plain 128-bit loads.
  uint32_t groups = (sz + (uint32_t)((uintptr_t)sp & 15u)) >> 9;  // full groups still to do, counted from the boundary
  // keep the trip count and the shared-memory bases in registers: under the kernel's register
  // cap the compiler would otherwise re-derive them from the slice geometry and %warpid every trip
  asm volatile("" : "+r"(groups), "+r"(sb), "+r"(enc_addr));
  {
    constexpr bool kAligned = true;
    const uint32_t mis = (uint32_t)((uintptr_t)sp & 15u);
    const uint8_t* p = sp - mis + lane * 16;  // this lane's 16 symbols of the next group
    auto load = [&](const uint8_t* q) {
      return kAligned ? *reinterpret_cast<const uint4*>(q) : load16(q, 0, 16, false, lim);
    };
    uint4 va = make_uint4(0, 0, 0, 0), vb = va;
    if (groups) va = load(p);
    if (groups >= 2) vb = load(p + 512);
    uint32_t E[16];
    if (mis != 0 && groups) {  // the first group of a slice that starts off a boundary: lane 0 leaves out what is not ours
      gather(va, std::false_type{}, 16, E, lane == 0 ? mis : 0u);
      va = vb;
      if (groups > 2) vb = load(p + 1024);
      finish(E);
      p += 512;
      groups -= 1;
    }
    while (groups >= 2) {
      gather(va, std::true_type{}, 16, E);
      if (groups > 2) va = load(p + 1024);
      finish(E);
      gather(vb, std::true_type{}, 16, E);
      if (groups > 3) vb = load(p + 1536);
      finish(E);
      p += 1024;
      groups -= 2;
    }
    if (groups) {
      gather(va, std::true_type{}, 16, E);
      finish(E);
    }
  }
  // the slice's last, partial group (also its first if the slice is that short)
  {
    const uint32_t mis = (uint32_t)((uintptr_t)sp & 15u);
    const uint32_t szb = sz + mis;  // counted from the boundary
    if (szb & 511u) {
      const uint32_t off = lane * 16 + ((szb >> 9) << 9);  // this lane's symbols of the group
      const uint32_t valid = off < szb ? (szb - off < 16 ? szb - off : 16) : 0;
      uint32_t E[16];
      gather(load16(sp - mis, off, valid, aligned, lim), std::false_type{}, valid, E, (szb >> 9) == 0 && lane == 0 ? mis : 0u);
      finish(E);
    }
  }
  __syncwarp();
  *overflow = over;
  if (!kPiece && __any_sync(0xffffffffu, (flag & 0x1f000u) != 0)) return ~0ull;  // a symbol without a code
  return bitpos;
}

// Staged mode, step 2: the staged stream words go out as aligned 128-byte rows below e_off
// (same alignment algebra as encode_stream_warp): output word m = top half of
// (W[m-1] : W[m]) << 8r with W[-1] = 0 (the zero word in front of the stream) and W[m] = 0 from
// the stream's end on.  A lane reads both words itself (two conflict-free loads; no shuffles),
// and the buffer is zeroed afterwards with 128-bit stores.
__device__ inline void copy_stream_out_warp(uint32_t stage_base, unsigned long long bits, uint8_t* dst,
                                            uint32_t e_off, uint32_t region) {
  const int lane = lane_id();
  const uint32_t r = ((e_off - 1u) & 3u) + 1u;
  const uint32_t sh = 8u * r;
  const uint32_t e_al = e_off - r;
  const uint32_t s_al = (e_off - region + 3u) & ~3u;
  const uint32_t m_last = (e_al - s_al) >> 2;  // inclusive
  const uint32_t wtot = (uint32_t)((bits + 31) >> 5);
  uint32_t* out = reinterpret_cast<uint32_t*>(dst + e_al) - lane;  // this lane's word of the current row
  uint32_t sa = stage_base + 4u * (uint32_t)(kStageFront + lane);  // stream word `lane`
  uint32_t rows = wtot >> 5;  // rows of 32 data words: no predicates
  const uint32_t m_tail = rows << 5;
  for (; rows >= 2; rows -= 2) {
    const uint32_t lo0 = lds_u32(sa), hi0 = lds_u32(sa - 4u);
    const uint32_t lo1 = lds_u32(sa + 128u), hi1 = lds_u32(sa + 124u);
    stg_out(out, __funnelshift_lc(lo0, hi0, sh));
    stg_out(out - 32, __funnelshift_lc(lo1, hi1, sh));
    out -= 64;
    sa += 256;
  }
  if (rows) {
    const uint32_t lo0 = lds_u32(sa), hi0 = lds_u32(sa - 4u);
    stg_out(out, __funnelshift_lc(lo0, hi0, sh));
    out -= 32;
    sa += 128;
  }
  // the rest: remaining data words, the partial word, padding and the zero slop
  for (uint32_t m0 = m_tail; m0 <= m_last; m0 += 32) {
    const uint32_t m = m0 + lane;
    const uint32_t lo = m < wtot ? lds_u32(sa) : 0u;
    const uint32_t hi = m <= wtot ? lds_u32(sa - 4u) : 0u;
    if (m <= m_last) stg_out(out, __funnelshift_lc(lo, hi, sh));
    out -= 32;
    sa += 128;
  }
  __syncwarp();
  // zero what the stream used: 16-byte stores from the buffer's start (two words before the
  // stream) up to the 512-byte line that holds its last word -- never past the warp's buffer
  static_assert(kStageWords % 128 == 0 && kStageFront == 2, "zeroing rows cover whole 512-byte lines of the buffer");
  for (uint32_t i = 4u * (uint32_t)lane; i < wtot + kStageFront; i += 128u)
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(stage_base + 4u * i), "r"(0u) : "memory");
  __syncwarp();
}

// Long streams (a slice that does not fit a staging buffer): the stream's bit total and with it
// its place are known from a first pass, and the stream is staged and written out piece by piece.
// A piece of kPieceSyms symbols fits the buffer whatever its codes are; it is encoded starting at
// the bit position the previous piece ended on (mod 32), on top of that piece's last, partial word
// which stays behind as word 0 -- so the buffer always holds whole stream words, and the row
// formula of copy_stream_out_warp applies with a word offset.
constexpr uint32_t kPieceSyms = 3072;
static_assert(kPieceSyms % 512 == 0 && kPieceSyms * 12 + 31 <= (uint32_t)(kStageWords - kStageFront - 2) * 32,
              "a piece always fits");

// Buffer words [0, nw_emit) go out as stream words m_base.. (words from nw_data on are zero:
// padding and slop of the last piece); the words read are zeroed.  wend = word address of the
// region's aligned end, carry = stream word m_base - 1.
__device__ inline void copy_words_out_warp(uint32_t sb, uint32_t nw_data, uint32_t nw_emit, uint32_t* wend,
                                               uint32_t m_base, uint32_t sh, uint32_t carry) {
  const int lane = lane_id();
  for (uint32_t i0 = 0; i0 < nw_emit; i0 += 32) {
    const uint32_t i = i0 + lane;
    uint32_t lo = 0;
    if (i < nw_data) {
      lo = lds_u32(sb + 4u * i);
      sts_u32(sb + 4u * i, 0);
    }
    uint32_t hi = __shfl_up_sync(0xffffffffu, lo, 1);
    if (lane == 0) hi = carry;
    if (i < nw_emit) stg_out(wend - (m_base + i), __funnelshift_lc(lo, hi, sh));
    carry = __shfl_sync(0xffffffffu, lo, 31);  // only full rows are followed by another row
  }
}

__device__ inline void encode_long_stream_warp(uint32_t enc_addr, uint32_t stage_base, const uint8_t* sp, uint32_t sz,
                                               uint8_t* dst, uint32_t e_off, uint32_t region, const uint8_t* lim) {
  const int lane = lane_id();
  const uint32_t r = ((e_off - 1u) & 3u) + 1u;
  const uint32_t sh = 8u * r;
  const uint32_t e_al = e_off - r;
  const uint32_t s_al = (e_off - region + 3u) & ~3u;
  const uint32_t m_last = (e_al - s_al) >> 2;  // inclusive
  uint32_t* wend = reinterpret_cast<uint32_t*>(dst + e_al);
  const uint32_t sb = stage_base + 4u * kStageFront;
  uint32_t b = 0, m_base = 0, carry = 0;
  uint32_t off = 0;
  do {
    const uint32_t piece = sz - off < kPieceSyms ? sz - off : kPieceSyms;
    const bool last = off + piece == sz;
    bool over;
    const uint32_t endbit = (uint32_t)encode_stream_staged_warp<true>(enc_addr, stage_base, sp + off, piece, &over, lim, b);
    if (!last) {
      const uint32_t nfull = endbit >> 5;
      // the last data word of this piece is needed as the next row's carry: read it before it is zeroed
      const uint32_t tail_word = nfull ? lds_u32(sb + 4u * (nfull - 1)) : carry;
      __syncwarp();
      copy_words_out_warp(sb, nfull, nfull, wend, m_base, sh, carry);
      carry = tail_word;
      __syncwarp();
      if (lane == 0 && nfull != 0) {  // the partial word becomes word 0 of the next piece
        const uint32_t pw = lds_u32(sb + 4u * nfull);
        sts_u32(sb + 4u * nfull, 0);
        sts_u32(sb, pw);
      }
      __syncwarp();
      m_base += nfull;
      b = endbit & 31u;
    } else {
      copy_words_out_warp(sb, (endbit + 31u) >> 5, m_last - m_base + 1u, wend, m_base, sh, carry);
    }
    off += piece;
  } while (off < sz);
  __syncwarp();
}

// Worker-only barrier (named barrier 1): the table-builder warp never takes part in it.
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkThreads) : "memory"); }

// mbarrier (shared memory, CTA scope): one arrival completes a phase; a waiter is suspended by
// the hardware until the phase of the given parity is complete -- no polling loop that would
// compete for issue slots with the warps it waits for.
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {  // release: what was written before is visible to the waiter
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {  // acquire
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// Everything the workers do for one block once its table is in `tab`: header, streams, sizes.
// kLongSlices = false: every slice has at most kStageSlice symbols (decided at launch from the
// geometry), so only the staged mode and its per-stream fallback are compiled in -- the kernel of
// the common shapes carries no code for long streams.
// Returns whether the block went through the staged mode (the caller counts those blocks: the
// count's parity is the phase the placement barriers are in).
template <bool kLongSlices, class Smem>
__device__ inline bool encode_block_workers(Smem& sm, const HufTable& tab, const uint8_t* raw, uint64_t n,
                                            const uint8_t* src, uint32_t bn, uint32_t block_size, int K,
                                            uint8_t* dst, uint32_t* comp_size_out, uint32_t* status,
                                            uint32_t staged_iter) {
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  // Staged mode?  Slices of up to kStageSlice symbols always (a stream that does not fit its
  // staging buffer falls back by itself); longer slices when the table's mean code length says
  // that the streams fit with a margin -- compressible data does at twice the length,
  // incompressible data would only waste the attempt.  The same for every worker: geometry and table only.
  const uint32_t slice_max = (block_size + (uint32_t)K - 1) / (uint32_t)K;
  bool staged = !kLongSlices || slice_max <= (uint32_t)kStageSlice;
  if (kLongSlices && !staged && tab.avg_bits_x256 != 0) {
    const unsigned long long est = ((unsigned long long)slice_max * tab.avg_bits_x256) >> 8;
    staged = est <= (unsigned long long)(kStageWords - kStageFront - 2) * 28;  // 7/8 of the buffer's bits
  }
  // ---- header prefix (:799-808)
  const uint32_t hdr = tab.hdr_len;
  const uint32_t mask = tab.len_mask;
  const uint32_t npop = (uint32_t)__popc(mask);
  for (uint32_t i = tid; i < hdr; i += kWorkThreads) {
    uint8_t v;
    if (i < 4) v = (uint8_t)(bn >> (8 * i));
    else if (i < 8) v = (uint8_t)(mask >> (8 * (i - 4)));
    else if (i < 8 + npop) {
      int bit = 0;  // position of the (i-8)-th set bit of the mask
      for (uint32_t seen = 0;; ++bit)
        if ((mask >> bit) & 1u) {
          if (seen == i - 8) break;
          ++seen;
        }
      v = (uint8_t)tab.len_count[bit];  // 256 wraps to 0 (:804)
    } else v = tab.sorted_syms[i - 8 - npop];
    dst[i] = v;
  }
  const uint32_t hdr_total = hdr + 4u * (uint32_t)(K - 1);
  // SliceSizes<K> (:98-108) without a division per stream
  const uint32_t sl_q = bn / (uint32_t)K, sl_r = bn % (uint32_t)K;
  auto geom = [&](int s, uint32_t& st, uint32_t& sz) {
    const uint32_t us = (uint32_t)s;
    sz = sl_q + (us < sl_r ? 1u : 0u);
    st = us * sl_q + (us < sl_r ? us : sl_r);
  };
  if (tid == 0)
    for (uint32_t a = hdr_total; a & 3u; ++a) dst[a] = 0;  // slop bytes sharing a word with the header

  if (staged) {
    // ---- staged mode: a warp draws the next stream of the block (a ticket), encodes it into its
    // staging buffer -- which gives the stream's size -- and places it behind its predecessor:
    // region ends are cumulative (:772-786), so stream s waits until the end of stream s-1 is
    // published (mbarrier placed[s-1]), adds its own size, publishes, and copies its words out.
    // Tickets go out in stream order, so a predecessor has always started earlier and nobody waits
    // for a warp that waits for it; there is no barrier inside the block.  All K tickets are
    // always drawn (a block flagged bad only skips its global writes): every placement barrier
    // completes exactly once per staged block.
    const uint32_t parity = staged_iter & 1u;
    const uint32_t stage_base = smem_u32(&sm.u.stage[warp][0]);
    for (;;) {
      uint32_t s = 0;
      if (lane == 0) s = atomicAdd(&sm.ticket, 1u);
      s = __shfl_sync(0xffffffffu, s, 0);
      if (s >= (uint32_t)K) break;
      uint32_t st, sz;
      geom((int)s, st, sz);
      bool over = false;
      const unsigned long long bits = encode_stream_staged_warp(smem_u32(tab.enc), stage_base, src + st, sz, &over, raw + n);
      if (bits > 12ull * sz && lane == 0) atomicOr(&sm.bad, 1u);  // a symbol without a code
      const uint32_t my_region = (uint32_t)((bits + 7) >> 3) + kSlop;
      uint32_t prev_end = 0;
      if (s != 0) {
        mbar_wait(smem_u32(&sm.placed[s - 1]), parity);
        prev_end = *reinterpret_cast<volatile uint32_t*>(&sm.region_end[s - 1]);
      }
      const uint32_t my_end = prev_end + my_region;
      if (lane == 0) {
        *reinterpret_cast<volatile uint32_t*>(&sm.region_end[s]) = my_end;
        mbar_arrive(smem_u32(&sm.placed[s]));
      }
      const bool bad_now = *reinterpret_cast<volatile uint32_t*>(&sm.bad) != 0;  // covers every stream up to this one
      if (!bad_now) {
        const uint32_t e_off = hdr_total + my_end;
        if (!over) {
          copy_stream_out_warp(stage_base, bits, dst, e_off, my_region);
        } else {  // rare: more than 10 bits/symbol in this slice -> ring path, now that e_off is known
          for (int j = lane; j < kStageWords; j += 32) sts_u32(stage_base + 4u * j, 0);
          __syncwarp();
          encode_stream_warp(tab.enc, stage_base, src + st, sz, bits, dst, e_off, my_region, raw + n);
        }
      } else {  // leave the staging buffer clean for the next block
        for (int j = lane; j < kStageWords; j += 32) sts_u32(stage_base + 4u * j, 0);
        __syncwarp();
      }
    }
    worker_sync();
  } else if constexpr (kLongSlices) {
    // ---- long slices: per-stream bit totals first (:772-782)
    for (int s = warp; s < K; s += kCompWarps) {
      uint32_t st, sz;
      geom(s, st, sz);
      const unsigned long long bits = stream_length_warp(tab.enc, src + st, sz, &sm.bad, raw + n);
      if (lane == 0) sm.stream_bits[s] = bits;
    }
    worker_sync();
    if (tid == 0) {  // region offsets (:783-786)
      uint32_t pos = 0;
      for (int s = 0; s < K; ++s) {
        pos += (uint32_t)((sm.stream_bits[s] + 7) >> 3) + kSlop;
        sm.region_end[s] = pos;
      }
    }
    worker_sync();
    if (sm.bad == 0) {  // encode, one warp per stream
      for (int s = warp; s < K; s += kCompWarps) {
        uint32_t st, sz;
        geom(s, st, sz);
        const uint32_t e_off = hdr_total + sm.region_end[s];
        const uint32_t region = sm.region_end[s] - (s ? sm.region_end[s - 1] : 0u);
        encode_long_stream_warp(smem_u32(tab.enc), smem_u32(&sm.u.stage[warp][0]), src + st, sz, dst, e_off, region,
                                raw + n);
      }
    }
    worker_sync();
  }
  // ---- end_offset table (:809-811) and the block's size
  if (sm.bad != 0) {
    if (tid == 0) {
      *comp_size_out = 0;
      if (status) atomicOr(status, 1u);
    }
  } else {
    for (int s = tid; s < K - 1; s += kWorkThreads) {
      const uint32_t e = sm.region_end[s];
      uint8_t* p = dst + hdr + 4 * s;
      p[0] = (uint8_t)e;
      p[1] = (uint8_t)(e >> 8);
      p[2] = (uint8_t)(e >> 16);
      p[3] = (uint8_t)(e >> 24);
    }
    if (tid == 0) *comp_size_out = hdr_total + sm.region_end[K - 1];
  }
  return staged;
}

// Work counters of the compress launches.  Every launch takes the next pair round-robin; a pair
// is all-zero when idle (the last CTA of a launch re-arms it), and the ring is far longer than
// the number of launches a device can have in flight.
constexpr uint32_t kCounterSlots = 4096;
__device__ uint32_t g_comp_counters[kCounterSlots][2];

// Warp-specialised CTA: warps 0..7 are workers (histogram, encode), warp 8 builds tables.  While
// the workers encode block b with table[cur], the builder turns the histogram of the CTA's next
// block (counted by the workers just before) into table[cur^1]; its ~6k serial instructions
// disappear behind the encode.
template <bool kLongSlices>
__global__ void __launch_bounds__(kCompThreads, kCompCtasPerSm)
k_compress_blocks(const uint8_t* __restrict__ raw, uint64_t n, uint32_t block_size, int K,
                  uint32_t n_blocks, uint8_t* __restrict__ out, uint64_t slot_stride,
                  uint32_t* __restrict__ comp_sizes, const HufTable* __restrict__ shared_tab,
                  int check_presence, uint32_t* __restrict__ status, uint32_t counter_slot) {
  uint32_t* const counters = g_comp_counters[counter_slot];  // [0] next block, [1] CTAs finished
  extern __shared__ __align__(1024) uint8_t comp_smem[];  // dynamic: the layout exceeds the 48 KiB static limit
  CompSmem& sm = *reinterpret_cast<CompSmem*>(comp_smem);
  if (smem_u32(comp_smem) & 1023u) __trap();  // ring_put/stage buffers rely on 1 KiB alignment
  const int tid = threadIdx.x;
  const bool builder = tid >= kWorkThreads;  // the last warp
  const bool own_tables = shared_tab == nullptr;
  const bool need_hist = own_tables || check_presence;

  auto block_len = [&](uint32_t blk) -> uint32_t {
    const uint64_t o = (uint64_t)blk * block_size;
    return (uint32_t)((n - o) < (uint64_t)block_size ? (n - o) : (uint64_t)block_size);
  };
  // workers: histogram of block `blk` into sm.hist[slot]
  // (the union is all-zero on entry and left all-zero)
  auto histogram_block = [&](uint32_t blk, int slot) {
    bins_accumulate(sm.u.bins, raw + (uint64_t)blk * block_size, block_len(blk), tid, kWorkThreads);
    worker_sync();
    if (tid < 256) sm.hist[slot][tid] = bins_reduce_clear(sm.u.bins, tid);  // one thread per bin
  };
  auto build_table = [&](int slot) {
    if (block_size < (1u << 24)) build_table_warp<uint32_t, uint32_t>(sm.hist[slot], &sm.tab[slot], &sm.sc);
    else build_table_warp<uint32_t, unsigned long long>(sm.hist[slot], &sm.tab[slot], &sm.sc);
  };

  {  // the bins / staging union starts all-zero; every phase leaves it that way
    uint4* z = reinterpret_cast<uint4*>(&sm.u);
    for (int i = tid; i < (int)(sizeof(sm.u) / 16); i += kCompThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (tid < kMaxK) mbar_init(smem_u32(&sm.placed[tid]), 1);  // made visible by the __syncthreads below
  uint32_t staged_iter = 0;  // staged blocks this CTA has encoded (workers only; the same in each)
  // Blocks are handed out dynamically (one atomic per block on a per-launch counter): the CTAs
  // that share an SM do not progress at the same rate, and with a static split the SM would
  // run the last quarter of the kernel with one or two CTAs left.  A CTA always knows its
  // current and its next block (the next one's histogram is pipelined).
  if (tid == 0) {
    sm.blk_new = atomicAdd(&counters[0], 1u);
    sm.blk_new2 = atomicAdd(&counters[0], 1u);
  }
  __syncthreads();
  uint32_t b = sm.blk_new, nxt = sm.blk_new2;
  if (b < n_blocks) {
    // ---- prologue: first block's histogram and table (or the shared table into both slots)
    if (!builder && need_hist) histogram_block(b, 0);
    if (!own_tables) {
      const uint32_t* s = reinterpret_cast<const uint32_t*>(shared_tab);
      for (int t = 0; t < 2; ++t) {
        uint32_t* d = reinterpret_cast<uint32_t*>(&sm.tab[t]);
        for (int i = tid; i < (int)(sizeof(HufTable) / 4); i += kCompThreads) d[i] = s[i];
      }
    }
    __syncthreads();
    if (builder && own_tables) build_table(0);
    __syncthreads();

    int cur = 0;
    while (b < n_blocks) {
      const bool have_next = need_hist && nxt < n_blocks;
      // ---- phase A: the workers count the next block (the staging buffers of the previous block
      //      are dead, the bins alias them); the builder has nothing to do yet
      if (!builder) {
        if (tid == 0) {
          sm.bad = 0;
          sm.ticket = 0;
          sm.blk_new = atomicAdd(&counters[0], 1u);  // the block after the next one
        }
        if (have_next) histogram_block(nxt, cur ^ 1);
      }
      __syncthreads();
      const uint32_t after_next = sm.blk_new;
      // ---- phase B: workers encode block b with table[cur] || builder makes table[cur^1]
      if (builder) {
        if (have_next && own_tables) build_table(cur ^ 1);
      } else {
        if (!own_tables && check_presence) {  // every symbol of the block needs a code in the supplied table
          if (tid < 256 && sm.hist[cur][tid] != 0 && sm.tab[cur].enc[tid] == kEncInvalid) atomicOr(&sm.bad, 1u);
        }
        const uint64_t boff = (uint64_t)b * block_size;
        if (encode_block_workers<kLongSlices>(sm, sm.tab[cur], raw, n, raw + boff, block_len(b), block_size, K,
                                              out + (uint64_t)b * slot_stride, comp_sizes + b, status, staged_iter))
          ++staged_iter;
      }
      __syncthreads();
      b = nxt;
      nxt = after_next;
      cur ^= 1;
    }
  }
  // the last CTA to finish re-arms the counters for the next launch that uses this pair
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&counters[1], 1u) == gridDim.x - 1) {
      counters[0] = 0;
      counters[1] = 0;
      __threadfence();
    }
  }
}

// Small blocks (at most 32 KiB): the table build of a block (12-14 us alone, one warp, serial)
// takes several times longer than the block's histogram and encode (2-5 us on eight warps), so
// k_compress_blocks -- one builder warp per CTA, hidden behind ONE block's encode -- runs at the
// builder's pace.  Here a CTA takes kBatch consecutive blocks per iteration: the workers count
// them one after the other, then kBatch warps build the kBatch tables AT THE SAME TIME (their
// scratch areas lie in the bins / staging union, which is idle then), then the workers encode
// the blocks one after the other.  No warp specialisation beyond the ninth warp taking one of the
// builds; the CTAs that share an SM are in different phases, so the serial builds of one overlap
// the histograms and encodes of the others.
__global__ void __launch_bounds__(kCompThreads, kCompCtasPerSm)
k_compress_small_blocks(const uint8_t* __restrict__ raw, uint64_t n, uint32_t block_size, int K,
                        uint32_t n_blocks, uint8_t* __restrict__ out, uint64_t slot_stride,
                        uint32_t* __restrict__ comp_sizes, uint32_t* __restrict__ status, uint32_t counter_slot) {
  uint32_t* const counters = g_comp_counters[counter_slot];  // [0] next block, [1] CTAs finished
  extern __shared__ __align__(1024) uint8_t comp_smem[];
  CompSmemBatch& sm = *reinterpret_cast<CompSmemBatch*>(comp_smem);
  if (smem_u32(comp_smem) & 1023u) __trap();
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const bool builder = tid >= kWorkThreads;  // the ninth warp: no histogram or encode work
  auto block_len = [&](uint32_t blk) -> uint32_t {
    const uint64_t o = (uint64_t)blk * block_size;
    return (uint32_t)((n - o) < (uint64_t)block_size ? (n - o) : (uint64_t)block_size);
  };
  auto zero_union = [&](int bytes) {
    uint4* z = reinterpret_cast<uint4*>(&sm.u);
    for (int i = tid; i < bytes / 16; i += kCompThreads) z[i] = make_uint4(0, 0, 0, 0);
  };
  zero_union((int)sizeof(sm.u));
  if (tid < kMaxK) mbar_init(smem_u32(&sm.placed[tid]), 1);
  uint32_t staged_iter = 0;
  for (;;) {
    __syncthreads();  // the previous batch is done with sm.blk_new; the union is all-zero
    if (tid == 0) sm.blk_new = atomicAdd(&counters[0], (uint32_t)kBatch);
    __syncthreads();
    const uint32_t base = sm.blk_new;
    if (base >= n_blocks) break;
    const int nb = (int)(n_blocks - base < (uint32_t)kBatch ? n_blocks - base : (uint32_t)kBatch);
    // ---- histograms, one block after the other (the workers)
    if (!builder) {
      for (int j = 0; j < nb; ++j) {
        const uint32_t blk = base + (uint32_t)j;
        bins_accumulate(sm.u.bins, raw + (uint64_t)blk * block_size, block_len(blk), tid, kWorkThreads);
        worker_sync();
        if (tid < 256) sm.hist[j][tid] = bins_reduce_clear(sm.u.bins, tid);
        worker_sync();
      }
    }
    __syncthreads();
    // ---- the tables, all at once: warp 8 builds block 0's, warps 0.. the others'
    {
      const int j = builder ? 0 : warp + 1;
      if (j < nb) build_table_warp<uint32_t, uint32_t>(sm.hist[j], &sm.tab[j], &sm.u.sc[j]);
    }
    __syncthreads();
    zero_union((int)(sizeof(TableScratch) * kBatch + 15) & ~15);
    __syncthreads();
    // ---- encode, one block after the other (the workers)
    if (!builder) {
      for (int j = 0; j < nb; ++j) {
        const uint32_t blk = base + (uint32_t)j;
        if (tid == 0) {
          sm.bad = 0;
          sm.ticket = 0;
        }
        worker_sync();
        if (encode_block_workers<false>(sm, sm.tab[j], raw, n, raw + (uint64_t)blk * block_size, block_len(blk), block_size,
                                        K, out + (uint64_t)blk * slot_stride, comp_sizes + blk, status, staged_iter))
          ++staged_iter;
        worker_sync();  // everybody has read sm.bad / sm.region_end of this block
      }
    }
  }
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&counters[1], 1u) == gridDim.x - 1) {
      counters[0] = 0;
      counters[1] = 0;
      __threadfence();
    }
  }
}

// ===========================================================================
// Decompress kernel: one lane per stream (DecompressMultiImpl<K, Decoder2x>,
// codec/huffman.cpp:892-955; ParseCompressedHeader :714-736; Decoder2x :642-704)
// ===========================================================================
struct DecBlockHdr {
  uint32_t ok;
  uint32_t raw_size;
  uint32_t comp_size;
  uint32_t payload_off;  // header bytes incl. the end_offset table
  uint32_t ends_off;     // offset of the end_offset table
  uint32_t syms_off;
  uint32_t num_syms;
  uint32_t code_end[16];   // left-aligned (12-bit) end of the code range of each length
  uint32_t first_idx[16];  // index into sorted_syms of the first code of each length
};
struct DecBlockInfo : DecBlockHdr {
  uint8_t syms[256];       // sorted_syms copied out of the header: the table builder reads shared memory
};

// Decode table over the next BITS bits of the stream, up to MAXSYM symbols per entry
// (generalises Decoder1x / Decoder2x, codec/huffman.cpp:594-704).
//   entry: byte0..2 = symbols, bits 24..27 = stream bits consumed, bits 30..31 = symbol count.
//   BITS = 12, MAXSYM = 2: exactly the reference's two-symbol table (pair iff l1+l2 <= 12, :653);
//   BITS = 12, MAXSYM = 1: the reference's Decoder1x.  Both are only dumped for parity tests.
//   BITS = 9..11, MAXSYM = 3, EXT: what the decode kernel uses (BITS chosen from the stream count:
//     few streams per block mean few lanes per table, so the table must be small to keep lanes
//     resident).  Up to three symbols per lookup.  BITS bits cannot resolve a longer code; the
//     canonical code orders codes by length, so the codes of each length l > BITS own one
//     contiguous range of the code space, and behind the BITS-bit part (prefixes below P = first
//     longer code >> (12 - BITS)) the table continues with one entry per longer code, length by
//     length.  With p_l = the window's first l bits, level l addresses entry
//     base_l + p_l - (first l-bit code >> (12 - l)), and the right level is the one with the
//     LARGEST index: a coarser level counts the longer codes in fractions of its unit (never more
//     than their number), a finer one is negative below its range.  So the lookup index is
//     max over l = BITS..12 of (p_l + c_l), c_l per block -- no branch, nothing after the load.
//     At most 256 codes in all, so the table has fewer than 2^BITS + 256 entries (2^11 + 128 for
//     BITS = 11, where every dead prefix holds exactly two codes).
// L1 (u8 per entry: the first code's own length, 15 = none) is scratch that may be reused
// afterwards; the first symbol sits in byte 0 of T from the first pass on and never changes.
template <int BITS, int MAXSYM, bool EXT = false>
__device__ inline void build_dtable(const DecBlockHdr* bi, const uint8_t* syms, uint32_t* T, uint8_t* L1,
                                    int tid, int nthreads) {
  constexpr int N = 1 << BITS;
  constexpr int SH = kMaxCodeLen - BITS;
  for (int e = tid; e < N; e += nthreads) {
    const uint32_t v = (uint32_t)e << SH;  // left-aligned 12-bit value of this prefix
    int l = 0;
    while (l <= BITS && v >= bi->code_end[l]) ++l;
    if (l <= BITS) {
      const uint32_t lo = l ? bi->code_end[l - 1] : 0u;
      const uint32_t idx = bi->first_idx[l] + ((v - lo) >> (kMaxCodeLen - l));
      T[e] = idx < bi->num_syms ? syms[idx] : 0u;
      L1[e] = (uint8_t)l;
    } else {
      L1[e] = 15;  // never fits behind another symbol
      // no code of at most BITS bits starts here: a malformed (incomplete) table, or -- EXT --
      // the prefix of 12-bit codes, which the extension below resolves.  One symbol, 12 bits:
      // whatever reaches such an entry makes progress.
      T[e] = MAXSYM == 1 ? 0u : (12u << 24) | (1u << 30);
    }
  }
  __syncthreads();
  for (int e = tid; e < N; e += nthreads) {
    uint32_t nb = L1[e];
    if (nb > (uint32_t)BITS) continue;  // written above
    uint32_t n = 1, out = T[e] & 0xffu;
#pragma unroll
    for (int k = 1; k < MAXSYM; ++k) {
      const uint32_t rest = ((uint32_t)e << nb) & (uint32_t)(N - 1);
      const uint32_t l2 = L1[rest];
      if (nb + l2 > (uint32_t)BITS) break;  // :653
      out |= (T[rest] & 0xffu) << (8 * k);  // byte 0 is stable under concurrent rewrites
      nb += l2;
      ++n;
    }
    T[e] = out | (nb << 24) | (n << 30);
  }
  __syncthreads();
  if (EXT) {  // the codes longer than BITS bits, one entry each, from index P on (over the dead entries)
    uint32_t base = bi->code_end[BITS] >> SH;  // P: the first prefix no code of at most BITS bits owns
#pragma unroll
    for (int l = BITS + 1; l <= kMaxCodeLen; ++l) {
      const uint32_t nl = (bi->code_end[l] - bi->code_end[l - 1]) >> (kMaxCodeLen - l);
      for (uint32_t j = tid; j < nl; j += nthreads) {
        const uint32_t idx = bi->first_idx[l] + j;
        T[base + j] = (idx < bi->num_syms ? syms[idx] : 0u) | ((uint32_t)l << 24) | (1u << 30);
      }
      base += nl;
    }
    __syncthreads();
  }
}

// 2^bits entries + what the longer codes add beyond the dead prefixes they replace (see
// build_dtable): at most 256 codes in all; with 11 bits every dead prefix holds exactly two
__host__ __device__ constexpr int dec_entries(int bits) { return (1 << bits) + (bits == 11 ? 128 : 256); }
// Index bits of the decode table.  A bigger table yields more symbols per lookup but costs 2^bits
// entries to build per block and shared memory that could hold more resident streams: the best
// size grows with the block and with the streams that share a table (measured on the K x block
// grid of BASELINE config 5, profiles/r2_decode_table_bits.md).
inline int dec_bits_for(int K, uint32_t block_size) {
  static const int forced = [] {  // tuning aid: HUFB200_DEC_BITS=9|10|11 overrides the choice
    const char* e = getenv("HUFB200_DEC_BITS");
    const int v = e ? atoi(e) : 0;
    return v >= 9 && v <= 11 ? v : 0;
  }();
  if (forced) return forced;
  if (block_size <= (32u << 10)) return 9;
  if (block_size <= (64u << 10)) return K <= 16 ? 9 : 10;
  if (block_size <= (128u << 10)) return 10;
  if (block_size <= (256u << 10)) return (K > 16 && K <= 40) ? 11 : 10;
  return 11;
}
#ifndef HUF_DEC_LOOKUPS
#define HUF_DEC_LOOKUPS 10
#endif
#ifndef HUF_DEC_ROW
#define HUF_DEC_ROW 64
#endif
constexpr int kDecRow = HUF_DEC_ROW;          // bytes of output staging per lane: a ring of 32-byte chunks
constexpr int kDecChunk = 32;                 // one 256-bit store
constexpr int kDecLookups = HUF_DEC_LOOKUPS;  // table lookups per lane per round (fixed: no lane waits for another inside a round)
static_assert(kDecRow == 64 || kDecRow == 128, "ring of 2 or 4 chunks");
static_assert(kDecChunk - 1 + 3 * kDecLookups < kDecRow, "a round must not overrun the unwritten part of the ring");
static_assert(12 * kDecLookups + 31 < 6 * 32, "a round must not consume more than 5 input words");
constexpr uint32_t kRowWrap = (kDecRow / 4 - 1) * 128;
constexpr int kDecEmits = (kDecChunk - 1 + 3 * kDecLookups) / kDecChunk;  // most complete chunks a round can leave behind
// Ring positions inside the lookup loop.  A lane's ring word i sits 128 * i bytes behind its
// first one (lane-private bank).  The position is kept as a counter in the TOP bits of a register
// (4 bits for the 16-word input ring, 4 or 5 for the output ring): `+= one` wraps by itself, and
// the byte offset is one shift that the address add absorbs -- no AND per step.
constexpr int kInRingBits = 4;
constexpr uint32_t kInRingBytes = 16 * 32 * 4;  // per warp
constexpr int kOutRingBits = kDecRow == 64 ? 4 : 5;
template <int BITS>
__device__ __forceinline__ uint32_t ring_ofs(uint32_t cnt) { return cnt >> (32 - BITS - 7); }  // 128 * (cnt >> (32 - BITS))

// One thread.  The fixed part of the header is fetched with independent loads up front (a loop
// that reads a count byte, then decides where the next one is, would chain up to 13 global-memory
// latencies); the symbols are copied by the caller's threads (copy_header_syms) unless copy_syms.
__device__ inline void parse_header(const uint8_t* blk, uint32_t comp_size, int K, uint32_t expect_raw,
                                    DecBlockHdr* bi, uint8_t* syms_out) {  // syms_out: nullptr = the caller copies them
  bi->ok = 0;
  bi->comp_size = comp_size;
  bi->num_syms = 0;
  if (comp_size < 8u + 4u * (uint32_t)(K - 1)) return;
  // raw_size, mask and the (at most 13) count bytes
  unsigned long long f0 = 0, f1 = 0, f2 = 0;
  const uint32_t avail = comp_size < 24u ? comp_size : 24u;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f0 |= (unsigned long long)blk[i] << (8 * i);
    if ((uint32_t)(8 + i) < avail) f1 |= (unsigned long long)blk[8 + i] << (8 * i);
    if ((uint32_t)(16 + i) < avail) f2 |= (unsigned long long)blk[16 + i] << (8 * i);
  }
  const uint32_t raw_size = (uint32_t)f0, mask = (uint32_t)(f0 >> 32);
  bi->raw_size = raw_size;
  if (mask >> (kMaxCodeLen + 1)) return;
  if (raw_size != expect_raw) return;
  const uint32_t npop = (uint32_t)__popc(mask);
  if (8u + npop > comp_size) return;
  uint32_t pos = 8, nsyms = 0, code = 0;
  for (int l = 0; l <= kMaxCodeLen; ++l) {
    uint32_t cnt = 0;
    if (mask & (1u << l)) {
      const uint32_t q = pos - 8u;  // < 13
      cnt = (uint32_t)((q < 8u ? f1 >> (8u * q) : f2 >> (8u * (q - 8u))) & 0xffu);
      ++pos;
      if (npop == 1 && cnt == 0) cnt = 256;  // :724-728
    }
    bi->first_idx[l] = nsyms;
    nsyms += cnt;
    code += cnt << (kMaxCodeLen - l);
    bi->code_end[l] = code;
  }
  if (nsyms > 256) return;
  if (raw_size != 0 && nsyms == 0) return;
  // The codes must tile the 12-bit code space exactly (Kraft sum == 1): every Huffman code does,
  // LimitCodeLengths keeps it so (:297-327), a lone symbol has the empty code (4096 >> 0).  An
  // over- or under-subscribed length table is corrupt; the decode table relies on completeness.
  if (nsyms != 0 && code != (1u << kMaxCodeLen)) return;
  bi->syms_off = pos;
  bi->ends_off = pos + nsyms;
  bi->payload_off = pos + nsyms + 4u * (uint32_t)(K - 1);
  if (bi->payload_off > comp_size) return;
  bi->num_syms = nsyms;
  if (syms_out)
    for (uint32_t i = 0; i < nsyms; ++i) syms_out[i] = blk[pos + i];
  bi->ok = 1;
}
// sorted_syms of a parsed header into bi->syms, by `nthreads` threads (bi->num_syms is 0 for a malformed header)
__device__ __forceinline__ void copy_header_syms(const uint8_t* blk, const DecBlockHdr* bi, uint8_t* syms, int tid,
                                                 int nthreads) {
  const uint32_t n = bi->num_syms, off = bi->syms_off;
  for (uint32_t i = (uint32_t)tid; i < n; i += (uint32_t)nthreads) syms[i] = blk[off + i];
}
__device__ inline void parse_header(const uint8_t* blk, uint32_t comp_size, int K, uint32_t expect_raw,
                                    DecBlockInfo* bi, bool copy_syms = true) {
  parse_header(blk, comp_size, K, expect_raw, static_cast<DecBlockHdr*>(bi), copy_syms ? bi->syms : nullptr);
}
__device__ __forceinline__ void copy_header_syms(const uint8_t* blk, DecBlockInfo* bi, int tid, int nthreads) {
  copy_header_syms(blk, bi, bi->syms, tid, nthreads);
}

// 32 bytes per lane in one instruction (256-bit global accesses, sm_100+): a lane's sector or
// output chunk costs one trip through the load/store pipe instead of two
#ifndef HUF_DEC_STMOD
#define HUF_DEC_STMOD ".L1::no_allocate"  // the decoder's 32-byte output stores leave the L1 to its input sectors
#endif
#ifndef HUF_DEC_LDMOD
#define HUF_DEC_LDMOD ""  // tuning aid: cache modifiers of the decoder's sector loads (".L1::no_allocate", ".L2::128B" ...)
#endif
struct Sector {
  uint32_t w[8];
};
__device__ __forceinline__ Sector ld_sector(uintptr_t addr, uintptr_t lo_lim) {
  Sector s;
  if (addr >= lo_lim) {
    asm volatile("ld.global" HUF_DEC_LDMOD ".v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(s.w[0]), "=r"(s.w[1]), "=r"(s.w[2]), "=r"(s.w[3]), "=r"(s.w[4]), "=r"(s.w[5]), "=r"(s.w[6]),
                   "=r"(s.w[7])
                 : "l"(addr));
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) s.w[i] = 0;
  }
  return s;
}
__device__ __forceinline__ void st_chunk32(uint8_t* addr, const uint32_t (&v)[8]) {
  asm volatile("st.global" HUF_DEC_STMOD ".v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// per-CTA scratch behind the tables: the T1 build scratch, later the lane rings and output rows
__host__ __device__ inline size_t dec_region_bytes(int bpc, int nthreads, int entries) {
  const size_t a = (size_t)bpc * (entries + 256);  // build scratch: first-code lengths + the header's symbols
  const size_t b = (size_t)(nthreads >> 5) * kInRingBytes + (((size_t)nthreads * kDecRow + 15) & ~(size_t)15);
  return ((a > b ? a : b) + 15) & ~(size_t)15;
}

// ---- split decode (see "Split decode" below): the streams of a block are cut into items of
// sub_bits stream bits; an item is one lane's work.  Device workspace, laid out by split_layout().
constexpr int kSplitThreads = 128;  // items per CTA (all of one block: they share its decode table)
constexpr int kSplitBits = 10;      // index bits of the decode table the split kernels use
struct SplitArgs {
  uint32_t sub_bits;           // item length in stream bits (a multiple of 32)
  uint32_t max_block_ctas;     // most CTAs a block may need (the arrays are sized for n_blocks times that)
  unsigned long long* cta_first;   // [n_blocks] exclusive prefix sum of block_ctas
  unsigned long long* total_ctas;  // [1]
  uint32_t* block_ctas;        // [n_blocks] CTAs of kSplitThreads items the block takes (0: empty or malformed)
  uint32_t* s_first;           // [n_blocks * K] first item of the stream, counted within its block
  uint32_t* s_nitems;          // [n_blocks * K]
  uint32_t* s_eoff;            // [n_blocks * K] end of the stream's region relative to the payload start
  uint32_t* s_bits;            // [n_blocks * K] 8 * (region bytes - 8): upper bound of the stream's bits
  uint32_t* start;             // [items] bit the item's decode starts on (final after k_split_scan)
  uint32_t* cnt;               // [items] symbols the item yields (final after k_split_scan)
  uint32_t* exit_bit[2];       // [items] first code boundary at or behind the item's end; ping-pong over the passes
  uint32_t* out_off;           // [items] offset of the item's first symbol within its slice
};

// The CTA's block in a launch whose CTAs are dealt out per block: the last b with cta_first[b] <= c
// (blocks without CTAs share their successor's prefix value and are skipped by taking the last).
__device__ __forceinline__ uint32_t split_block_of(const unsigned long long* cta_first, uint32_t n_blocks,
                                                   unsigned long long c) {
  uint32_t lo = 0, hi = n_blocks;  // invariant: cta_first[lo] <= c, cta_first[hi] > c (hi == n_blocks: total > c)
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (cta_first[mid] <= c) lo = mid;
    else hi = mid;
  }
  return lo;
}

// kItems: lane t decodes item t of the CTA's block (start bit, symbol count and output offset from
// the workspace) instead of stream t % K from its first bit -- the write pass of the split decode.
template <int kDecBits, bool kItems = false>
__global__ void __launch_bounds__(kDecMaxThreads)
k_decompress_blocks(const uint8_t* __restrict__ comp, const unsigned long long* __restrict__ offsets,
                    const uint32_t* __restrict__ comp_sizes, uint32_t n_blocks, int K, int bpc,
                    uint8_t* __restrict__ raw, uint64_t raw_n, uint32_t block_size,
                    uint32_t* __restrict__ status, SplitArgs sa) {
  extern __shared__ __align__(16) uint8_t dsm[];
  constexpr int kDecEntries = dec_entries(kDecBits);
  // layout: tables [bpc][kDecEntries] u32 | region (T1 scratch, then rings [nwarps][16][32] u32 + rows) | infos [bpc]
  const int nthreads = blockDim.x;
  const int nwarps = nthreads >> 5;
  uint32_t* tables = reinterpret_cast<uint32_t*>(dsm);
  uint8_t* region = dsm + (size_t)bpc * kDecEntries * 4;
  DecBlockHdr* infos = reinterpret_cast<DecBlockHdr*>(region + dec_region_bytes(bpc, nthreads, kDecEntries));
  // sorted_syms of the CTA's blocks: only the table build reads them, so they sit in its scratch
  // (behind the bpc first-code-length arrays) and not in the resident part of the CTA's shared memory
  uint8_t* const hdr_syms = region + (size_t)bpc * kDecEntries;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  uint32_t b0 = blockIdx.x * (uint32_t)bpc;
  uint32_t item0 = 0;  // kItems: first item of this CTA within its block
  unsigned long long gitem0 = 0;  // ... and in the workspace arrays
  if constexpr (kItems) {  // bpc == 1, blockDim.x == kSplitThreads
    const unsigned long long c = blockIdx.x;
    if (c >= *sa.total_ctas) return;
    b0 = split_block_of(sa.cta_first, n_blocks, c);
    const unsigned long long first = sa.cta_first[b0];
    item0 = (uint32_t)(c - first) * (uint32_t)kSplitThreads;
    gitem0 = c * (unsigned long long)kSplitThreads;
  }

  // ---- header parse, one thread per block
  if (tid < bpc) {
    const uint32_t b = b0 + tid;
    DecBlockHdr* bi = &infos[tid];
    bi->ok = 0;
    if (b < n_blocks) {
      const uint64_t roff = (uint64_t)b * block_size;
      const uint32_t expect = (uint32_t)((raw_n - roff) < (uint64_t)block_size ? (raw_n - roff) : (uint64_t)block_size);
      parse_header(comp + offsets[b], comp_sizes[b], K, expect, bi, nullptr);
      if (!bi->ok && status) atomicOr(status, 1u);
    } else {
      bi->num_syms = 0;
    }
  }
  __syncthreads();
  for (int lb = 0; lb < bpc; ++lb)
    if (b0 + lb < n_blocks) copy_header_syms(comp + offsets[b0 + lb], &infos[lb], hdr_syms + 256 * lb, tid, nthreads);
  __syncthreads();
  // ---- decode tables
  for (int lb = 0; lb < bpc; ++lb) {
    const DecBlockHdr* bi = &infos[lb];
    if (b0 + lb < n_blocks && bi->ok && bi->raw_size != 0) {
      build_dtable<kDecBits, 3, true>(bi, hdr_syms + 256 * lb, tables + (size_t)lb * kDecEntries,
                                region + (size_t)lb * kDecEntries, tid, nthreads);
    }
  }
  __syncthreads();

  // ---- decode: thread t <-> (local block t / K, stream t % K)
  int lb = tid / K;
  int s = tid - lb * K;
  uint32_t it_start = 0, it_cnt = 0, it_off = 0;
  if constexpr (kItems) {  // thread t <-> item item0 + t of the block: find its stream
    __shared__ uint32_t sf[kMaxK], sn[kMaxK];
    if (tid < K) {
      sf[tid] = sa.s_first[(size_t)b0 * K + tid];
      sn[tid] = sa.s_nitems[(size_t)b0 * K + tid];
    }
    __syncthreads();
    lb = bpc;  // no stream unless found below
    s = 0;
    const uint32_t li = item0 + (uint32_t)tid;
    for (int q = 0; q < K; ++q)
      if (li >= sf[q] && li - sf[q] < sn[q]) {
        lb = 0;
        s = q;
      }
    if (lb == 0) {
      const unsigned long long gi = gitem0 + (unsigned long long)tid;
      it_start = sa.start[gi];
      it_cnt = sa.cnt[gi];
      it_off = sa.out_off[gi];
      if (it_cnt == 0) lb = bpc;
    }
  }
  const uint32_t b = b0 + (uint32_t)(kItems ? 0 : lb);
  bool active = (lb < bpc) && (b < n_blocks);
  const DecBlockHdr* bi = &infos[active ? lb : 0];
  active = active && bi->ok && bi->raw_size != 0;

  uint32_t left = 0;           // symbols still to produce
  uint8_t* outp = nullptr;     // next output byte
  uintptr_t e16 = 16, lo_lim = 0;
  uint32_t acc = 0;            // bits 0..5: bits consumed from the window, bits 6..: symbols in the row
  uint32_t rd = 0, staged = 0, cidx = 0;
  uint32_t ob = 0;             // pending output symbols (fewer than 4), first symbol in the low byte
  bool bad_lane = false;
  uint32_t region_bytes = 0, pad_bits = 0;  // for the end-of-stream check
  if (active) {
    const uint8_t* blk = comp + offsets[b];
    uint32_t st, sz;
    slice_geom(bi->raw_size, K, s, st, sz);
    uint32_t e_off;
    const uint32_t payload = bi->comp_size - bi->payload_off;
    if (s == K - 1) e_off = payload;  // :901
    else {
      const uint8_t* p = blk + bi->ends_off + 4 * s;
      e_off = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    }
    // the region in front: ends are cumulative and every region holds at least its 8 slop bytes (:783-786)
    uint32_t e_prev = 0;
    if (s != 0) {
      const uint8_t* p = blk + bi->ends_off + 4 * (s - 1);
      e_prev = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    }
    if (e_off > payload || e_prev > e_off || e_off - e_prev < (uint32_t)kSlop) {
      // corrupt end offsets: produce nothing and read nothing through them
      bad_lane = true;
      sz = 0;
      e_off = 0;
      e_prev = 0;
      it_cnt = 0;
    }
    region_bytes = e_off - e_prev;
    pad_bits = 0;
    left = sz;
    outp = raw + (uint64_t)b * block_size + st;
    uintptr_t end_addr = (uintptr_t)(blk + bi->payload_off) + e_off;  // exclusive
    if constexpr (kItems) {  // the item's first bit is bit 7 - (start % 8) of the byte start / 8 below the region's end
      left = it_cnt <= sz - (it_off < sz ? it_off : sz) ? it_cnt : 0u;  // (k_split_scan keeps the items inside the slice)
      outp += it_off;
      end_addr -= it_start >> 3;
    }
    e16 = (end_addr + 31) & ~(uintptr_t)31;  // input is fetched in whole 32-byte sectors
    const uint32_t pad = (uint32_t)(e16 - end_addr);
    lo_lim = (uintptr_t)blk & ~(uintptr_t)31;
    rd = pad >> 2;
    acc = 8 * (pad & 3) + (kItems ? (it_start & 7u) : 0u);
    pad_bits = 8 * pad;
  }
  if (bad_lane && status) atomicOr(status, 1u);

  // shared-space addresses, kept in registers
  const uint32_t t_addr = smem_u32(tables + (size_t)(lb < bpc ? lb : 0) * kDecEntries);
  // bases of the finer levels (see build_dtable): level l reads entry p_l + c_l; lanes without a
  // stream keep c_l = 0 (their window is zero)
  uint32_t tl_addr[kMaxCodeLen - kDecBits];
  {
    uint32_t base = active ? (bi->code_end[kDecBits] >> (kMaxCodeLen - kDecBits)) : 0u;
#pragma unroll
    for (int l = kDecBits + 1; l <= kMaxCodeLen; ++l) {
      const uint32_t first = active ? (bi->code_end[l - 1] >> (kMaxCodeLen - l)) : 0u;
      tl_addr[l - kDecBits - 1] = t_addr + 4u * (base - first);
      if (active) base += (bi->code_end[l] - bi->code_end[l - 1]) >> (kMaxCodeLen - l);
    }
  }
  const uint32_t col = smem_u32(region) + (uint32_t)warp * kInRingBytes + 4u * (uint32_t)lane;  // word i at col + (i & 15) * 128
  // output staging ring: word j (4 symbols) of this lane at row + (j & 15) * 128 -- lane-private bank
  const uint32_t row = smem_u32(region) + (uint32_t)nwarps * kInRingBytes + (uint32_t)warp * (32 * kDecRow) + 4u * (uint32_t)lane;
  // make the three addresses opaque so that they stay in registers instead of being recomputed
  // from tid / %ctaid inside the lookup loop
  asm volatile("" : "+r"(const_cast<uint32_t&>(t_addr)), "+r"(const_cast<uint32_t&>(col)), "+r"(const_cast<uint32_t&>(row)));
#pragma unroll
  for (int i = 0; i < kMaxCodeLen - kDecBits; ++i) asm volatile("" : "+r"(tl_addr[i]));
  const uint32_t one = K >= 1 ? 1u : 0u;  // 1 (K is validated by the entry points), but not a constant the assembler could fold the multiply with

  // prime: stage two 32-byte sectors (the whole 16-word ring) and keep the next one in registers.
  // Each top-up takes a full sector, so the kernel does not depend on L1 to serve the other half
  // of a sector later (L1 is nearly gone once shared memory is carved out for the tables).
  Sector pf = {};  // words in address order: the stream runs from w[7] down to w[0]
  auto load_sector = [&]() {
    const uintptr_t a = e16 - 32 * (uintptr_t)(cidx + 1);
    if (active) pf = ld_sector(a, lo_lim);
    ++cidx;
  };
  auto stage_sector = [&]() {
    const uint32_t o = (staged & 15) * 128;  // staged % 8 == 0: the eight words do not wrap
#pragma unroll
    for (int i = 0; i < 8; ++i) sts_u32(col + o + 128 * i, pf.w[7 - i]);
    staged += 8;
  };
  load_sector();
  stage_sector();
  load_sector();
  stage_sector();
  load_sector();
  uint32_t hi = lds_u32(col + (rd & 15) * 128);
  uint32_t lo = lds_u32(col + ((rd + 1) & 15) * 128);
  rd += 2;

  // Rounds of kDecLookups lookups per lane.  Every lane does the same number of lookups per
  // round (the symbols they yield differ, 1..3 each); the symbols go into a 64-byte ring per lane
  // whose byte positions are congruent to the output addresses mod 16, so every complete 16-byte
  // ring chunk leaves as one aligned 128-bit store.  acc: bits 0..5 = bits consumed from the
  // window, bits 6.. = ring write position in bytes (starts at the slice's misalignment h0).
  // Positions live in bits 6..31 of acc, so they are kept below 2^26: a stream longer than
  // kPosWindow symbols is decoded through a window of positions that is moved along (see the
  // rebase at the top of the round loop); `beyond` = symbols past the window's end.
  constexpr uint32_t kPosWindow = 1u << 25;
  uint32_t beyond = left > kPosWindow ? left - kPosWindow : 0u;
  uint32_t h0 = (uint32_t)((uintptr_t)outp & (kDecChunk - 1));
  uint8_t* out_al = outp - h0;                           // chunk-aligned; ring byte p <-> out_al[p]
  uint32_t end_pos = h0 + (left - beyond);               // ring position one past the last symbol (of the window)
  uint32_t end_acc = end_pos << 6;
  uint32_t full_chunks = end_pos >> 5;                   // chunks that lie entirely inside the slice
  acc |= h0 << 6;
  uint32_t wofs = (h0 >> 2) << (32 - kOutRingBits);     // ring word being filled (top-bits counter, see ring_ofs)
  uint32_t rdo = rd << (32 - kInRingBits);              // next input ring word
  uint32_t chunk = 0;                                   // next 32-byte chunk to write out
  // One round = kDecLookups straight-line lookups (no branch per lookup, so consecutive lookups
  // overlap).  CHECKED rounds are the last few of a warp, when some lane may run out of symbols:
  // a lane that has them all sees an all-zero entry, which changes nothing.
  auto lookups = [&](auto checked) {
#pragma unroll
    for (int it = 0; it < kDecLookups; ++it) {
      const uint32_t win = __funnelshift_l(lo, hi, acc);  // shift amount = acc & 31
      const uint32_t nxw = lds_u32(col + ring_ofs<kInRingBits>(rdo));  // next ring word, needed only if this lookup crosses a word
      // index = max over the levels, formed on the byte addresses (see build_dtable); signed: a
      // level's base may lie below zero as a shared-window offset
      // Level l's entry sits at base_l + 4 * (win >> (32 - l)).  With W = win >> (30 - BITS) the
      // first three levels are W, 2W and 4W up to the bits below an entry's four bytes (W & 3,
      // 2 * (W & 1), none): true addresses are multiples of four apart, so those bits never
      // change which level wins (equal true addresses are the same entry), and the winner is
      // rounded down to its entry.  One shift (ALU pipe) for three levels instead of one each;
      // the three multiply-adds run on the FMA pipe.
      const uint32_t W = win >> (30 - kDecBits);
      int ea = (int)mad_u32_rr(W, one, t_addr);  // (an add would be fused into the max and put it on the ALU pipe)
      if constexpr (kDecBits + 1 <= kMaxCodeLen) ea = max(ea, (int)mad_u32<2>(W, tl_addr[0]));
      if constexpr (kDecBits + 2 <= kMaxCodeLen) ea = max(ea, (int)mad_u32<4>(W, tl_addr[1]));
#pragma unroll
      for (int l = kDecBits + 3; l <= kMaxCodeLen; ++l)
        ea = max(ea, (int)entry_addr(tl_addr[l - kDecBits - 1], win >> (32 - l)));
      uint32_t e = lds_u32_ro((uint32_t)ea & ~3u);
      if (decltype(checked)::value && acc >= end_acc) e = 0;
      const uint32_t sh = (acc >> 3) & 0x18u;  // 8 * (position in the ring word being filled)
      const uint32_t old = acc;
      acc += e >> 24;  // bits consumed into bits 0..5, symbol count into bits 6..: the loop-carried chain
      // everything below hangs off that chain
      const uint32_t v = e & 0xffffffu;  // the entry's symbols, unused bytes are zero
      ob = mad_u32_rr(v, 1u << sh, ob);  // ob | v << sh (the bytes are free) on the FMA pipe
      if ((acc ^ old) & 0x100u) {  // the write position crossed a multiple of 4: one ring word is complete
        sts_u32(row + ring_ofs<kOutRingBits>(wofs), ob);
        wofs += 1u << (32 - kOutRingBits);
        ob = __funnelshift_l(v, 0, sh);  // the symbols that did not fit: v >> (32 - sh), 0 for sh == 0
      }
      if (acc & 32u) {
        hi = lo;
        lo = nxw;
        rdo += 1u << (32 - kInRingBits);
        acc -= 32;
      }
    }
  };
  // a lane is "far" from its end if a whole round cannot exhaust it; lanes without a stream run
  // harmless lookups on a zero window in unchecked rounds (private rings, nothing is emitted)
  uint32_t far_acc = end_acc >= ((3u * kDecLookups) << 6) ? end_acc - ((3u * kDecLookups) << 6) : 0u;
  const bool no_stream = !active || left == 0;
  const bool long_streams = __any_sync(0xffffffffu, left >= kPosWindow / 2 - 64);  // can a position reach the rebase point?
  // Every lookup of an unfinished lane yields at least one symbol, so the warp needs at most
  // max(left) / kDecLookups + 1 rounds; the count is enforced, so that no table or payload,
  // however corrupt, can keep the kernel running.
  uint32_t rounds_left = left / (uint32_t)kDecLookups + 4u;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) rounds_left = max(rounds_left, __shfl_xor_sync(0xffffffffu, rounds_left, d));
  while (__any_sync(0xffffffffu, acc < end_acc)) {
    if (rounds_left-- == 0) {
      if (status) atomicOr(status, 1u);
      break;
    }
    // rare (streams of more than 16 Mi symbols): move the position window.  Everything already
    // written out is dropped from the positions -- a multiple of the 64-byte ring, so ring
    // offsets keep their meaning -- and the window's end moves on by as much as is left.
    if (long_streams) {
      if (acc >= ((kPosWindow / 2) << 6)) {
        const uint32_t delta = (chunk & ~(uint32_t)(kDecRow / kDecChunk - 1)) * (uint32_t)kDecChunk;
        const uint32_t add = beyond < delta ? beyond : delta;
        beyond -= add;
        acc -= delta << 6;
        chunk -= delta / (uint32_t)kDecChunk;
        out_al += delta;
        h0 = 0;  // the slice's head lies behind
        end_pos = end_pos - delta + add;
        end_acc = end_pos << 6;
        full_chunks = end_pos >> 5;
        far_acc = end_acc >= ((3u * kDecLookups) << 6) ? end_acc - ((3u * kDecLookups) << 6) : 0u;
      }
    }
    // Top up the input ring: a round consumes at most 5 words (+1 looked ahead).  The sector
    // stored now was requested at the previous top-up, so its latency is hidden.
    if (staged - rd <= 8) {  // room for a sector; afterwards at least 9 words are staged
      stage_sector();
      load_sector();
    }
    const uint32_t rdo0 = rdo;
    if (__all_sync(0xffffffffu, no_stream || acc <= far_acc)) lookups(std::false_type{});
    else lookups(std::true_type{});
    rd += (rdo - rdo0) >> (32 - kInRingBits);  // words consumed by this round (at most 4)
    // write out the complete chunks (a round adds at most 3 * kDecLookups bytes)
#pragma unroll
    for (int t = 0; t < kDecEmits; ++t) {
      uint32_t avail = acc >> 11;  // complete chunks by write position ...
      if (avail > full_chunks) avail = full_chunks;  // ... that do not reach past the slice
      if (chunk < avail) {
        const uint32_t w0 = (chunk * 1024) & kRowWrap;  // eight ring words per chunk
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = lds_u32(row + w0 + 128 * i);
        if (chunk == 0 && h0 != 0) {  // the slice starts inside this chunk
          for (uint32_t i = h0; i < (uint32_t)kDecChunk; ++i) {
            uint32_t w = v[0];
#pragma unroll
            for (int j = 1; j < 8; ++j)
              if ((i >> 2) == (uint32_t)j) w = v[j];
            out_al[i] = (uint8_t)(w >> (8 * (i & 3)));
          }
        } else {
          st_chunk32(out_al + kDecChunk * (size_t)chunk, v);
        }
        ++chunk;
      }
    }
  }
  // tail: the bytes of the last, partial chunk
  if (left) {
    sts_u32(row + ring_ofs<kOutRingBits>(wofs), ob);  // the ring word still being filled
    uint32_t p = kDecChunk * chunk;
    if (p < h0) p = h0;
    for (; p < end_pos; ++p) out_al[p] = (uint8_t)lds_u8(row + (((p >> 2) * 128) & kRowWrap) + (p & 3));
  }
  // A stream of b bits sits in a region of ceil(b / 8) + 8 bytes (:783-786), so the bits this
  // lane consumed for its symbols must end inside the region's first stream byte: anything else
  // means a corrupt table, payload or end offset, even if every access stayed inside the block.
  // Position = 32 * (ring words behind the window) + bits consumed from the window.  The last
  // lookup may have decoded up to two symbols more than the slice has (from the zero padding
  // behind the stream; they are not written): each of them accounts for at most 12 bits more.
  if (!kItems && active && !bad_lane) {  // (items: k_split_scan has checked the stream's symbol count)
    const uint32_t consumed = 32u * (rd - 2u) + (acc & 63u) - pad_bits;
    const uint32_t extra = left ? (acc >> 6) - end_pos : 0u;  // 0..2 symbols past the slice's end
    const uint32_t stream_bits_max = 8u * (region_bytes - (uint32_t)kSlop);
    const bool ok = consumed + 7u >= stream_bits_max && consumed <= stream_bits_max + (uint32_t)kMaxCodeLen * extra &&
                    extra <= 2u;
    if (!ok && status) atomicOr(status, 1u);
  }
}

// ===========================================================================
// Split decode: more lanes than streams.  DecompressMulti<K> (codec/huffman.cpp:892-955) walks K
// streams, and a Huffman stream can only be read from a code boundary -- but prefix codes
// re-synchronise: a decoder started on an arbitrary bit falls into step with the true code
// boundaries after a few codes.  So every stream is cut into items of sub_bits bits; a lane decodes
// from its item's start to the first code boundary at or behind the item's end (its exit) and
// counts the symbols on the way:
//   k_split_plan   per block: header geometry -> items per stream, CTAs per block
//   k_split_sync   pass 0: a short warm-up decode in front of every item gives its probable start,
//                  from which it decodes to its exit and counts;
//                  pass p: an item whose predecessor's exit differs from the start it used decodes
//                  again from there (a warm-up that had not fallen into step yet).
//   k_split_scan   per stream: the items are consistent up to the first one whose start is not its
//                  predecessor's exit; offsets = prefix sums of the counts; what lies behind an
//                  inconsistent item (codes that never re-synchronise, e.g. all of one length)
//                  becomes ONE item that runs to the stream's end, so the result never depends on
//                  re-synchronisation, only the speed does.  The symbol total is checked against
//                  the slice (only the last item may over-count, by what the < 8 pad bits decode to).
//   k_decompress_blocks<.., true>   the write pass: one lane per item.
// Two decodes of every bit instead of one, on as many lanes as the device holds: it pays when the
// streams are too few to fill the device (one buffer of K streams: K lanes otherwise).
// ===========================================================================
// One warp per block: lane 0 parses the header, the lanes take the streams.
__global__ void __launch_bounds__(32)
k_split_plan(const uint8_t* __restrict__ comp, const unsigned long long* __restrict__ offsets,
             const uint32_t* __restrict__ comp_sizes, uint32_t n_blocks, int K, uint64_t raw_n, uint32_t block_size,
             SplitArgs sa, uint32_t* __restrict__ status) {
  __shared__ DecBlockInfo info;
  const uint32_t b = blockIdx.x;
  const int lane = threadIdx.x;
  DecBlockInfo* bi = &info;
  const uint64_t roff = (uint64_t)b * block_size;
  const uint32_t expect = (uint32_t)((raw_n - roff) < (uint64_t)block_size ? (raw_n - roff) : (uint64_t)block_size);
  const uint8_t* blk = comp + offsets[b];
  if (lane == 0) parse_header(blk, comp_sizes[b], K, expect, bi, false);
  __syncwarp();
  bool ok = bi->ok != 0;
  uint32_t items = 0;
  if (ok && bi->raw_size != 0) {
    const uint32_t payload = bi->comp_size - bi->payload_off;
    auto end_of = [&](int s) -> uint32_t {  // :901
      if (s == K - 1) return payload;
      const uint8_t* p = blk + bi->ends_off + 4 * s;
      return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    };
    for (int s0 = 0; s0 < K; s0 += 32) {
      const int s = s0 + lane;
      uint32_t n = 0, bits = 0, e_off = 0;
      bool good = true;
      if (s < K) {
        e_off = end_of(s);
        const uint32_t e_prev = s ? end_of(s - 1) : 0u;
        good = !(e_off > payload || e_prev > e_off || e_off - e_prev < (uint32_t)kSlop || e_off - e_prev >= (1u << 29));
        if (good) {
          bits = 8u * (e_off - e_prev - (uint32_t)kSlop);
          n = (bits + sa.sub_bits - 1) / sa.sub_bits;
          if (n == 0) n = 1;
        }
      }
      if (!__all_sync(0xffffffffu, good)) ok = false;
      const uint32_t incl = warp_incl_scan(n);
      if (s < K) {
        const size_t g = (size_t)b * K + s;
        sa.s_first[g] = items + incl - n;
        sa.s_nitems[g] = n;
        sa.s_eoff[g] = e_off;
        sa.s_bits[g] = bits;
      }
      items += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  uint32_t ctas = ok ? (items + kSplitThreads - 1) / kSplitThreads : 0u;
  if (ctas > sa.max_block_ctas) {  // more compressed bytes than a block of this size can have
    ok = false;
    ctas = 0;
  }
  if (lane == 0) {
    sa.block_ctas[b] = ctas;
    if (!ok && status) atomicOr(status, 1u);
  }
}

// The stream as 32-bit words from the region's end down, seen through an aligned window: virtual
// word v is the aligned little-endian u32 at wend - 4 - 4v, virtual bit = 8 * pad + stream bit.
struct SplitReader {
  uintptr_t wend, lo_lim;
  __device__ __forceinline__ uint32_t word(uint32_t v) const {
    const uintptr_t a = wend - 4u - 4u * (uintptr_t)v;
    return a >= lo_lim ? __ldg(reinterpret_cast<const uint32_t*>(a)) : 0u;
  }
};

// Walks the codes of a stream from bit `pos` to the first code boundary at or behind bit `limit`;
// returns that boundary and adds the number of codes passed to cnt.  T: the block's decode table
// (build_dtable<BITS, 3, true>), c_l: its level bases.
template <int BITS>
__device__ __forceinline__ uint32_t split_walk(const uint32_t* T, const uint8_t* L1, const int (&c_l)[kMaxCodeLen - BITS + 1],
                                               const SplitReader& rd, uint32_t pad_bits, uint32_t pos, uint32_t limit,
                                               uint32_t& cnt) {
  if (pos >= limit) return pos;
  uint32_t P = pad_bits + pos;  // virtual bit position
  const uint32_t Plimit = pad_bits + limit;
  uint32_t wi = P >> 5;
  uint32_t hi = rd.word(wi), lo = rd.word(wi + 1), nx = rd.word(wi + 2);  // nx: requested one word ahead of its use
  // far from the end: whole table entries (up to three symbols); an entry never spans more than
  // BITS bits and a longer code has an entry of its own, so 3 * 12 bits of margin keep every
  // code that starts in front of the limit counted one by one below
  const uint32_t Pfar = Plimit > 36u ? Plimit - 36u : 0u;
  while (P < Pfar) {
    const uint32_t win = __funnelshift_l(lo, hi, P);
    int idx = (int)(win >> (32 - BITS));
#pragma unroll
    for (int l = BITS + 1; l <= kMaxCodeLen; ++l) idx = max(idx, (int)(win >> (32 - l)) + c_l[l - BITS - 1]);
    const uint32_t e = T[idx];
    P += (e >> 24) & 15u;
    cnt += e >> 30;
    if ((P >> 5) != wi) {
      ++wi;
      hi = lo;
      lo = nx;
      nx = rd.word(wi + 2);
    }
  }
  // near the end: one code at a time.  Its length: what build_dtable left in L1 for the window's
  // first BITS bits (the first code's own length), or -- 15 there: a longer code -- the length in
  // that code's own entry
  while (P < Plimit) {
    const uint32_t win = __funnelshift_l(lo, hi, P);
    uint32_t l = L1[win >> (32 - BITS)];
    if (l > (uint32_t)BITS) {
      int idx = (int)(win >> (32 - BITS));
#pragma unroll
      for (int q = BITS + 1; q <= kMaxCodeLen; ++q) idx = max(idx, (int)(win >> (32 - q)) + c_l[q - BITS - 1]);
      l = (T[idx] >> 24) & 15u;
    }
    P += l;
    ++cnt;
    if ((P >> 5) != wi) {
      ++wi;
      hi = lo;
      lo = nx;
      nx = rd.word(wi + 2);
    }
  }
  return P - pad_bits;
}

// pass 0: item j > 0 first runs a short warm-up: it decodes from `warm` bits in front of its first
//         bit and takes the first code boundary at or behind that bit as its start -- the true one
//         as soon as the warm-up has fallen into step (item 0 starts on bit 0).  Then it decodes
//         from there to the first boundary at or behind its end (its exit) and counts its symbols.
// pass p >= 1: an item whose predecessor's exit is not the start it used decodes again from there.
template <int BITS>
__global__ void __launch_bounds__(kSplitThreads)
k_split_sync(const uint8_t* __restrict__ comp, const unsigned long long* __restrict__ offsets,
             const uint32_t* __restrict__ comp_sizes, uint32_t n_blocks, int K, uint64_t raw_n, uint32_t block_size,
             SplitArgs sa, int pass) {
  constexpr int kEntries = dec_entries(BITS);
  __shared__ uint32_t T[kEntries];
  __shared__ __align__(16) uint8_t L1[kEntries];
  __shared__ DecBlockInfo bi;
  __shared__ uint32_t sf[kMaxK], sn[kMaxK];
  const unsigned long long c = blockIdx.x;
  if (c >= *sa.total_ctas) return;
  const int tid = threadIdx.x;
  const uint32_t b = split_block_of(sa.cta_first, n_blocks, c);
  const uint32_t li = (uint32_t)(c - sa.cta_first[b]) * (uint32_t)kSplitThreads + (uint32_t)tid;
  const unsigned long long gi = c * (unsigned long long)kSplitThreads + (unsigned long long)tid;
  if (tid < K) {
    sf[tid] = sa.s_first[(size_t)b * K + tid];
    sn[tid] = sa.s_nitems[(size_t)b * K + tid];
  }
  __syncthreads();
  // this lane's item: stream s, item j of it
  int s = -1;
  uint32_t j = 0;
  for (int q = 0; q < K; ++q)
    if (li >= sf[q] && li - sf[q] < sn[q]) {
      s = q;
      j = li - sf[q];
    }
  const bool active = s >= 0;
  const uint32_t* exit_prev = (pass & 1) ? sa.exit_bit[0] : sa.exit_bit[1];
  uint32_t* exit_cur = (pass & 1) ? sa.exit_bit[1] : sa.exit_bit[0];
  uint32_t start = 0;
  bool need = active;
  if (active && pass != 0) {
    start = j ? exit_prev[gi - 1] : 0u;
    need = start != sa.start[gi];
  }
  if (!__syncthreads_or(need)) {  // nothing in this CTA moves: the exits carry over
    if (active) exit_cur[gi] = exit_prev[gi];
    return;
  }
  const uint8_t* blk = comp + offsets[b];
  if (tid == 0) {
    const uint64_t roff = (uint64_t)b * block_size;
    const uint32_t expect = (uint32_t)((raw_n - roff) < (uint64_t)block_size ? (raw_n - roff) : (uint64_t)block_size);
    parse_header(blk, comp_sizes[b], K, expect, &bi, false);  // valid: k_split_plan gave the block CTAs
  }
  __syncthreads();
  copy_header_syms(blk, &bi, tid, kSplitThreads);
  __syncthreads();
  build_dtable<BITS, 3, true>(&bi, bi.syms, T, L1, tid, kSplitThreads);
  if (!active) return;
  if (!need) {
    exit_cur[gi] = exit_prev[gi];
    return;
  }
  const uint32_t bits_max = sa.s_bits[(size_t)b * K + s];
  uint32_t limit = (j + 1) * sa.sub_bits;
  if (limit > bits_max || j + 1 == sn[s]) limit = bits_max;
  uint32_t pos = start, cnt = 0;
  if (bi.code_end[0] == 0) {  // (a lone symbol has the empty code: no bits, nothing to walk)
    const uintptr_t end_addr = (uintptr_t)(blk + bi.payload_off) + sa.s_eoff[(size_t)b * K + s];
    SplitReader rd;
    rd.wend = (end_addr + 3) & ~(uintptr_t)3;
    rd.lo_lim = (uintptr_t)blk & ~(uintptr_t)3;
    const uint32_t pad_bits = 8u * (uint32_t)(rd.wend - end_addr);
    // level bases of the table extension (see build_dtable)
    int c_l[kMaxCodeLen - BITS + 1];
    {
      uint32_t base = bi.code_end[BITS] >> (kMaxCodeLen - BITS);
#pragma unroll
      for (int l = BITS + 1; l <= kMaxCodeLen; ++l) {
        const uint32_t first = bi.code_end[l - 1] >> (kMaxCodeLen - l);
        c_l[l - BITS - 1] = (int)base - (int)first;
        base += (bi.code_end[l] - bi.code_end[l - 1]) >> (kMaxCodeLen - l);
      }
    }
    if (pass == 0 && j != 0) {  // warm-up
      const uint32_t first_bit = j * sa.sub_bits;
      const uint32_t warm = sa.sub_bits / 2 < 512u ? sa.sub_bits / 2 : 512u;
      uint32_t dummy = 0;
      start = split_walk<BITS>(T, L1, c_l, rd, pad_bits, first_bit - warm, first_bit, dummy);
    }
    pos = split_walk<BITS>(T, L1, c_l, rd, pad_bits, start, limit, cnt);
  }
  sa.start[gi] = start;
  sa.cnt[gi] = cnt;
  exit_cur[gi] = pos;
}

// ---- the whole split decode of ONE small buffer in ONE launch instead of seven: the reference's
// own benchmark unit is a 100 KiB buffer per call (codec/huffman_benchmark.cpp:61-81), where
// launches and table builds are most of a call.  The K streams are dealt out to up to eight CTAs
// (whole streams: nothing crosses a CTA); one item per thread; the item arrays live in registers
// and shared memory; the phases of k_split_plan / k_split_sync / k_split_scan and the write pass
// are separated by __syncthreads.  The host sizes sub_bits for a CTA that gets 1.5 times its share
// of the bits; a CTA whose streams hold more than kSmallItems items reports kSplitRetry and the
// host takes the spread form.
constexpr int kSmallItems = 1024;
constexpr int kSmallMaxCtas = 8;
constexpr uint32_t kSplitRetry = 2u;  // status bit: the items did not fit, nothing usable was written
// Decodes `cnt` symbols from stream bit `pos` on and stores them at out (byte stores up to a
// 4-byte boundary, then words).
template <int BITS>
__device__ __forceinline__ void split_write_lane(const uint32_t* T, const int (&c_l)[kMaxCodeLen - BITS + 1],
                                                 const SplitReader& rd, uint32_t pad_bits, uint32_t pos, uint32_t cnt,
                                                 uint8_t* out) {
  if (cnt == 0) return;
  uint32_t P = pad_bits + pos;
  uint32_t wi = P >> 5;
  uint32_t hi = rd.word(wi), lo = rd.word(wi + 1), nx = rd.word(wi + 2);  // nx: requested one word ahead of its use
  unsigned long long buf = 0;  // decoded symbols not stored yet, first symbol in the low byte
  uint32_t fill = 0;
  while (cnt) {
    const uint32_t win = __funnelshift_l(lo, hi, P);
    int idx = (int)(win >> (32 - BITS));
#pragma unroll
    for (int l = BITS + 1; l <= kMaxCodeLen; ++l) idx = max(idx, (int)(win >> (32 - l)) + c_l[l - BITS - 1]);
    const uint32_t e = T[idx];
    P += (e >> 24) & 15u;
    uint32_t n = e >> 30;  // 1..3 symbols (the last entry of an item may hold more than are left)
    if (n > cnt) n = cnt;
    cnt -= n;
    buf |= (unsigned long long)(e & (0xffffffu >> (8u * (3u - n)))) << (8u * fill);
    fill += n;
    if ((P >> 5) != wi) {
      ++wi;
      hi = lo;
      lo = nx;
      nx = rd.word(wi + 2);
    }
    while (fill >= 4u || (fill && ((uintptr_t)out & 3u))) {
      if (((uintptr_t)out & 3u) == 0 && fill >= 4u) {
        *reinterpret_cast<uint32_t*>(out) = (uint32_t)buf;
        out += 4;
        buf >>= 32;
        fill -= 4;
      } else {
        *out++ = (uint8_t)buf;
        buf >>= 8;
        fill -= 1;
      }
    }
  }
  for (; fill; --fill) {
    *out++ = (uint8_t)buf;
    buf >>= 8;
  }
}

template <int BITS>
__global__ void __launch_bounds__(kSmallItems)
k_split_small(const uint8_t* __restrict__ comp, const uint32_t* __restrict__ comp_size_p, int K, uint8_t* __restrict__ raw,
              uint32_t raw_n, uint32_t sub_bits, uint32_t* __restrict__ status) {
  constexpr int kEntries = dec_entries(BITS);
  __shared__ uint32_t T[kEntries];
  __shared__ __align__(16) uint8_t L1[kEntries];
  __shared__ DecBlockInfo bi;
  __shared__ uint32_t sf[kMaxK + 1], sn[kMaxK], seoff[kMaxK], sbits[kMaxK], sjbad[kMaxK], sbase[kMaxK], scounted[kMaxK];
  __shared__ uint32_t exit_a[kSmallItems], exit_b[kSmallItems];
  __shared__ uint32_t warp_tot[kSmallItems / 32];
  __shared__ uint32_t bad_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint8_t* blk = comp;
  if (tid == 0) {
    bad_s = 0;
    parse_header(blk, *comp_size_p, K, raw_n, &bi, false);
  }
  __syncthreads();
  if (!bi.ok) {
    if (tid == 0 && status) atomicOr(status, 1u);
    return;
  }
  if (bi.raw_size == 0) return;
  copy_header_syms(blk, &bi, tid, kSmallItems);
  // ---- plan (k_split_plan): thread s takes stream s
  if (tid < K) {
    const uint32_t payload = bi.comp_size - bi.payload_off;
    auto end_of = [&](int s) -> uint32_t {  // :901
      if (s == K - 1) return payload;
      const uint8_t* p = blk + bi.ends_off + 4 * s;
      return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    };
    const uint32_t e_off = end_of(tid), e_prev = tid ? end_of(tid - 1) : 0u;
    uint32_t n = 0, bits = 0;
    if (e_off > payload || e_prev > e_off || e_off - e_prev < (uint32_t)kSlop) {
      bad_s = 1;
    } else {
      bits = 8u * (e_off - e_prev - (uint32_t)kSlop);
      n = (bits + sub_bits - 1) / sub_bits;
      if (n == 0) n = 1;
    }
    sn[tid] = n;
    seoff[tid] = e_off;
    sbits[tid] = bits;
    sjbad[tid] = n;
  }
  __syncthreads();
  build_dtable<BITS, 3, true>(&bi, bi.syms, T, L1, tid, kSmallItems);  // (ends with a barrier)
  // this CTA's streams
  const int s_lo = (int)(((long long)blockIdx.x * K) / (int)gridDim.x), s_hi = (int)(((long long)(blockIdx.x + 1) * K) / (int)gridDim.x);
  if (tid == 0) {
    uint32_t tot = 0;
    for (int s = s_lo; s < s_hi; ++s) {
      sf[s] = tot;
      tot += sn[s];
    }
    if (tot > (uint32_t)kSmallItems && !bad_s) bad_s = kSplitRetry;  // more than 1.5 times this CTA's share
  }
  __syncthreads();
  if (bad_s) {
    if (tid == 0 && status) atomicOr(status, bad_s);
    return;
  }
  // ---- this thread's item
  int s = -1;
  uint32_t j = 0;
  for (int q = s_lo; q < s_hi; ++q)
    if ((uint32_t)tid >= sf[q] && (uint32_t)tid - sf[q] < sn[q]) {
      s = q;
      j = (uint32_t)tid - sf[q];
    }
  const bool active = s >= 0;
  const bool lone = bi.code_end[0] != 0;  // the empty code: no bits, one item per stream yields its whole slice
  SplitReader rd{};
  uint32_t pad_bits = 0, limit = 0;
  int c_l[kMaxCodeLen - BITS + 1];
  {
    uint32_t base = bi.code_end[BITS] >> (kMaxCodeLen - BITS);
#pragma unroll
    for (int l = BITS + 1; l <= kMaxCodeLen; ++l) {
      const uint32_t first = bi.code_end[l - 1] >> (kMaxCodeLen - l);
      c_l[l - BITS - 1] = (int)base - (int)first;
      base += (bi.code_end[l] - bi.code_end[l - 1]) >> (kMaxCodeLen - l);
    }
  }
  if (active) {
    const uintptr_t end_addr = (uintptr_t)(blk + bi.payload_off) + seoff[s];
    rd.wend = (end_addr + 3) & ~(uintptr_t)3;
    rd.lo_lim = (uintptr_t)blk & ~(uintptr_t)3;
    pad_bits = 8u * (uint32_t)(rd.wend - end_addr);
    limit = (j + 1) * sub_bits;
    if (limit > sbits[s] || j + 1 == sn[s]) limit = sbits[s];
  }
  // ---- pass 0 (k_split_sync): warm-up, then decode to the exit and count
  uint32_t start = 0, cnt = 0, ex = 0;
  if (active && !lone) {
    if (j != 0) {
      const uint32_t first_bit = j * sub_bits;
      const uint32_t warm = sub_bits / 2 < 512u ? sub_bits / 2 : 512u;
      uint32_t dummy = 0;
      start = split_walk<BITS>(T, L1, c_l, rd, pad_bits, first_bit - warm, first_bit, dummy);
    }
    ex = split_walk<BITS>(T, L1, c_l, rd, pad_bits, start, limit, cnt);
  }
  exit_a[tid] = ex;
  __syncthreads();
  // ---- two repair passes
  uint32_t* e_prev_buf = exit_a;
  uint32_t* e_cur_buf = exit_b;
#pragma unroll 1
  for (int pass = 1; pass < 3; ++pass) {
    if (active && !lone && j != 0) {
      const uint32_t want = e_prev_buf[tid - 1];
      if (want != start) {
        start = want;
        cnt = 0;
        ex = split_walk<BITS>(T, L1, c_l, rd, pad_bits, start, limit, cnt);
      }
    }
    e_cur_buf[tid] = ex;
    __syncthreads();
    uint32_t* t = e_prev_buf;
    e_prev_buf = e_cur_buf;
    e_cur_buf = t;
  }
  const uint32_t* exf = e_prev_buf;  // the exits of the last pass
  // ---- scan (k_split_scan): consistent prefix of every stream, offsets, the serial tail, the totals
  if (active && j != 0 && start != exf[tid - 1]) atomicMin(&sjbad[s], j);
  __syncthreads();
  const uint32_t jbad = active ? sjbad[s] : 0u;
  const uint32_t c = (active && j < jbad) ? cnt : 0u;
  const uint32_t incl_w = warp_incl_scan(c);
  if (lane == 31) warp_tot[warp] = incl_w;
  __syncthreads();
  uint32_t before = 0;
  for (int w = 0; w < warp; ++w) before += warp_tot[w];
  const uint32_t incl = before + incl_w;
  if (active && j == 0) sbase[s] = incl - c;
  if (active && j + 1 == sn[s]) scounted[s] = incl;  // minus sbase[s] below
  __syncthreads();
  uint32_t st = 0, sz = 0, off = 0;
  if (active) {
    slice_geom(bi.raw_size, K, s, st, sz);
    off = incl - c - sbase[s];
    const uint32_t counted = scounted[s] - sbase[s];
    const uint32_t n = sn[s];
    bool bad = false;
    if (lone) {
      cnt = j == 0 ? sz : 0u;
      off = 0;
    } else if (jbad < n) {  // what lies behind the last consistent item: one serial item
      if (counted > sz) bad = true;
      if (j > jbad) cnt = 0;
      if (j == jbad) {
        start = exf[tid - 1];
        cnt = sz - counted;
        off = counted;
      }
    } else {  // only the last item can have counted too much (the < 8 pad bits behind the stream)
      if (counted < sz || counted - sz > 7u) bad = true;
      else if (j + 1 == n) {
        if (counted - sz > cnt) bad = true;
        else cnt -= counted - sz;
      }
    }
    if (bad) {
      cnt = 0;
      atomicOr(&bad_s, 1u);
    }
    if (off > sz || cnt > sz - off) cnt = 0;  // (never outside the slice, whatever the input)
  }
  // ---- write pass
  if (active) split_write_lane<BITS>(T, c_l, rd, pad_bits, start, cnt, raw + st + off);
  __syncthreads();
  if (tid == 0 && bad_s && status) atomicOr(status, 1u);
}

// One CTA per stream.  final_pass: the pass whose exits are current.
__global__ void __launch_bounds__(256)
k_split_scan(const uint8_t* __restrict__ comp, const unsigned long long* __restrict__ offsets, uint32_t n_blocks, int K,
             uint64_t raw_n, uint32_t block_size, SplitArgs sa, int final_pass, uint32_t* __restrict__ status) {
  __shared__ uint32_t jbad_s, warp_tot[8], running_s;
  const uint32_t g = blockIdx.x;
  const uint32_t b = g / (uint32_t)K, s = g % (uint32_t)K;
  if (b >= n_blocks || sa.block_ctas[b] == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = sa.s_nitems[g];
  const unsigned long long gi0 = sa.cta_first[b] * (unsigned long long)kSplitThreads + sa.s_first[g];
  const uint32_t* ex = (final_pass & 1) ? sa.exit_bit[1] : sa.exit_bit[0];
  const uint64_t roff = (uint64_t)b * block_size;
  const uint32_t bn = (uint32_t)((raw_n - roff) < (uint64_t)block_size ? (raw_n - roff) : (uint64_t)block_size);
  uint32_t st, sz;
  slice_geom(bn, K, (int)s, st, sz);
  // a lone symbol (empty code): the stream has no bits and one item, which yields the whole slice
  const uint8_t* blk = comp + offsets[b];
  const uint32_t mask = (uint32_t)blk[4] | ((uint32_t)blk[5] << 8) | ((uint32_t)blk[6] << 16) | ((uint32_t)blk[7] << 24);
  const bool lone = (mask & 1u) != 0;
  if (tid == 0) {
    jbad_s = n;
    running_s = 0;
  }
  __syncthreads();
  // first item whose start is not its predecessor's exit (item 0 starts on bit 0 by construction)
  for (uint32_t j = 1 + (uint32_t)tid; j < n; j += 256)
    if (sa.start[gi0 + j] != ex[gi0 + j - 1]) atomicMin(&jbad_s, j);
  __syncthreads();
  const uint32_t jbad = jbad_s;
  // offsets of the consistent items: exclusive prefix sums of their counts
  for (uint32_t j0 = 0; j0 < jbad; j0 += 256) {
    const uint32_t j = j0 + (uint32_t)tid;
    const uint32_t c = j < jbad ? sa.cnt[gi0 + j] : 0u;
    const uint32_t incl = warp_incl_scan(c);
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t before = running_s;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    if (j < jbad) sa.out_off[gi0 + j] = before + incl - c;
    __syncthreads();
    if (tid == 255) running_s = before + incl;
    __syncthreads();
  }
  const uint32_t counted = running_s;
  bool bad = false;
  if (jbad < n) {  // the rest of the stream as one item from the last good exit on
    for (uint32_t j = jbad + 1 + (uint32_t)tid; j < n; j += 256) sa.cnt[gi0 + j] = 0;
    if (tid == 0) {
      if (counted > sz) bad = true;
      sa.start[gi0 + jbad] = ex[gi0 + jbad - 1];
      sa.cnt[gi0 + jbad] = bad ? 0u : sz - counted;
      sa.out_off[gi0 + jbad] = bad ? 0u : counted;
    }
  } else if (tid == 0) {
    // every item is in step.  Only the last one can have counted too much: the < 8 pad bits behind
    // the stream decode to at most 7 symbols.
    const uint32_t last = sa.cnt[gi0 + n - 1];
    if (lone) {
      sa.cnt[gi0] = sz;
      sa.out_off[gi0] = 0;
    } else if (counted < sz || counted - sz > 7u || counted - sz > last) {
      bad = true;
    } else {
      sa.cnt[gi0 + n - 1] = last - (counted - sz);
    }
  }
  if (bad) {  // malformed: the stream yields nothing, and the call reports it
    for (uint32_t j = 0; j < n; ++j) sa.cnt[gi0 + j] = 0;
    if (status) atomicOr(status, 1u);
  }
}

// Dumps the reference-format decode tables (BITS = 12) built by the decode kernel's builder:
// one_symbol == 0: DecodedSym2x {num_bits_decoded, syms[2], num_syms} (:634-640), 4 bytes per entry;
// one_symbol != 0: DecodedSym {code_len, sym} of Decoder1x (:588-632), 2 bytes per entry.
__global__ void __launch_bounds__(256) k_dump_dtable(const uint16_t* __restrict__ len_count,
                                                     const uint8_t* __restrict__ syms, int num_syms,
                                                     int one_symbol, uint8_t* __restrict__ out) {
  __shared__ uint32_t T[4096];
  __shared__ uint8_t T1[4096];
  __shared__ DecBlockInfo bi;
  __shared__ uint8_t sy[256];
  if (threadIdx.x == 0) {
    uint32_t nsyms = 0, code = 0;
    for (int l = 0; l <= kMaxCodeLen; ++l) {
      bi.first_idx[l] = nsyms;
      nsyms += len_count[l];
      code += (uint32_t)len_count[l] << (kMaxCodeLen - l);
      bi.code_end[l] = code;
    }
    bi.num_syms = (uint32_t)num_syms;
  }
  for (int i = threadIdx.x; i < num_syms; i += blockDim.x) sy[i] = syms[i];
  __syncthreads();
  if (one_symbol) {
    build_dtable<kMaxCodeLen, 1>(&bi, sy, T, T1, threadIdx.x, blockDim.x);
    for (int e = threadIdx.x; e < 4096; e += blockDim.x) {
      const uint32_t v = T[e];  // 0 where no code starts (the reference leaves {0, 0} there)
      out[2 * e + 0] = (uint8_t)((v >> 24) & 15u);
      out[2 * e + 1] = (uint8_t)v;
    }
    return;
  }
  build_dtable<kMaxCodeLen, 2>(&bi, sy, T, T1, threadIdx.x, blockDim.x);
  for (int e = threadIdx.x; e < 4096; e += blockDim.x) {
    const uint32_t v = T[e];
    out[4 * e + 0] = (uint8_t)((v >> 24) & 15u);
    out[4 * e + 1] = (uint8_t)v;
    out[4 * e + 2] = (uint8_t)(v >> 8);
    out[4 * e + 3] = (uint8_t)(v >> 30);
  }
}

// ===========================================================================
// One large buffer (CompressMulti<K> of a single string_view, codec/huffman.cpp:738-846) spread
// over the whole device.  The format gives K streams and nothing finer, so the encoder makes its
// own units: every stream is cut into pieces of kPieceSyms symbols.
//   k_histogram_streams  per-stream histograms, as the reference takes them (:758-766)
//   k_single_plan        total histogram -> table; stream bit totals = sum over symbols of
//                        count x code length (:776-782); region ends; header; the plan
//   k_piece_lengths      bit total of every piece (one warp per piece)
//   k_encode_pieces      one warp per piece: its start bit is the sum of the pieces in front of
//                        it; staged in shared memory like a block's stream and ORed into the
//                        zero-filled output (pieces meet inside bytes)
// ===========================================================================
struct SinglePlan {
  uint32_t total_size;           // size of the compressed buffer
  uint32_t hdr_total;            // header incl. the end-offset table
  uint32_t bad;                  // a symbol without a code (caller-supplied table)
  uint32_t pad;
  uint32_t region_end[kMaxK];    // cumulative, relative to hdr_total (:783-786)
  uint32_t piece_first[kMaxK + 1];  // index of each stream's first piece
};

__host__ __device__ inline uint32_t pieces_of(uint32_t sz) { return (sz + kPieceSyms - 1) / kPieceSyms; }

__global__ void __launch_bounds__(kHistThreads)
k_histogram_streams(const uint8_t* __restrict__ in, uint32_t n, int K, uint32_t* __restrict__ hist /* [K][256] */) {
  __shared__ uint32_t bins[256 * 32];
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) bins[i] = 0;
  __syncthreads();
  const uint32_t nvec_total = (n + 15) >> 4;
  const uint32_t per = (nvec_total + gridDim.x - 1) / gridDim.x;
  uint64_t pos = (uint64_t)blockIdx.x * per * 16;
  uint64_t end = pos + (uint64_t)per * 16;
  if (end > n) end = n;
  const uint32_t q = n / (uint32_t)K, r = n % (uint32_t)K;
  const uint64_t big = (uint64_t)r * (q + 1);  // the first r slices have q + 1 bytes (:98-108)
  while (pos < end) {  // the CTA's share, slice by slice
    uint32_t s;
    uint64_t s_end;
    if (pos < big) {
      s = (uint32_t)(pos / (q + 1));
      s_end = (uint64_t)(s + 1) * (q + 1);
    } else {
      s = r + (uint32_t)((pos - big) / q);
      s_end = big + (uint64_t)(s - r + 1) * q;
    }
    const uint64_t seg_end = s_end < end ? s_end : end;
    bins_accumulate(bins, in + pos, seg_end - pos, threadIdx.x, blockDim.x);
    __syncthreads();
    if (threadIdx.x < 256) {
      const uint32_t c = bins_reduce_clear(bins, threadIdx.x);
      if (c) atomicAdd(hist + (size_t)s * 256 + threadIdx.x, c);
    }
    __syncthreads();
    pos = seg_end;
  }
}

// Header prefix of a compressed buffer (:799-808): raw size, length mask, counts, symbols.
__device__ inline void write_header_prefix(const HufTable& tab, uint32_t raw_size, uint8_t* dst, int tid, int nthreads) {
  const uint32_t hdr = tab.hdr_len;
  const uint32_t mask = tab.len_mask;
  const uint32_t npop = (uint32_t)__popc(mask);
  for (uint32_t i = tid; i < hdr; i += nthreads) {
    uint8_t v;
    if (i < 4) v = (uint8_t)(raw_size >> (8 * i));
    else if (i < 8) v = (uint8_t)(mask >> (8 * (i - 4)));
    else if (i < 8 + npop) {
      int bit = 0;  // position of the (i-8)-th set bit of the mask
      for (uint32_t seen = 0;; ++bit)
        if ((mask >> bit) & 1u) {
          if (seen == i - 8) break;
          ++seen;
        }
      v = (uint8_t)tab.len_count[bit];  // 256 wraps to 0 (:804)
    } else v = tab.sorted_syms[i - 8 - npop];
    dst[i] = v;
  }
}

// One CTA of 256 threads.  shared_tab: nullptr = build the table from the total histogram.
__global__ void __launch_bounds__(256) k_single_plan(const uint32_t* __restrict__ hist, uint32_t n, int K,
                                                     const HufTable* __restrict__ shared_tab,
                                                     HufTable* __restrict__ tab_out, SinglePlan* __restrict__ plan,
                                                     uint8_t* __restrict__ dst) {
  __shared__ HufTable tab;
  __shared__ TableScratch sc;
  __shared__ uint32_t tot[256];
  __shared__ unsigned long long sbits[kMaxK];
  __shared__ uint32_t bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) bad = 0;
  {
    uint32_t t = 0;
    for (int s = 0; s < K; ++s) t += hist[(size_t)s * 256 + tid];
    tot[tid] = t;
  }
  __syncthreads();
  if (shared_tab) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(shared_tab);
    uint32_t* d = reinterpret_cast<uint32_t*>(&tab);
    for (int i = tid; i < (int)(sizeof(HufTable) / 4); i += 256) d[i] = src[i];
  } else if (warp == 0) {
    if (n < (1u << 24)) build_table_warp<uint32_t, uint32_t>(tot, &tab, &sc);
    else build_table_warp<uint32_t, unsigned long long>(tot, &tab, &sc);
  }
  __syncthreads();
  if (tot[tid] != 0 && tab.enc[tid] == kEncInvalid) atomicOr(&bad, 1u);  // only possible with a supplied table
  // stream bit totals (:776-782)
  for (int s = warp; s < K; s += 8) {
    unsigned long long b = 0;
    for (int c = lane; c < 256; c += 32) {
      const uint32_t e = tab.enc[c];
      if (e != kEncInvalid) b += (unsigned long long)hist[(size_t)s * 256 + c] * (e >> 16);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) b += __shfl_xor_sync(0xffffffffu, b, d);
    if (lane == 0) sbits[s] = b;
  }
  __syncthreads();
  const uint32_t hdr_total = tab.hdr_len + 4u * (uint32_t)(K - 1);
  if (tid == 0) {
    const uint32_t q = n / (uint32_t)K, r = n % (uint32_t)K;
    uint32_t pos = 0, pc = 0;
    for (int s = 0; s < K; ++s) {
      pos += (uint32_t)((sbits[s] + 7) >> 3) + kSlop;
      plan->region_end[s] = pos;
      plan->piece_first[s] = pc;
      pc += pieces_of(q + ((uint32_t)s < r ? 1u : 0u));
    }
    plan->piece_first[K] = pc;
    plan->hdr_total = hdr_total;
    plan->total_size = bad ? 0u : hdr_total + pos;
    plan->bad = bad;
  }
  write_header_prefix(tab, n, dst, tid, 256);
  __syncthreads();
  for (int s = tid; s < K - 1; s += 256) {  // end_offset table (:809-811)
    const uint32_t e = plan->region_end[s];
    uint8_t* p = dst + tab.hdr_len + 4 * s;
    p[0] = (uint8_t)e;
    p[1] = (uint8_t)(e >> 8);
    p[2] = (uint8_t)(e >> 16);
    p[3] = (uint8_t)(e >> 24);
  }
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&tab);
    uint32_t* d = reinterpret_cast<uint32_t*>(tab_out);
    for (int i = tid; i < (int)(sizeof(HufTable) / 4); i += 256) d[i] = src[i];
  }
}

// (stream, piece index inside the stream) of the global piece p; warp-uniform
__device__ __forceinline__ void piece_of(const SinglePlan* plan, int K, uint32_t p, uint32_t& s, uint32_t& j) {
  uint32_t lo = 0, hi = (uint32_t)K;  // piece_first[lo] <= p < piece_first[hi]
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (plan->piece_first[mid] <= p) lo = mid;
    else hi = mid;
  }
  s = lo;
  j = p - plan->piece_first[lo];
}

constexpr int kPieceWarps = 8;

__global__ void __launch_bounds__(32 * kPieceWarps)
k_piece_lengths(const uint8_t* __restrict__ raw, uint32_t n, int K, const HufTable* __restrict__ tab,
                const SinglePlan* __restrict__ plan, uint32_t* __restrict__ piece_bits) {
  __shared__ uint32_t enc[256];
  __shared__ uint32_t bad;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) enc[i] = tab->enc[i];
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  const uint32_t p = blockIdx.x * kPieceWarps + (threadIdx.x >> 5);
  if (p >= plan->piece_first[K]) return;
  uint32_t s, j;
  piece_of(plan, K, p, s, j);
  uint32_t st, sz;
  slice_geom(n, K, (int)s, st, sz);
  const uint32_t off = j * kPieceSyms;
  const uint32_t len = sz - off < kPieceSyms ? sz - off : kPieceSyms;
  const unsigned long long b = stream_length_warp(enc, raw + st + off, len, &bad, raw + n);
  if ((threadIdx.x & 31) == 0) piece_bits[p] = (uint32_t)b;
}

__global__ void __launch_bounds__(32 * kPieceWarps)
k_encode_pieces(const uint8_t* __restrict__ raw, uint32_t n, int K, const HufTable* __restrict__ tab_g,
                const SinglePlan* __restrict__ plan, const uint32_t* __restrict__ piece_bits,
                uint8_t* __restrict__ dst) {
  extern __shared__ __align__(1024) uint8_t esm[];
  // layout: stage[kPieceWarps][kStageWords] u32 (each 1 KiB-aligned: 5 KiB) | HufTable
  uint32_t* stage = reinterpret_cast<uint32_t*>(esm);
  HufTable* tab = reinterpret_cast<HufTable*>(esm + (size_t)kPieceWarps * kStageWords * 4);
  {
    uint4* z = reinterpret_cast<uint4*>(esm);
    for (int i = threadIdx.x; i < kPieceWarps * kStageWords / 4; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tab_g);
    uint32_t* d = reinterpret_cast<uint32_t*>(tab);
    for (int i = threadIdx.x; i < (int)(sizeof(HufTable) / 4); i += blockDim.x) d[i] = src[i];
  }
  __syncthreads();
  if (plan->bad) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t p = blockIdx.x * kPieceWarps + warp;
  if (p >= plan->piece_first[K]) return;
  uint32_t s, j;
  piece_of(plan, K, p, s, j);
  uint32_t st, sz;
  slice_geom(n, K, (int)s, st, sz);
  const uint32_t off = j * kPieceSyms;
  const uint32_t len = sz - off < kPieceSyms ? sz - off : kPieceSyms;
  // start bit of the piece inside its stream
  unsigned long long b0 = 0;
  {
    const uint32_t* pb = piece_bits + plan->piece_first[s];
    for (uint32_t i = lane; i < j; i += 32) b0 += pb[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) b0 += __shfl_xor_sync(0xffffffffu, b0, d);
  }
  const uint32_t stage_base = smem_u32(stage + (size_t)warp * kStageWords);
  bool over;
  const uint32_t endbit = (uint32_t)encode_stream_staged_warp<true>(smem_u32(tab->enc), stage_base, raw + st + off, len,
                                                                    &over, raw + n, (uint32_t)(b0 & 31u));
  // Buffer word i is stream word m_base + i; output word m = top half of (W[m-1] : W[m]) << 8r
  // (see copy_stream_out_warp) is linear in the stream's bits, so every piece ORs in what ITS
  // bits contribute: words m_base .. m_base + nw (the last one only through W[m-1]).  The two
  // words at either end may also get bits of the neighbouring pieces: atomics; the rest is ours.
  const uint32_t e_off = plan->hdr_total + plan->region_end[s];
  const uint32_t r = ((e_off - 1u) & 3u) + 1u;
  const uint32_t sh = 8u * r;
  uint32_t* wend = reinterpret_cast<uint32_t*>(dst + (e_off - r));
  const uint32_t m_base = (uint32_t)(b0 >> 5);
  const uint32_t nw = (endbit + 31u) >> 5;
  const uint32_t sb = stage_base + 4u * kStageFront;
  for (uint32_t i = lane; i <= nw; i += 32) {
    const uint32_t lo = i < nw ? lds_u32(sb + 4u * i) : 0u;
    const uint32_t hi = i >= 1 ? lds_u32(sb + 4u * (i - 1)) : 0u;
    const uint32_t v = __funnelshift_lc(lo, hi, sh);
    uint32_t* a = wend - (m_base + i);
    if (i < 2 || i + 2 > nw) {
      if (v) atomicOr(a, v);
    } else {
      *a = v;
    }
  }
}

// ===========================================================================
// Slot layout -> packed layout
// ===========================================================================
__global__ void __launch_bounds__(1024) k_scan_sizes(const uint32_t* __restrict__ sizes, uint32_t n,
                                                     unsigned long long* __restrict__ offsets,
                                                     unsigned long long* __restrict__ total) {
  __shared__ unsigned long long warp_sums[32];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    unsigned long long v = i < n ? sizes[i] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
      if ((threadIdx.x & 31) >= d) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned long long woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += warp_sums[w];
    const unsigned long long c = carry;
    if (i < n) offsets[i] = c + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(256) k_pack(const uint8_t* __restrict__ slots, uint64_t slot_stride,
                                              const uint32_t* __restrict__ sizes,
                                              const unsigned long long* __restrict__ offsets,
                                              uint8_t* __restrict__ packed) {
  const uint32_t b = blockIdx.x;
  const uint8_t* src = slots + (uint64_t)b * slot_stride;
  uint8_t* dst = packed + offsets[b];
  const uint32_t sz = sizes[b];
  // destination-aligned 16-byte stores, byte-granular edges
  const uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
  const uint32_t h = head < sz ? head : sz;
  for (uint32_t i = threadIdx.x; i < h; i += blockDim.x) dst[i] = src[i];
  const uint32_t nvec = (sz - h) >> 4;
  for (uint32_t v = threadIdx.x; v < nvec; v += blockDim.x) {
    const uint8_t* s = src + h + 16 * v;
    uint32_t w[4];
    if ((h & 3) == 0) {
      const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s);
      w[0] = s32[0]; w[1] = s32[1]; w[2] = s32[2]; w[3] = s32[3];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        w[j] = (uint32_t)s[4 * j] | ((uint32_t)s[4 * j + 1] << 8) | ((uint32_t)s[4 * j + 2] << 16) |
               ((uint32_t)s[4 * j + 3] << 24);
    }
    *reinterpret_cast<uint4*>(dst + h + 16 * v) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (uint32_t i = h + 16 * nvec + threadIdx.x; i < sz; i += blockDim.x) dst[i] = src[i];
}

// ===========================================================================
// Launchers (called from huf_api.cu)
// ===========================================================================
size_t table_bytes() { return sizeof(HufTable); }

cudaError_t launch_histogram(const uint8_t* d_in, uint64_t n, unsigned long long* d_out, int grid,
                             cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(d_out, 0, 256 * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  if (n == 0) return cudaSuccess;
  k_histogram<<<grid, kHistThreads, 0, st>>>(d_in, n, d_out);
  return cudaGetLastError();
}

cudaError_t launch_build_table(const unsigned long long* d_hist, void* d_table, cudaStream_t st) {
  k_build_table<<<1, 32, 0, st>>>(d_hist, reinterpret_cast<HufTable*>(d_table));
  return cudaGetLastError();
}

cudaError_t launch_make_table(const uint32_t* d_hist, const uint16_t* d_len_count, const uint8_t* d_syms,
                              int n, int mode, void* d_table, cudaStream_t st) {
  k_make_table<<<1, 32, 0, st>>>(d_hist, d_len_count, d_syms, n, mode, reinterpret_cast<HufTable*>(d_table));
  return cudaGetLastError();
}

static std::atomic<uint32_t> g_counter_slot{0};

cudaError_t launch_compress(const uint8_t* d_raw, uint64_t n, uint32_t block_size, int K, uint32_t n_blocks,
                            uint8_t* d_out, uint64_t slot_stride, uint32_t* d_sizes, const void* d_table,
                            int check_presence, uint32_t* d_status, int grid, cudaStream_t st) {
  if (n_blocks == 0) return cudaSuccess;
  // two builds of the kernel: slices of at most kStageSlice symbols (no long-stream code), and the rest
  const bool long_slices = (block_size + (uint32_t)K - 1) / (uint32_t)K > (uint32_t)kStageSlice;
  auto kernel = long_slices ? k_compress_blocks<true> : k_compress_blocks<false>;
  // the attribute is per device and per kernel: remember where it has been set (bit per ordinal)
  static std::atomic<unsigned long long> configured[3];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  // small blocks with tables of their own: batches of blocks with their tables built side by side
  static const uint32_t batch_max = [] {  // tuning aid: HUFB200_COMP_BATCH_MAX=largest block size of the batch kernel (0 = never)
    const char* v = getenv("HUFB200_COMP_BATCH_MAX");
    return v ? (uint32_t)atoi(v) : (32u << 10);
  }();
  if (!long_slices && d_table == nullptr && block_size <= batch_max && K <= 32) {  // (K = 48: measured 2 % slower)
    if (dev >= 64 || !((configured[2].load(std::memory_order_acquire) >> dev) & 1ull)) {
      e = cudaFuncSetAttribute(k_compress_small_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CompSmemBatch));
      if (e != cudaSuccess) return e;
      if (dev < 64) configured[2].fetch_or(1ull << dev, std::memory_order_release);
    }
    k_compress_small_blocks<<<grid, kCompThreads, sizeof(CompSmemBatch), st>>>(
        d_raw, n, block_size, K, n_blocks, d_out, slot_stride, d_sizes, d_status,
        g_counter_slot.fetch_add(1, std::memory_order_relaxed) % kCounterSlots);
    return cudaGetLastError();
  }
  if (dev >= 64 || !((configured[long_slices].load(std::memory_order_acquire) >> dev) & 1ull)) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CompSmem));
    if (e != cudaSuccess) return e;
    if (dev < 64) configured[long_slices].fetch_or(1ull << dev, std::memory_order_release);
  }
  kernel<<<grid, kCompThreads, sizeof(CompSmem), st>>>(d_raw, n, block_size, K, n_blocks, d_out, slot_stride, d_sizes,
                                                       reinterpret_cast<const HufTable*>(d_table), check_presence, d_status,
                                                       g_counter_slot.fetch_add(1, std::memory_order_relaxed) % kCounterSlots);
  return cudaGetLastError();
}

static size_t decompress_smem_bytes_threads(int bpc, int nthreads, int K, uint32_t block_size) {
  const int entries = dec_entries(dec_bits_for(K, block_size));
  return (size_t)bpc * entries * 4 + dec_region_bytes(bpc, nthreads, entries) + (size_t)bpc * sizeof(DecBlockHdr);
}
size_t decompress_smem_bytes(int K, int bpc, uint32_t block_size) {
  return decompress_smem_bytes_threads(bpc, ((K * bpc + 31) / 32) * 32, K, block_size);
}

cudaError_t launch_decompress(const uint8_t* d_comp, const unsigned long long* d_offsets,
                              const uint32_t* d_sizes, uint32_t n_blocks, int K, int bpc, uint8_t* d_raw,
                              uint64_t raw_n, uint32_t block_size, uint32_t* d_status, cudaStream_t st) {
  if (n_blocks == 0) return cudaSuccess;
  int nthreads = ((K * bpc + 31) / 32) * 32;
  const int bits = dec_bits_for(K, block_size);
  const uint32_t grid = (n_blocks + bpc - 1) / bpc;
  // A launch that leaves most of the device empty (a single buffer, a few blocks) gets full-size
  // CTAs: the extra warps have no stream, they only help build the tables and leave.
  if (grid <= 64 && nthreads < kDecMaxThreads) nthreads = kDecMaxThreads;
  size_t smem = decompress_smem_bytes_threads(bpc, nthreads, K, block_size);
  static const size_t pad = [] {  // tuning aid: HUFB200_DEC_PAD=bytes of unused shared memory per CTA (lowers residency)
    const char* e = getenv("HUFB200_DEC_PAD");
    return e ? (size_t)atoi(e) : (size_t)0;
  }();
  smem += pad;
  auto kernel = bits == 9 ? k_decompress_blocks<9> : (bits == 10 ? k_decompress_blocks<10> : k_decompress_blocks<11>);
  {  // opt-in beyond the 48 KiB default: the attribute is per device and per kernel and shared by
     // all host threads, so it is set once per device, to the device's limit -- never per launch
     // (two threads decoding with different K would otherwise lower it under each other)
    static std::atomic<unsigned long long> configured[3];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !((configured[bits - 9].load(std::memory_order_acquire) >> dev) & 1ull)) {
      int optin = 0;
      e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
      if (e != cudaSuccess) return e;
      if (const char* c = getenv("HUFB200_DEC_CARVEOUT")) {  // tuning aid: shared-memory carve-out in percent
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(c));
        if (e != cudaSuccess) return e;
      }
      if (dev < 64) configured[bits - 9].fetch_or(1ull << dev, std::memory_order_release);
    }
  }
  kernel<<<grid, nthreads, smem, st>>>(d_comp, d_offsets, d_sizes, n_blocks, K, bpc, d_raw, raw_n, block_size, d_status,
                                       SplitArgs{});
  return cudaGetLastError();
}

// ---- split decode: workspace layout and launches
uint32_t split_sub_bits(uint64_t raw_n, uint32_t n_blocks, int K, int sms) {
  // items of about (compressed bits) / (a quarter of the lanes the device holds), between 1024 and 8192 bits
  (void)n_blocks;
  (void)K;
  static const uint32_t forced = [] {  // tuning / test aid: HUFB200_SPLIT_BITS=64..65536 fixes the item length
    const char* e = getenv("HUFB200_SPLIT_BITS");
    const int v = e ? atoi(e) : 0;
    return v >= 64 && v <= 65536 ? (uint32_t)v : 0u;
  }();
  if (forced) return forced;
  // (measured, profiles/r2_split_decode.md: 8192-bit items are as fast as any for a full device;
  // a 64 MiB buffer does best with 4096, small buffers need every lane they can get)
  const uint64_t lanes = (uint64_t)sms * 512u;
  uint64_t want = raw_n * 4u / (lanes ? lanes : 1u);  // ~4 bits per symbol
  uint32_t sub = 1024;
  while (sub < 8192u && sub < want) sub <<= 1;
  return sub;
}
static uint32_t split_max_block_ctas(uint32_t block_size, int K, uint32_t sub_bits) {
  // no block has more compressed bytes than 12 bits per symbol plus the header and the slop
  const uint64_t bound_bits = (uint64_t)block_size * kMaxCodeLen + 8u * (uint64_t)(8 + 13 + 256 + 4 * K + kSlop * K);
  const uint64_t items = bound_bits / sub_bits + (uint64_t)K + 1u;
  return (uint32_t)((items + kSplitThreads - 1) / kSplitThreads);
}
static SplitArgs split_layout(uint8_t* base, uint32_t n_blocks, int K, uint32_t block_size, uint32_t sub_bits,
                              size_t* total) {
  SplitArgs a{};
  a.sub_bits = sub_bits;
  a.max_block_ctas = split_max_block_ctas(block_size, K, sub_bits);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  };
  const size_t ns = (size_t)n_blocks * K;
  const size_t items = (size_t)n_blocks * a.max_block_ctas * kSplitThreads;
  a.cta_first = reinterpret_cast<unsigned long long*>(take(8 * ((size_t)n_blocks + 1)));
  a.total_ctas = reinterpret_cast<unsigned long long*>(take(8));
  a.block_ctas = reinterpret_cast<uint32_t*>(take(4 * (size_t)n_blocks));
  a.s_first = reinterpret_cast<uint32_t*>(take(4 * ns));
  a.s_nitems = reinterpret_cast<uint32_t*>(take(4 * ns));
  a.s_eoff = reinterpret_cast<uint32_t*>(take(4 * ns));
  a.s_bits = reinterpret_cast<uint32_t*>(take(4 * ns));
  a.start = reinterpret_cast<uint32_t*>(take(4 * items));
  a.cnt = reinterpret_cast<uint32_t*>(take(4 * items));
  a.exit_bit[0] = reinterpret_cast<uint32_t*>(take(4 * items));
  a.exit_bit[1] = reinterpret_cast<uint32_t*>(take(4 * items));
  a.out_off = reinterpret_cast<uint32_t*>(take(4 * items));
  *total = off;
  return a;
}
size_t decompress_split_work_bytes(uint32_t n_blocks, int K, uint32_t block_size, uint32_t sub_bits) {
  size_t total = 0;
  split_layout(nullptr, n_blocks, K, block_size, sub_bits, &total);
  return total;
}

// One small buffer (compressed size known to the host): the whole split decode as one CTA.
// (one SM: 45 us per 100 KiB of raw bytes; the seven launches of the spread form take about 110 us
// for anything up to a few MiB -- profiles/r2_split_decode.md)
bool split_small_fits(uint64_t comp_bytes, int K) { return comp_bytes <= (1u << 20) && K < kSmallItems / 2; }
cudaError_t launch_decompress_split_small(const uint8_t* d_comp, const uint32_t* d_size, uint64_t comp_bytes, int K,
                                          uint8_t* d_raw, uint32_t raw_n, uint32_t* d_status, cudaStream_t st) {
  const int ctas = K < kSmallMaxCtas ? K : kSmallMaxCtas;
  // items so that a CTA with 1.5 times its share of the bits still fits: a multiple of 32, at least 384
  const uint64_t cap = (uint64_t)(kSmallItems - (K + ctas - 1) / ctas);
  uint64_t sub = (12u * comp_bytes + cap * ctas - 1) / (cap * ctas);
  sub = (sub + 31) & ~(uint64_t)31;
  if (sub < 384) sub = 384;
  k_split_small<kSplitBits><<<ctas, kSmallItems, 0, st>>>(d_comp, d_size, K, d_raw, raw_n, (uint32_t)sub, d_status);
  return cudaGetLastError();
}

constexpr int kSplitPasses = 3;  // pass 0 and two repairs; what is still out of step decodes serially (k_split_scan)

cudaError_t launch_decompress_split(const uint8_t* d_comp, const unsigned long long* d_offsets,
                                    const uint32_t* d_sizes, uint32_t n_blocks, int K, uint8_t* d_raw, uint64_t raw_n,
                                    uint32_t block_size, uint32_t sub_bits, void* d_work, uint32_t* d_status,
                                    int* launches, cudaStream_t st) {
  *launches = 0;
  if (n_blocks == 0) return cudaSuccess;
  size_t total = 0;
  const SplitArgs sa = split_layout(reinterpret_cast<uint8_t*>(d_work), n_blocks, K, block_size, sub_bits, &total);
  const unsigned long long max_ctas = (unsigned long long)n_blocks * sa.max_block_ctas;
  if (max_ctas >> 31) return cudaErrorInvalidValue;
  k_split_plan<<<n_blocks, 32, 0, st>>>(d_comp, d_offsets, d_sizes, n_blocks, K, raw_n, block_size, sa, d_status);
  k_scan_sizes<<<1, 1024, 0, st>>>(sa.block_ctas, n_blocks, sa.cta_first, sa.total_ctas);
  for (int pass = 0; pass < kSplitPasses; ++pass)
    k_split_sync<kSplitBits><<<(uint32_t)max_ctas, kSplitThreads, 0, st>>>(d_comp, d_offsets, d_sizes, n_blocks, K, raw_n,
                                                                           block_size, sa, pass);
  k_split_scan<<<n_blocks * (uint32_t)K, 256, 0, st>>>(d_comp, d_offsets, n_blocks, K, raw_n, block_size, sa,
                                                        kSplitPasses - 1, d_status);
  *launches = 3 + kSplitPasses;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // the write pass: the decode kernel with one lane per item
  auto kernel = k_decompress_blocks<kSplitBits, true>;
  const int entries = dec_entries(kSplitBits);
  const size_t smem = (size_t)entries * 4 + dec_region_bytes(1, kSplitThreads, entries) + sizeof(DecBlockHdr);
  {
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !((configured.load(std::memory_order_acquire) >> dev) & 1ull)) {
      // (this instance always takes the same amount: one table of kSplitBits bits, kSplitThreads lanes)
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      if (dev < 64) configured.fetch_or(1ull << dev, std::memory_order_release);
    }
  }
  kernel<<<(uint32_t)max_ctas, kSplitThreads, smem, st>>>(d_comp, d_offsets, d_sizes, n_blocks, K, 1, d_raw, raw_n,
                                                          block_size, d_status, sa);
  *launches += 1;
  return cudaGetLastError();
}

size_t single_plan_bytes() { return sizeof(SinglePlan); }

uint32_t single_piece_count(uint32_t n, int K) {
  const uint32_t q = n / (uint32_t)K, r = n % (uint32_t)K;
  return r * pieces_of(q + 1) + ((uint32_t)K - r) * pieces_of(q);
}

// d_hist: K x 256 u32 (zeroed here), d_table_out: HufTable, d_plan: SinglePlan, d_piece_bits: one u32
// per piece, d_out: zero-filled by the caller up to hufb200_compress_bound(n, K) + 8 bytes.
cudaError_t launch_compress_single(const uint8_t* d_raw, uint32_t n, int K, const void* d_shared_table,
                                   uint32_t* d_hist, void* d_table_out, void* d_plan, uint32_t* d_piece_bits,
                                   uint8_t* d_out, int sms, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(d_hist, 0, (size_t)K * 256 * sizeof(uint32_t), st);
  if (e != cudaSuccess) return e;
  const uint32_t want = (n + 65535) / 65536;
  const uint32_t grid = want < (uint32_t)sms * 4 ? (want ? want : 1) : (uint32_t)sms * 4;
  k_histogram_streams<<<grid, kHistThreads, 0, st>>>(d_raw, n, K, d_hist);
  k_single_plan<<<1, 256, 0, st>>>(d_hist, n, K, reinterpret_cast<const HufTable*>(d_shared_table),
                                   reinterpret_cast<HufTable*>(d_table_out), reinterpret_cast<SinglePlan*>(d_plan), d_out);
  const uint32_t pieces = single_piece_count(n, K);
  const uint32_t pgrid = (pieces + kPieceWarps - 1) / kPieceWarps;
  k_piece_lengths<<<pgrid, 32 * kPieceWarps, 0, st>>>(d_raw, n, K, reinterpret_cast<const HufTable*>(d_table_out),
                                                      reinterpret_cast<const SinglePlan*>(d_plan), d_piece_bits);
  const size_t smem = (size_t)kPieceWarps * kStageWords * 4 + sizeof(HufTable);
  static std::atomic<unsigned long long> configured{0};
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 64 || !((configured.load(std::memory_order_acquire) >> dev) & 1ull)) {
    e = cudaFuncSetAttribute(k_encode_pieces, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev < 64) configured.fetch_or(1ull << dev, std::memory_order_release);
  }
  k_encode_pieces<<<pgrid, 32 * kPieceWarps, smem, st>>>(d_raw, n, K, reinterpret_cast<const HufTable*>(d_table_out),
                                                         reinterpret_cast<const SinglePlan*>(d_plan), d_piece_bits, d_out);
  return cudaGetLastError();
}

cudaError_t launch_dump_dtable(const uint16_t* d_len_count, const uint8_t* d_syms, int num_syms, int one_symbol,
                               uint8_t* d_out, cudaStream_t st) {
  k_dump_dtable<<<1, 256, 0, st>>>(d_len_count, d_syms, num_syms, one_symbol, d_out);
  return cudaGetLastError();
}

cudaError_t launch_pack(const uint8_t* d_slots, uint64_t slot_stride, const uint32_t* d_sizes, uint32_t n_blocks,
                        uint8_t* d_packed, unsigned long long* d_offsets, unsigned long long* d_total,
                        cudaStream_t st) {
  k_scan_sizes<<<1, 1024, 0, st>>>(d_sizes, n_blocks, d_offsets, d_total);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || n_blocks == 0 || d_packed == nullptr) return e;
  k_pack<<<n_blocks, 256, 0, st>>>(d_slots, slot_stride, d_sizes, d_offsets, d_packed);
  return cudaGetLastError();
}

cudaError_t launch_scan_sizes(const uint32_t* d_sizes, uint32_t n_blocks, unsigned long long* d_offsets,
                              unsigned long long* d_total, cudaStream_t st) {
  k_scan_sizes<<<1, 1024, 0, st>>>(d_sizes, n_blocks, d_offsets, d_total);
  return cudaGetLastError();
}

}  // namespace hufb200
