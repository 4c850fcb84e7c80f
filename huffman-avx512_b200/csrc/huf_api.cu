// huf_api.cu -- the extern "C" ABI declared in include/hufb200.h.
//
// Host entry points move the caller's HOST buffers through a per-thread device
// workspace (grow-only) and run the same kernels as the *_dev entry points.
// There is no CPU implementation of any codec step in this file: without a
// device every compute call fails.
#include "../../include/hufb200.h"

#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <vector>

#include "huf_kernels.h"

namespace {

using namespace hufb200;

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(expr)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver           \
                      ? HUFB200_E_NODEVICE                                               \
                      : HUFB200_E_CUDA,                                                  \
                  "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);  \
  } while (0)

// Grow-only device buffer owned by the calling host thread.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int dev = -1;
  cudaError_t reserve(size_t n) {
    int cur = 0;
    cudaError_t e = cudaGetDevice(&cur);
    if (e != cudaSuccess) return e;
    if (p && (cur != dev || n > cap)) {
      cudaFree(p);
      p = nullptr;
      cap = 0;
    }
    if (!p) {
      size_t want = n < 256 ? 256 : n;
      want = (want + 255) & ~(size_t)255;
      e = cudaMalloc(&p, want);
      if (e != cudaSuccess) {
        p = nullptr;
        return e;
      }
      cap = want;
      dev = cur;
    }
    return cudaSuccess;
  }
  template <typename T>
  T* as() { return reinterpret_cast<T*>(p); }
  void release() {
    if (p) cudaFree(p);  // (fails harmlessly once the runtime is shut down)
    p = nullptr;
    cap = 0;
    dev = -1;
  }
  ~DevBuf() { release(); }
};

// Grow-only pinned host buffer (staging of the small single-buffer calls).
struct PinBuf {
  uint8_t* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (p && n <= cap) return cudaSuccess;
    release();
    size_t want = n < (256u << 10) ? (256u << 10) : n;
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&p), want, cudaHostAllocDefault);
    if (e != cudaSuccess) {
      p = nullptr;
      return e;
    }
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  ~PinBuf() { release(); }
};
// Single-buffer calls of up to this many bytes (either side) go through pinned staging: a copy
// from or to pageable memory is staged by the driver and blocks the calling thread per copy; with
// pinned staging a call is one memcpy in, asynchronous copies and kernels, ONE synchronisation and
// one memcpy out.  (BASELINE config 1 is 100 KiB per call.)
constexpr size_t kSmallCall = (size_t)2 << 20;

// Per-thread state of the single-buffer host calls: device buffers and one non-blocking stream
// (never the legacy default stream, which would serialise against everything else the host
// application runs on the device).
struct Workspace {
  DevBuf in, out, sizes, offsets, misc, table;
  PinBuf pin_in, pin_out;
  cudaStream_t st = nullptr;
  int dev = -1;
  cudaError_t ready() {
    int cur = 0;
    cudaError_t e = cudaGetDevice(&cur);
    if (e != cudaSuccess) return e;
    if (st && dev != cur) {
      cudaStreamDestroy(st);
      st = nullptr;
    }
    if (!st) {
      e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
      if (e != cudaSuccess) {
        st = nullptr;
        return e;
      }
      dev = cur;
    }
    return cudaSuccess;
  }
  void release() {
    in.release(); out.release(); sizes.release(); offsets.release(); misc.release(); table.release();
    pin_in.release(); pin_out.release();
    if (st) cudaStreamDestroy(st);
    st = nullptr;
    dev = -1;
  }
  ~Workspace() { release(); }
};
thread_local Workspace g_ws;

// Two-deep copy/compute pipeline for the host-pointer container calls: chunk c+1's host->device
// copy and kernels overlap chunk c's device->host copy (PCIe is full duplex; the kernels are
// ~20x faster than either copy).
struct Pipe {
  cudaStream_t st = nullptr;
  int dev = -1;
  DevBuf in, out, sizes, offsets, packed, misc;
  // pinned landing zone for the per-chunk results (a pageable destination would make the
  // device->host copy synchronous and stall the pipeline): [0] chunk bytes (u64), [1] status
  unsigned long long* meta = nullptr;
  cudaError_t ready() {
    int cur = 0;
    cudaError_t e = cudaGetDevice(&cur);
    if (e != cudaSuccess) return e;
    if (st && dev != cur) {
      cudaStreamDestroy(st);
      st = nullptr;
    }
    if (!st) {
      e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
      if (e != cudaSuccess) return e;
      dev = cur;
    }
    if (!meta) {
      e = cudaHostAlloc(reinterpret_cast<void**>(&meta), 64, cudaHostAllocDefault);
      if (e != cudaSuccess) {
        meta = nullptr;
        return e;
      }
    }
    return cudaSuccess;
  }
  void release() {
    in.release(); out.release(); sizes.release(); offsets.release(); packed.release(); misc.release();
    if (st) cudaStreamDestroy(st);
    st = nullptr;
    if (meta) cudaFreeHost(meta);
    meta = nullptr;
    dev = -1;
  }
  ~Pipe() { release(); }
};
thread_local Pipe g_pipe[2];
constexpr size_t kChunkBytes = (size_t)32 << 20;  // raw bytes per pipeline chunk

// Every exit of a pipelined call -- the error returns included -- waits for both pipe streams:
// copies queued against the caller's buffers must not outlive the call, and the next call must
// find the pipes idle.
struct PipeDrain {
  ~PipeDrain() {
    for (int i = 0; i < 2; ++i)
      if (g_pipe[i].st) cudaStreamSynchronize(g_pipe[i].st);
  }
};

struct DevInfo {
  int dev = -1;
  int sms = 0;
};
thread_local DevInfo g_dev;

int sm_count(int* sms) {
  int cur = 0;
  CU(cudaGetDevice(&cur));
  if (g_dev.dev != cur) {
    int n = 0;
    CU(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, cur));
    g_dev.dev = cur;
    g_dev.sms = n;
  }
  *sms = g_dev.sms;
  return HUFB200_OK;
}

bool valid_k(int k) { return k >= 1 && k <= HUFB200_MAX_K; }

constexpr size_t kMaxBlock = (size_t)1 << 30;  // offsets are 32-bit inside a buffer (codec/huffman.cpp:772)
constexpr uint32_t kMagic = 0x32424648u;        // "HFB2"
constexpr size_t kContainerHeader = 32;

// Blocks per decode CTA: one lane per stream, so bpc * k lanes.  About 64 lanes (2 warps for
// K = 32) measured best on B200: enough CTAs stay resident while the shared-memory total stays
// under the 196 KiB carve-out, which leaves the L1 its larger size.  Among 64..128 lanes the
// count that fills its warps best wins (K = 48: two blocks = three full warps).
#ifndef HUF_DEC_LANES
#define HUF_DEC_LANES 64
#endif
// Split decode or one lane per stream?  The split path decodes every bit three times but on as
// many lanes as the device holds; one lane per stream needs more than about 110 streams per SM (16384 in all) to be faster
// (measured on the K x block grid, profiles/r2_split_decode.md).  HUFB200_SPLIT=0|1 forces it.
constexpr size_t kMaxSplitSlice = (size_t)256 << 20;  // split decode: 12 bits per symbol of a stream must stay below 2^32
bool prefer_split(int k, size_t n_blocks, size_t raw_n) {
  static const int forced = [] {
    const char* e = getenv("HUFB200_SPLIT");
    return e ? (atoi(e) ? 1 : 0) : -1;
  }();
  if (raw_n == 0 || n_blocks == 0) return false;
  if (forced >= 0) return forced == 1;
  if (raw_n / n_blocks / (size_t)k > kMaxSplitSlice) return false;  // (bit positions inside a stream are 32-bit there)
  const size_t streams = n_blocks * (size_t)k;
  return streams <= 16384 && raw_n / streams >= 4096;
}

// (HUFB200_SPLIT=0 also switches the one-CTA form off; HUFB200_SPLIT_SMALL=0 only that)
bool split_small_allowed() {
  static const bool on = [] {
    const char* a = getenv("HUFB200_SPLIT");
    const char* b = getenv("HUFB200_SPLIT_SMALL");
    return !(a && atoi(a) == 0) && !(b && atoi(b) == 0);
  }();
  return on;
}

int decode_bpc(int k) {
  static const int forced = [] {  // tuning aid: HUFB200_DEC_LANES overrides the lanes per decode CTA
    const char* e = getenv("HUFB200_DEC_LANES");
    const int v = e ? atoi(e) : 0;
    return v >= 32 && v <= 256 ? v : 0;
  }();
  // with 32 or more streams per table, four-warp CTAs measured best (profiles/r2_decode_table_bits.md)
  const int lanes_target = forced ? forced : (k >= 32 ? 2 * HUF_DEC_LANES : HUF_DEC_LANES);
  int best = 1;
  double best_util = 0.0;
  for (int b = 1; b <= 16; ++b) {
    const int lanes = b * k;
    if (b > 1 && (lanes > 2 * lanes_target || lanes > 256)) break;
    const double util = (double)lanes / (double)((lanes + 31) / 32 * 32);
    const bool big_enough = lanes >= lanes_target;
    // until the target size is reached more lanes are always better; beyond it only a better fill counts
    if (b == 1 || util > best_util + 1e-9 || (best * k < lanes_target && util >= best_util - 0.02)) {
      best = b;
      best_util = util;
    }
    if (big_enough && util > 0.999) break;
  }
  return best;
}

int compress_grid(uint32_t n_blocks, int sms) {
  const uint32_t cap = (uint32_t)sms * (uint32_t)kCompCtasPerSm;
  return (int)(n_blocks < cap ? n_blocks : cap);
}

int do_compress_dev(int k, size_t block_size, const uint8_t* d_raw, size_t n, uint32_t n_blocks,
                    uint8_t* d_out, size_t slot_stride, uint32_t* d_sizes, const void* d_table,
                    int check_presence, uint32_t* d_status, cudaStream_t st) {
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  CU(launch_compress(d_raw, n, (uint32_t)block_size, k, n_blocks, d_out, slot_stride, d_sizes, d_table,
                     check_presence, d_status, compress_grid(n_blocks, sms), st));
  if (n_blocks) g_launches.fetch_add(1, std::memory_order_relaxed);
  return HUFB200_OK;
}

// compress one host buffer as `n_blocks` blocks into the workspace; sizes come back in `sizes`
// staged (one block only): the input goes through ws.pin_in, and size, status and the block's slot
// up to its bound come back through ws.pin_out ([0] size, [4] status, [64..] bytes) under the one
// synchronisation -- the caller copies what it needs out of ws.pin_out
int compress_host(int k, size_t block_size, const uint8_t* raw, size_t n, uint32_t n_blocks,
                  const void* d_table, int check_presence, std::vector<uint32_t>* sizes,
                  size_t* slot_stride_out, bool staged = false) {
  Workspace& ws = g_ws;
  CU(ws.ready());
  const size_t stride = hufb200_slot_stride(block_size, k);
  CU(ws.in.reserve(n + 16));
  CU(ws.out.reserve(stride * n_blocks));
  CU(ws.sizes.reserve(sizeof(uint32_t) * (n_blocks + 1)));
  CU(ws.misc.reserve(256));
  const size_t bound = hufb200_compress_bound(n, k);
  if (staged) {
    CU(ws.pin_in.reserve(n + 64));
    CU(ws.pin_out.reserve(bound + 128));
    if (n) memcpy(ws.pin_in.p, raw, n);
    raw = ws.pin_in.p;
  }
  if (n) CU(cudaMemcpyAsync(ws.in.p, raw, n, cudaMemcpyHostToDevice, ws.st));
  CU(cudaMemsetAsync(ws.misc.p, 0, 4, ws.st));
  int rc = do_compress_dev(k, block_size, ws.in.as<uint8_t>(), n, n_blocks, ws.out.as<uint8_t>(), stride,
                           ws.sizes.as<uint32_t>(), d_table, check_presence, ws.misc.as<uint32_t>(), ws.st);
  if (rc) return rc;
  sizes->resize(n_blocks);
  uint32_t status = 0;
  if (staged) {
    CU(cudaMemcpyAsync(ws.pin_out.p, ws.sizes.p, 4, cudaMemcpyDeviceToHost, ws.st));
    CU(cudaMemcpyAsync(ws.pin_out.p + 4, ws.misc.p, 4, cudaMemcpyDeviceToHost, ws.st));
    CU(cudaMemcpyAsync(ws.pin_out.p + 64, ws.out.p, bound, cudaMemcpyDeviceToHost, ws.st));
    CU(cudaStreamSynchronize(ws.st));
    memcpy(sizes->data(), ws.pin_out.p, 4);
    memcpy(&status, ws.pin_out.p + 4, 4);
  } else {
    CU(cudaMemcpyAsync(sizes->data(), ws.sizes.p, sizeof(uint32_t) * n_blocks, cudaMemcpyDeviceToHost, ws.st));
    CU(cudaMemcpyAsync(&status, ws.misc.p, 4, cudaMemcpyDeviceToHost, ws.st));
    CU(cudaStreamSynchronize(ws.st));
  }
  if (status) return fail(HUFB200_E_CORRUPT, "a symbol of the input has no code in the supplied table");
  *slot_stride_out = stride;
  return HUFB200_OK;
}

// A single buffer of at least this many bytes is spread over the whole device (per-stream
// histograms, one plan, pieces of 3072 symbols: launch_compress_single); smaller ones go to one
// CTA, which needs a single launch.
#ifndef HUF_SINGLE_MULTI_MIN
#define HUF_SINGLE_MULTI_MIN (256u << 10)
#endif
// (with 8 streams or fewer one CTA has idle warps: there the spread form wins from 64 KiB on,
// measured on the 100 KiB buffer of BASELINE config 1)
inline bool single_goes_wide(size_t n, int k) { return n >= HUF_SINGLE_MULTI_MIN || (k <= 8 && n >= (64u << 10)); }

// compresses one host buffer into ws.out; *size = compressed size
int compress_single_host(int k, const uint8_t* raw, size_t n, const void* d_table, uint32_t* size, bool staged = false) {
  Workspace& ws = g_ws;
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  const size_t bound = hufb200_compress_bound(n, k) + 16;
  const uint32_t pieces = single_piece_count((uint32_t)n, k);
  CU(ws.in.reserve(n + 16));
  CU(ws.out.reserve(bound));
  const size_t hist_b = (size_t)k * 256 * sizeof(uint32_t), plan_b = (single_plan_bytes() + 255) & ~(size_t)255,
               tab_b = (table_bytes() + 255) & ~(size_t)255;
  CU(ws.sizes.reserve(hist_b + plan_b + tab_b + (size_t)pieces * 4 + 256));
  uint8_t* m = ws.sizes.as<uint8_t>();
  uint32_t* d_hist = reinterpret_cast<uint32_t*>(m);
  void* d_plan = m + hist_b;
  void* d_tab = m + hist_b + plan_b;
  uint32_t* d_pieces = reinterpret_cast<uint32_t*>(m + hist_b + plan_b + tab_b);
  if (staged) {  // see compress_host
    CU(ws.pin_in.reserve(n + 64));
    CU(ws.pin_out.reserve(bound + 128));
    memcpy(ws.pin_in.p, raw, n);
    raw = ws.pin_in.p;
  }
  CU(cudaMemcpyAsync(ws.in.p, raw, n, cudaMemcpyHostToDevice, ws.st));
  CU(cudaMemsetAsync(ws.out.p, 0, bound, ws.st));  // pieces OR their bits into it
  CU(launch_compress_single(ws.in.as<uint8_t>(), (uint32_t)n, k, d_table, d_hist, d_tab, d_plan, d_pieces,
                            ws.out.as<uint8_t>(), sms, ws.st));
  g_launches.fetch_add(4, std::memory_order_relaxed);
  uint32_t head[4] = {0, 0, 0, 0};  // total_size, hdr_total, bad
  if (staged) {
    CU(cudaMemcpyAsync(ws.pin_out.p, d_plan, sizeof(head), cudaMemcpyDeviceToHost, ws.st));
    CU(cudaMemcpyAsync(ws.pin_out.p + 64, ws.out.p, bound - 16, cudaMemcpyDeviceToHost, ws.st));
    CU(cudaStreamSynchronize(ws.st));
    memcpy(head, ws.pin_out.p, sizeof(head));
  } else {
    CU(cudaMemcpyAsync(head, d_plan, sizeof(head), cudaMemcpyDeviceToHost, ws.st));
    CU(cudaStreamSynchronize(ws.st));
  }
  if (head[2]) return fail(HUFB200_E_CORRUPT, "a symbol of the input has no code in the supplied table");
  *size = head[0];
  return HUFB200_OK;
}

}  // namespace

extern "C" {

int hufb200_version(void) { return 100; }
const char* hufb200_last_error(void) { return g_err; }
uint64_t hufb200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void hufb200_release_workspace(void) {
  for (int i = 0; i < 2; ++i) {
    if (g_pipe[i].st) cudaStreamSynchronize(g_pipe[i].st);
    g_pipe[i].release();
  }
  if (g_ws.st) cudaStreamSynchronize(g_ws.st);
  g_ws.release();
}

int hufb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

/* ------------------------------------------------------------------ histogram */

int hufb200_histogram_dev(const uint8_t* d_in, size_t n, uint64_t* d_out, void* stream) {
  if (!d_out || (!d_in && n)) return fail(HUFB200_E_INVALID, "null pointer");
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  // 32 KB of bins + 512 threads per CTA: 4 CTAs per SM saturate the shared-memory atomics
  uint64_t want = (n + 65535) / 65536;
  int grid = (int)(want < (uint64_t)sms * 4 ? (want ? want : 1) : (uint64_t)sms * 4);
  CU(launch_histogram(d_in, n, reinterpret_cast<unsigned long long*>(d_out), grid, (cudaStream_t)stream));
  if (n) g_launches.fetch_add(1, std::memory_order_relaxed);
  return HUFB200_OK;
}

int hufb200_histogram64(const uint8_t* in, size_t n, uint64_t out[256]) {
  if (!out || (!in && n)) return fail(HUFB200_E_INVALID, "null pointer");
  Workspace& ws = g_ws;
  CU(ws.ready());
  CU(ws.in.reserve(n + 16));
  CU(ws.misc.reserve(256 * sizeof(uint64_t)));
  if (n) CU(cudaMemcpyAsync(ws.in.p, in, n, cudaMemcpyHostToDevice, ws.st));
  int rc = hufb200_histogram_dev(ws.in.as<uint8_t>(), n, ws.misc.as<uint64_t>(), ws.st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(out, ws.misc.p, 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ws.st));
  CU(cudaStreamSynchronize(ws.st));
  return HUFB200_OK;
}

int hufb200_histogram(const uint8_t* in, size_t n, uint32_t out[256]) {
  if (n >> 32) return fail(HUFB200_E_INVALID, "n >= 2^32 does not fit ByteHistogram; use hufb200_histogram64");
  uint64_t h[256];
  int rc = hufb200_histogram64(in, n, h);
  if (rc) return rc;
  for (int i = 0; i < 256; ++i) out[i] = (uint32_t)h[i];
  return HUFB200_OK;
}

/* ---------------------------------------------------------------- table build */

size_t hufb200_table_bytes(void) { return table_bytes(); }

int hufb200_build_table_dev(const uint64_t* d_hist, void* d_table, void* stream) {
  if (!d_hist || !d_table) return fail(HUFB200_E_INVALID, "null pointer");
  CU(launch_build_table(reinterpret_cast<const unsigned long long*>(d_hist), d_table, (cudaStream_t)stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return HUFB200_OK;
}

namespace {
// mirrors hufb200::HufTable (huf_device.cuh) for the host-side dump
struct HostTable {
  uint32_t enc[256];
  uint32_t enc2[256];
  uint8_t sorted_syms[256];
  uint16_t len_count[16];
  uint32_t len_mask;
  int32_t num_syms;
  uint32_t hdr_len;
  uint32_t avg_bits_x256;
};
}  // namespace

int hufb200_make_table(const uint32_t hist[256], uint16_t len_count[13], uint8_t sorted_syms[256],
                       int* num_syms, uint32_t* len_mask, uint16_t code_bits[256], uint16_t code_len[256]) {
  if (!hist) return fail(HUFB200_E_INVALID, "null pointer");
  if (sizeof(HostTable) != table_bytes()) return fail(HUFB200_E_INVALID, "table layout mismatch");
  Workspace& ws = g_ws;
  CU(ws.ready());
  CU(ws.misc.reserve(256 * sizeof(uint32_t)));
  CU(ws.table.reserve(table_bytes()));
  CU(cudaMemcpyAsync(ws.misc.p, hist, 256 * sizeof(uint32_t), cudaMemcpyHostToDevice, ws.st));
  CU(launch_make_table(ws.misc.as<uint32_t>(), nullptr, nullptr, 0, 0, ws.table.p, ws.st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  HostTable t;
  CU(cudaMemcpyAsync(&t, ws.table.p, sizeof(t), cudaMemcpyDeviceToHost, ws.st));
  CU(cudaStreamSynchronize(ws.st));
  if (len_count)
    for (int i = 0; i <= HUFB200_MAX_CODE_LEN; ++i) len_count[i] = t.len_count[i];
  if (sorted_syms) {
    memset(sorted_syms, 0, 256);
    memcpy(sorted_syms, t.sorted_syms, (size_t)t.num_syms);
  }
  if (num_syms) *num_syms = t.num_syms;
  if (len_mask) *len_mask = t.len_mask;
  for (int c = 0; c < 256; ++c) {
    const uint32_t e = t.enc[c];
    const bool present = e != 0x10000000u;
    const uint32_t l = present ? (e >> 16) : 0;
    // BitCode.bits is left-aligned in 12 bits (codec/huffman.cpp:214-224)
    if (code_bits) code_bits[c] = present ? (uint16_t)((e & 0xffffu) << (HUFB200_MAX_CODE_LEN - l)) : 0;
    if (code_len) code_len[c] = (uint16_t)l;
  }
  return HUFB200_OK;
}

namespace {
int decode_table_dump(const uint16_t len_count[13], const uint8_t* sorted_syms, int num_syms, int one_symbol,
                      uint8_t* out);
}
int hufb200_decode_table(const uint16_t len_count[13], const uint8_t* sorted_syms, int num_syms,
                         uint8_t out[4096 * 4]) {
  return decode_table_dump(len_count, sorted_syms, num_syms, 0, out);
}
int hufb200_decode_table1x(const uint16_t len_count[13], const uint8_t* sorted_syms, int num_syms,
                           uint8_t out[4096 * 2]) {
  return decode_table_dump(len_count, sorted_syms, num_syms, 1, out);
}
namespace {
int decode_table_dump(const uint16_t len_count[13], const uint8_t* sorted_syms, int num_syms, int one_symbol,
                      uint8_t* out) {
  if (!len_count || !out || num_syms < 0 || num_syms > 256 || (!sorted_syms && num_syms))
    return fail(HUFB200_E_INVALID, "bad table arguments");
  Workspace& ws = g_ws;
  CU(ws.ready());
  CU(ws.misc.reserve(1024));
  CU(ws.out.reserve(4096 * 4));
  uint8_t* d_lc = ws.misc.as<uint8_t>();
  uint8_t* d_sy = d_lc + 64;
  CU(cudaMemcpyAsync(d_lc, len_count, 13 * sizeof(uint16_t), cudaMemcpyHostToDevice, ws.st));
  if (num_syms) CU(cudaMemcpyAsync(d_sy, sorted_syms, (size_t)num_syms, cudaMemcpyHostToDevice, ws.st));
  CU(launch_dump_dtable(reinterpret_cast<const uint16_t*>(d_lc), d_sy, num_syms, one_symbol, ws.out.as<uint8_t>(), ws.st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CU(cudaMemcpyAsync(out, ws.out.p, one_symbol ? 4096 * 2 : 4096 * 4, cudaMemcpyDeviceToHost, ws.st));
  CU(cudaStreamSynchronize(ws.st));
  return HUFB200_OK;
}
}  // namespace

/* -------------------------------------------------------------- single buffer */

size_t hufb200_compress_bound(size_t n, int k) {
  if (k < 1) k = 1;
  return 8 + 13 + 256 + 4 * (size_t)(k - 1) + (n * HUFB200_MAX_CODE_LEN + 7) / 8 + (size_t)k * 9;
}

size_t hufb200_slot_stride(size_t block_size, int k) {
  return (hufb200_compress_bound(block_size, k) + 4 + 255) & ~(size_t)255;
}

int hufb200_compress(int k, const uint8_t* raw, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
  if (!valid_k(k)) return fail(HUFB200_E_INVALID, "k=%d outside 1..%d", k, HUFB200_MAX_K);
  if (n > kMaxBlock) return fail(HUFB200_E_INVALID, "n=%zu exceeds the 2^30 single-buffer limit", n);
  if (!out_len || (!raw && n) || (!out && cap)) return fail(HUFB200_E_INVALID, "null pointer");
  std::vector<uint32_t> sizes(1);
  size_t stride = 0;
  int rc;
  const bool staged = n <= kSmallCall;  // small calls: pinned staging, one synchronisation
  if (single_goes_wide(n, k)) {
    CU(g_ws.ready());
    rc = compress_single_host(k, raw, n, nullptr, &sizes[0], staged);
  } else {
    rc = compress_host(k, n ? n : 1, raw, n, 1, nullptr, 0, &sizes, &stride, staged);
  }
  if (rc) return rc;
  *out_len = sizes[0];
  if (sizes[0] > cap) return fail(HUFB200_E_NOSPACE, "need %u bytes, have %zu", sizes[0], cap);
  if (staged) {
    if (sizes[0] > hufb200_compress_bound(n, k)) return fail(HUFB200_E_CUDA, "compressed size beyond its bound");
    memcpy(out, g_ws.pin_out.p + 64, sizes[0]);
    return HUFB200_OK;
  }
  CU(cudaMemcpyAsync(out, g_ws.out.p, sizes[0], cudaMemcpyDeviceToHost, g_ws.st));
  CU(cudaStreamSynchronize(g_ws.st));
  return HUFB200_OK;
}

int hufb200_compress_with_table(int k, const uint8_t* raw, size_t n, const uint16_t len_count[13],
                                const uint8_t* sorted_syms, int num_syms, uint8_t* out, size_t cap,
                                size_t* out_len) {
  if (!valid_k(k)) return fail(HUFB200_E_INVALID, "k=%d outside 1..%d", k, HUFB200_MAX_K);
  if (n > kMaxBlock) return fail(HUFB200_E_INVALID, "n=%zu exceeds the 2^30 single-buffer limit", n);
  if (!out_len || (!raw && n) || !len_count || num_syms < 0 || num_syms > 256 || (!sorted_syms && num_syms))
    return fail(HUFB200_E_INVALID, "bad arguments");
  Workspace& ws = g_ws;
  CU(ws.ready());
  CU(ws.misc.reserve(1024));
  CU(ws.table.reserve(table_bytes()));
  uint8_t* d_lc = ws.misc.as<uint8_t>() + 256;
  uint8_t* d_sy = d_lc + 64;
  CU(cudaMemcpyAsync(d_lc, len_count, 13 * sizeof(uint16_t), cudaMemcpyHostToDevice, ws.st));
  if (num_syms) CU(cudaMemcpyAsync(d_sy, sorted_syms, (size_t)num_syms, cudaMemcpyHostToDevice, ws.st));
  CU(launch_make_table(nullptr, reinterpret_cast<const uint16_t*>(d_lc), d_sy, num_syms, 1, ws.table.p, ws.st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  std::vector<uint32_t> sizes(1);
  size_t stride = 0;
  int rc;
  if (single_goes_wide(n, k)) rc = compress_single_host(k, raw, n, ws.table.p, &sizes[0]);
  else rc = compress_host(k, n ? n : 1, raw, n, 1, ws.table.p, 1, &sizes, &stride);
  if (rc) return rc;
  *out_len = sizes[0];
  if (sizes[0] > cap) return fail(HUFB200_E_NOSPACE, "need %u bytes, have %zu", sizes[0], cap);
  CU(cudaMemcpyAsync(out, g_ws.out.p, sizes[0], cudaMemcpyDeviceToHost, g_ws.st));
  CU(cudaStreamSynchronize(g_ws.st));
  return HUFB200_OK;
}

int hufb200_raw_size(const uint8_t* comp, size_t n, size_t* raw_size) {
  if (!comp || n < 8 || !raw_size) return fail(HUFB200_E_CORRUPT, "compressed buffer shorter than its header");
  uint32_t v;
  memcpy(&v, comp, 4);
  *raw_size = v;
  return HUFB200_OK;
}

int hufb200_decompress(int k, const uint8_t* comp, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
  if (!valid_k(k)) return fail(HUFB200_E_INVALID, "k=%d outside 1..%d", k, HUFB200_MAX_K);
  if (!out_len) return fail(HUFB200_E_INVALID, "null pointer");
  size_t raw_size = 0;
  int rc = hufb200_raw_size(comp, n, &raw_size);
  if (rc) return rc;
  if (n >> 32) return fail(HUFB200_E_INVALID, "compressed size does not fit 32 bits");
  *out_len = raw_size;
  if (raw_size > cap) return fail(HUFB200_E_NOSPACE, "need %zu bytes, have %zu", raw_size, cap);
  if (raw_size > kMaxBlock) return fail(HUFB200_E_INVALID, "raw_size exceeds the 2^30 single-buffer limit");
  Workspace& ws = g_ws;
  CU(ws.ready());
  CU(ws.in.reserve(n + 32));
  CU(ws.out.reserve(raw_size + 16));
  CU(ws.misc.reserve(256));
  struct Meta {
    unsigned long long off;
    uint32_t size;
    uint32_t status;
  } meta = {0ull, (uint32_t)n, 0u};
  const bool small = n <= kSmallCall && raw_size <= kSmallCall;
  if (small) {  // pinned staging: [meta | compressed bytes] in, [status | raw bytes] out
    CU(ws.pin_in.reserve(n + 64));
    CU(ws.pin_out.reserve(raw_size + 64));
    memcpy(ws.pin_in.p, &meta, sizeof(meta));
    memcpy(ws.pin_in.p + 32, comp, n);
    CU(cudaMemcpyAsync(ws.in.p, ws.pin_in.p + 32, n, cudaMemcpyHostToDevice, ws.st));
    CU(cudaMemcpyAsync(ws.misc.p, ws.pin_in.p, sizeof(meta), cudaMemcpyHostToDevice, ws.st));
  } else {
    CU(cudaMemcpyAsync(ws.in.p, comp, n, cudaMemcpyHostToDevice, ws.st));
    CU(cudaMemcpyAsync(ws.misc.p, &meta, sizeof(meta), cudaMemcpyHostToDevice, ws.st));
  }
  uint8_t* m = ws.misc.as<uint8_t>();
  if (small && raw_size / (size_t)k >= 1024 && split_small_fits(n, k) && split_small_allowed()) {
    // a small buffer: the whole split decode in one CTA, one launch
    CU(launch_decompress_split_small(ws.in.as<uint8_t>(), reinterpret_cast<uint32_t*>(m + 8), n, k,
                                     ws.out.as<uint8_t>(), (uint32_t)raw_size, reinterpret_cast<uint32_t*>(m + 12),
                                     ws.st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  } else if (prefer_split(k, 1, raw_size)) {  // K streams cannot fill the device: cut them into items
    int sms = 0;
    rc = sm_count(&sms);
    if (rc) return rc;
    const uint32_t sub = split_sub_bits(raw_size, 1, k, sms);
    CU(ws.offsets.reserve(decompress_split_work_bytes(1, k, (uint32_t)raw_size, sub)));
    int launches = 0;
    CU(launch_decompress_split(ws.in.as<uint8_t>(), reinterpret_cast<unsigned long long*>(m),
                               reinterpret_cast<uint32_t*>(m + 8), 1, k, ws.out.as<uint8_t>(), raw_size,
                               (uint32_t)raw_size, sub, ws.offsets.p, reinterpret_cast<uint32_t*>(m + 12), &launches,
                               ws.st));
    g_launches.fetch_add((uint64_t)launches, std::memory_order_relaxed);
  } else {
    CU(launch_decompress(ws.in.as<uint8_t>(), reinterpret_cast<unsigned long long*>(m),
                         reinterpret_cast<uint32_t*>(m + 8), 1, k, 1, ws.out.as<uint8_t>(), raw_size,
                         (uint32_t)(raw_size ? raw_size : 1), reinterpret_cast<uint32_t*>(m + 12), ws.st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  uint32_t status = 0;
  if (small) {
    CU(cudaMemcpyAsync(ws.pin_out.p, m + 12, 4, cudaMemcpyDeviceToHost, ws.st));
    if (raw_size) CU(cudaMemcpyAsync(ws.pin_out.p + 32, ws.out.p, raw_size, cudaMemcpyDeviceToHost, ws.st));
    CU(cudaStreamSynchronize(ws.st));
    memcpy(&status, ws.pin_out.p, 4);
    if (status == kSplitSmallRetry) {
      // the one-launch form found a CTA with more than 1.5 times its share of the bits (streams of
      // very unequal length): the spread form has no such limit
      int sms = 0;
      rc = sm_count(&sms);
      if (rc) return rc;
      const uint32_t sub = split_sub_bits(raw_size, 1, k, sms);
      CU(ws.offsets.reserve(decompress_split_work_bytes(1, k, (uint32_t)raw_size, sub)));
      CU(cudaMemsetAsync(m + 12, 0, 4, ws.st));
      int launches = 0;
      CU(launch_decompress_split(ws.in.as<uint8_t>(), reinterpret_cast<unsigned long long*>(m),
                                 reinterpret_cast<uint32_t*>(m + 8), 1, k, ws.out.as<uint8_t>(), raw_size,
                                 (uint32_t)raw_size, sub, ws.offsets.p, reinterpret_cast<uint32_t*>(m + 12), &launches,
                                 ws.st));
      g_launches.fetch_add((uint64_t)launches, std::memory_order_relaxed);
      CU(cudaMemcpyAsync(ws.pin_out.p, m + 12, 4, cudaMemcpyDeviceToHost, ws.st));
      CU(cudaMemcpyAsync(ws.pin_out.p + 32, ws.out.p, raw_size, cudaMemcpyDeviceToHost, ws.st));
      CU(cudaStreamSynchronize(ws.st));
      memcpy(&status, ws.pin_out.p, 4);
    }
    if (!status && raw_size) memcpy(out, ws.pin_out.p + 32, raw_size);
  } else {
    CU(cudaMemcpyAsync(&status, m + 12, 4, cudaMemcpyDeviceToHost, ws.st));
    if (raw_size) CU(cudaMemcpyAsync(out, ws.out.p, raw_size, cudaMemcpyDeviceToHost, ws.st));
    CU(cudaStreamSynchronize(ws.st));
  }
  if (status) return fail(HUFB200_E_CORRUPT, "malformed compressed buffer");
  return HUFB200_OK;
}

/* ------------------------------------------------------------ block container */

size_t hufb200_blocks_count(size_t n, size_t block_size) {
  return block_size ? (n + block_size - 1) / block_size : 0;
}

size_t hufb200_container_bound(size_t n, size_t block_size, int k) {
  const size_t nb = hufb200_blocks_count(n, block_size);
  return kContainerHeader + 4 * nb + nb * hufb200_compress_bound(block_size, k);
}

int hufb200_compress_blocks(int k, size_t block_size, const uint8_t* raw, size_t n, uint8_t* out, size_t cap,
                            size_t* out_len) {
  if (!valid_k(k)) return fail(HUFB200_E_INVALID, "k=%d outside 1..%d", k, HUFB200_MAX_K);
  if (block_size == 0 || block_size > kMaxBlock) return fail(HUFB200_E_INVALID, "block_size outside 1..2^30");
  if (!out_len || (!raw && n) || (!out && cap)) return fail(HUFB200_E_INVALID, "null pointer");
  const size_t nb = hufb200_blocks_count(n, block_size);
  if (nb >> 31) return fail(HUFB200_E_INVALID, "too many blocks");
  const size_t index_end = kContainerHeader + 4 * nb;
  if (index_end <= cap) {
    // header: magic, version|k, block_size, n_blocks, raw_size(u64), reserved
    memset(out, 0, kContainerHeader);
    const uint32_t magic = kMagic, vk = 1u | ((uint32_t)k << 16), bs = (uint32_t)block_size, nb32 = (uint32_t)nb;
    const uint64_t rs = n;
    memcpy(out + 0, &magic, 4);
    memcpy(out + 4, &vk, 4);
    memcpy(out + 8, &bs, 4);
    memcpy(out + 12, &nb32, 4);
    memcpy(out + 16, &rs, 8);
  }
  const size_t stride = hufb200_slot_stride(block_size, k);
  size_t per = kChunkBytes / block_size;  // blocks per chunk
  if (per < 1) per = 1;
  const size_t n_chunks = (nb + per - 1) / per;
  PipeDrain drain;
  size_t total = index_end;  // running container size
  bool fits = index_end <= cap;
  struct Pending {
    bool live = false;
    size_t blk0 = 0, nblk = 0;
  } pend[2];
  // second half of a chunk: its sizes and total are on the host -> payload device->host
  auto finish = [&](int slot) -> int {
    Pending& pd = pend[slot];
    if (!pd.live) return HUFB200_OK;
    Pipe& pp = g_pipe[slot];
    CU(cudaStreamSynchronize(pp.st));
    pd.live = false;
    const unsigned long long chunk_total = pp.meta[0];
    if (pp.meta[1]) return fail(HUFB200_E_CORRUPT, "kernel reported a malformed block");
    if (fits && total + chunk_total <= cap) {
      CU(cudaMemcpyAsync(out + total, pp.packed.p, chunk_total, cudaMemcpyDeviceToHost, pp.st));
    } else {
      fits = false;
    }
    total += chunk_total;
    return HUFB200_OK;
  };
  for (size_t c = 0; c < n_chunks; ++c) {
    const int slot = (int)(c & 1);
    Pipe& pp = g_pipe[slot];
    int rc = finish(slot);  // the chunk that used this slot two iterations ago
    if (rc) return rc;
    CU(pp.ready());
    CU(cudaStreamSynchronize(pp.st));  // its payload copy must be done before the buffers are reused
    const size_t blk0 = c * per, nblk = (nb - blk0 < per) ? nb - blk0 : per;
    const size_t off = blk0 * block_size, len = (n - off < nblk * block_size) ? n - off : nblk * block_size;
    CU(pp.in.reserve(len + 16));
    CU(pp.out.reserve(stride * nblk));
    CU(pp.sizes.reserve(4 * (nblk + 1)));
    CU(pp.offsets.reserve(8 * (nblk + 2)));
    CU(pp.packed.reserve(hufb200_compress_bound(block_size, k) * nblk + 32));
    CU(pp.misc.reserve(256));
    CU(cudaMemcpyAsync(pp.in.p, raw + off, len, cudaMemcpyHostToDevice, pp.st));
    CU(cudaMemsetAsync(pp.misc.p, 0, 4, pp.st));
    rc = do_compress_dev(k, block_size, pp.in.as<uint8_t>(), len, (uint32_t)nblk, pp.out.as<uint8_t>(), stride,
                         pp.sizes.as<uint32_t>(), nullptr, 0, pp.misc.as<uint32_t>(), pp.st);
    if (rc) return rc;
    unsigned long long* d_off = pp.offsets.as<unsigned long long>();
    CU(launch_pack(pp.out.as<uint8_t>(), stride, pp.sizes.as<uint32_t>(), (uint32_t)nblk, pp.packed.as<uint8_t>(),
                   d_off, d_off + nblk, pp.st));
    g_launches.fetch_add(2, std::memory_order_relaxed);
    Pending& pd = pend[slot];
    pd.live = true;
    pd.blk0 = blk0;
    pd.nblk = nblk;
    if (index_end <= cap) {
      CU(cudaMemcpyAsync(out + kContainerHeader + 4 * blk0, pp.sizes.p, 4 * nblk, cudaMemcpyDeviceToHost, pp.st));
    }
    pp.meta[1] = 0;
    CU(cudaMemcpyAsync(&pp.meta[0], d_off + nblk, 8, cudaMemcpyDeviceToHost, pp.st));
    CU(cudaMemcpyAsync(&pp.meta[1], pp.misc.p, 4, cudaMemcpyDeviceToHost, pp.st));
  }
  for (size_t c = n_chunks; c < n_chunks + 2; ++c) {  // drain in issue order
    int rc = finish((int)(c & 1));
    if (rc) return rc;
  }
  for (int slot = 0; slot < 2; ++slot)
    if (g_pipe[slot].st) CU(cudaStreamSynchronize(g_pipe[slot].st));
  *out_len = total;
  if (!fits) return fail(HUFB200_E_NOSPACE, "need %zu bytes, have %zu", total, cap);
  return HUFB200_OK;
}

int hufb200_container_info(const uint8_t* c, size_t n, int* k, size_t* block_size, size_t* raw_size,
                           size_t* n_blocks) {
  if (!c || n < kContainerHeader) return fail(HUFB200_E_CORRUPT, "container shorter than its header");
  uint32_t magic, vk, bs, nb;
  uint64_t rs;
  memcpy(&magic, c, 4);
  memcpy(&vk, c + 4, 4);
  memcpy(&bs, c + 8, 4);
  memcpy(&nb, c + 12, 4);
  memcpy(&rs, c + 16, 8);
  if (magic != kMagic || (vk & 0xffff) != 1) return fail(HUFB200_E_CORRUPT, "bad container magic/version");
  const int kk = (int)(vk >> 16);
  if (!valid_k(kk) || bs == 0 || bs > kMaxBlock) return fail(HUFB200_E_CORRUPT, "bad container parameters");
  if (hufb200_blocks_count(rs, bs) != nb) return fail(HUFB200_E_CORRUPT, "block count does not match raw size");
  if (n < kContainerHeader + 4 * (size_t)nb) return fail(HUFB200_E_CORRUPT, "truncated block index");
  if (k) *k = kk;
  if (block_size) *block_size = bs;
  if (raw_size) *raw_size = rs;
  if (n_blocks) *n_blocks = nb;
  return HUFB200_OK;
}

int hufb200_decompress_blocks(const uint8_t* c, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
  int k = 0;
  size_t bs = 0, rs = 0, nb = 0;
  int rc = hufb200_container_info(c, n, &k, &bs, &rs, &nb);
  if (rc) return rc;
  if (!out_len) return fail(HUFB200_E_INVALID, "null pointer");
  *out_len = rs;
  if (rs > cap) return fail(HUFB200_E_NOSPACE, "need %zu bytes, have %zu", rs, cap);
  if (nb == 0) return HUFB200_OK;
  const uint8_t* index = c + kContainerHeader;
  const uint8_t* payload = index + 4 * nb;
  const size_t payload_n = n - kContainerHeader - 4 * nb;
  uint64_t sum = 0;
  for (size_t b = 0; b < nb; ++b) {
    uint32_t s;
    memcpy(&s, index + 4 * b, 4);
    sum += s;
  }
  if (sum != payload_n) return fail(HUFB200_E_CORRUPT, "block sizes do not add up to the payload size");
  // chunked two-stream pipeline: payload host->device, decode, raw device->host
  PipeDrain drain;
  size_t per = kChunkBytes / bs;
  if (per < 1) per = 1;
  const size_t n_chunks = (nb + per - 1) / per;
  bool used[2] = {false, false};
  size_t pay_off = 0;
  for (size_t ci = 0; ci < n_chunks; ++ci) {
    const int slot = (int)(ci & 1);
    Pipe& pp = g_pipe[slot];
    CU(pp.ready());
    CU(cudaStreamSynchronize(pp.st));  // buffers of chunk ci-2 are free again
    if (used[slot] && pp.meta[1]) return fail(HUFB200_E_CORRUPT, "malformed block in container");
    const size_t blk0 = ci * per, nblk = (nb - blk0 < per) ? nb - blk0 : per;
    const size_t roff = blk0 * bs, rlen = (rs - roff < nblk * bs) ? rs - roff : nblk * bs;
    size_t clen = 0;
    for (size_t j = 0; j < nblk; ++j) {
      uint32_t sz;
      memcpy(&sz, index + 4 * (blk0 + j), 4);
      clen += sz;
    }
    CU(pp.in.reserve(clen + 32));
    CU(pp.out.reserve(rlen + 16));
    CU(pp.sizes.reserve(4 * (nblk + 1)));
    CU(pp.offsets.reserve(8 * (nblk + 1)));
    CU(pp.misc.reserve(256));
    CU(cudaMemcpyAsync(pp.in.p, payload + pay_off, clen, cudaMemcpyHostToDevice, pp.st));
    CU(cudaMemcpyAsync(pp.sizes.p, index + 4 * blk0, 4 * nblk, cudaMemcpyHostToDevice, pp.st));
    CU(launch_scan_sizes(pp.sizes.as<uint32_t>(), (uint32_t)nblk, pp.offsets.as<unsigned long long>(), nullptr, pp.st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CU(cudaMemsetAsync(pp.misc.p, 0, 4, pp.st));
    if (prefer_split(k, nblk, rlen)) {
      const size_t wb = hufb200_decompress_split_work_bytes(k, bs, nblk, rlen);
      CU(pp.packed.reserve(wb));
      rc = hufb200_decompress_split_dev(k, bs, pp.in.as<uint8_t>(), pp.offsets.as<uint64_t>(), pp.sizes.as<uint32_t>(),
                                        nblk, pp.out.as<uint8_t>(), rlen, pp.packed.p, wb, pp.misc.as<uint32_t>(), pp.st);
    } else {
      rc = hufb200_decompress_blocks_dev(k, bs, pp.in.as<uint8_t>(), pp.offsets.as<uint64_t>(), pp.sizes.as<uint32_t>(),
                                         nblk, pp.out.as<uint8_t>(), rlen, pp.misc.as<uint32_t>(), pp.st);
    }
    if (rc) return rc;
    pp.meta[1] = 0;
    CU(cudaMemcpyAsync(&pp.meta[1], pp.misc.p, 4, cudaMemcpyDeviceToHost, pp.st));
    CU(cudaMemcpyAsync(out + roff, pp.out.p, rlen, cudaMemcpyDeviceToHost, pp.st));
    used[slot] = true;
    pay_off += clen;
  }
  for (int slot = 0; slot < 2; ++slot)
    if (used[slot]) CU(cudaStreamSynchronize(g_pipe[slot].st));
  for (int slot = 0; slot < 2; ++slot)
    if (used[slot] && g_pipe[slot].meta[1]) return fail(HUFB200_E_CORRUPT, "malformed block in container");
  return HUFB200_OK;
}

/* ------------------------------------------------------------- device-resident */

int hufb200_compress_blocks_dev(int k, size_t block_size, const uint8_t* d_raw, size_t n, uint8_t* d_out,
                                size_t slot_stride, uint32_t* d_comp_sizes, const void* d_table,
                                uint32_t* d_status, void* stream) {
  if (!valid_k(k)) return fail(HUFB200_E_INVALID, "k=%d outside 1..%d", k, HUFB200_MAX_K);
  if (block_size == 0 || block_size > kMaxBlock) return fail(HUFB200_E_INVALID, "block_size outside 1..2^30");
  if (((uintptr_t)d_raw & 15) || ((uintptr_t)d_out & 15) || slot_stride % 16)
    return fail(HUFB200_E_INVALID, "slot_stride, d_raw and d_out must be 16-byte aligned");
  if (slot_stride < hufb200_slot_stride(block_size, k)) return fail(HUFB200_E_INVALID, "slot_stride too small");
  const size_t nb = hufb200_blocks_count(n, block_size);
  if (nb >> 31) return fail(HUFB200_E_INVALID, "too many blocks");
  if (nb && (!d_raw || !d_out || !d_comp_sizes)) return fail(HUFB200_E_INVALID, "null pointer");
  return do_compress_dev(k, block_size, d_raw, n, (uint32_t)nb, d_out, slot_stride, d_comp_sizes, d_table, 0,
                         d_status, (cudaStream_t)stream);
}

int hufb200_decompress_blocks_dev(int k, size_t block_size, const uint8_t* d_comp, const uint64_t* d_offsets,
                                  const uint32_t* d_comp_sizes, size_t n_blocks, uint8_t* d_raw, size_t raw_n,
                                  uint32_t* d_status, void* stream) {
  if (!valid_k(k)) return fail(HUFB200_E_INVALID, "k=%d outside 1..%d", k, HUFB200_MAX_K);
  if (block_size == 0 || block_size > kMaxBlock) return fail(HUFB200_E_INVALID, "block_size outside 1..2^30");
  if (n_blocks >> 31) return fail(HUFB200_E_INVALID, "too many blocks");
  if (hufb200_blocks_count(raw_n, block_size) != n_blocks)
    return fail(HUFB200_E_INVALID, "n_blocks does not match raw_n / block_size");
  if (n_blocks && (!d_comp || !d_offsets || !d_comp_sizes || !d_raw)) return fail(HUFB200_E_INVALID, "null pointer");
  CU(launch_decompress(d_comp, reinterpret_cast<const unsigned long long*>(d_offsets), d_comp_sizes,
                       (uint32_t)n_blocks, k, decode_bpc(k), d_raw, raw_n, (uint32_t)block_size, d_status,
                       (cudaStream_t)stream));
  if (n_blocks) g_launches.fetch_add(1, std::memory_order_relaxed);
  return HUFB200_OK;
}

size_t hufb200_decompress_split_work_bytes(int k, size_t block_size, size_t n_blocks, size_t raw_n) {
  if (!valid_k(k) || block_size == 0 || block_size > kMaxBlock || (n_blocks >> 31)) return 0;
  int sms = 0;
  if (sm_count(&sms)) return 0;
  return decompress_split_work_bytes((uint32_t)n_blocks, k, (uint32_t)block_size,
                                     split_sub_bits(raw_n, (uint32_t)n_blocks, k, sms));
}

int hufb200_decompress_prefers_split(int k, size_t n_blocks, size_t raw_n) { return prefer_split(k, n_blocks, raw_n) ? 1 : 0; }

int hufb200_decompress_split_dev(int k, size_t block_size, const uint8_t* d_comp, const uint64_t* d_offsets,
                                 const uint32_t* d_comp_sizes, size_t n_blocks, uint8_t* d_raw, size_t raw_n,
                                 void* d_work, size_t work_bytes, uint32_t* d_status, void* stream) {
  if (!valid_k(k)) return fail(HUFB200_E_INVALID, "k=%d outside 1..%d", k, HUFB200_MAX_K);
  if (block_size == 0 || block_size > kMaxBlock) return fail(HUFB200_E_INVALID, "block_size outside 1..2^30");
  if (n_blocks >> 31) return fail(HUFB200_E_INVALID, "too many blocks");
  if (hufb200_blocks_count(raw_n, block_size) != n_blocks)
    return fail(HUFB200_E_INVALID, "n_blocks does not match raw_n / block_size");
  if (n_blocks && (!d_comp || !d_offsets || !d_comp_sizes || !d_raw || !d_work)) return fail(HUFB200_E_INVALID, "null pointer");
  if ((n_blocks * (size_t)k) >> 31) return fail(HUFB200_E_INVALID, "too many streams for the split decode");
  if (block_size / (size_t)k > kMaxSplitSlice) return fail(HUFB200_E_INVALID, "more than 256 MiB per stream: above the split decode's limit");
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  const uint32_t sub = split_sub_bits(raw_n, (uint32_t)n_blocks, k, sms);
  const size_t need = decompress_split_work_bytes((uint32_t)n_blocks, k, (uint32_t)block_size, sub);
  if (work_bytes < need) return fail(HUFB200_E_NOSPACE, "workspace: need %zu bytes, have %zu", need, work_bytes);
  if (((uintptr_t)d_work & 255) != 0) return fail(HUFB200_E_INVALID, "d_work must be 256-byte aligned");
  int launches = 0;
  CU(launch_decompress_split(d_comp, reinterpret_cast<const unsigned long long*>(d_offsets), d_comp_sizes,
                             (uint32_t)n_blocks, k, d_raw, raw_n, (uint32_t)block_size, sub, d_work, d_status, &launches,
                             (cudaStream_t)stream));
  g_launches.fetch_add((uint64_t)launches, std::memory_order_relaxed);
  return HUFB200_OK;
}

int hufb200_pack_blocks_dev(const uint8_t* d_slots, size_t slot_stride, const uint32_t* d_comp_sizes,
                            size_t n_blocks, uint8_t* d_packed, uint64_t* d_offsets, uint64_t* d_total,
                            void* stream) {
  if (n_blocks >> 31) return fail(HUFB200_E_INVALID, "too many blocks");
  if (!d_offsets || (n_blocks && (!d_slots || !d_comp_sizes))) return fail(HUFB200_E_INVALID, "null pointer");
  CU(launch_pack(d_slots, slot_stride, d_comp_sizes, (uint32_t)n_blocks, d_packed,
                 reinterpret_cast<unsigned long long*>(d_offsets), reinterpret_cast<unsigned long long*>(d_total),
                 (cudaStream_t)stream));
  g_launches.fetch_add(n_blocks && d_packed ? 2 : 1, std::memory_order_relaxed);
  return HUFB200_OK;
}

}  // extern "C"
