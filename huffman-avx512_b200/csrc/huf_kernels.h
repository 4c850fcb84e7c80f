// huf_kernels.h -- launcher interface between huf_kernels.cu and huf_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace hufb200 {

constexpr int kHistThreads = 512;
// compress CTA: kWorkThreads worker threads (histogram, encode; must be 256 = one per bin) plus
// one table-builder warp
#ifndef HUF_COMP_MINB
#define HUF_COMP_MINB 4
#endif
#ifndef HUF_COMP_WARPS
#define HUF_COMP_WARPS 8
#endif
constexpr int kWorkThreads = 32 * HUF_COMP_WARPS;  // >= 256: one thread per bin in the reduction
constexpr int kCompThreads = kWorkThreads + 32;
constexpr int kCompCtasPerSm = HUF_COMP_MINB;
constexpr int kCompWarps = kWorkThreads / 32;  // worker warps
constexpr int kDecMaxThreads = 256;

size_t table_bytes();
size_t decompress_smem_bytes(int K, int bpc, uint32_t block_size);

cudaError_t launch_histogram(const uint8_t* d_in, uint64_t n, unsigned long long* d_out, int grid,
                             cudaStream_t st);
cudaError_t launch_build_table(const unsigned long long* d_hist, void* d_table, cudaStream_t st);
cudaError_t launch_make_table(const uint32_t* d_hist, const uint16_t* d_len_count, const uint8_t* d_syms,
                              int n, int mode, void* d_table, cudaStream_t st);
cudaError_t launch_compress(const uint8_t* d_raw, uint64_t n, uint32_t block_size, int K, uint32_t n_blocks,
                            uint8_t* d_out, uint64_t slot_stride, uint32_t* d_sizes, const void* d_table,
                            int check_presence, uint32_t* d_status, int grid, cudaStream_t st);
cudaError_t launch_decompress(const uint8_t* d_comp, const unsigned long long* d_offsets,
                              const uint32_t* d_sizes, uint32_t n_blocks, int K, int bpc, uint8_t* d_raw,
                              uint64_t raw_n, uint32_t block_size, uint32_t* d_status, cudaStream_t st);
// Split decode (streams cut into items of sub_bits bits, one lane per item; see huf_kernels.cu):
// plan, scan, three sync passes, item scan, write pass -- no synchronisation.  d_work:
// decompress_split_work_bytes() bytes.
uint32_t split_sub_bits(uint64_t raw_n, uint32_t n_blocks, int K, int sms);
size_t decompress_split_work_bytes(uint32_t n_blocks, int K, uint32_t block_size, uint32_t sub_bits);
cudaError_t launch_decompress_split(const uint8_t* d_comp, const unsigned long long* d_offsets,
                                    const uint32_t* d_sizes, uint32_t n_blocks, int K, uint8_t* d_raw, uint64_t raw_n,
                                    uint32_t block_size, uint32_t sub_bits, void* d_work, uint32_t* d_status,
                                    int* launches, cudaStream_t st);
// The same for ONE buffer of at most 1 MiB compressed as a single launch (up to eight CTAs, whole
// streams each).  A status of kSplitSmallRetry (bit 1) means that a CTA's streams did not fit its
// items: nothing usable was written, take launch_decompress_split.
constexpr uint32_t kSplitSmallRetry = 2u;
bool split_small_fits(uint64_t comp_bytes, int K);
cudaError_t launch_decompress_split_small(const uint8_t* d_comp, const uint32_t* d_size, uint64_t comp_bytes, int K,
                                          uint8_t* d_raw, uint32_t raw_n, uint32_t* d_status, cudaStream_t st);
// One large buffer over the whole device (see huf_kernels.cu): five launches, no synchronisation.
size_t single_plan_bytes();  // the plan starts with u32 total_size (0 = a symbol without a code), u32 hdr_total
uint32_t single_piece_count(uint32_t n, int K);
cudaError_t launch_compress_single(const uint8_t* d_raw, uint32_t n, int K, const void* d_shared_table,
                                   uint32_t* d_hist, void* d_table_out, void* d_plan, uint32_t* d_piece_bits,
                                   uint8_t* d_out, int sms, cudaStream_t st);
cudaError_t launch_dump_dtable(const uint16_t* d_len_count, const uint8_t* d_syms, int num_syms, int one_symbol,
                               uint8_t* d_out, cudaStream_t st);
cudaError_t launch_pack(const uint8_t* d_slots, uint64_t slot_stride, const uint32_t* d_sizes, uint32_t n_blocks,
                        uint8_t* d_packed, unsigned long long* d_offsets, unsigned long long* d_total,
                        cudaStream_t st);
cudaError_t launch_scan_sizes(const uint32_t* d_sizes, uint32_t n_blocks, unsigned long long* d_offsets,
                              unsigned long long* d_total, cudaStream_t st);

}  // namespace hufb200
