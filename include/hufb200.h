/* hufb200.h -- C ABI of the B200-native multi-stream Huffman codec.
 *
 * Drop-in boundary for the hot path of ahartik/huffman-avx512
 * (histogram -> canonical table build -> N-stream encode -> N-stream decode).
 * Every compressed buffer produced here is byte-identical to what the
 * reference's scalar encoder emits for the same input and stream count K, and
 * every buffer the reference emits is decoded to the original bytes.
 *
 * Each entry point names the reference interface it replaces (paths relative
 * to the reference repository).  All functions return HUFB200_OK (0) or a
 * negative HUFB200_E_* code; none aborts (the reference asserts/aborts instead,
 * codec/huffman.cpp:55-59).  There is NO CPU fallback: without a usable CUDA
 * device every compute entry point fails with HUFB200_E_NODEVICE/E_CUDA.
 *
 * Pointers named d_* are DEVICE pointers; everything else is HOST memory.
 * `stream` is a cudaStream_t passed as void* (NULL = default stream).  The
 * *_dev entry points neither allocate nor synchronise.
 */
#ifndef HUFB200_H_
#define HUFB200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define HUFB200_API __attribute__((visibility("default")))
#else
#define HUFB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define HUFB200_OK 0
#define HUFB200_E_INVALID (-1)  /* bad argument (k, block_size, null pointer, size limits) */
#define HUFB200_E_NOSPACE (-2)  /* output capacity too small; *out_len holds the size needed */
#define HUFB200_E_CUDA (-3)     /* CUDA runtime error; see hufb200_last_error() */
#define HUFB200_E_CORRUPT (-4)  /* malformed compressed input / symbol without a code */
#define HUFB200_E_NODEVICE (-5) /* no CUDA device visible */

#define HUFB200_MAX_CODE_LEN 12 /* kMaxCodeLength, codec/huffman.cpp:38 */
#define HUFB200_MAX_K 64        /* stream counts 1..64; the reference ships 1,2,4,8,16,24,32,40,48 */

/* ---- library ---- */
HUFB200_API int hufb200_version(void);
HUFB200_API const char* hufb200_last_error(void); /* thread-local message of the last failing call */
HUFB200_API int hufb200_device_count(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
HUFB200_API uint64_t hufb200_launch_count(void);
/* The host-pointer entry points keep per-thread device buffers, pinned words and streams between
 * calls (grow-only, sized by the largest call so far).  This frees the calling thread's; they are
 * also freed when the thread exits.  The reference keeps no state: call this
 * where that matters. */
HUFB200_API void hufb200_release_workspace(void);

/* ---- histogram: huffman::MakeHistogram, codec/histogram.h:12, codec/histogram.cpp:193-201 ---- */
/* n < 2^32 (ByteHistogram is u32, codec/histogram.h:10). */
HUFB200_API int hufb200_histogram(const uint8_t* in, size_t n, uint32_t out[256]);
/* 64-bit bins for inputs beyond what ByteHistogram can count. */
HUFB200_API int hufb200_histogram64(const uint8_t* in, size_t n, uint64_t out[256]);
/* d_out: 256 x u64, overwritten.  Launches on `stream`. */
HUFB200_API int hufb200_histogram_dev(const uint8_t* d_in, size_t n, uint64_t* d_out, void* stream);

/* ---- table build: MakeCanonicalCoding, codec/huffman.cpp:339-437 (incl. LimitCodeLengths
 * :297-327 and ForallCodes :260-284).  Runs the table-build kernel on the device. ---- */
HUFB200_API int hufb200_make_table(const uint32_t hist[256], uint16_t len_count[13], uint8_t sorted_syms[256],
                       int* num_syms, uint32_t* len_mask, uint16_t code_bits[256],
                       uint16_t code_len[256]);
/* Two-symbol decode table: Decoder2x, codec/huffman.cpp:642-681.  out = 4096 entries of
 * {num_bits, sym0, sym1, num_syms} (DecodedSym2x, :634-640).  Runs the decode kernel's builder. */
HUFB200_API int hufb200_decode_table(const uint16_t len_count[13], const uint8_t* sorted_syms, int num_syms,
                         uint8_t out[4096 * 4]);
/* One-symbol decode table: Decoder1x, codec/huffman.cpp:594-632.  out = 4096 entries of
 * {code_len, sym} (DecodedSym, :588-592); entries no code reaches stay {0, 0}.  Same builder. */
HUFB200_API int hufb200_decode_table1x(const uint16_t len_count[13], const uint8_t* sorted_syms, int num_syms,
                           uint8_t out[4096 * 2]);

/* ---- single buffer: huffman::CompressMulti<K> / DecompressMulti<K>,
 * codec/huffman.h:9-12, codec/huffman.cpp:738-846 / :892-960 ---- */
/* Upper bound of the compressed size of n raw bytes with k streams. */
HUFB200_API size_t hufb200_compress_bound(size_t n, int k);
/* out receives exactly the bytes CompressMulti<k>(raw) returns.  n <= 2^30: the format's offsets
 * are 32-bit and the reference computes them as int (codec/huffman.cpp:772); larger inputs get
 * HUFB200_E_INVALID -- cut them into blocks (hufb200_compress_blocks). */
HUFB200_API int hufb200_compress(int k, const uint8_t* raw, size_t n, uint8_t* out, size_t cap,
                     size_t* out_len);
/* out receives exactly the bytes DecompressMulti<k>(comp) returns.  The size comes from the
 * buffer's untrusted raw_size field: it is checked against cap before anything is allocated. */
HUFB200_API int hufb200_decompress(int k, const uint8_t* comp, size_t n, uint8_t* out, size_t cap,
                       size_t* out_len);
/* raw_size field of a compressed buffer (ParseCompressedHeader, codec/huffman.cpp:714-718). */
HUFB200_API int hufb200_raw_size(const uint8_t* comp, size_t n, size_t* raw_size);
/* Same wire format with a caller-supplied table ("same table" parity mode). Every symbol of
 * raw must have a code, else HUFB200_E_CORRUPT. */
HUFB200_API int hufb200_compress_with_table(int k, const uint8_t* raw, size_t n, const uint16_t len_count[13],
                                const uint8_t* sorted_syms, int num_syms, uint8_t* out,
                                size_t cap, size_t* out_len);

/* ---- block container (ours; the reference has no framing, codec/huffman.cpp:794-811):
 * the input is cut into blocks of block_size bytes (the last may be shorter), each block is an
 * independent CompressMulti<k> buffer.  Container = 32-byte header, u32 comp_size[n_blocks],
 * then the block buffers back to back.  See DESIGN.md "Block container". ---- */
HUFB200_API size_t hufb200_blocks_count(size_t n, size_t block_size);
HUFB200_API size_t hufb200_container_bound(size_t n, size_t block_size, int k);
HUFB200_API int hufb200_compress_blocks(int k, size_t block_size, const uint8_t* raw, size_t n, uint8_t* out,
                            size_t cap, size_t* out_len);
HUFB200_API int hufb200_decompress_blocks(const uint8_t* container, size_t n, uint8_t* out, size_t cap,
                              size_t* out_len);
HUFB200_API int hufb200_container_info(const uint8_t* container, size_t n, int* k, size_t* block_size,
                           size_t* raw_size, size_t* n_blocks);

/* ---- device-resident batched path (what bench.py times) ---- */
/* Bytes between consecutive output slots: compress_bound(block_size,k) rounded up to 256. */
HUFB200_API size_t hufb200_slot_stride(size_t block_size, int k);
/* Compresses ceil(n/block_size) blocks of d_raw (16-byte aligned).  Block b is written at
 * d_out + b*slot_stride (d_out 16-byte aligned, slot_stride >= hufb200_slot_stride()), its size
 * to d_comp_sizes[b].  Up to 3 zero bytes may be written past a block's end inside its slot.
 * d_table: NULL = one table per block (the reference's behaviour); otherwise a shared table
 * built by hufb200_build_table_dev (every block still carries the full header).
 * d_status: one u32, set non-zero by the kernel on E_CORRUPT conditions (may be NULL).
 * Launches may overlap on different streams; each takes one of 4096 device-side work counters in
 * turn, so at most 4095 compress launches may be outstanding on a device at once. */
HUFB200_API int hufb200_compress_blocks_dev(int k, size_t block_size, const uint8_t* d_raw, size_t n,
                                uint8_t* d_out, size_t slot_stride, uint32_t* d_comp_sizes,
                                const void* d_table, uint32_t* d_status, void* stream);
/* Decodes n_blocks blocks; block b starts at d_comp + d_offsets[b], is d_comp_sizes[b] bytes long
 * and decodes to d_raw + b*block_size (raw_n total bytes; the last block may be shorter).  The
 * compressed bytes are fetched in whole aligned 32-byte sectors: the allocation that holds d_comp
 * must cover the sectors of its first and last byte (true for any pointer into cudaMalloc'ed
 * memory whose allocation extends to the next 32-byte boundary after the last block). */
HUFB200_API int hufb200_decompress_blocks_dev(int k, size_t block_size, const uint8_t* d_comp,
                                  const uint64_t* d_offsets, const uint32_t* d_comp_sizes,
                                  size_t n_blocks, uint8_t* d_raw, size_t raw_n,
                                  uint32_t* d_status, void* stream);
/* Split decode: the same result as hufb200_decompress_blocks_dev for inputs whose n_blocks * k
 * streams are too few to fill the device (one large buffer: k lanes).  DecompressMulti<K>
 * (codec/huffman.cpp:892-955) reads each stream from its first bit; here every stream is cut into
 * items of a few thousand bits, each decoded by its own lane from a guessed start -- prefix codes
 * fall into step after a few codes -- until the starts are consistent (what never falls into step
 * is decoded serially behind the last good item), then written out with one lane per item.
 * d_work: hufb200_decompress_split_work_bytes() bytes, 256-byte aligned, contents irrelevant.
 * block_size / k <= 256 MiB (bit positions inside a stream are 32-bit here), n_blocks * k < 2^31.
 * hufb200_decompress_prefers_split: the library's own choice between the two (the host-pointer
 * calls apply it themselves). */
HUFB200_API size_t hufb200_decompress_split_work_bytes(int k, size_t block_size, size_t n_blocks, size_t raw_n);
HUFB200_API int hufb200_decompress_prefers_split(int k, size_t n_blocks, size_t raw_n);
HUFB200_API int hufb200_decompress_split_dev(int k, size_t block_size, const uint8_t* d_comp,
                                 const uint64_t* d_offsets, const uint32_t* d_comp_sizes,
                                 size_t n_blocks, uint8_t* d_raw, size_t raw_n, void* d_work,
                                 size_t work_bytes, uint32_t* d_status, void* stream);
/* Gathers the slots into one packed byte string: block b is copied to d_packed + d_offsets[b],
 * where d_offsets = exclusive prefix sum of d_comp_sizes (computed here, on the device). */
HUFB200_API int hufb200_pack_blocks_dev(const uint8_t* d_slots, size_t slot_stride,
                            const uint32_t* d_comp_sizes, size_t n_blocks, uint8_t* d_packed,
                            uint64_t* d_offsets, uint64_t* d_total, void* stream);
/* Shared table: sizeof the opaque device table, and its construction from a 256 x u64 histogram
 * (e.g. the all-reduced histogram of every shard). */
HUFB200_API size_t hufb200_table_bytes(void);
HUFB200_API int hufb200_build_table_dev(const uint64_t* d_hist, void* d_table, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HUFB200_H_ */
