/* oracle/huf_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's multi-stream Huffman codec hot path
 * (ahartik/huffman-avx512, codec/huffman.cpp + codec/histogram.cpp).  It is the
 * CHECKER for the CUDA path; nothing under huffman-avx512_b200/ may call it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * legs load it.  Parity is PINNED: tests/test_oracle_vs_ref.py compares every
 * function here byte-for-byte with the unmodified reference (oracle/_ref) and
 * tests/test_oracle_golden.py against fixtures generated from it
 * (tests/golden/, made by tests/golden/make_golden.py).
 */
#ifndef HUF_ORACLE_H_
#define HUF_ORACLE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HUFO_MAX_CODE_LEN 12
#define HUFO_MAX_K 64

typedef struct {
  uint16_t code_bits[256]; /* left-aligned in 12 bits (BitCode.bits, huffman.cpp:214-224) */
  uint16_t code_len[256];
  uint8_t sorted_syms[256];
  int num_syms;
  uint16_t len_count[33];
  uint32_t len_mask;
} hufo_coding;

/* MakeHistogramSimple, codec/histogram.cpp:184-191 (all variants agree, histogram_test.cpp:29-42). */
void hufo_histogram(const uint8_t* in, size_t n, uint32_t out[256]);
void hufo_histogram64(const uint8_t* in, size_t n, uint64_t out[256]);
/* SliceSizes<K>, codec/huffman.cpp:98-108. */
void hufo_slice_sizes(size_t len, int k, size_t* sizes);
/* libstdc++ 13 std::sort with the comparator of codec/huffman.cpp:353-354. */
void hufo_sort_syms(const uint32_t hist[256], uint8_t* syms, int n);
/* LimitCodeLengths, codec/huffman.cpp:297-327. */
void hufo_limit_code_lengths(uint16_t len_count[33]);
/* MakeCanonicalCoding, codec/huffman.cpp:339-437. */
void hufo_make_coding(const uint32_t hist[256], hufo_coding* out);
/* Code assignment only (ForallCodes, :260-284) from (len_count, sorted_syms). */
void hufo_assign_codes(const uint16_t len_count[13], const uint8_t* syms, int num_syms,
                       uint16_t code_bits[256], uint16_t code_len[256]);

/* >= size of CompressMulti<k>(n bytes). */
size_t hufo_compress_bound(size_t n, int k);
/* CompressMulti<k>, codec/huffman.cpp:738-846. 0 ok, -1 cap too small, -2 bad k. */
int hufo_compress(int k, const uint8_t* raw, size_t n, uint8_t* out, size_t cap, size_t* out_len);
/* Same wire format but with an externally supplied table (shared-table mode):
 * every symbol present in raw must have a code in the table. */
int hufo_compress_with_table(int k, const uint8_t* raw, size_t n, const uint16_t len_count[13],
                             const uint8_t* sorted_syms, int num_syms, uint8_t* out, size_t cap,
                             size_t* out_len);
/* DecompressMulti<k>, codec/huffman.cpp:892-960. 0 ok, -1 cap too small, -2 bad k, -3 malformed. */
int hufo_decompress(int k, const uint8_t* comp, size_t n, uint8_t* out, size_t cap,
                    size_t* out_len);
/* Decoder1x / Decoder2x tables, codec/huffman.cpp:594-704; 4096 entries each. */
void hufo_dtable1x(const uint16_t len_count[13], const uint8_t* syms, int num_syms,
                   uint8_t out[4096 * 2]);
void hufo_dtable2x(const uint16_t len_count[13], const uint8_t* syms, int num_syms,
                   uint8_t out[4096 * 4]);

#ifdef __cplusplus
}
#endif
#endif /* HUF_ORACLE_H_ */
