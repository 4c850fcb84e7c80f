/* oracle/huf_oracle.c -- TEST INFRASTRUCTURE ONLY (see huf_oracle.h).
 *
 * Plain-C, single-threaded restatement of the reference codec's algorithm.  It
 * is written from the wire format rather than from the
 * reference's control flow: streams are produced as ordinary forward MSB-first
 * bitstreams and then laid down backwards into their regions, which yields the
 * same bytes as the reference's CodeWriter (codec/huffman.cpp:439-500).
 * Every function cites the reference lines it restates.  Parity pinned by
 * tests/test_oracle_vs_ref.py (against oracle/_ref) and tests/golden/.
 */
#include "huf_oracle.h"

#include <stdlib.h>
#include <string.h>

#define MAXLEN HUFO_MAX_CODE_LEN
#define SLOP 8 /* kSlop, codec/huffman.cpp:770 */

/* ---------------------------------------------------------------- histogram */

/* codec/histogram.cpp:184-191 */
void hufo_histogram(const uint8_t* in, size_t n, uint32_t out[256]) {
  memset(out, 0, 256 * sizeof(uint32_t));
  for (size_t i = 0; i < n; ++i) out[in[i]]++;
}

void hufo_histogram64(const uint8_t* in, size_t n, uint64_t out[256]) {
  memset(out, 0, 256 * sizeof(uint64_t));
  for (size_t i = 0; i < n; ++i) out[in[i]]++;
}

/* codec/huffman.cpp:98-108 */
void hufo_slice_sizes(size_t len, int k, size_t* sizes) {
  for (int i = 0; i < k; ++i) sizes[i] = len / (size_t)k + ((size_t)i < len % (size_t)k ? 1 : 0);
}

/* ------------------------------------------------- libstdc++ std::sort clone
 * codec/huffman.cpp:353-354 sorts the present symbols with std::sort and a
 * comparator that only looks at the counts, so the order of equal-count symbols
 * is whatever libstdc++'s introsort leaves.  This restates GCC
 * 13's bits/stl_algo.h + bits/stl_heap.h: introsort with depth limit 2*floor(log2 n),
 * median-of-three pivot moved to the front, unguarded Hoare partition, recursion
 * on the right part, threshold 16, heapsort fallback, final insertion sort with a
 * guarded 16-element prefix.  less(a,b) <=> hist[a] > hist[b].
 */
typedef struct {
  const uint32_t* hist;
} sort_ctx;

#define LESS(c, a, b) ((c)->hist[(a)] > (c)->hist[(b)])

static void sw(uint8_t* a, uint8_t* b) {
  uint8_t t = *a;
  *a = *b;
  *b = t;
}

/* std::__push_heap */
static void push_heap_(const sort_ctx* c, uint8_t* first, long hole, long top, uint8_t value) {
  long parent = (hole - 1) / 2;
  while (hole > top && LESS(c, first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

/* std::__adjust_heap */
static void adjust_heap_(const sort_ctx* c, uint8_t* first, long hole, long len, uint8_t value) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (LESS(c, first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  push_heap_(c, first, hole, top, value);
}

/* std::__partial_sort(first, last, last) == __heap_select + __sort_heap */
static void heap_sort_(const sort_ctx* c, uint8_t* first, uint8_t* last) {
  long len = last - first;
  if (len >= 2) { /* __make_heap */
    long parent = (len - 2) / 2;
    for (;;) {
      uint8_t v = first[parent];
      adjust_heap_(c, first, parent, len, v);
      if (parent == 0) break;
      parent--;
    }
  }
  /* __heap_select's loop over [middle,last) is empty because middle == last */
  while (last - first > 1) { /* __sort_heap / __pop_heap */
    --last;
    uint8_t v = *last;
    *last = *first;
    adjust_heap_(c, first, 0, last - first, v);
  }
}

/* std::__move_median_to_first */
static void median_to_first_(const sort_ctx* c, uint8_t* result, uint8_t* a, uint8_t* b,
                             uint8_t* cc) {
  if (LESS(c, *a, *b)) {
    if (LESS(c, *b, *cc)) sw(result, b);
    else if (LESS(c, *a, *cc)) sw(result, cc);
    else sw(result, a);
  } else if (LESS(c, *a, *cc)) sw(result, a);
  else if (LESS(c, *b, *cc)) sw(result, cc);
  else sw(result, b);
}

/* std::__unguarded_partition */
static uint8_t* partition_(const sort_ctx* c, uint8_t* first, uint8_t* last, uint8_t* pivot) {
  for (;;) {
    while (LESS(c, *first, *pivot)) ++first;
    --last;
    while (LESS(c, *pivot, *last)) --last;
    if (!(first < last)) return first;
    sw(first, last);
    ++first;
  }
}

/* std::__introsort_loop */
static void introsort_loop_(const sort_ctx* c, uint8_t* first, uint8_t* last, long depth) {
  while (last - first > 16) {
    if (depth == 0) {
      heap_sort_(c, first, last);
      return;
    }
    --depth;
    uint8_t* mid = first + (last - first) / 2;
    median_to_first_(c, first, first + 1, mid, last - 1);
    uint8_t* cut = partition_(c, first + 1, last, first);
    introsort_loop_(c, cut, last, depth);
    last = cut;
  }
}

/* std::__unguarded_linear_insert */
static void linear_insert_(const sort_ctx* c, uint8_t* last) {
  uint8_t val = *last;
  uint8_t* next = last - 1;
  while (LESS(c, val, *next)) {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}

/* std::__insertion_sort */
static void insertion_sort_(const sort_ctx* c, uint8_t* first, uint8_t* last) {
  if (first == last) return;
  for (uint8_t* i = first + 1; i != last; ++i) {
    if (LESS(c, *i, *first)) {
      uint8_t val = *i;
      memmove(first + 1, first, (size_t)(i - first));
      *first = val;
    } else {
      linear_insert_(c, i);
    }
  }
}

void hufo_sort_syms(const uint32_t hist[256], uint8_t* syms, int n) {
  sort_ctx c = {hist};
  if (n <= 0) return;
  long lg = 0;
  for (long t = n; t > 1; t >>= 1) lg++;
  introsort_loop_(&c, syms, syms + n, 2 * lg);
  if (n > 16) { /* std::__final_insertion_sort */
    insertion_sort_(&c, syms, syms + 16);
    for (uint8_t* i = syms + 16; i != syms + n; ++i) linear_insert_(&c, i);
  } else {
    insertion_sort_(&c, syms, syms + n);
  }
}

/* ------------------------------------------------------------- table build */

/* codec/huffman.cpp:297-327 */
void hufo_limit_code_lengths(uint16_t len_count[33]) {
  for (int i = MAXLEN + 1; i <= 32; ++i) {
    len_count[MAXLEN] = (uint16_t)(len_count[MAXLEN] + len_count[i]);
    len_count[i] = 0;
  }
  uint32_t kraft = 0;
  for (int i = 0; i <= MAXLEN; ++i) kraft += (uint32_t)len_count[i] << (MAXLEN - i);
  const uint32_t one = 1u << MAXLEN;
  while (kraft > one) {
    len_count[MAXLEN]--;
    for (int j = MAXLEN - 1; j >= 0; --j) {
      if (len_count[j] > 0) {
        len_count[j]--;
        len_count[j + 1] = (uint16_t)(len_count[j + 1] + 2);
        break;
      }
    }
    kraft--;
  }
}

/* ForallCodes, codec/huffman.cpp:260-284 */
void hufo_assign_codes(const uint16_t len_count[13], const uint8_t* syms, int num_syms,
                       uint16_t code_bits[256], uint16_t code_len[256]) {
  memset(code_bits, 0, 256 * sizeof(uint16_t));
  memset(code_len, 0, 256 * sizeof(uint16_t));
  uint32_t code = 0, inc = 1u << MAXLEN;
  int i = 0;
  for (int len = 0; len <= MAXLEN; ++len) {
    for (int j = 0; j < len_count[len] && i < num_syms; ++j) {
      code_bits[syms[i]] = (uint16_t)code;
      code_len[syms[i]] = (uint16_t)len;
      ++i;
      code += inc;
    }
    inc >>= 1;
  }
}

/* codec/huffman.cpp:339-437.  The depth histogram is obtained by propagating
 * depths from the root down the created internal nodes (children always have a
 * smaller index than their parent), which counts the same leaves per depth as
 * the reference's recursive CollectCodeLen (:329-337). */
void hufo_make_coding(const uint32_t hist[256], hufo_coding* out) {
  memset(out, 0, sizeof(*out));
  int n = 0;
  for (int c = 0; c < 256; ++c)
    if (hist[c] != 0) out->sorted_syms[n++] = (uint8_t)c;
  out->num_syms = n;
  if (n == 0) return;
  hufo_sort_syms(hist, out->sorted_syms, n);

  uint32_t tree_count[256];
  int child[256][2];
  int next_sym = n - 1, next_node = 0, tree_size = 0;
  while ((tree_size - next_node) + (next_sym + 1) > 1) {
    int picked[2];
    uint32_t sum = 0;
    for (int t = 0; t < 2; ++t) {
      int take_leaf = 0;
      if (next_sym >= 0) {
        if (next_node == tree_size) take_leaf = 1;
        else take_leaf = hist[out->sorted_syms[next_sym]] <= tree_count[next_node]; /* :375 */
      }
      if (take_leaf) {
        sum += hist[out->sorted_syms[next_sym--]];
        picked[t] = -1;
      } else {
        sum += tree_count[next_node];
        picked[t] = next_node++;
      }
    }
    child[tree_size][0] = picked[0];
    child[tree_size][1] = picked[1];
    tree_count[tree_size] = sum; /* u32 wrap-around like the reference (:365, :414) */
    tree_size++;
  }
  if (tree_size == 0) {
    out->len_count[0] = 1; /* single symbol: the root is a leaf at depth 0 (:417-418) */
  } else {
    int depth[256];
    depth[tree_size - 1] = 0;
    for (int node = tree_size - 1; node >= 0; --node) {
      for (int t = 0; t < 2; ++t) {
        int ch = child[node][t];
        int d = depth[node] + 1;
        if (ch < 0) out->len_count[d > 32 ? 32 : d]++;
        else depth[ch] = d;
      }
    }
  }
  hufo_limit_code_lengths(out->len_count);
  for (int i = 0; i <= MAXLEN; ++i)
    if (out->len_count[i]) out->len_mask |= 1u << i;
  hufo_assign_codes(out->len_count, out->sorted_syms, n, out->code_bits, out->code_len);
}

/* ------------------------------------------------------------------ encode */

static void put_u32(uint8_t* p, uint32_t x) { /* write_u32, :242-248 (little endian) */
  p[0] = (uint8_t)x;
  p[1] = (uint8_t)(x >> 8);
  p[2] = (uint8_t)(x >> 16);
  p[3] = (uint8_t)(x >> 24);
}
static uint32_t get_u32(const uint8_t* p) {
  return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}

size_t hufo_compress_bound(size_t n, int k) {
  return 8 + 13 + 256 + 4 * (size_t)(k - 1) + (n * MAXLEN + 7) / 8 + (size_t)k * (SLOP + 1);
}

static int compress_with_coding(int k, const uint8_t* raw, size_t n, const hufo_coding* cd,
                                uint8_t* out, size_t cap, size_t* out_len) {
  if (k < 1 || k > HUFO_MAX_K) return -2;
  size_t sizes[HUFO_MAX_K];
  hufo_slice_sizes(n, k, sizes);

  /* per-stream bit totals and region sizes (:772-786) */
  size_t region[HUFO_MAX_K];
  uint64_t bits[HUFO_MAX_K];
  size_t payload = 0;
  {
    const uint8_t* p = raw;
    for (int s = 0; s < k; ++s) {
      uint64_t b = 0;
      for (size_t i = 0; i < sizes[s]; ++i) b += cd->code_len[p[i]];
      p += sizes[s];
      bits[s] = b;
      region[s] = (size_t)((b + 7) / 8) + SLOP;
      payload += region[s];
    }
  }
  int npop = 0;
  for (int i = 0; i <= MAXLEN; ++i) npop += (cd->len_mask >> i) & 1;
  const size_t header = 8 + (size_t)npop + (size_t)cd->num_syms + 4 * (size_t)(k - 1);
  *out_len = header + payload;
  if (*out_len > cap) return -1;

  /* header (:794-811) */
  uint8_t* w = out;
  put_u32(w, (uint32_t)n);
  put_u32(w + 4, cd->len_mask);
  w += 8;
  for (int len = 0; len < 32; ++len)
    if (cd->len_count[len] != 0) *w++ = (uint8_t)cd->len_count[len]; /* 256 wraps to 0 (:804) */
  memcpy(w, cd->sorted_syms, (size_t)cd->num_syms);
  w += cd->num_syms;
  {
    size_t end = 0;
    for (int s = 0; s < k - 1; ++s) {
      end += region[s];
      put_u32(w, (uint32_t)end);
      w += 4;
    }
  }
  /* payload: region s = 8 zero bytes + ... + first stream byte at the very end */
  memset(w, 0, payload);
  const uint8_t* p = raw;
  for (int s = 0; s < k; ++s) {
    uint8_t* last = w + region[s] - 1; /* stream byte j lives at last - j */
    uint64_t pos = 0;                  /* bit position in the forward stream */
    for (size_t i = 0; i < sizes[s]; ++i) {
      unsigned len = cd->code_len[p[i]];
      unsigned code = (unsigned)cd->code_bits[p[i]] >> (MAXLEN - len); /* right-aligned */
      for (unsigned b = 0; b < len; ++b) {
        if ((code >> (len - 1 - b)) & 1u) last[-(long)(pos >> 3)] |= (uint8_t)(0x80u >> (pos & 7));
        pos++;
      }
    }
    (void)bits;
    p += sizes[s];
    w += region[s];
  }
  return 0;
}

/* codec/huffman.cpp:738-846 */
int hufo_compress(int k, const uint8_t* raw, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
  uint32_t hist[256];
  hufo_histogram(raw, n, hist);
  hufo_coding cd;
  hufo_make_coding(hist, &cd);
  return compress_with_coding(k, raw, n, &cd, out, cap, out_len);
}

int hufo_compress_with_table(int k, const uint8_t* raw, size_t n, const uint16_t len_count[13],
                             const uint8_t* sorted_syms, int num_syms, uint8_t* out, size_t cap,
                             size_t* out_len) {
  hufo_coding cd;
  memset(&cd, 0, sizeof(cd));
  for (int i = 0; i <= MAXLEN; ++i) {
    cd.len_count[i] = len_count[i];
    if (len_count[i]) cd.len_mask |= 1u << i;
  }
  memcpy(cd.sorted_syms, sorted_syms, (size_t)num_syms);
  cd.num_syms = num_syms;
  hufo_assign_codes(len_count, sorted_syms, num_syms, cd.code_bits, cd.code_len);
  return compress_with_coding(k, raw, n, &cd, out, cap, out_len);
}

/* ------------------------------------------------------------------ decode */

/* Decoder1x, codec/huffman.cpp:594-609: entry {code_len, sym} for every 12-bit prefix. */
void hufo_dtable1x(const uint16_t len_count[13], const uint8_t* syms, int num_syms,
                   uint8_t out[4096 * 2]) {
  memset(out, 0, 4096 * 2);
  uint32_t code = 0;
  int i = 0;
  for (int len = 0; len <= MAXLEN; ++len) {
    uint32_t span = 1u << (MAXLEN - len);
    for (int j = 0; j < len_count[len] && i < num_syms; ++j, ++i) {
      for (uint32_t e = code; e < code + span && e < 4096; ++e) {
        out[2 * e] = (uint8_t)len;
        out[2 * e + 1] = syms[i];
      }
      code += span;
    }
  }
}

/* Decoder2x, codec/huffman.cpp:642-681, built index-wise: look the
 * first symbol up, shift it out, look the second up in the zero-padded rest; the
 * pair is taken when both lengths fit in 12 bits. */
void hufo_dtable2x(const uint16_t len_count[13], const uint8_t* syms, int num_syms,
                   uint8_t out[4096 * 4]) {
  uint8_t t1[4096 * 2];
  hufo_dtable1x(len_count, syms, num_syms, t1);
  memset(out, 0, 4096 * 4);
  if (num_syms == 0) return;
  for (uint32_t e = 0; e < 4096; ++e) {
    unsigned l1 = t1[2 * e];
    uint32_t rest = (e << l1) & 0xfffu;
    unsigned l2 = t1[2 * rest];
    if (l1 + l2 <= MAXLEN) {
      out[4 * e] = (uint8_t)(l1 + l2);
      out[4 * e + 1] = t1[2 * e + 1];
      out[4 * e + 2] = t1[2 * rest + 1];
      out[4 * e + 3] = 2;
    } else {
      out[4 * e] = (uint8_t)l1;
      out[4 * e + 1] = t1[2 * e + 1];
      out[4 * e + 2] = 0;
      out[4 * e + 3] = 1;
    }
  }
}

/* Next 12 bits of the forward stream whose byte j is region_last[-j]; bytes
 * below `floor` read as zero (CodeReader::FillBuffer's slow path, :540-549). */
static uint32_t peek12(const uint8_t* last, const uint8_t* floor_, uint64_t pos) {
  uint32_t v = 0;
  for (int t = 0; t < 3; ++t) {
    const uint8_t* p = last - (long)(pos >> 3) - t;
    v = (v << 8) | (p >= floor_ ? *p : 0u);
  }
  return (v >> (12 - (pos & 7))) & 0xfffu;
}

/* codec/huffman.cpp:892-960 (+ header parse :714-736).  Each stream yields
 * exactly sizes[k] symbols: pairs from the two-symbol table while at least two
 * remain, then single symbols (:870-878). */
int hufo_decompress(int k, const uint8_t* comp, size_t n, uint8_t* out, size_t cap,
                    size_t* out_len) {
  if (k < 1 || k > HUFO_MAX_K) return -2;
  if (n < 8) return -3;
  const uint32_t raw_size = get_u32(comp);
  const uint32_t mask = get_u32(comp + 4);
  size_t pos = 8;
  uint16_t len_count[13] = {0};
  int num_syms = 0, npop = 0;
  for (int i = 0; i <= MAXLEN; ++i) npop += (mask >> i) & 1;
  for (int i = 0; i <= MAXLEN; ++i) {
    if (mask & (1u << i)) {
      if (pos >= n) return -3;
      len_count[i] = comp[pos];
      if (npop == 1 && comp[pos] == 0) len_count[i] = 256; /* :724-728 */
      pos++;
      num_syms += len_count[i];
    }
  }
  if (num_syms > 256 || pos + (size_t)num_syms + 4 * (size_t)(k - 1) > n) return -3;
  const uint8_t* syms = comp + pos;
  pos += (size_t)num_syms;
  size_t end_off[HUFO_MAX_K];
  for (int s = 0; s < k - 1; ++s) {
    end_off[s] = get_u32(comp + pos);
    pos += 4;
  }
  const uint8_t* payload = comp + pos;
  end_off[k - 1] = n - pos;
  *out_len = raw_size;
  if (raw_size > cap) return -1;

  uint8_t t1[4096 * 2], t2[4096 * 4];
  hufo_dtable1x(len_count, syms, num_syms, t1);
  hufo_dtable2x(len_count, syms, num_syms, t2);
  size_t sizes[HUFO_MAX_K];
  hufo_slice_sizes(raw_size, k, sizes);
  uint8_t* o = out;
  for (int s = 0; s < k; ++s) {
    if (end_off[s] > n - pos || end_off[s] == 0) {
      if (sizes[s]) return -3;
      continue;
    }
    const uint8_t* last = payload + end_off[s] - 1;
    uint64_t bit = 0;
    size_t left = sizes[s];
    while (left >= 2) {
      const uint8_t* e = t2 + 4 * peek12(last, payload, bit);
      o[0] = e[1];
      o[1] = e[2];
      o += e[3];
      left -= e[3];
      bit += e[0];
      if (e[3] == 0) return -3;
    }
    while (left >= 1) {
      const uint8_t* e = t1 + 2 * peek12(last, payload, bit);
      *o++ = e[1];
      left--;
      bit += e[0];
    }
  }
  return 0;
}
