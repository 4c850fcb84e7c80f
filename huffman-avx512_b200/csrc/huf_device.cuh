// huf_device.cuh -- shared device-side definitions for the sm_100a Huffman kernels.
//
// Wire format and algorithm follow ahartik/huffman-avx512 (codec/huffman.cpp); the
// decomposition into CTA/warp phases is ours.  Citations are
// `file:line` in the reference repository.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace hufb200 {

constexpr int kMaxCodeLen = 12;   // kMaxCodeLength, codec/huffman.cpp:38
constexpr int kSlop = 8;          // kSlop, codec/huffman.cpp:770
constexpr int kMaxK = 64;
constexpr uint32_t kEncInvalid = 0x40000000u;  // enc[] entry of a symbol without a code

// Encode-side table.  Lives in shared memory (per-block tables) or in global
// memory (shared-table mode, built by k_build_table).
struct HufTable {
  uint32_t enc[256];         // code right-aligned in bits 0..11, length in bits 16..19
  uint8_t sorted_syms[256];  // canonical order (CanonicalCoding::sorted_syms, :288)
  uint16_t len_count[16];    // [0..12] used (CanonicalCoding::len_count, :290)
  uint32_t len_mask;         // bit i set <=> len_count[i] != 0 (:422-426)
  int32_t num_syms;
  uint32_t hdr_len;          // 8 + popcount(len_mask) + num_syms
  uint32_t pad_;
};

// Scratch for the table build (shared memory, one per CTA).
struct TableScratch {
  unsigned long long keys[256];        // (count << 8) | symbol, sorted by count descending
  unsigned long long tree_count[256];  // internal-node weights (tree_count, :365)
  uint16_t node_parent[256];
  uint16_t leaf_parent[256];
  uint16_t node_depth[256];
  uint32_t len_count33[36];            // depth histogram before limiting (:290, :329-337)
  uint32_t cum[16];                    // inclusive prefix of len_count
  uint32_t start_code[16];             // first left-aligned code of each length
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// SliceSizes<K>, codec/huffman.cpp:98-108: start offset and size of slice s.
__device__ __forceinline__ void slice_geom(uint32_t n, int K, int s, uint32_t& start, uint32_t& size) {
  const uint32_t q = n / (uint32_t)K, r = n % (uint32_t)K;
  const uint32_t us = (uint32_t)s;
  size = q + (us < r ? 1u : 0u);
  start = us * q + (us < r ? us : r);
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane_id() >= d) v += t;
  }
  return v;
}

// ---------------------------------------------------------------------------
// libstdc++ std::sort clone (GCC 13 bits/stl_algo.h, bits/stl_heap.h) run by ONE
// thread on the key array.  The reference sorts the present symbols with a
// comparator that ignores the symbol value (codec/huffman.cpp:353-354), so the
// order of equal-count symbols -- and with it sorted_syms, every code and every
// compressed byte -- is whatever introsort leaves.  Reproducing
// it exactly is the only way to emit the reference's bytes.
// less(a, b) <=> count(a) > count(b), count = key >> 8.
// ---------------------------------------------------------------------------
namespace sortclone {
typedef unsigned long long key_t;
__device__ __forceinline__ bool less(key_t a, key_t b) { return (a >> 8) > (b >> 8); }
__device__ __forceinline__ void swp(key_t* a, int i, int j) {
  key_t t = a[i];
  a[i] = a[j];
  a[j] = t;
}

__device__ inline void push_heap(key_t* a, int first, int hole, int top, key_t value) {
  int parent = (hole - 1) / 2;
  while (hole > top && less(a[first + parent], value)) {
    a[first + hole] = a[first + parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  a[first + hole] = value;
}

__device__ inline void adjust_heap(key_t* a, int first, int hole, int len, key_t value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (less(a[first + child], a[first + child - 1])) child--;
    a[first + hole] = a[first + child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    a[first + hole] = a[first + child - 1];
    hole = child - 1;
  }
  push_heap(a, first, hole, top, value);
}

// std::__partial_sort(first, last, last): make_heap + sort_heap
__device__ inline void heap_sort(key_t* a, int first, int last) {
  int len = last - first;
  if (len >= 2) {
    int parent = (len - 2) / 2;
    for (;;) {
      key_t v = a[first + parent];
      adjust_heap(a, first, parent, len, v);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {
    --last;
    key_t v = a[last];
    a[last] = a[first];
    adjust_heap(a, first, 0, last - first, v);
  }
}

__device__ inline void linear_insert(key_t* a, int last) {
  key_t val = a[last];
  int next = last - 1;
  while (less(val, a[next])) {
    a[last] = a[next];
    last = next;
    --next;
  }
  a[last] = val;
}

__device__ inline void insertion_sort(key_t* a, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (less(a[i], a[first])) {
      key_t val = a[i];
      for (int j = i; j > first; --j) a[j] = a[j - 1];
      a[first] = val;
    } else {
      linear_insert(a, i);
    }
  }
}

__device__ inline void sort(key_t* a, int n) {
  if (n <= 1) return;
  if (n > 16) {
    int lg = 31 - __clz(n);
    // explicit stack instead of the recursion on the right part; the two parts are
    // disjoint, so the order in which they are processed does not change the result
    int st_first[40], st_last[40], st_depth[40];
    int sp = 0;
    int first = 0, last = n, depth = 2 * lg;
    for (;;) {
      while (last - first > 16) {
        if (depth == 0) {
          heap_sort(a, first, last);
          break;
        }
        --depth;
        // __unguarded_partition_pivot
        const int mid = first + (last - first) / 2;
        {  // __move_median_to_first(first, first+1, mid, last-1)
          const int ia = first + 1, ib = mid, ic = last - 1;
          if (less(a[ia], a[ib])) {
            if (less(a[ib], a[ic])) swp(a, first, ib);
            else if (less(a[ia], a[ic])) swp(a, first, ic);
            else swp(a, first, ia);
          } else if (less(a[ia], a[ic])) swp(a, first, ia);
          else if (less(a[ib], a[ic])) swp(a, first, ic);
          else swp(a, first, ib);
        }
        int lo = first + 1, hi = last;
        const key_t pivot_unused = 0;
        (void)pivot_unused;
        for (;;) {  // __unguarded_partition(first+1, last, first); the pivot stays at a[first]
          while (less(a[lo], a[first])) ++lo;
          --hi;
          while (less(a[first], a[hi])) --hi;
          if (!(lo < hi)) break;
          swp(a, lo, hi);
          ++lo;
        }
        const int cut = lo;
        // right part [cut, last) is pushed, left part [first, cut) continues
        st_first[sp] = cut;
        st_last[sp] = last;
        st_depth[sp] = depth;
        ++sp;
        last = cut;
      }
      if (sp == 0) break;
      --sp;
      first = st_first[sp];
      last = st_last[sp];
      depth = st_depth[sp];
    }
    // __final_insertion_sort
    insertion_sort(a, 0, 16);
    for (int i = 16; i != n; ++i) linear_insert(a, i);
  } else {
    insertion_sort(a, 0, n);
  }
}
}  // namespace sortclone

// ---------------------------------------------------------------------------
// Table build, executed by ONE WARP (all 32 lanes must call it).
//   hist  : 256 counts (shared or global memory), CountT = uint32_t or uint64_t
//   tab   : output (shared memory)
// Restates MakeCanonicalCoding (codec/huffman.cpp:339-437): present symbols ->
// std::sort by count descending -> two-queue Huffman merge (leaf preferred on
// ties, :375) -> depth histogram -> LimitCodeLengths (:297-327) -> canonical
// codes (ForallCodes, :260-284).
// ---------------------------------------------------------------------------
template <typename CountT>
__device__ inline void build_table_warp(const CountT* hist, HufTable* tab, TableScratch* sc) {
  const int lane = lane_id();
  // 1. present symbols in ascending symbol order (:342-347)
  int n = 0;
#pragma unroll 1
  for (int base = 0; base < 256; base += 32) {
    const int c = base + lane;
    const unsigned long long cnt = (unsigned long long)hist[c];
    const unsigned m = __ballot_sync(0xffffffffu, cnt != 0);
    if (cnt != 0) sc->keys[n + __popc(m & ((1u << lane) - 1))] = (cnt << 8) | (unsigned)c;
    n += __popc(m);
  }
  for (int i = lane; i < 256; i += 32) tab->enc[i] = kEncInvalid;
  for (int i = lane; i < 36; i += 32) sc->len_count33[i] = 0;
  __syncwarp();

  if (n > 0) {
    // 2. sort + 3. Huffman merge: inherently serial, lane 0 only
    if (lane == 0) {
      sortclone::sort(sc->keys, n);
      int next_sym = n - 1, next_node = 0, tree_size = 0;
      while ((tree_size - next_node) + (next_sym + 1) > 1) {
        unsigned long long sum = 0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          bool take_leaf = false;
          if (next_sym >= 0) {
            if (next_node == tree_size) take_leaf = true;
            else take_leaf = (sc->keys[next_sym] >> 8) <= sc->tree_count[next_node];  // :375
          }
          if (take_leaf) {
            sum += sc->keys[next_sym] >> 8;
            sc->leaf_parent[next_sym] = (uint16_t)tree_size;
            --next_sym;
          } else {
            sum += sc->tree_count[next_node];
            sc->node_parent[next_node] = (uint16_t)tree_size;
            ++next_node;
          }
        }
        sc->tree_count[tree_size] = sum;
        ++tree_size;
      }
      if (tree_size == 0) {
        sc->len_count33[0] = 1;  // one symbol: the root is a leaf at depth 0 (:417-418)
      } else {
        sc->node_depth[tree_size - 1] = 0;
        for (int node = tree_size - 2; node >= 0; --node)
          sc->node_depth[node] = sc->node_depth[sc->node_parent[node]] + 1;
      }
    }
    __syncwarp();
    // 4. depth histogram over the leaves (CollectCodeLen, :329-337)
    if (n > 1) {
      for (int i = lane; i < n; i += 32) {
        int d = sc->node_depth[sc->leaf_parent[i]] + 1;
        atomicAdd(&sc->len_count33[d > 32 ? 32 : d], 1u);
      }
    }
    __syncwarp();
    // 5. LimitCodeLengths (:297-327), serial and tiny
    if (lane == 0) {
      uint32_t* lc = sc->len_count33;
      for (int i = kMaxCodeLen + 1; i <= 32; ++i) {
        lc[kMaxCodeLen] += lc[i];
        lc[i] = 0;
      }
      uint32_t kraft = 0;
      for (int i = 0; i <= kMaxCodeLen; ++i) kraft += lc[i] << (kMaxCodeLen - i);
      const uint32_t one = 1u << kMaxCodeLen;
      while (kraft > one) {
        --lc[kMaxCodeLen];
        for (int j = kMaxCodeLen - 1; j >= 0; --j) {
          if (lc[j] > 0) {
            --lc[j];
            lc[j + 1] += 2;
            break;
          }
        }
        --kraft;
      }
      // prefix tables for the canonical assignment
      uint32_t cum = 0, code = 0, mask = 0;
      for (int l = 0; l <= kMaxCodeLen; ++l) {
        sc->start_code[l] = code;
        code += lc[l] << (kMaxCodeLen - l);
        cum += lc[l];
        sc->cum[l] = cum;
        if (lc[l]) mask |= 1u << l;
        tab->len_count[l] = (uint16_t)lc[l];
      }
      tab->len_mask = mask;
      tab->hdr_len = 8u + (uint32_t)__popc(mask) + (uint32_t)n;
    }
    __syncwarp();
    // 6. canonical codes (ForallCodes, :260-284), one symbol per lane
    for (int i = lane; i < n; i += 32) {
      const unsigned sym = (unsigned)(sc->keys[i] & 0xffu);
      tab->sorted_syms[i] = (uint8_t)sym;
      int l = 0;
      while (l < kMaxCodeLen && (uint32_t)i >= sc->cum[l]) ++l;
      const uint32_t first_idx = l ? sc->cum[l - 1] : 0u;
      const uint32_t left = sc->start_code[l] + (((uint32_t)i - first_idx) << (kMaxCodeLen - l));
      tab->enc[sym] = (left >> (kMaxCodeLen - l)) | ((uint32_t)l << 16);
    }
  } else if (lane == 0) {
    for (int l = 0; l < 16; ++l) tab->len_count[l] = 0;
    tab->len_mask = 0;
    tab->hdr_len = 8;
  }
  if (lane == 0) {
    tab->num_syms = n;
    tab->pad_ = 0;
  }
  __syncwarp();
}

// Builds enc[] etc. from a caller-supplied (len_count, sorted_syms) pair; one warp.
__device__ inline void table_from_lengths_warp(const uint16_t* len_count, const uint8_t* syms,
                                               int n, HufTable* tab, TableScratch* sc) {
  const int lane = lane_id();
  for (int i = lane; i < 256; i += 32) tab->enc[i] = kEncInvalid;
  if (lane == 0) {
    uint32_t cum = 0, code = 0, mask = 0;
    for (int l = 0; l <= kMaxCodeLen; ++l) {
      sc->start_code[l] = code;
      code += (uint32_t)len_count[l] << (kMaxCodeLen - l);
      cum += len_count[l];
      sc->cum[l] = cum;
      if (len_count[l]) mask |= 1u << l;
      tab->len_count[l] = len_count[l];
    }
    tab->len_mask = mask;
    tab->num_syms = n;
    tab->hdr_len = 8u + (uint32_t)__popc(mask) + (uint32_t)n;
    tab->pad_ = 0;
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const unsigned sym = syms[i];
    tab->sorted_syms[i] = (uint8_t)sym;
    int l = 0;
    while (l < kMaxCodeLen && (uint32_t)i >= sc->cum[l]) ++l;
    const uint32_t first_idx = l ? sc->cum[l - 1] : 0u;
    const uint32_t left = sc->start_code[l] + (((uint32_t)i - first_idx) << (kMaxCodeLen - l));
    tab->enc[sym] = (left >> (kMaxCodeLen - l)) | ((uint32_t)l << 16);
  }
  __syncwarp();
}

}  // namespace hufb200
