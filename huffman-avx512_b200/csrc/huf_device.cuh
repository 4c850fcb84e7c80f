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
constexpr uint32_t kEncInvalid = 0x10000000u;  // enc[] entry of a symbol without a code
constexpr uint32_t kEnc2Invalid = 1u << 12;    // enc2[] entry of such a symbol: no bits, a marker that survives the entry sums

// Encode-side table.  Lives in shared memory (per-block tables) or in global
// memory (shared-table mode, built by k_build_table).
#ifndef HUF_MERGE_SIMPLE
#define HUF_MERGE_SIMPLE 1
#endif
struct HufTable {
  uint32_t enc[256];         // code right-aligned in bits 0..11, length in bits 16..19
  uint32_t enc2[256];        // the same codes for the staged encoder: code << (32 - len) | len (code top-aligned,
                             // length in bits 0..3, bits 4..19 zero); kEnc2Invalid for a symbol without a code.
                             // Must follow enc directly (the encoder reads it at enc + 1 KiB).
  uint8_t sorted_syms[256];  // canonical order (CanonicalCoding::sorted_syms, :288)
  uint16_t len_count[16];    // [0..12] used (CanonicalCoding::len_count, :290)
  uint32_t len_mask;         // bit i set <=> len_count[i] != 0 (:422-426)
  int32_t num_syms;
  uint32_t hdr_len;          // 8 + popcount(len_mask) + num_syms
  uint32_t avg_bits_x256;    // mean code length over the histogram the table was built from, 8.8 fixed point (0 = unknown)
};

// Scratch for the table build (shared memory, one per CTA).
struct TableScratch {
  unsigned long long keys[256];        // (count << 8) | symbol; u32 view when every count < 2^24
  unsigned long long tree_count[256];  // sorted keys (u64 mode) / internal-node weights (:365)
  uint16_t leaf_parent[256];
  uint16_t par[2][256];                // pointer-jumping ping-pong: ancestor link
  uint16_t dep[2][256];                // pointer-jumping ping-pong: distance to that ancestor
  int32_t st_first[20], st_last[20], st_depth[20];  // explicit introsort stack
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

// Inclusive warp scan.  shfl.up reports through its predicate whether the source lane exists, so a
// step is two instructions (shuffle, predicated add) and needs no lane-id compares.
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 t;\n\t"
        "shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t"
        "@p add.u32 %0, %0, t;\n\t"
        "}"
        : "+r"(v)
        : "r"(d));
  }
  return v;
}

// ---------------------------------------------------------------------------
// Exact emulation of libstdc++'s std::sort (GCC 13 bits/stl_algo.h, bits/stl_heap.h).
// The reference sorts the present symbols with a comparator that ignores the symbol
// value (codec/huffman.cpp:353-354), so the order of equal-count symbols -- and with it
// sorted_syms, every code and every compressed byte -- is whatever introsort leaves
// (SURVEY.md H1).  std::sort = __introsort_loop (partitions down to ranges of <= 16,
// heapsort when the depth limit 2*floor(log2 n) is exhausted) followed by
// __final_insertion_sort.  The insertion sort never moves an element past an equal one,
// i.e. it is a STABLE sort of whatever the partition phase left.  So only the partition
// phase (inherently serial, ~n log(n/16) steps, skipped for n <= 16) is run by one
// thread; the final pass is replaced by a parallel stable rank computation.
// less(a, b) <=> count(a) > count(b), count = key >> 8.
// ---------------------------------------------------------------------------
namespace sortclone {
template <typename K> __device__ __forceinline__ bool less(K a, K b) { return (a >> 8) > (b >> 8); }
template <typename K> __device__ __forceinline__ void swp(K* a, int i, int j) {
  K t = a[i];
  a[i] = a[j];
  a[j] = t;
}

template <typename K>
__device__ inline void push_heap(K* a, int first, int hole, int top, K value) {
  int parent = (hole - 1) / 2;
  while (hole > top && less(a[first + parent], value)) {
    a[first + hole] = a[first + parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  a[first + hole] = value;
}

template <typename K>
__device__ inline void adjust_heap(K* a, int first, int hole, int len, K value) {
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
template <typename K>
__device__ inline void heap_sort(K* a, int first, int last) {
  int len = last - first;
  if (len >= 2) {
    int parent = (len - 2) / 2;
    for (;;) {
      K v = a[first + parent];
      adjust_heap(a, first, parent, len, v);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {
    --last;
    K v = a[last];
    a[last] = a[first];
    adjust_heap(a, first, 0, last - first, v);
  }
}

// std::__introsort_loop on a[0..n), n > 16 (the recursion on the right part is an explicit stack;
// the two parts are disjoint so their processing order does not matter), run by the whole warp.
// What one __unguarded_partition call does to its range is
// fixed by two lists taken from the range as it is before the call: A = positions (ascending)
// whose key does not sort before the pivot -- where the left scan stops -- and B = positions
// (descending) whose key does not sort after it -- where the right scan stops.  The i-th
// iteration swaps a[A_i] and a[B_i] as long as A_i < B_i (a swapped element lands behind the
// scan pointers, so the lists never need updating), and the left scan that ends the loop stops
// at the next A position or at B_s, which holds an A-type element after the last swap.  So the
// lists are built with ballots, all swaps happen at once, and the cut is min(A_{s+1}, B_s).
// The heapsort fallback stays serial (lane 0).  listA / listB: 256 entries each.
template <typename K>
__device__ inline void partition_phase_warp(K* a, int n, int32_t* st_first, int32_t* st_last, int32_t* st_depth,
                                            uint16_t* listA, uint16_t* listB) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  int sp = 0;
  int first = 0, last = n, depth = 2 * (31 - __clz(n));
  for (;;) {
    while (last - first > 16) {
      if (depth == 0) {
        if (lane == 0) heap_sort(a, first, last);
        __syncwarp();
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      K pivot;
      {  // __move_median_to_first(first, first+1, mid, last-1); every lane reads the same words
        const K ka = a[first + 1], kb = a[mid], kc = a[last - 1];
        int pick;
        if (less(ka, kb)) {
          if (less(kb, kc)) pick = mid;
          else if (less(ka, kc)) pick = last - 1;
          else pick = first + 1;
        } else if (less(ka, kc)) pick = first + 1;
        else if (less(kb, kc)) pick = last - 1;
        else pick = mid;
        pivot = a[pick];
        const K old_first = a[first];
        __syncwarp();
        if (lane == 0) {
          a[pick] = old_first;
          a[first] = pivot;
        }
        __syncwarp();
      }
      // the two stop lists over [first+1, last)
      int nA = 0, nB = 0;
      for (int base = first + 1; base < last; base += 32) {
        const int p = base + lane;
        const bool f = p < last && !less(a[p < last ? p : first], pivot);
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (f) listA[nA + __popc(m & lt)] = (uint16_t)p;
        nA += __popc(m);
      }
      for (int top = last - 1; top > first; top -= 32) {
        const int p = top - lane;
        const bool f = p > first && !less(pivot, a[p > first ? p : first]);
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (f) listB[nB + __popc(m & lt)] = (uint16_t)p;
        nB += __popc(m);
      }
      __syncwarp();
      // s = number of swaps; A ascends and B descends, so the valid pairs are a prefix
      const int nm = nA < nB ? nA : nB;
      int s = 0;
      for (int i0 = 0; i0 < nm; i0 += 32) {
        const int i = i0 + lane;
        const unsigned m = __ballot_sync(0xffffffffu, i < nm && listA[i < nm ? i : 0] < listB[i < nm ? i : 0]);
        s += __popc(m);
        if (m != 0xffffffffu) break;
      }
      for (int i = lane; i < s; i += 32) {
        const int pa = listA[i], pb = listB[i];
        const K t = a[pa];
        a[pa] = a[pb];
        a[pb] = t;
      }
      int cut;
      if (s == 0) cut = listA[0];  // the median move guarantees nA >= 1 and nB >= 1
      else {
        cut = listB[s - 1];
        if (s < nA && (int)listA[s] < cut) cut = listA[s];
      }
      __syncwarp();
      st_first[sp] = cut;  // right part [cut, last) for later, left part [first, cut) now (same value from every lane)
      st_last[sp] = last;
      st_depth[sp] = depth;
      ++sp;
      last = cut;
    }
    if (sp == 0) break;
    --sp;
    __syncwarp();
    first = st_first[sp];
    last = st_last[sp];
    depth = st_depth[sp];
  }
  __syncwarp();
}
}  // namespace sortclone

// ---------------------------------------------------------------------------
// Table build, executed by ONE WARP (all 32 lanes must call it).
//   hist  : 256 counts (shared memory), CountT = uint32_t or uint64_t
//   KeyT  : uint32_t when every count < 2^24, else unsigned long long
//   tab   : output (shared memory)
// Restates MakeCanonicalCoding (codec/huffman.cpp:339-437): present symbols ->
// std::sort by count descending -> two-queue Huffman merge (leaf preferred on
// ties, :375) -> depth histogram -> LimitCodeLengths (:297-327) -> canonical
// codes (ForallCodes, :260-284).
// ---------------------------------------------------------------------------
template <typename CountT, typename KeyT>
__device__ __noinline__ void build_table_warp(const CountT* hist, HufTable* tab, TableScratch* sc) {
  const int lane = lane_id();
  // u32 keys: keys[] in the first, the sorted copy in the second half of sc->keys, node
  // weights in sc->tree_count.  u64 keys: keys[] = sc->keys, sorted copy = sc->tree_count,
  // node weights overwrite sc->keys (free once the sorted copy exists).
  KeyT* keys = reinterpret_cast<KeyT*>(sc->keys);
  KeyT* sorted = sizeof(KeyT) == 8 ? reinterpret_cast<KeyT*>(sc->tree_count) : keys + 256;
  CountT* tree = sizeof(KeyT) == 8 ? reinterpret_cast<CountT*>(sc->keys) : reinterpret_cast<CountT*>(sc->tree_count);
  // 1. present symbols in ascending symbol order (:342-347)
  int n = 0;
#pragma unroll 1
  for (int base = 0; base < 256; base += 32) {
    const int c = base + lane;
    const CountT cnt = hist[c];
    const unsigned m = __ballot_sync(0xffffffffu, cnt != 0);
    if (cnt != 0) keys[n + __popc(m & ((1u << lane) - 1))] = ((KeyT)cnt << 8) | (KeyT)c;
    n += __popc(m);
  }
  for (int i = lane; i < 256; i += 32) {
    tab->enc[i] = kEncInvalid;
    tab->enc2[i] = kEnc2Invalid;
  }
  for (int i = lane; i < 36; i += 32) sc->len_count33[i] = 0;
  __syncwarp();

  if (n > 0) {
    // 2a. introsort partition phase (only for n > 16); it leaves runs of at most 16 elements (or
    //     heap-sorted ranges) that are in order among each other
    if (n > 16) sortclone::partition_phase_warp(keys, n, sc->st_first, sc->st_last, sc->st_depth, sc->par[0], sc->par[1]);
    // 2b. final insertion sort == stable sort of what 2a left: rank = #greater + #equal-before.
    //     No element moves out of its run, so only the 16 neighbours on either side can change
    //     places with it; everything further left ranks before it, everything further right after.
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      const KeyT mine = i < n ? keys[i] : 0;
      const KeyT mc = mine >> 8;
      const int w0 = i - 16;
      int rank = w0 > 0 ? w0 : 0;
      // straight-line: 33 independent loads (index clamped into the array, the in-range test
      // decides whether the comparison counts)
#pragma unroll
      for (int t = 0; t < 33; ++t) {
        const int j = w0 + t;
        const int jc = j < 0 ? 0 : (j < n ? j : n - 1);
        const KeyT oc = keys[jc] >> 8;
        const bool before = (oc > mc) || (oc == mc && t < 16);  // t < 16 <=> j < i
        rank += ((unsigned)j < (unsigned)n && before) ? 1 : 0;
      }
      if (i < n) sorted[rank] = mine;
    }
    __syncwarp();
    // 3. two-queue Huffman merge with the reference's tie rule (a leaf is preferred, :370-376):
    //    inherently serial, one lane.  Whether a queue head exists is decided by the queue
    //    indices, never by its value (u32 weights may wrap like the reference's, :365, :414).
    //    Two forms of the loop (HUF_MERGE_SIMPLE): heads in registers and reloaded after the pick
    //    that used them (the default: fewest instructions), or heads AND their successors in
    //    registers with a branch-free body (loads off the dependent chain, three times the
    //    instructions).
    const int n_nodes = n - 1;
    // Many leaves of similar weight (incompressible input: 256 symbols, all near n/256) make the
    // merge batchable without changing a single decision: while a node t is waiting, every leaf
    // not heavier than t is taken before t (the rule compares the leaf with the queue's head
    // only, and new nodes queue up behind it), so those leaves pair up among themselves, all
    // pairs at once; and once the leaves are gone the waiting nodes pair up in queue order.
    // Skewed inputs give batches of one, for which the serial loop below is quicker.
    const bool batched = n >= 32 && (CountT)(sorted[n / 2] >> 8) <= 4 * (CountT)(sorted[n - 1] >> 8);
    if (batched) {
      auto W = [&](int i) { return (CountT)(sorted[n - 1 - i] >> 8); };  // leaves in ascending order
      int q = 0, nh = 0, m = 0;  // next leaf, head of the node queue, nodes made (all warp-uniform)
      while (m < n_nodes) {
        if (nh == m) {  // no node waiting: two leaves
          if (lane == 0) {
            tree[m] = W(q) + W(q + 1);
            sc->leaf_parent[n - 1 - q] = (uint16_t)m;
            sc->leaf_parent[n - 2 - q] = (uint16_t)m;
          }
          q += 2;
          m += 1;
        } else if (q == n) {  // leaves used up: the waiting nodes pair up in order
          int p = (m - nh) >> 1;
          p = p < 32 ? p : 32;
          if (lane < p) {
            tree[m + lane] = tree[nh + 2 * lane] + tree[nh + 2 * lane + 1];
            sc->par[0][nh + 2 * lane] = (uint16_t)(m + lane);
            sc->par[0][nh + 2 * lane + 1] = (uint16_t)(m + lane);
          }
          nh += 2 * p;
          m += p;
        } else {
          const CountT t = tree[nh];
          const int idx = q + lane;
          const int c = __popc(__ballot_sync(0xffffffffu, idx < n && W(idx < n ? idx : 0) <= t));  // a prefix: the leaves ascend
          if (c >= 2) {  // c / 2 leaf pairs at once
            const int p = c >> 1;
            if (lane < p) {
              tree[m + lane] = W(q + 2 * lane) + W(q + 2 * lane + 1);
              sc->leaf_parent[n - 1 - (q + 2 * lane)] = (uint16_t)(m + lane);
              sc->leaf_parent[n - 2 - (q + 2 * lane)] = (uint16_t)(m + lane);
            }
            q += 2 * p;
            m += p;
          } else if (c == 1) {  // the last such leaf, then t itself
            if (lane == 0) {
              tree[m] = W(q) + t;
              sc->leaf_parent[n - 1 - q] = (uint16_t)m;
              sc->par[0][nh] = (uint16_t)m;
            }
            q += 1;
            nh += 1;
            m += 1;
          } else {  // t first, then whatever the rule picks (:370-376)
            const int nh1 = nh + 1;
            const bool leaf2 = nh1 == m || W(q) <= tree[nh1 < n_nodes ? nh1 : 0];
            if (lane == 0) {
              sc->par[0][nh] = (uint16_t)m;
              if (leaf2) sc->leaf_parent[n - 1 - q] = (uint16_t)m;
              else sc->par[0][nh1] = (uint16_t)m;
              tree[m] = t + (leaf2 ? W(q) : tree[nh1]);
            }
            q += leaf2 ? 1 : 0;
            nh += leaf2 ? 1 : 2;
            m += 1;
          }
        }
        __syncwarp();
      }
      if (lane == 0) sc->par[0][n_nodes - 1] = (uint16_t)(n_nodes - 1);  // the root points at itself
#if HUF_MERGE_SIMPLE
    } else if (lane == 0 && n > 1) {
      // the plain form: queue heads in two registers, reloaded after the pick that used them.  A
      // third of the instructions of the look-ahead form below at about the same latency per step
      // (two dependent shared-memory loads instead of a long select chain); what counts once
      // the build shares its SM's issue slots with thirty-two worker warps.
      int ls = n - 1, nh = 0;
      CountT L = (CountT)(sorted[ls] >> 8), N = 0;
      for (int m = 0; m < n_nodes; ++m) {
        CountT sum = 0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const bool leaf = (ls >= 0) && (nh == m || L <= N);  // nodes nh..m-1 are waiting; a leaf wins a tie (:375)
          if (leaf) {
            sum += L;
            sc->leaf_parent[ls] = (uint16_t)m;
            --ls;
            L = (CountT)(sorted[ls > 0 ? ls : 0] >> 8);
          } else {
            sum += N;
            sc->par[0][nh] = (uint16_t)m;
            ++nh;
            N = tree[nh < n_nodes ? nh : n_nodes - 1];  // stale when nh == m: set below
          }
        }
        tree[m] = sum;
        if (nh == m) N = sum;  // the new node is the only one waiting
      }
      sc->par[0][n_nodes - 1] = (uint16_t)(n_nodes - 1);  // the root points at itself
#else
    } else if (lane == 0 && n > 1) {
      int ls = n - 1, nh = 0;
      CountT L0 = (CountT)(sorted[ls] >> 8);
      CountT L1 = (CountT)(sorted[ls - 1] >> 8);
      CountT N0 = 0, N1 = 0;
      for (int m = 0; m < n_nodes; ++m) {
        CountT sum = 0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int li = ls - 2, ni = nh + 2;
          const CountT l_next = (CountT)(sorted[li > 0 ? li : 0] >> 8);  // head after next, if a leaf goes
          const CountT n_next = tree[ni < n_nodes ? ni : n_nodes - 1];   // same for the nodes (stale if >= m: fixed below)
          const bool leaf = (ls >= 0) && (nh == m || L0 <= N0);
          sum += leaf ? L0 : N0;
          uint16_t* link = leaf ? &sc->leaf_parent[ls > 0 ? ls : 0] : &sc->par[0][nh];
          *link = (uint16_t)m;
          ls -= leaf ? 1 : 0;
          nh += leaf ? 0 : 1;
          L0 = leaf ? L1 : L0;
          L1 = leaf ? l_next : L1;
          N0 = leaf ? N0 : N1;
          N1 = leaf ? N1 : n_next;
        }
        tree[m] = sum;
        N0 = nh == m ? sum : N0;  // the new node is the only / the second one waiting
        N1 = nh + 1 == m ? sum : N1;
      }
      sc->par[0][n_nodes - 1] = (uint16_t)(n_nodes - 1);  // the root points at itself
#endif
    }
    __syncwarp();
    // 4. node depths by pointer jumping (8 synchronous rounds cover any depth <= 255), then
    //    the leaf depth histogram (CollectCodeLen, :329-337)
    if (n > 1) {
      for (int i = lane; i < n_nodes; i += 32) sc->dep[0][i] = (i == n_nodes - 1) ? 0 : 1;
      __syncwarp();
      int cur = 0;
#pragma unroll 1
      for (int round = 0; round < 8; ++round) {
        bool open = false;  // some link of this lane does not point at the root yet
        for (int i = lane; i < n_nodes; i += 32) {
          const int p = sc->par[cur][i];
          const int pp = sc->par[cur][p];
          sc->dep[cur ^ 1][i] = sc->dep[cur][i] + sc->dep[cur][p];
          sc->par[cur ^ 1][i] = (uint16_t)pp;
          open |= pp != n_nodes - 1;
        }
        cur ^= 1;
        __syncwarp();
        if (!__any_sync(0xffffffffu, open)) break;  // every node has its full depth
      }
      for (int i = lane; i < n; i += 32) {
        const int d = sc->dep[cur][sc->leaf_parent[i]] + 1;
        atomicAdd(&sc->len_count33[d > 32 ? 32 : d], 1u);
      }
    } else if (lane == 0) {
      sc->len_count33[0] = 1;  // one symbol: the root is a leaf at depth 0 (:417-418)
    }
    __syncwarp();
    // 5. LimitCodeLengths (:297-327): lane l holds len_count[l] (l <= 12; deeper levels folded
    //    into 12 first), the Kraft sum is a warp reduction, and each repair step -- one leaf off
    //    level 12, the deepest non-empty shallower level j gives one to j+1 twice -- is a ballot
    //    and three predicated register updates instead of a walk over shared memory
    {
      uint32_t* lc = sc->len_count33;
      uint32_t mine = lane <= 32 ? lc[lane] : 0u;  // lane 32 does not exist: levels 0..31 here, level 32 below
      uint32_t deep = (lane > kMaxCodeLen) ? mine : 0u;
      if (lane == 0) deep += lc[32];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) deep += __shfl_xor_sync(0xffffffffu, deep, d);
      if (lane == kMaxCodeLen) mine += deep;
      if (lane > kMaxCodeLen) mine = 0;
      uint32_t kraft = lane <= kMaxCodeLen ? (mine << (kMaxCodeLen - lane)) : 0u;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) kraft += __shfl_xor_sync(0xffffffffu, kraft, d);
      const uint32_t one = 1u << kMaxCodeLen;
      while (kraft > one) {
        if (lane == kMaxCodeLen) --mine;
        const unsigned nonempty = __ballot_sync(0xffffffffu, lane < kMaxCodeLen && mine > 0);
        if (nonempty) {
          const int j = 31 - __clz(nonempty);
          if (lane == j) --mine;
          if (lane == j + 1) mine += 2;
        }
        --kraft;
      }
      if (lane <= kMaxCodeLen) lc[lane] = mine;
      else if (lane < 32) lc[lane] = 0;
      if (lane == 0) lc[32] = 0;
    }
    __syncwarp();
    if (lane == 0) {
      uint32_t* lc = sc->len_count33;
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
    unsigned long long bits = 0, total = 0;  // for the mean code length
    for (int i = lane; i < n; i += 32) {
      const unsigned sym = (unsigned)(sorted[i] & 0xffu);
      tab->sorted_syms[i] = (uint8_t)sym;
      int l = 0;
      while (l < kMaxCodeLen && (uint32_t)i >= sc->cum[l]) ++l;
      const uint32_t first_idx = l ? sc->cum[l - 1] : 0u;
      const uint32_t left = sc->start_code[l] + (((uint32_t)i - first_idx) << (kMaxCodeLen - l));
      tab->enc[sym] = (left >> (kMaxCodeLen - l)) | ((uint32_t)l << 16);
      tab->enc2[sym] = (left << (32 - kMaxCodeLen)) | (uint32_t)l;
      const unsigned long long cnt = (unsigned long long)hist[sym];
      bits += cnt * (unsigned)l;
      total += cnt;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      bits += __shfl_xor_sync(0xffffffffu, bits, d);
      total += __shfl_xor_sync(0xffffffffu, total, d);
    }
    if (lane == 0) tab->avg_bits_x256 = total ? (uint32_t)((bits << 8) / total) + 1u : 0u;  // rounded up, never 0 when known
  } else if (lane == 0) {
    for (int l = 0; l < 16; ++l) tab->len_count[l] = 0;
    tab->len_mask = 0;
    tab->hdr_len = 8;
  }
  if (lane == 0) {
    tab->num_syms = n;
    if (n == 0) tab->avg_bits_x256 = 0;
  }
  __syncwarp();
}

// Builds enc[] etc. from a caller-supplied (len_count, sorted_syms) pair; one warp.
__device__ inline void table_from_lengths_warp(const uint16_t* len_count, const uint8_t* syms,
                                               int n, HufTable* tab, TableScratch* sc) {
  const int lane = lane_id();
  for (int i = lane; i < 256; i += 32) {
    tab->enc[i] = kEncInvalid;
    tab->enc2[i] = kEnc2Invalid;
  }
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
    tab->avg_bits_x256 = 0;  // no histogram here
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
    tab->enc2[sym] = (left << (32 - kMaxCodeLen)) | (uint32_t)l;
  }
  __syncwarp();
}

}  // namespace hufb200
