// tree.cuh -- device SumTree: binary tree in the reference's own flat layout (2N-1 doubles, leaf j at j+N-1, children
// 2i+1 / 2i+2; srl/rl/memories/priority_memories/proportional_memory.py:13-47), so backup()/restore() interchange with
// the reference and the descent makes the identical "<=" decisions on identical node values.  CPU twin: oracle/sumtree.py.
#pragma once
#include "philox.cuh"

namespace srlx {

// SumTree._retrieve (proportional_memory.py:57-66): walk down from the root.
// Five levels are resolved per memory round trip: the 62 descendants of the current node down to depth +5 sit in five
// contiguous runs of the BFS array (2,4,8,16,32 doubles), the warp fetches them with independent coalesced loads and
// then replays the five "val <= tree[left]" decisions out of registers with shuffles -- same comparisons on the same
// stored values as the sequential walk, but ~4 dependent L2 latencies for a 2M-leaf tree instead of 21.
__device__ inline int64_t tree_retrieve_warp(const double* __restrict__ tree, int64_t n_nodes, double val, int64_t idx = 0,
                                             double* node_val = nullptr) {
  const int lane = threadIdx.x & 31;  // (idx, val): where an earlier part of the same walk -- a cached top of the tree -- arrived
  // node_val (optional): receives tree[returned index] as fetched during the walk (the chosen child is always among the values
  // of the last round), which saves the caller a dependent load of the leaf priority; NaN if the walk took no step
  double chosen = __longlong_as_double(0x7ff8000000000000ll);
  while (true) {
    // level k (1..5) of the subtree rooted at idx occupies [(idx+1)*2^k - 1, (idx+1)*2^k - 1 + 2^k)
    // lane l loads: k=1: l<2, k=2: l<4, k=3: l<8, k=4: l<16, k=5: all 32
    double v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0;
    const int64_t b1 = (idx + 1) * 2 - 1, b2 = (idx + 1) * 4 - 1, b3 = (idx + 1) * 8 - 1, b4 = (idx + 1) * 16 - 1,
                  b5 = (idx + 1) * 32 - 1;
    if (lane < 2 && b1 + lane < n_nodes) v1 = __ldcg(tree + b1 + lane);
    if (lane < 4 && b2 + lane < n_nodes) v2 = __ldcg(tree + b2 + lane);
    if (lane < 8 && b3 + lane < n_nodes) v3 = __ldcg(tree + b3 + lane);
    if (lane < 16 && b4 + lane < n_nodes) v4 = __ldcg(tree + b4 + lane);
    if (b5 + lane < n_nodes) v5 = __ldcg(tree + b5 + lane);
    int rel = 0;  // position of the current node inside its level of the subtree
    bool leaf = false;
#pragma unroll
    for (int k = 1; k <= 5; ++k) {
      const int64_t left = 2 * idx + 1;
      if (left >= n_nodes) { leaf = true; break; }
      const double vk = (k == 1) ? v1 : (k == 2) ? v2 : (k == 3) ? v3 : (k == 4) ? v4 : v5;
      const double tl = __shfl_sync(0xffffffffu, vk, 2 * rel);
      const double tr = __shfl_sync(0xffffffffu, vk, 2 * rel + 1);
      if (val <= tl) {
        idx = left;
        rel = 2 * rel;
        chosen = tl;
      } else {
        idx = left + 1;
        val -= tl;
        rel = 2 * rel + 1;
        chosen = tr;
      }
    }
    if (leaf || 2 * idx + 1 >= n_nodes) {
      if (node_val) *node_val = chosen;
      return idx;
    }
  }
}

// single-thread walk (used by the sequential no-duplicate path and as the reference for the warp version)
__device__ inline int64_t tree_retrieve_seq(const double* __restrict__ tree, int64_t n_nodes, double val) {
  int64_t idx = 0;
  while (true) {
    const int64_t left = 2 * idx + 1;
    if (left >= n_nodes) return idx;
    const double tl = __ldcg(tree + left);
    if (val <= tl) idx = left;
    else { idx = left + 1; val -= tl; }
  }
}

// ProportionalMemory.update for a batch (proportional_memory.py:171-177), executed by one thread block.
// The reference applies the items one after the other: change = p_i - tree[leaf_i]; tree[leaf_i] = p_i; every ancestor +=
// change.  fp64 addition is not associative, so a node's value depends on the ORDER of its additions; to stay bit-identical
// the items are applied in item order -- but out of shared memory, where one item costs a few dozen cycles instead of
// depth x an L2 round trip:
//   1. every (item, level) pair hashes its node id into a shared-memory table (atomicCAS, linear probing): all pairs that
//      name the same node get the same slot; the pair that claims a slot loads the node's current value (one round trip,
//      all nodes in parallel)
//   2. ONE warp walks the items in order, lane = level above the leaf: lane 0 forms the change and overwrites the leaf,
//      the other lanes add the change to their ancestor -- the levels of one item are distinct nodes, so the lanes never
//      collide, and consecutive items see each other's results as the sequential loop does (duplicate leaves included)
//   3. every claimed slot is written back.
// Batches are cut into passes of kTreeHashChunk items (table load factor <= 1/2 for trees of up to 2^31 nodes).
//   idx[i]   tree index of item i (leaf + capacity - 1), pri[i] its final priority (already (|td|+eps)^alpha)
// Requires blockDim-wide participation (blockDim >= 64); ends with __syncthreads().
constexpr int kTreeHashSlots = 4096;
constexpr int kTreeHashChunk = 64;
constexpr uint32_t kTreeHashEmpty = 0xFFFFFFFFu;
struct TreeHashScratch {
  double vals[kTreeHashSlots];
  uint32_t keys[kTreeHashSlots];
  unsigned short slot[kTreeHashChunk * 32];  // [item][level] -> table slot, 0xFFFF above the root
};

// The three steps for one pass of m <= kTreeHashChunk items.  `prepare` needs only the tree indices, so a caller that knows them
// early (the replay CTA of learner_small.cu: the batch was sampled an update ago) runs it before the priorities exist.
__device__ inline void tree_update_prepare(const double* __restrict__ tree, const int64_t* idx, int m, TreeHashScratch* hs) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int h = tid; h < kTreeHashSlots; h += nt) hs->keys[h] = kTreeHashEmpty;
  __syncthreads();
  // claim a slot per distinct node; the claimer fetches the node
  for (int w = tid; w < m * 32; w += nt) {
    const int i = w >> 5, l = w & 31;
    const uint64_t ip1 = (uint64_t)idx[i] + 1;
    unsigned short s = 0xFFFF;
    if ((ip1 >> l) != 0) {
      const uint32_t node = (uint32_t)((ip1 >> l) - 1);
      uint32_t h = (node * 2654435761u) >> 20;  // Fibonacci hash -> 12 bits
      while (true) {
        const uint32_t old = atomicCAS(&hs->keys[h], kTreeHashEmpty, node);
        if (old == kTreeHashEmpty) { hs->vals[h] = __ldcg(tree + node); break; }
        if (old == node) break;
        h = (h + 1) & (kTreeHashSlots - 1);
      }
      s = (unsigned short)h;
    }
    hs->slot[w] = s;
  }
  __syncthreads();
}

__device__ inline void tree_update_apply(double* __restrict__ tree, const double* pri, int m, TreeHashScratch* hs, double* cache, int n_cache) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  // the items in order, one warp, lane = level
  if (tid < 32) {
    for (int i = 0; i < m; ++i) {
      const unsigned short s = hs->slot[i * 32 + lane];
      double c = 0.0;
      if (lane == 0) {
        const double p = pri[i];
        c = p - hs->vals[s];
        hs->vals[s] = p;
      }
      c = __shfl_sync(0xffffffffu, c, 0);
      if (lane > 0 && s != 0xFFFF) hs->vals[s] += c;
      __syncwarp();
    }
  }
  __syncthreads();
  // write back
  for (int h = tid; h < kTreeHashSlots; h += nt) {
    const uint32_t node = hs->keys[h];
    if (node != kTreeHashEmpty) {
      __stcg(tree + node, hs->vals[h]);
      if (node < (uint32_t)n_cache) cache[node] = hs->vals[h];  // the caller's shared-memory copy of the top levels
    }
  }
  __syncthreads();
}

__device__ inline void tree_update_batch(double* __restrict__ tree, const int64_t* idx, const double* pri, int n, TreeHashScratch* hs,
                                         double* cache = nullptr, int n_cache = 0) {
  for (int base = 0; base < n; base += kTreeHashChunk) {
    const int m = min(kTreeHashChunk, n - base);
    tree_update_prepare(tree, idx + base, m, hs);
    tree_update_apply(tree, pri + base, m, hs, cache, n_cache);
  }
}

// ProportionalMemory.sample draw for a whole batch (proportional_memory.py:142-157), executed by one thread block.
// Attempt k of sample i uses uniform u(i,k): injected (u01 != NULL, row-major [B][max_tries]) or
// Philox(seed, STREAM_SAMPLE, (i | k<<16, rng_step_lo, rng_step_hi)).  An attempt is rejected when the leaf priority is 0
// or, with has_duplicate == 0, when the leaf was already picked by an earlier sample -- evaluated in sample order, so
// the result equals the reference's sequential loop.  One warp walks one sample (tree_retrieve_warp).
// Outputs (shared memory): s_idx tree indices, s_pri leaf priorities, s_att scratch; *retries accumulates rejections.
__device__ inline void per_sample_block(const double* __restrict__ tree, int64_t n_nodes, double total, int B,
                                        uint64_t seed, uint64_t rng_step, const double* __restrict__ u01, int max_tries,
                                        int has_duplicate, int64_t* s_idx, double* s_pri, double* s_att,
                                        unsigned long long* retries, const double* cache = nullptr, int n_cache = 0) {
  // cache: optional copy of tree[0 .. n_cache) (n_cache = 2^k - 1 whole top levels) in shared memory, kept equal to the tree by
  // the caller; the walk takes its first k-1 decisions out of it (same values, same comparisons) and continues in the tree
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  auto draw = [&](int i, int k) -> double {
    if (u01) return u01[(size_t)i * max_tries + k];
    const uint4 w = philox(seed, STREAM_SAMPLE, (uint32_t)i | ((uint32_t)k << 16), (uint32_t)rng_step, (uint32_t)(rng_step >> 32));
    return u01_f64(w.x, w.y);
  };
  for (int i = warp; i < B; i += nwarps) {
    int64_t idx = 0;
    double p = 0.0;
    int k = 0;
    for (; k < max_tries; ++k) {
      double r = draw(i, k) * total;
      int64_t top = 0;
      while (2 * top + 1 < (int64_t)n_cache) {  // warp-uniform
        const double tl = cache[2 * top + 1];
        if (r <= tl) top = 2 * top + 1;
        else { r -= tl; top = 2 * top + 2; }
      }
      idx = tree_retrieve_warp(tree, n_nodes, r, top, &p);  // p = tree[idx] as read by the walk's last round
      if (p != p) p = __ldcg(tree + idx);                   // (no step taken: a one-leaf tree)
      if (p != 0.0) break;
    }
    if (lane == 0) {
      s_idx[i] = idx;
      s_pri[i] = p;
      s_att[i] = (double)k;
      if (k) atomicAdd(retries, (unsigned long long)k);
    }
  }
  __syncthreads();
  if (!has_duplicate && tid == 0) {
    for (int i = 1; i < B; ++i) {
      int k = (int)s_att[i];
      while (k < max_tries) {
        bool dup = false;
        for (int j = 0; j < i; ++j) dup |= (s_idx[j] == s_idx[i]);
        if (!dup && s_pri[i] != 0.0) break;
        ++k;
        *retries += 1;
        if (k >= max_tries) break;
        s_idx[i] = tree_retrieve_seq(tree, n_nodes, draw(i, k) * total);
        s_pri[i] = __ldcg(tree + s_idx[i]);
      }
    }
  }
  __syncthreads();
}

// IS weights (proportional_memory.py:159-167): w_i = (size * p_i / total)^-beta, divided by the batch max, as float32.
__device__ inline void per_weights_block(double total, double size, double beta, int B, const double* s_pri, double* s_tmp,
                                         float* s_w) {
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < B; i += blockDim.x) s_tmp[i] = pow(size * (s_pri[i] / total), -beta);
  __syncthreads();
  if (tid < 32) {
    double mx = 0.0;
    for (int i = lane; i < B; i += 32) mx = fmax(mx, s_tmp[i]);
    for (int s = 16; s > 0; s >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    for (int i = lane; i < B; i += 32) s_w[i] = (float)(s_tmp[i] / mx);
  }
  __syncthreads();
}

}  // namespace srlx
