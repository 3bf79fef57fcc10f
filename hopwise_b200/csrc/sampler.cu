// KG / rec negative sampler on the GPU, bit-exact with hopwise's NumPy path.
//
// Replaces (paths under /root/reference/hopwise/):
//   sampler/sampler.py:140-183   AbstractSampler.sample_by_key_ids (rejection rounds)
//   sampler/sampler.py:315-316   KGSampler._uni_sampling = np.random.randint(1, entity_num, n)
//   sampler/sampler.py:226-227   Sampler._uni_sampling   = np.random.randint(1, item_num, n)
//   sampler/sampler.py:321-336   used_ids[head] = set(tails of head)   (here: sorted CSR)
// and numpy's legacy global MT19937 stream underneath (third party; see oracle/mt19937.py for the
// restated algorithm): np.random.randint(low, high, n) = masked rejection over tempered 32-bit
// words: mask = 2^ceil(log2(rng+1)) - 1, keep (w & mask) when it is <= rng = high-1-low.
//
// Stream order is inherently sequential (word i of round r sits behind every word of earlier
// rounds and the next 624-word block depends on the previous one), so one CTA walks the stream
// and everything inside a block of 624 words is data-parallel:
//   twist in three dependent phases (words [0,227), [227,454), [454,624)), temper + mask + accept
//   flag per word, block-wide exclusive scan of the flags, scatter the j-th accepted value to
//   the j-th open slot; then every drawn slot is checked against its key's sorted forbidden
//   list (binary search) and the failing slots, compacted in ascending order like the reference's
//   list comprehension, form the next round.
// Popularity-biased candidates (sampler.py:68-116, AbstractSampler._build_alias_table / _pop_sampling): a round
// of L open slots consumes randint(0, n_keys, L) -- the same masked rejection -- and then np.random.random(L):
// two words per double, (w0 >> 5) * 2^26 + (w1 >> 6) over 2^53, a pair may straddle a 624-word block; slot j
// takes keys[idx_j] when prob[idx_j] > p_j, else alias[idx_j].  The table itself is built on the host with the
// reference's arithmetic (hopwise_b200/sampler.py build_alias_table).
// The advanced state (624 words + pos) is written back so a later call -- or NumPy on the host
// after a copy -- continues the same stream.
#include "common.cuh"

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr int SAMPLER_THREADS = 1024;

__device__ __forceinline__ uint32_t mt_mix(uint32_t u, uint32_t v, uint32_t far) {
  const uint32_t y = (u & 0x80000000u) | (v & 0x7FFFFFFFu);
  return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9D2C5680u;
  y ^= (y << 15) & 0xEFC60000u;
  y ^= y >> 18;
  return y;
}

// In-place generation of the next 624 words (all threads of the CTA call this).
__device__ __forceinline__ void mt_twist(uint32_t* mt) {
  const int t = threadIdx.x;
  uint32_t nv = 0;
  // phase 1: new[0,227) from old[0,228) and old[397,624)
  if (t < 227) nv = mt_mix(mt[t], mt[t + 1], mt[t + MT_M]);
  __syncthreads();
  if (t < 227) mt[t] = nv;
  __syncthreads();
  // phase 2: new[227,454) from old[227,455) and new[0,227)
  if (t < 227) nv = mt_mix(mt[227 + t], mt[228 + t], mt[t]);
  __syncthreads();
  if (t < 227) mt[227 + t] = nv;
  __syncthreads();
  // phase 3: new[454,623) from old[454,624) and new[227,396); word 623 needs new[0] and new[396]
  if (t < 169) nv = mt_mix(mt[454 + t], mt[455 + t], mt[227 + t]);
  else if (t == 169) nv = mt_mix(mt[623], mt[0], mt[396]);
  __syncthreads();
  if (t < 170) mt[454 + t] = nv;
  __syncthreads();
}

// Exclusive scan of one int per thread over the CTA; returns the prefix, *total gets the sum.
__device__ __forceinline__ int block_exclusive_scan(int x, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];  // SAMPLER_THREADS / 32 == 32 warps
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += y;
    }
    warp_sums[lane] = wi - w;  // exclusive prefix of the warp sums
    if (lane == 31) warp_sums[32] = wi;
  }
  __syncthreads();
  const int prefix = warp_sums[warp] + incl - x;
  *total = warp_sums[32];
  __syncthreads();  // warp_sums may be reused right away
  return prefix;
}

struct SampleArgs {
  uint32_t* state;  // [626]: 624 key words, pos, status
  const int64_t* keys;
  int64_t n;
  int num;
  const int64_t* used_off;
  const int64_t* used_vals;
  int64_t low;
  uint32_t rng, mask;
  int64_t* out;
  int32_t* check_a;
  int32_t* check_b;
  int max_rounds;
  // popularity mode (pop_n > 0): alias table over pop_n keys, scratch for one round's indices and raw words
  int64_t pop_n;
  const int64_t* pop_keys;
  const double* pop_prob;
  const int64_t* pop_alias;
  int32_t* idx_tmp;    // [total]
  uint32_t* word_tmp;  // [2 * total]
};

__global__ void __launch_bounds__(SAMPLER_THREADS, 1) sample_kernel(const SampleArgs a) {
  __shared__ uint32_t mt[MT_N];
  __shared__ int warp_sums[33];
  __shared__ int s_pos, s_filled, s_len, s_next_len, s_newpos;
  const int t = threadIdx.x;
  const int64_t total = a.n * a.num;

  for (int i = t; i < MT_N; i += blockDim.x) mt[i] = a.state[i];
  if (t == 0) {
    s_pos = (int)a.state[MT_N];
    s_len = (int)total;
  }
  __syncthreads();

  int32_t* check = a.check_a;
  int32_t* check_next = a.check_b;
  bool identity = true;  // round 0: the open slots are 0..total-1
  int rounds = 0;

  while (true) {
    const int L = s_len;
    if (L == 0) break;
    if (++rounds > a.max_rounds) break;  // a key whose forbidden list covers the whole range
    // ---- draw L accepted values, in stream order, into the open slots -----------------------
    const bool pop = a.pop_n > 0;
    if (t == 0) s_filled = 0;
    __syncthreads();
    while (!(pop && a.rng == 0u)) {  // randint(0, 1, L) is all zeros and consumes nothing
      const int filled = s_filled;
      if (filled >= L) break;
      if (s_pos >= MT_N) {
        __syncthreads();
        mt_twist(mt);
        if (t == 0) s_pos = 0;
        __syncthreads();
      }
      const int pos = s_pos;
      const int i = pos + t;
      uint32_t w = 0;
      int ok = 0;
      if (i < MT_N) {
        w = mt_temper(mt[i]) & a.mask;
        ok = w <= a.rng;
      }
      int tot;
      const int rank = block_exclusive_scan(ok, warp_sums, &tot);
      const int need = L - filled;
      if (t == 0) s_newpos = MT_N;
      __syncthreads();
      if (ok && rank < need) {
        if (pop) {
          a.idx_tmp[filled + rank] = (int32_t)w;  // the j-th index of this round, j = position in the open list
        } else {
          const int slot = identity ? (filled + rank) : check[filled + rank];
          a.out[slot] = a.low + (int64_t)w;
        }
        if (rank == need - 1) s_newpos = i + 1;  // the word that produced the last needed value
      }
      __syncthreads();
      if (t == 0) {
        s_pos = s_newpos;  // MT_N when the whole rest of the block was consumed
        s_filled = filled + (tot < need ? tot : need);
      }
      __syncthreads();
    }
    if (pop) {
      // ---- np.random.random(L): 2L raw words in stream order, then one double per slot ------------------
      const int want = 2 * L;
      int done = 0;
      while (done < want) {
        if (s_pos >= MT_N) {
          __syncthreads();
          mt_twist(mt);
          if (t == 0) s_pos = 0;
          __syncthreads();
        }
        const int pos = s_pos;
        const int take = min(want - done, MT_N - pos);
        if (t < take) a.word_tmp[done + t] = mt_temper(mt[pos + t]);
        __syncthreads();
        if (t == 0) s_pos = pos + take;
        done += take;
        __syncthreads();
      }
      for (int j = t; j < L; j += blockDim.x) {
        const int slot = identity ? j : check[j];
        const int32_t idx = a.rng == 0u ? 0 : a.idx_tmp[j];
        const double hi = (double)(a.word_tmp[2 * j] >> 5), lo = (double)(a.word_tmp[2 * j + 1] >> 6);
        const double p = (hi * 67108864.0 + lo) / 9007199254740992.0;
        a.out[slot] = a.pop_prob[idx] > p ? a.pop_keys[idx] : a.pop_alias[idx];
      }
      __syncthreads();
    }
    // ---- membership test + order-preserving compaction of the failing slots -------------------
    if (t == 0) s_next_len = 0;
    __syncthreads();
    for (int base = 0; base < L; base += blockDim.x) {
      const int j = base + t;
      int bad = 0, slot = 0;
      if (j < L) {
        slot = identity ? j : check[j];
        const int64_t v = a.out[slot];
        const int64_t key = a.keys[slot % a.n];
        int64_t lo = a.used_off[key], hi = a.used_off[key + 1];
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          const int64_t x = a.used_vals[mid];
          if (x == v) { bad = 1; break; }
          if (x < v) lo = mid + 1; else hi = mid;
        }
      }
      int tot;
      const int rank = block_exclusive_scan(bad, warp_sums, &tot);
      const int nb = s_next_len;
      if (bad) check_next[nb + rank] = slot;
      __syncthreads();
      if (t == 0) s_next_len = nb + tot;
      __syncthreads();
    }
    if (t == 0) s_len = s_next_len;
    __syncthreads();
    int32_t* tmp = check;
    check = check_next;
    check_next = tmp;
    identity = false;
    __threadfence_block();
  }

  for (int i = t; i < MT_N; i += blockDim.x) a.state[i] = mt[i];
  if (t == 0) {
    a.state[MT_N] = (uint32_t)s_pos;
    if (s_len != 0) a.state[MT_N + 1] = 1u;  // sticky: the caller raises (sampler.py:318-336 raises up front)
  }
}

// np.random.seed(seed): init_genrand, pos = 624.  Sequential by construction (624 steps).
__global__ void mt_seed_kernel(uint32_t* state, uint32_t seed) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t prev = seed;
  state[0] = prev;
  for (int i = 1; i < MT_N; ++i) {
    prev = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)i;
    state[i] = prev;
  }
  state[MT_N] = MT_N;
  state[MT_N + 1] = 0u;
}

}  // namespace

extern "C" int64_t kge_sample_workspace_bytes(int64_t total) { return total < 0 ? -1 : 2 * total * (int64_t)sizeof(int32_t); }

extern "C" int64_t kge_sample_alias_workspace_bytes(int64_t total) { return total < 0 ? -1 : 5 * total * (int64_t)sizeof(int32_t); }

namespace {

int launch_sampler(uint32_t* mt_state, const int64_t* keys, int64_t n, int32_t num, const int64_t* used_off,
                   const int64_t* used_vals, int64_t low, int64_t high, int64_t pop_n, const int64_t* pop_keys,
                   const double* pop_prob, const int64_t* pop_alias, int64_t* out, void* workspace,
                   kge_stream_t stream) {
  KGE_REQUIRE(n >= 0 && num >= 1, KGE_E_ARG, "bad n / num");
  if (n == 0) return 0;
  KGE_REQUIRE(mt_state && keys && used_off && used_vals && out && workspace, KGE_E_ARG, "NULL argument");
  KGE_REQUIRE(n * (int64_t)num < 0x3FFFFFFFll, KGE_E_UNSUPPORTED, "more than 2^30-1 samples in one call");
  const int64_t rng = high - 1 - low;
  // randint(low, low + 1) is the constant low and consumes no words: only the alias mode (one key) can ask for it
  KGE_REQUIRE(rng >= (pop_n > 0 ? 0 : 1) && rng < 0xFFFFFFFFll, KGE_E_UNSUPPORTED,
              "randint range %lld outside the 32-bit masked path", (long long)rng);
  uint32_t mask = (uint32_t)rng;
  mask |= mask >> 1;
  mask |= mask >> 2;
  mask |= mask >> 4;
  mask |= mask >> 8;
  mask |= mask >> 16;
  const int64_t total = n * (int64_t)num;
  SampleArgs a;
  a.state = mt_state;
  a.keys = keys;
  a.n = n;
  a.num = num;
  a.used_off = used_off;
  a.used_vals = used_vals;
  a.low = low;
  a.rng = (uint32_t)rng;
  a.mask = mask;
  a.out = out;
  a.check_a = reinterpret_cast<int32_t*>(workspace);
  a.check_b = a.check_a + total;
  a.max_rounds = 4096;
  a.pop_n = pop_n;
  a.pop_keys = pop_keys;
  a.pop_prob = pop_prob;
  a.pop_alias = pop_alias;
  a.idx_tmp = pop_n > 0 ? a.check_b + total : nullptr;
  a.word_tmp = pop_n > 0 ? reinterpret_cast<uint32_t*>(a.idx_tmp + total) : nullptr;
  sample_kernel<<<1, SAMPLER_THREADS, 0, (cudaStream_t)stream>>>(a);
  KGE_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int kge_sample_negatives(uint32_t* mt_state, const int64_t* keys, int64_t n, int32_t num,
                                    const int64_t* used_off, const int64_t* used_vals, int64_t low, int64_t high,
                                    int64_t* out, void* workspace, kge_stream_t stream) {
  return launch_sampler(mt_state, keys, n, num, used_off, used_vals, low, high, 0, nullptr, nullptr, nullptr, out,
                        workspace, stream);
}

extern "C" int kge_sample_negatives_alias(uint32_t* mt_state, const int64_t* keys, int64_t n, int32_t num,
                                          const int64_t* used_off, const int64_t* used_vals, int64_t pop_n,
                                          const int64_t* pop_keys, const double* pop_prob, const int64_t* pop_alias,
                                          int64_t* out, void* workspace, kge_stream_t stream) {
  KGE_REQUIRE(pop_n >= 1 && pop_keys && pop_prob && pop_alias, KGE_E_ARG, "empty alias table");
  return launch_sampler(mt_state, keys, n, num, used_off, used_vals, 0, pop_n, pop_n, pop_keys, pop_prob, pop_alias,
                        out, workspace, stream);
}

extern "C" int kge_mt19937_seed(uint32_t* mt_state, uint32_t seed, kge_stream_t stream) {
  KGE_REQUIRE(mt_state, KGE_E_ARG, "NULL state");
  mt_seed_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(mt_state, seed);
  KGE_LAUNCH_CHECK();
  return 0;
}
