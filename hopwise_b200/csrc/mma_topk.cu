// Full-sort top-k on the 5th-generation tensor cores (tcgen05 + TMEM) of sm_100a, exact in fp32.
//
// Replaces the same reference code as score.cu's fused top-k
// (model/knowledge_graph_embedding_recommender/{distmult.py:119-133, complex.py:162-190,
// transe.py:112-126, rotate.py:161-190}, trainer/trainer.py:716-735, evaluator/collector.py:176-177)
// for the scorers that are dense contractions: DistMult / ComplEx directly (score = q . t) and
// the L2 models through  -||q - t||^2 = 2 (q . t - ||t||^2 / 2) - ||q||^2  (the per-row constant
// does not change the order; -||t||^2/2 rides in three extra K columns as a bf16 hi/mid/lo split).
//
// The tensor cores only FILTER; the answer is exact:
//   1. bf16 GEMM  a^_ij ~ q_i . t_j  with fp32 accumulation in TMEM.  |a^ - a| <= eps_i =
//      1.02 * 2^-8 * ||q_i|| * max_j ||t_j||   (two bf16 roundings per product, Cauchy-Schwarz).
//   2. epilogue, one thread per query row: maxima of groups of 8 adjacent targets are compared with
//      a running threshold thr_i = tau_i - 2 eps_i, where tau_i is the k-th largest group maximum
//      among groups without masked targets seen so far (a lower bound of the k-th largest
//      unmasked approximate score).  Every member of the exact top-k has a^ >= tau_final - 2 eps,
//      so its group survives.  Surviving group ids go to a 64-entry list per row (warp-cooperative
//      compaction when it fills; a row whose list cannot be compacted is flagged).
//   3. rescore kernel: the <= 512 targets of a row's surviving groups are scored in fp32 with the
//      same sequential FMA chain as the CUDA-core kernel (bit-identical scores), masked, and the
//      top-k is selected under (score desc, id asc).  Rows flagged in 2 or with fewer than k valid
//      candidates are reported to the caller, which runs them through kge_full_sort_topk.
//
// Kernel shape: CTA = 256 query rows = two M=128 accumulators against a target tile of N=128
// (so every B tile feeds two MMAs and halves the L2 traffic per flop); TMEM holds 2 x 2
// accumulators of 128 columns, double-buffered so the epilogue of tile i overlaps the MMAs of tile
// i+1.  Warp 0 streams pre-tiled bf16 target images with cp.async.bulk into a ring of
// shared-memory stages (mbarrier expect_tx), warp 1 issues tcgen05.mma (one thread), warps 4..11
// run the epilogue from tcgen05.ld.  Operands use the no-swizzle K-major canonical layout
// [K/8][rows][8 bf16]: 8x16-byte core matrices, SBO = 128 B, LBO = rows * 16 B.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "score_common.cuh"

namespace {

constexpr int MM = 256;   // query rows per CTA
constexpr int TN = 128;   // targets per tile (UMMA N)
constexpr int GRP = 8;    // targets per candidate group
constexpr int CAND = 64;  // candidate groups kept per row
constexpr int MMA_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int IMG_HEADER = 128;  // bytes before the first tile image: {float tmax}
constexpr uint32_t SPIN_LIMIT = 1u << 26;

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// A wait that cannot hang the GPU: a protocol bug traps instead of spinning forever.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
#define KGE_TMEM_LD32_ASM(R, ADDR)                                                                                  \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                     \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                     \
      : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]),  \
        "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), "=r"(R[16]),      \
        "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]),     \
        "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]), "=r"(R[31])                   \
      : "r"(ADDR)                                                                                                  \
      : "memory")
// 64 consecutive accumulator columns of this thread's TMEM lane: two loads in flight, one wait.
__device__ __forceinline__ void tmem_ld32x2(uint32_t taddr, float (&v)[64]) {
  uint32_t r0[32], r1[32];
  KGE_TMEM_LD32_ASM(r0, taddr);
  KGE_TMEM_LD32_ASM(r1, taddr + 32u);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    v[i] = __uint_as_float(r0[i]);
    v[32 + i] = __uint_as_float(r1[i]);
  }
}
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// No-swizzle K-major shared-memory matrix descriptor (bits: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type 0 = SWIZZLE_NONE [61,64)).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, N = 128, M = 128.
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint16_t bf16_bits(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf16_back(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }

// ---- target image ---------------------------------------------------------------------------------
struct PrepArgs {
  kge_model_t m;
  int64_t n_targets;
  int parts, kp, dist;
  float* tn2;        // [n_targets] squared norms (scratch inside the image buffer's tail)
  float* header;     // {tmax}
  uint16_t* tiles;   // [n_tiles][kp/8][TN][8]
};

__global__ void __launch_bounds__(256) target_norm_kernel(const PrepArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int d = a.m.d;
  float wmax = 0.f;
  for (int64_t j = warp; j < a.n_targets; j += n_warps) {
    float s = 0.f;
    for (int p = 0; p < a.parts; ++p)
      for (int c = lane; c < d; c += 32) {
        const float x = __ldg(a.m.entity.w[p] + j * d + c);
        s = fmaf(x, x, s);
      }
    s = warp_sum(s);
    if (lane == 0) a.tn2[j] = s;
    wmax = fmaxf(wmax, s);
  }
  if (lane == 0 && wmax > 0.f) atomicMax(reinterpret_cast<int*>(a.header), __float_as_int(sqrtf(wmax) * 1.0001f));
}

__global__ void __launch_bounds__(256) target_image_kernel(const PrepArgs a) {
  const int d = a.m.d;
  const int kchunks = a.kp / 8;
  const int64_t n_tiles = (a.n_targets + TN - 1) / TN;
  const int64_t total = n_tiles * kchunks * TN;
  const int kd = a.parts * d;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx % TN);
    const int kc = (int)((idx / TN) % kchunks);
    const int64_t tile = idx / ((int64_t)TN * kchunks);
    const int64_t j = tile * TN + r;
    uint16_t out[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 8 + e;
      float x = 0.f;
      if (j < a.n_targets) {
        if (k < kd) {
          const int p = k / d, c = k - p * d;
          x = __ldg(a.m.entity.w[p] + j * d + c);
          out[e] = bf16_bits(x);
          continue;
        }
        if (a.dist && k < kd + 3) {  // bf16 hi / mid / lo of -||t||^2 / 2
          const float s = -0.5f * a.tn2[j];
          const float hi = bf16_back(bf16_bits(s));
          const float mid = bf16_back(bf16_bits(s - hi));
          x = (k == kd) ? hi : (k == kd + 1 ? mid : (s - hi - mid));
        }
      }
      out[e] = bf16_bits(x);
    }
    uint4 v;
    v.x = out[0] | ((uint32_t)out[1] << 16);
    v.y = out[2] | ((uint32_t)out[3] << 16);
    v.z = out[4] | ((uint32_t)out[5] << 16);
    v.w = out[6] | ((uint32_t)out[7] << 16);
    reinterpret_cast<uint4*>(a.tiles)[idx] = v;
  }
}

// ---- main kernel -------------------------------------------------------------------------------------
struct MmaArgs {
  ScoreArgs s;
  int64_t n_targets;
  int64_t n_tiles;
  int parts, kp, dist, stages;
  const float* header;
  const uint16_t* tiles;
  const int64_t* hist_off;
  const int64_t* hist_items;
  int mask_first;
  int k;
  uint2* cand;        // [n][CAND]: {approx value bits, group id | unsafe << 31}
  int32_t* cand_cnt;  // [n]: entries, or -1 when the list could not be compacted
  float* dbg_out;     // optional dense approximate scores [n, n_tiles * TN]
};

// Entry of a row's candidate list: x = group maximum (fp32 bits); y = group id (bits 0..28),
// bit 30 = "safety known", bit 31 = "unsafe" (the group holds a masked / out-of-range target, so
// its maximum must not feed the threshold).  Safety is resolved lazily, at compaction time, by the
// whole warp (one binary search per lane instead of one per push).
constexpr uint32_t GID_MASK = 0x1FFFFFFFu, F_KNOWN = 0x40000000u, F_UNSAFE = 0x80000000u;

__device__ __forceinline__ uint32_t orderable(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t b) {
  return __uint_as_float((b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b);
}

struct MaskInfo {
  const int64_t* hist_items;
  int64_t h_lo, h_hi;
  int64_t n_targets;
  int mask_first;
};

__device__ __forceinline__ bool group_unsafe(const MaskInfo& mi, uint32_t gid) {
  const int64_t j0 = (int64_t)gid * GRP;
  if (j0 + GRP > mi.n_targets || (mi.mask_first && gid == 0)) return true;
  int64_t lo = mi.h_lo, hi = mi.h_hi;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (mi.hist_items[mid] < j0) lo = mid + 1; else hi = mid;
  }
  return lo < mi.h_hi && mi.hist_items[lo] < j0 + GRP;
}

// Warp-cooperative compaction of one row's candidate list: tau = k-th largest maximum among the
// safe groups (radix select on the order-preserving bit pattern), keep every entry >= tau - 2 eps.
// Returns the new count (lane-uniform), -1 when the list cannot be shrunk; thr_out = new threshold.
__device__ __forceinline__ int compact_row(uint2* buf, int cnt, int k, float eps, const MaskInfo& mi, int lane,
                                           float& thr_out) {
  uint2 e[CAND / 32];
  bool valid[CAND / 32];
  uint32_t key[CAND / 32];
#pragma unroll
  for (int q = 0; q < CAND / 32; ++q) {
    const int i = q * 32 + lane;
    valid[q] = i < cnt;
    e[q] = valid[q] ? buf[i] : make_uint2(0u, 0u);
    if (valid[q] && !(e[q].y & F_KNOWN)) e[q].y |= F_KNOWN | (group_unsafe(mi, e[q].y & GID_MASK) ? F_UNSAFE : 0u);
    key[q] = (valid[q] && !(e[q].y & F_UNSAFE)) ? orderable(__uint_as_float(e[q].x)) : 0u;  // orderable() > 0
  }
  // largest T with #(key >= T) >= k  ==  the k-th largest safe key (0 when fewer than k are safe)
  uint32_t T = 0u;
#pragma unroll 4
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t c = T | (1u << bit);
    int n = 0;
#pragma unroll
    for (int q = 0; q < CAND / 32; ++q) n += __popc(__ballot_sync(0xffffffffu, key[q] >= c));
    if (n >= k) T = c;
  }
  const float tau = T ? from_orderable(T) : -INFINITY;
  const float thr = tau - 2.f * eps;
  __syncwarp();
  int base = 0;
#pragma unroll
  for (int q = 0; q < CAND / 32; ++q) {
    const bool keep = valid[q] && __uint_as_float(e[q].x) >= thr;
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (keep) buf[base + __popc(b & ((1u << lane) - 1u))] = e[q];
    base += __popc(b);
  }
  __syncwarp();
  thr_out = thr;
  return (base > CAND - 12) ? -1 : base;
}

__global__ void __launch_bounds__(MMA_THREADS, 1) fullsort_mma_kernel(const MmaArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int kp = a.kp;
  const int kchunks = kp / 8;
  const uint32_t a_bytes = (uint32_t)MM * kp * 2;
  const uint32_t b_bytes = (uint32_t)TN * kp * 2;
  uint16_t* As = reinterpret_cast<uint16_t*>(smem);
  unsigned char* Bs = smem + a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bs + (size_t)a.stages * b_bytes);
  // bars: full[stages], empty[stages], tfull[2], tempty[2]
  float* eps_row = reinterpret_cast<float*>(bars + 2 * a.stages + 4);  // [MM]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(eps_row + MM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row0 = (int64_t)blockIdx.x * MM;
  const int nrows = (int)min((int64_t)MM, a.s.n - row0);
  const int d = a.s.m.d;
  const int kd = a.parts * d;

  // ---- setup: zero A, build bf16 queries, barriers, TMEM ---------------------------------------------
  for (uint32_t i = threadIdx.x; i < a_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(As)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const float tmax = a.header[0];
  for (int u = warp; u < MM; u += MMA_THREADS / 32) {
    float nq = 0.f;
    if (u < nrows) {
      for (int c = lane; c < d; c += 32) {
        float q0, q1;
        query_value(a.s, row0 + u, c, q0, q1);
        nq = fmaf(q0, q0, nq);
        As[((size_t)(c >> 3) * MM + u) * 8 + (c & 7)] = bf16_bits(q0);
        if (a.parts == 2) {
          nq = fmaf(q1, q1, nq);
          const int k1 = d + c;
          As[((size_t)(k1 >> 3) * MM + u) * 8 + (k1 & 7)] = bf16_bits(q1);
        }
      }
      if (a.dist && lane < 3) {
        const int k1 = kd + lane;
        As[((size_t)(k1 >> 3) * MM + u) * 8 + (k1 & 7)] = bf16_bits(1.0f);
      }
    }
    nq = warp_sum(nq);
    if (lane == 0) {
      float eps = 1.02f * 0.00390625f * sqrtf(nq) * tmax;                // 2^-8 ||q|| max||t||
      if (a.dist) eps += 9.5367431640625e-7f * 0.5f * tmax * tmax;        // 2^-20 * ||t||^2 / 2 (split remainder)
      eps_row[u] = eps;
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);              // full: producer's expect_tx arrive
      mbar_init(smem_u32(&bars[a.stages + s]), 1);   // empty: one tcgen05.commit
    }
    mbar_init(smem_u32(&bars[2 * a.stages + 0]), 1);  // tfull[0]: one commit
    mbar_init(smem_u32(&bars[2 * a.stages + 1]), 1);
    mbar_init(smem_u32(&bars[2 * a.stages + 2]), 8);  // tempty[0]: one arrive per epilogue warp
    mbar_init(smem_u32(&bars[2 * a.stages + 3]), 8);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async();  // the generic-proxy writes of A must be visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[a.stages]);
  const uint32_t tfull0 = smem_u32(&bars[2 * a.stages]), tempty0 = smem_u32(&bars[2 * a.stages + 2]);

  if (warp == 0) {
    // ===== producer: stream target tiles into the ring =====
    if (lane == 0) {
      for (int64_t t = 0; t < a.n_tiles; ++t) {
        const int s = (int)(t % a.stages);
        const uint32_t ph = (uint32_t)((t / a.stages) & 1);
        mbar_wait(empty0 + 8 * s, ph ^ 1u);
        mbar_arrive_expect_tx(full0 + 8 * s, b_bytes);
        bulk_g2s(smem_u32(Bs + (size_t)s * b_bytes), reinterpret_cast<const unsigned char*>(a.tiles) + (size_t)t * b_bytes,
                 b_bytes, full0 + 8 * s);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread =====
    if (lane == 0) {
      const uint32_t a_lbo = MM * 16, b_lbo = TN * 16, sbo = 128;
      const uint32_t a_addr = smem_u32(As);
      for (int64_t t = 0; t < a.n_tiles; ++t) {
        const int s = (int)(t % a.stages);
        const uint32_t ph = (uint32_t)((t / a.stages) & 1);
        const int buf = (int)(t & 1);
        const uint32_t tph = (uint32_t)((t >> 1) & 1);
        mbar_wait(tempty0 + 8 * buf, tph ^ 1u);  // epilogue has drained this accumulator pair
        mbar_wait(full0 + 8 * s, ph);            // tile landed
        tc_fence_after();
        const uint32_t b_addr = smem_u32(Bs + (size_t)s * b_bytes);
        for (int h = 0; h < 2; ++h) {
          const uint32_t dcol = tmem_base + (uint32_t)(buf * 2 + h) * TN;
          for (int ks = 0; ks < kp / 16; ++ks) {
            const uint64_t adesc = make_smem_desc(a_addr + (uint32_t)h * 128 * 16 + (uint32_t)ks * 2 * a_lbo, a_lbo, sbo);
            const uint64_t bdesc = make_smem_desc(b_addr + (uint32_t)ks * 2 * b_lbo, b_lbo, sbo);
            umma_f16(dcol, adesc, bdesc, IDESC, ks > 0 ? 1u : 0u);
          }
        }
        umma_commit(empty0 + 8 * s);       // smem stage reusable once these MMAs have read it
        umma_commit(tfull0 + 8 * buf);     // accumulators ready
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===== epilogue: one thread per query row =====
    const int e = warp - EPI_WARP0;
    const int h = e >> 2, quad = e & 3;   // quad == warp % 4: the TMEM lane quadrant this warp may read
    const int u = h * 128 + quad * 32 + lane;
    const bool active = u < nrows;
    const int64_t qrow = row0 + u;
    const float eps = eps_row[u];
    float thr = active ? -INFINITY : INFINITY;
    int cnt = 0;
    bool overflow = false;
    uint2* mybuf = a.cand + (active ? qrow : row0) * CAND;
    int64_t h_lo = 0, h_hi = 0;
    if (active && a.hist_off) {
      h_lo = a.hist_off[qrow];
      h_hi = a.hist_off[qrow + 1];
    }
    for (int64_t t = 0; t < a.n_tiles; ++t) {
      const int buf = (int)(t & 1);
      const uint32_t tph = (uint32_t)((t >> 1) & 1);
      mbar_wait(tfull0 + 8 * buf, tph);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < TN; c0 += 64) {
        float v[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 2 + h) * TN + (uint32_t)c0;
        tmem_ld32x2(taddr, v);  // two 32-column loads in flight, one wait
        if (a.dbg_out && active) {
          float* o = a.dbg_out + qrow * (a.n_tiles * TN) + t * TN + c0;
#pragma unroll
          for (int i = 0; i < 64; ++i) o[i] = v[i];
        }
        const uint32_t g0 = (uint32_t)((t * TN + c0) / GRP);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float m = fmaxf(max3f(v[8 * g], v[8 * g + 1], v[8 * g + 2]),
                                max3f(max3f(v[8 * g + 3], v[8 * g + 4], v[8 * g + 5]), v[8 * g + 6], v[8 * g + 7]));
          if (m >= thr && (int64_t)(g0 + g) * GRP < a.n_targets) {  // rare after the first tiles
            mybuf[cnt] = make_uint2(__float_as_uint(m), g0 + g);
            ++cnt;
          }
        }
        unsigned full = __ballot_sync(0xffffffffu, cnt > CAND - 8);
        while (full) {
          const int r = __ffs(full) - 1;
          full &= full - 1;
          const int cnt_r = __shfl_sync(0xffffffffu, cnt, r);
          const float eps_r = __shfl_sync(0xffffffffu, eps, r);
          MaskInfo mi;
          mi.hist_items = a.hist_items;
          mi.h_lo = __shfl_sync(0xffffffffu, h_lo, r);
          mi.h_hi = __shfl_sync(0xffffffffu, h_hi, r);
          mi.n_targets = a.n_targets;
          mi.mask_first = a.mask_first;
          uint2* buf_r = a.cand + (row0 + h * 128 + quad * 32 + r) * CAND;
          __syncwarp();
          float thr_new;
          const int n_new = compact_row(buf_r, cnt_r, a.k, eps_r, mi, lane, thr_new);
          if (lane == r) {
            if (n_new < 0) {
              overflow = true;
              thr = INFINITY;  // stop collecting: the row goes to the exact path
              cnt = 0;
            } else {
              cnt = n_new;
              thr = thr_new;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
    }
    if (active) a.cand_cnt[qrow] = overflow ? -1 : cnt;
  }

  // ---- teardown ------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---- exact re-score + top-k ------------------------------------------------------------------------------
struct RescoreArgs {
  ScoreArgs s;
  int64_t n_targets;
  int parts, dist;
  const int64_t* hist_off;
  const int64_t* hist_items;
  int mask_first;
  int k;
  const uint2* cand;
  const int32_t* cand_cnt;
  int64_t* ids_out;
  float* scores_out;
  int32_t* row_flags;
};

constexpr int RS_WARPS = 4;

__global__ void __launch_bounds__(RS_WARPS * 32) rescore_topk_kernel(const RescoreArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = a.s.m.d;
  const int kd = a.parts * d;
  const int k = a.k;
  // per warp: q[kd] floats | keys[CAND * GRP] | list[k]
  const size_t per_warp = ((size_t)kd * 4 + 15) / 16 * 16 + (size_t)CAND * GRP * 8 + (size_t)k * 8;
  unsigned char* base = smem + warp * per_warp;
  float* q = reinterpret_cast<float*>(base);
  uint64_t* keys = reinterpret_cast<uint64_t*>(base + ((size_t)kd * 4 + 15) / 16 * 16);
  uint64_t* list = keys + CAND * GRP;
  const float margin = (a.s.m.model == KGE_ROTATE) ? a.s.m.margin : 0.f;

  for (int64_t row = (int64_t)blockIdx.x * RS_WARPS + warp; row < a.s.n; row += (int64_t)gridDim.x * RS_WARPS) {
    const int cnt = a.cand_cnt[row];
    if (cnt < 0) {
      if (lane == 0) a.row_flags[row] = 1;
      continue;
    }
    for (int c = lane; c < d; c += 32) {
      float q0, q1;
      query_value(a.s, row, c, q0, q1);
      q[c] = q0;
      if (a.parts == 2) q[d + c] = q1;
    }
    __syncwarp();
    int64_t h_lo = 0, h_hi = 0;
    if (a.hist_off) {
      h_lo = a.hist_off[row];
      h_hi = a.hist_off[row + 1];
    }
    const int total = cnt * GRP;
    int n_valid = 0;
    for (int idx = lane; idx < total; idx += 32) {
      const uint32_t gid = a.cand[row * CAND + idx / GRP].y & GID_MASK;
      const int64_t j = (int64_t)gid * GRP + (idx % GRP);
      uint64_t key = 0ull;
      bool ok = j < a.n_targets && !(a.mask_first && j == 0);
      if (ok && h_hi > h_lo) {
        int64_t lo = h_lo, hi = h_hi;
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if (a.hist_items[mid] < j) lo = mid + 1; else hi = mid;
        }
        ok = !(lo < h_hi && a.hist_items[lo] == j);
      }
      if (ok) {
        // the CUDA-core kernel's chain: parts in order, columns ascending, one fmaf per column
        float acc = 0.f;
        for (int p = 0; p < a.parts; ++p) {
          const float* t = a.s.m.entity.w[p] + j * d;
          const float* qp = q + p * d;
          if ((d & 3) == 0) {
            for (int c = 0; c < d; c += 4) {
              const float4 tv = __ldg(reinterpret_cast<const float4*>(t + c));
              if (a.dist) {
                const float e0 = qp[c] - tv.x, e1 = qp[c + 1] - tv.y, e2 = qp[c + 2] - tv.z, e3 = qp[c + 3] - tv.w;
                acc = fmaf(e0, e0, acc);
                acc = fmaf(e1, e1, acc);
                acc = fmaf(e2, e2, acc);
                acc = fmaf(e3, e3, acc);
              } else {
                acc = fmaf(qp[c], tv.x, acc);
                acc = fmaf(qp[c + 1], tv.y, acc);
                acc = fmaf(qp[c + 2], tv.z, acc);
                acc = fmaf(qp[c + 3], tv.w, acc);
              }
            }
          } else {
            for (int c = 0; c < d; ++c) {
              const float tv = __ldg(t + c);
              if (a.dist) {
                const float e0 = qp[c] - tv;
                acc = fmaf(e0, e0, acc);
              } else {
                acc = fmaf(qp[c], tv, acc);
              }
            }
          }
        }
        const float sc = a.dist ? (margin - sqrtf(acc)) : acc;
        key = make_key(sc, (uint32_t)j);
        ++n_valid;
      }
      keys[idx] = key;
    }
    n_valid = (int)warp_sum((float)n_valid);
    __syncwarp();
    int len = 0;
    for (int idx = 0; idx < total; ++idx) {
      const uint64_t key = keys[idx];
      if (key != 0ull) warp_insert(list, len, k, key, lane);
    }
    if (n_valid < k) {
      if (lane == 0) a.row_flags[row] = 1;
    } else {
      if (lane == 0) a.row_flags[row] = 0;
      for (int i = lane; i < k; i += 32) {
        const uint64_t key = list[i];
        a.ids_out[row * k + i] = key_id(key);
        if (a.scores_out) a.scores_out[row * k + i] = key_score(key);
      }
    }
    __syncwarp();
  }
}

struct MmaPlan {
  int parts, dist, kp, stages;
  size_t smem;
  int64_t n_tiles;
};

int plan_mma(const kge_model_t* m, int64_t n_targets, int k, MmaPlan& pl) {
  KGE_REQUIRE(m && m->model >= KGE_TRANSE && m->model <= KGE_COMPLEX, KGE_E_ARG, "bad model");
  pl.parts = (m->model == KGE_ROTATE || m->model == KGE_COMPLEX) ? 2 : 1;
  pl.dist = (m->model == KGE_TRANSE || m->model == KGE_ROTATE) ? 1 : 0;
  const int kd = pl.parts * m->d + (pl.dist ? 3 : 0);
  pl.kp = (kd + 15) / 16 * 16;
  KGE_REQUIRE(pl.kp <= 256, KGE_E_UNSUPPORTED, "K = %d too large for the tensor-core path", pl.kp);
  KGE_REQUIRE(k >= 1 && k <= 32, KGE_E_UNSUPPORTED, "k = %d too large for the tensor-core path (max 32)", k);
  KGE_REQUIRE(n_targets >= 1 && n_targets < (int64_t)0x1FFFFFFF * GRP, KGE_E_UNSUPPORTED, "bad n_targets");
  const size_t a_bytes = (size_t)MM * pl.kp * 2, b_bytes = (size_t)TN * pl.kp * 2;
  const size_t fixed = a_bytes + 64 * 8 + MM * 4 + 64;
  int stages = (int)((200 * 1024 - fixed) / b_bytes);
  if (stages > 6) stages = 6;
  KGE_REQUIRE(stages >= 2, KGE_E_UNSUPPORTED, "K = %d leaves no room for a pipeline", pl.kp);
  pl.stages = stages;
  pl.smem = a_bytes + (size_t)stages * b_bytes + (size_t)(2 * stages + 4) * 8 + MM * 4 + 64;
  pl.n_tiles = (n_targets + TN - 1) / TN;
  return 0;
}

}  // namespace

extern "C" int64_t kge_mma_image_bytes(const kge_model_t* model, int64_t n_targets) {
  MmaPlan pl;
  if (!model || plan_mma(model, n_targets, 1, pl)) return -1;
  return IMG_HEADER + pl.n_tiles * TN * (int64_t)pl.kp * 2 + ((n_targets * 4 + 127) / 128) * 128;
}

extern "C" int kge_mma_prepare_targets(const kge_model_t* model, int64_t n_targets, void* image, int64_t image_bytes,
                                       kge_stream_t stream) {
  MmaPlan pl;
  if (int e = plan_mma(model, n_targets, 1, pl)) return e;
  KGE_REQUIRE(n_targets <= model->entity.rows, KGE_E_ARG, "n_targets beyond the entity table");
  const int64_t need = kge_mma_image_bytes(model, n_targets);
  KGE_REQUIRE(image && image_bytes >= need, KGE_E_ARG, "image buffer too small: need %lld bytes", (long long)need);
  KGE_REQUIRE((reinterpret_cast<uintptr_t>(image) & 127) == 0, KGE_E_ARG, "image buffer must be 128-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  PrepArgs a;
  a.m = *model;
  a.n_targets = n_targets;
  a.parts = pl.parts;
  a.kp = pl.kp;
  a.dist = pl.dist;
  a.header = reinterpret_cast<float*>(image);
  a.tiles = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(image) + IMG_HEADER);
  a.tn2 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(image) + IMG_HEADER + pl.n_tiles * TN * (int64_t)pl.kp * 2);
  KGE_CUDA(cudaMemsetAsync(image, 0, IMG_HEADER, st));
  const int sms = kge_num_sms();
  target_norm_kernel<<<sms * 4, 256, 0, st>>>(a);
  KGE_LAUNCH_CHECK();
  target_image_kernel<<<sms * 8, 256, 0, st>>>(a);
  KGE_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t kge_full_sort_topk_mma_workspace_bytes(const kge_model_t* model, int64_t n, int64_t n_targets,
                                                          int32_t k) {
  MmaPlan pl;
  if (!model || n < 0 || plan_mma(model, n_targets, k, pl)) return -1;
  const int64_t rows = (n + MM - 1) / MM * MM;
  return rows * CAND * 8 + rows * 4;
}

extern "C" int kge_full_sort_topk_mma(const kge_model_t* model, const int64_t* heads, const int64_t* rels, int64_t n,
                                      int head_is_user, int64_t n_targets, const void* image, const int64_t* hist_off,
                                      const int64_t* hist_items, int mask_first, int32_t k, int64_t* ids_out,
                                      float* scores_out, int32_t* row_flags, void* workspace, int64_t workspace_bytes,
                                      float* debug_scores, kge_stream_t stream) {
  MmaPlan pl;
  if (int e = plan_mma(model, n_targets, k, pl)) return e;
  KGE_REQUIRE(n >= 0 && n_targets <= model->entity.rows && k <= n_targets, KGE_E_ARG, "bad n / n_targets / k");
  if (n == 0) return 0;
  KGE_REQUIRE(heads && image && ids_out && row_flags && workspace, KGE_E_ARG, "NULL argument");
  KGE_REQUIRE((hist_off == nullptr) == (hist_items == nullptr), KGE_E_ARG, "hist_off / hist_items must come together");
  const int64_t need = kge_full_sort_topk_mma_workspace_bytes(model, n, n_targets, k);
  KGE_REQUIRE(workspace_bytes >= need, KGE_E_ARG, "workspace too small: need %lld bytes", (long long)need);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = (n + MM - 1) / MM * MM;

  MmaArgs a = {};
  a.s.m = *model;
  a.s.heads = heads;
  a.s.rels = rels;
  a.s.tails = nullptr;
  a.s.n = n;
  a.s.head_is_user = head_is_user;
  a.s.rel_row = model->ui_relation_fullsort;
  a.n_targets = n_targets;
  a.n_tiles = pl.n_tiles;
  a.parts = pl.parts;
  a.kp = pl.kp;
  a.dist = pl.dist;
  a.stages = pl.stages;
  a.header = reinterpret_cast<const float*>(image);
  a.tiles = reinterpret_cast<const uint16_t*>(reinterpret_cast<const unsigned char*>(image) + IMG_HEADER);
  a.hist_off = hist_off;
  a.hist_items = hist_items;
  a.mask_first = mask_first;
  a.k = k;
  a.cand = reinterpret_cast<uint2*>(workspace);
  a.cand_cnt = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(workspace) + rows * CAND * 8);
  a.dbg_out = debug_scores;
  KGE_CUDA(cudaFuncSetAttribute(fullsort_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  fullsort_mma_kernel<<<(unsigned)(rows / MM), MMA_THREADS, pl.smem, st>>>(a);
  KGE_LAUNCH_CHECK();

  RescoreArgs r = {};
  r.s = a.s;
  r.n_targets = n_targets;
  r.parts = pl.parts;
  r.dist = pl.dist;
  r.hist_off = hist_off;
  r.hist_items = hist_items;
  r.mask_first = mask_first;
  r.k = k;
  r.cand = a.cand;
  r.cand_cnt = a.cand_cnt;
  r.ids_out = ids_out;
  r.scores_out = scores_out;
  r.row_flags = row_flags;
  const int kd = pl.parts * model->d;
  const size_t per_warp = ((size_t)kd * 4 + 15) / 16 * 16 + (size_t)CAND * GRP * 8 + (size_t)k * 8;
  const size_t rs_smem = per_warp * RS_WARPS;
  if (rs_smem > 48 * 1024)
    KGE_CUDA(cudaFuncSetAttribute(rescore_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem));
  int64_t g = (n + RS_WARPS - 1) / RS_WARPS;
  const int64_t cap = (int64_t)kge_num_sms() * 16;
  rescore_topk_kernel<<<(unsigned)(g < cap ? g : cap), RS_WARPS * 32, rs_smem, st>>>(r);
  KGE_LAUNCH_CHECK();
  return 0;
}
