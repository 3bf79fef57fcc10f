// Scoring side of the KGE hot path for sm_100a: predict, dense full-sort scores and the fused
// full-sort + mask + per-user top-k, all on the CUDA cores in exact fp32.
//
// Replaces (paths under /root/reference/hopwise/):
//   model/knowledge_graph_embedding_recommender/{transe.py:100-154, distmult.py:97-146,
//   rotate.py:133-220, complex.py:130-219}         (predict / full_sort_predict and _kg twins)
//   trainer/trainer.py:716-735                     (scores[:, 0] = -inf; scores[history] = -inf)
//   evaluator/collector.py:176-177                 (torch.topk over [n, I])
//
// All four scorers are "transform the (head, relation) pair into one query vector q of
// Kdim = parts*d floats, then reduce against the target row t":
//   TransE   q = h + r                      score = -||q - t||
//   RotatE   q = rot(h, theta) (re | im)    score = margin - ||q - t||   (one L2 norm, rotate.py:61-69)
//   DistMult q = h * r                      score = q . t
//   ComplEx  q = (hr*rr | hi*rr + hr*ri - hi*ri)   score = q . t        (complex.py:53-62 as written)
// so one tile kernel serves them all: a CTA owns BU=32 query rows (q kept in shared memory for
// the whole sweep) and streams target tiles of BT=128 rows through shared memory in K-chunks
// of 32 floats; each thread accumulates a 4x4 register block.  The epilogue either stores
// the dense [n, n_targets] scores (API parity: the unchanged trainer mutates that tensor) or
// keeps a per-row sorted top-k list of 64-bit (score, ~id) keys in shared memory, so the dense
// matrix never exists (800 GB at 1M x 200k).  Order: score descending, id ascending.
#include <math.h>

#include "common.cuh"
#include "score_common.cuh"

namespace {

constexpr int BU = 32;    // query rows per CTA
constexpr int BT = 128;   // target rows per tile
constexpr int KC = 32;    // K chunk (floats)
constexpr int TS = KC + 4;  // padded row stride of the target tile: conflict-free LDS.128
constexpr int TILE_THREADS = 256;

// ---- predict: one lane group per (head, relation, tail) -------------------------------------
template <int MODEL, int VEC, int G, int NCH>
__global__ void __launch_bounds__(256) predict_kernel(const ScoreArgs a, float* __restrict__ out) {
  constexpr int E = VEC * NCH;
  constexpr int PH = (MODEL == KGE_ROTATE || MODEL == KGE_COMPLEX) ? 2 : 1;
  constexpr int PR = (MODEL == KGE_COMPLEX || MODEL == KGE_TRANSH) ? 2 : 1;
  const int d = a.m.d;
  const int gl = (threadIdx.x & 31) % G;
  const int groups_per_cta = blockDim.x / G;
  const int64_t n_groups = (int64_t)gridDim.x * groups_per_cta;
  const kge_table_t& HT = a.head_is_user ? a.m.user : a.m.entity;
  for (int64_t i = (int64_t)blockIdx.x * groups_per_cta + threadIdx.x / G; i < a.n; i += n_groups) {
    const int64_t h_id = __ldg(a.heads + i);
    const int64_t r_id = a.rels ? __ldg(a.rels + i) : (int64_t)a.rel_row;
    const int64_t t_id = __ldg(a.tails + i);
    float h[PH][E], r[PR][E], t[PH][E];
#pragma unroll
    for (int p = 0; p < PH; ++p) {
      frag_load<VEC, G, NCH>(HT.w[p], h_id, d, gl, h[p]);
      frag_load<VEC, G, NCH>(a.m.entity.w[p], t_id, d, gl, t[p]);
    }
#pragma unroll
    for (int p = 0; p < PR; ++p) frag_load<VEC, G, NCH>(a.m.relation.w[p], r_id, d, gl, r[p]);
    float s = 0.f;
    if (MODEL == KGE_TRANSE) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float x = h[0][e] + r[0][e] - t[0][e];
        s += x * x;
      }
      s = -sqrtf(group_sum<G>(s));
    } else if (MODEL == KGE_TRANSH) {   // transh.py:53-58, 73-74: project head and tail, then TransE's norm
      float sw = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) sw += r[1][e];
      sw = group_sum<G>(sw);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float c = 1.f - sw * r[1][e];
        const float x = (h[0][e] * c + r[0][e]) - t[0][e] * c;
        s += x * x;
      }
      s = -sqrtf(group_sum<G>(s));
    } else if (MODEL == KGE_TORUSE) {   // toruse.py:66-76 (padding lanes hold zeros: min(0, 1) = 0)
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float x = (fracf_signed(h[0][e]) + fracf_signed(r[0][e])) - fracf_signed(t[0][e]);
        const float x2 = x * x;
        s += fminf(x2, 1.f - x2);
      }
      s = -(4.f * group_sum<G>(s));
    } else if (MODEL == KGE_DISTMULT) {
#pragma unroll
      for (int e = 0; e < E; ++e) s += h[0][e] * r[0][e] * t[0][e];
      s = group_sum<G>(s);
    } else if (MODEL == KGE_ROTATE) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        float sn, cs;
        sincosf(r[0][e], &sn, &cs);
        const float re = (cs * h[0][e] - sn * h[1][e]) - t[0][e];
        const float im = (cs * h[1][e] + sn * h[0][e]) - t[1][e];
        s += re * re + im * im;
      }
      s = a.m.margin - sqrtf(group_sum<G>(s));
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float A_ = h[0][e] * r[0][e];
        const float B_ = h[1][e] * r[0][e] + h[0][e] * r[1][e] - h[1][e] * r[1][e];
        s += t[0][e] * A_ + t[1][e] * B_;
      }
      s = group_sum<G>(s);
    }
    if (gl == 0) out[i] = s;
  }
}

// ---- TransD projection: out[i] = E[id_i] + RP[rel_i] * <E[id_i], EP[id_i]>  (transd.py:86-91) --------------------
template <int VEC, int G, int NCH>
__global__ void __launch_bounds__(256) transd_project_kernel(const float* __restrict__ emb, const float* __restrict__ vec,
                                                             const int64_t* __restrict__ ids, int64_t n, int d,
                                                             const float* __restrict__ rel_vec,
                                                             const int64_t* __restrict__ rel_ids, int64_t rel_row,
                                                             float* __restrict__ out) {
  constexpr int E = VEC * NCH;
  const int gl = (threadIdx.x & 31) % G;
  const int groups_per_cta = blockDim.x / G;
  const int64_t n_groups = (int64_t)gridDim.x * groups_per_cta;
  for (int64_t i = (int64_t)blockIdx.x * groups_per_cta + threadIdx.x / G; i < n; i += n_groups) {
    const int64_t row = ids ? __ldg(ids + i) : i;
    const int64_t rr = rel_ids ? __ldg(rel_ids + i) : rel_row;
    float e0[E], e1[E], rp[E];
    frag_load<VEC, G, NCH>(emb, row, d, gl, e0);
    frag_load<VEC, G, NCH>(vec, row, d, gl, e1);
    frag_load<VEC, G, NCH>(rel_vec, rr, d, gl, rp);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) s = __fmaf_rn(e0[e], e1[e], s);
    s = group_sum<G>(s);
#pragma unroll
    for (int e = 0; e < E; ++e) e0[e] = __fmaf_rn(rp[e], s, e0[e]);
    frag_store<VEC, G, NCH>(out, i, d, gl, e0);
  }
}

// ---- TransH projection: out[i] = E[id_i] * (1 - sum(w) * w), w = W[rel_i]  (transh.py:73-74) ---------------------
template <int VEC, int G, int NCH>
__global__ void __launch_bounds__(256) transh_project_kernel(const float* __restrict__ emb, const int64_t* __restrict__ ids,
                                                             int64_t n, int d, const float* __restrict__ norm_vec,
                                                             const int64_t* __restrict__ rel_ids, int64_t rel_row,
                                                             float* __restrict__ out) {
  constexpr int E = VEC * NCH;
  const int gl = (threadIdx.x & 31) % G;
  const int groups_per_cta = blockDim.x / G;
  const int64_t n_groups = (int64_t)gridDim.x * groups_per_cta;
  for (int64_t i = (int64_t)blockIdx.x * groups_per_cta + threadIdx.x / G; i < n; i += n_groups) {
    const int64_t row = ids ? __ldg(ids + i) : i;
    const int64_t rr = rel_ids ? __ldg(rel_ids + i) : rel_row;
    float e0[E], w[E];
    frag_load<VEC, G, NCH>(emb, row, d, gl, e0);
    frag_load<VEC, G, NCH>(norm_vec, rr, d, gl, w);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) s += w[e];
    s = group_sum<G>(s);
#pragma unroll
    for (int e = 0; e < E; ++e) e0[e] *= __fsub_rn(1.f, __fmul_rn(s, w[e]));
    frag_store<VEC, G, NCH>(out, i, d, gl, e0);
  }
}

// ---- the tile kernel ---------------------------------------------------------------------------
struct TileArgs {
  ScoreArgs s;
  int64_t n_targets;
  int kpad;            // padded floats per part = ceil(d / KC) * KC
  int tiles_per_split; // target tiles handled by one blockIdx.y
  // dense epilogue
  float* out;
  // top-k epilogue
  const int64_t* hist_off;
  const int64_t* hist_items;
  int mask_first;
  int k;
  uint64_t* part_keys;  // [n, n_splits, k]
  int n_splits;
  // Indirect rows (the tensor-core path's device-side fallback): the CTA works on entries [row_begin, min(*row_count,
  // row_end)) of row_map instead of rows [0, n); the row count is only known on the device, so the grid is sized for
  // a capacity, CTAs beyond the count leave at once and the others stride over the row blocks.  part_keys is
  // indexed by (entry - row_begin).
  const int32_t* row_map;
  const int32_t* row_count;
  int64_t row_begin, row_end;
};

__device__ __forceinline__ void build_queries(const ScoreArgs& a, const int64_t* rowid, int nrows, int kpad, float* Qs,
                                              const float* proj = nullptr) {
  const int d = a.m.d;
  const int parts = (a.m.model == KGE_ROTATE || a.m.model == KGE_COMPLEX) ? 2 : 1;
  const int qstride = parts * kpad;
  for (int idx = threadIdx.x; idx < BU * kpad; idx += blockDim.x) {
    const int u = idx / kpad, c = idx - u * kpad;
    float q0 = 0.f, q1 = 0.f;
    if (u < nrows && c < d) query_value(a, rowid[u], c, q0, q1, proj);
    Qs[u * qstride + c] = q0;
    if (parts == 2) Qs[u * qstride + kpad + c] = q1;
  }
}

// Stage target rows [t0, t0+BT) x columns [c0, c0+KC) of one part into Ts (zero padded).
template <bool FRAC, bool PROJ>
__device__ __forceinline__ void load_target_chunk(const float* __restrict__ W, int d, int64_t n_targets, int64_t t0,
                                                  int c0, float* Ts, const float* Cs) {
  if ((d & 3) == 0) {
    // 8 float4 per row chunk; consecutive lanes take consecutive rows -> conflict-free STS.128
    for (int idx = threadIdx.x; idx < BT * (KC / 4); idx += blockDim.x) {
      const int q = idx / BT, j = idx - q * BT;
      const int c = c0 + q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int64_t t = t0 + j;
      if (t < n_targets && c < d) v = __ldg(reinterpret_cast<const float4*>(W + t * d + c));
      if (FRAC) v = make_float4(fracf_signed(v.x), fracf_signed(v.y), fracf_signed(v.z), fracf_signed(v.w));
      if (PROJ) {   // (c < kpad always; Cs is zero beyond d like v)
        v = make_float4(v.x * Cs[c], v.y * Cs[c + 1], v.z * Cs[c + 2], v.w * Cs[c + 3]);
      }
      *reinterpret_cast<float4*>(Ts + j * TS + q * 4) = v;
    }
  } else {
    for (int idx = threadIdx.x; idx < BT * KC; idx += blockDim.x) {
      const int j = idx / KC, q = idx - j * KC;
      const int c = c0 + q;
      const int64_t t = t0 + j;
      const float v = (t < n_targets && c < d) ? __ldg(W + t * d + c) : 0.f;
      Ts[j * TS + q] = FRAC ? fracf_signed(v) : (PROJ ? v * Cs[c] : v);
    }
  }
}

// MODE: how a (query, target) pair of rows is contracted -- 0 dot product, 1 squared distance (score = margin - sqrt),
// 2 torus distance sum(min(x^2, 1 - x^2)) on frac()ed rows (score = -4 * sum).  Zero padding contributes 0 in all three.
// 3 = mode 1 on rows projected by one relation's factor c = 1 - sum(w) * w (TransH over items: every query row takes the
// user->item relation, so one factor vector serves the whole launch; it is built once per CTA in shared memory).
constexpr int MODE_DOT = 0, MODE_DIST = 1, MODE_TORUS = 2, MODE_PROJ = 3;
template <int MODE, bool TOPK>
__global__ void __launch_bounds__(TILE_THREADS) fullsort_tile_kernel(const TileArgs a) {
  constexpr bool DIST = MODE == MODE_DIST || MODE == MODE_PROJ;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int model = a.s.m.model;
  const int parts = (model == KGE_ROTATE || model == KGE_COMPLEX) ? 2 : 1;
  const int d = a.s.m.d;
  const int kpad = a.kpad;
  const int qstride = parts * kpad;
  const int k = a.k;

  float* Qs = reinterpret_cast<float*>(smem_raw);  // [BU][qstride]
  float* Ts = Qs + BU * qstride;                   // [BT][TS]
  // top-k only
  uint64_t* lists = reinterpret_cast<uint64_t*>(Ts + BT * TS);  // [BU][k]
  uint64_t* cand = lists + (TOPK ? BU * k : 0);                 // [BU][BT]
  uint64_t* thr = cand + (TOPK ? BU * BT : 0);                  // [BU]
  uint32_t* maskbits = reinterpret_cast<uint32_t*>(thr + (TOPK ? BU : 0));  // [BU][BT/32]
  int* cnt = reinterpret_cast<int*>(maskbits + (TOPK ? BU * (BT / 32) : 0));  // [BU]
  int* llen = cnt + (TOPK ? BU : 0);                                        // [BU]
  int64_t* cursor = reinterpret_cast<int64_t*>(llen + (TOPK ? BU : 0));     // [BU] (8-byte aligned by layout)
  int64_t* rowid = cursor + (TOPK ? BU : 0);                                // [BU] query row of each slot
  float* Cs = reinterpret_cast<float*>(rowid + BU);                         // [kpad] projection factor (MODE_PROJ)
  if (MODE == MODE_PROJ) {
    // c = 1 - sum(w) * w of the user->item relation's hyperplane vector (relation part 1), zero beyond d
    __shared__ float s_part[TILE_THREADS / 32];
    const float* wv = a.s.m.relation.w[1] + (int64_t)a.s.rel_row * d;
    float part = 0.f;
    for (int c = threadIdx.x; c < d; c += blockDim.x) part += __ldg(wv + c);
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    float sw = 0.f;
    for (int i = 0; i < TILE_THREADS / 32; ++i) sw += s_part[i];
    for (int c = threadIdx.x; c < kpad; c += blockDim.x) Cs[c] = c < d ? 1.f - sw * __ldg(wv + c) : 0.f;
    __syncthreads();
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tu = warp;  // users 4*tu .. 4*tu+3

  const int64_t n_tiles = (a.n_targets + BT - 1) / BT;
  const int64_t tile_lo = (int64_t)blockIdx.y * a.tiles_per_split;
  const int64_t tile_hi = min(n_tiles, tile_lo + a.tiles_per_split);
  const int64_t row_stop = a.row_map ? min((int64_t)__ldg(a.row_count), a.row_end) : a.s.n;

  for (int64_t row0 = a.row_begin + (int64_t)blockIdx.x * BU; row0 < row_stop; row0 += (int64_t)gridDim.x * BU) {
  const int nrows = (int)min((int64_t)BU, row_stop - row0);
  __syncthreads();   // (indirect rows: the previous row block of this CTA is done with shared memory)
  for (int u = threadIdx.x; u < BU; u += blockDim.x)
    rowid[u] = u < nrows ? (a.row_map ? (int64_t)__ldg(a.row_map + row0 + u) : row0 + u) : 0;
  __syncthreads();
  build_queries(a.s, rowid, nrows, kpad, Qs, MODE == MODE_PROJ ? Cs : nullptr);
  if (TOPK) {
    for (int u = threadIdx.x; u < BU; u += blockDim.x) {
      thr[u] = 0ull;
      llen[u] = 0;
      cnt[u] = 0;
      int64_t cur = 0;
      if (u < nrows && a.hist_off) {
        // first history entry >= tile_lo * BT (history sorted ascending per row)
        int64_t lo = a.hist_off[rowid[u]], hi = a.hist_off[rowid[u] + 1];
        const int64_t first = tile_lo * BT;
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if (a.hist_items[mid] < first) lo = mid + 1; else hi = mid;
        }
        cur = lo;
      }
      cursor[u] = cur;
    }
  }
  __syncthreads();

  for (int64_t tile = tile_lo; tile < tile_hi; ++tile) {
    const int64_t t0 = tile * BT;
    float acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;

    if (TOPK) {
      // history bits of this tile
      for (int i = threadIdx.x; i < BU * (BT / 32); i += blockDim.x) maskbits[i] = 0u;
      __syncthreads();
      if (a.hist_off) {
        for (int u = warp; u < nrows; u += TILE_THREADS / 32) {
          const int64_t end = a.hist_off[rowid[u] + 1];
          int64_t cur = cursor[u];
          const int64_t limit = t0 + BT;
          while (true) {
            const int64_t idx = cur + lane;
            int64_t item = limit;
            if (idx < end) item = a.hist_items[idx];
            const bool in = item < limit;
            if (in && item >= t0) {
              const int off = (int)(item - t0);
              atomicOr(&maskbits[u * (BT / 32) + (off >> 5)], 1u << (off & 31));
            }
            const unsigned b = __ballot_sync(0xffffffffu, in);
            cur += __popc(b);
            if (b != 0xffffffffu) break;
          }
          if (lane == 0) cursor[u] = cur;
        }
      }
    }

    for (int p = 0; p < parts; ++p) {
      const float* W = a.s.m.entity.w[p];
      for (int c0 = 0; c0 < kpad; c0 += KC) {
        __syncthreads();  // previous chunk consumed
        load_target_chunk<MODE == MODE_TORUS, MODE == MODE_PROJ>(W, d, a.n_targets, t0, c0, Ts, Cs);
        __syncthreads();
        const float* qbase = Qs + (4 * tu) * qstride + p * kpad + c0;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
          float4 qv[4], tv[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) qv[x] = *reinterpret_cast<const float4*>(qbase + x * qstride + kk);
#pragma unroll
          for (int y = 0; y < 4; ++y) tv[y] = *reinterpret_cast<const float4*>(Ts + (lane + 32 * y) * TS + kk);
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
              if (DIST) {
                const float e0 = qv[x].x - tv[y].x, e1 = qv[x].y - tv[y].y;
                const float e2 = qv[x].z - tv[y].z, e3 = qv[x].w - tv[y].w;
                acc[x][y] = fmaf(e0, e0, acc[x][y]);
                acc[x][y] = fmaf(e1, e1, acc[x][y]);
                acc[x][y] = fmaf(e2, e2, acc[x][y]);
                acc[x][y] = fmaf(e3, e3, acc[x][y]);
              } else if (MODE == MODE_TORUS) {
                const float e0 = qv[x].x - tv[y].x, e1 = qv[x].y - tv[y].y;
                const float e2 = qv[x].z - tv[y].z, e3 = qv[x].w - tv[y].w;
                const float s0 = e0 * e0, s1 = e1 * e1, s2 = e2 * e2, s3 = e3 * e3;
                acc[x][y] += fminf(s0, 1.f - s0);
                acc[x][y] += fminf(s1, 1.f - s1);
                acc[x][y] += fminf(s2, 1.f - s2);
                acc[x][y] += fminf(s3, 1.f - s3);
              } else {
                acc[x][y] = fmaf(qv[x].x, tv[y].x, acc[x][y]);
                acc[x][y] = fmaf(qv[x].y, tv[y].y, acc[x][y]);
                acc[x][y] = fmaf(qv[x].z, tv[y].z, acc[x][y]);
                acc[x][y] = fmaf(qv[x].w, tv[y].w, acc[x][y]);
              }
            }
        }
      }
    }

    // ---- epilogue of this tile ------------------------------------------------------------
    const float margin = (model == KGE_ROTATE) ? a.s.m.margin : 0.f;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int u = 4 * tu + x;
      if (u >= nrows) continue;
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        const int jo = lane + 32 * y;
        const int64_t j = t0 + jo;
        if (j >= a.n_targets) continue;
        float sc = DIST ? (margin - sqrtf(acc[x][y])) : (MODE == MODE_TORUS ? -(4.f * acc[x][y]) : acc[x][y]);
        if (!TOPK) {
          a.out[rowid[u] * a.n_targets + j] = sc;
        } else {
          const bool masked = (a.mask_first && j == 0) || ((maskbits[u * (BT / 32) + (jo >> 5)] >> (jo & 31)) & 1u);
          if (masked) sc = -INFINITY;
          const uint64_t key = make_key(sc, (uint32_t)j);
          if (key > thr[u]) {
            const int slot = atomicAdd(&cnt[u], 1);
            cand[u * BT + slot] = key;
          }
        }
      }
    }
    if (TOPK) {
      __syncthreads();
      for (int u = warp; u < nrows; u += TILE_THREADS / 32) {
        const int c = cnt[u];
        if (c == 0) continue;
        int len = llen[u];
        uint64_t* list = lists + u * k;
        for (int i = 0; i < c; ++i) warp_insert(list, len, k, cand[u * BT + i], lane);
        if (lane == 0) {
          llen[u] = len;
          cnt[u] = 0;
          thr[u] = (len == k) ? list[k - 1] : 0ull;
        }
      }
      // the next tile's first __syncthreads orders these writes before its reads
    }
  }

  if (TOPK) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < nrows * k; idx += blockDim.x) {
      const int u = idx / k, i = idx - u * k;
      const uint64_t key = (i < llen[u]) ? lists[u * k + i] : 0ull;
      a.part_keys[((row0 - a.row_begin + u) * a.n_splits + blockIdx.y) * k + i] = key;
    }
  }
  if (!a.row_map) break;   // direct rows: the grid covers every row block
  }
}

// Merge the per-split lists of one row (one warp per row) and decode.
// With row_map: entry e in [row_begin, min(*row_count, row_end)) of the map, lists at (e - row_begin), output row
// row_map[e]; warps stride over the entries (the count is only known on the device).
__global__ void __launch_bounds__(256) topk_merge_kernel(const uint64_t* __restrict__ part_keys, int64_t n,
                                                         int n_splits, int k, int64_t* __restrict__ ids_out,
                                                         float* __restrict__ scores_out,
                                                         const int32_t* __restrict__ row_map,
                                                         const int32_t* __restrict__ row_count, int64_t row_begin,
                                                         int64_t row_end) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw);  // [warps][k]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t stop = row_map ? min((int64_t)__ldg(row_count), row_end) : n;
  const int64_t stride = row_map ? (int64_t)gridDim.x * (blockDim.x >> 5) : stop;   // direct rows: one pass
  for (int64_t ent = row_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; ent < stop; ent += stride) {
  const int64_t row = row_map ? (int64_t)__ldg(row_map + ent) : ent;
  uint64_t* list = lists + warp * k;
  int len = 0;
  const uint64_t* src = part_keys + (ent - row_begin) * (int64_t)n_splits * k;
  if (n_splits == 1) {
    for (int i = lane; i < k; i += 32) list[i] = src[i];
    len = k;
    __syncwarp();
  } else {
    for (int i = 0; i < n_splits * k; ++i) {
      const uint64_t key = src[i];
      if (key != 0ull) warp_insert(list, len, k, key, lane);
    }
    for (int i = len + lane; i < k; i += 32) list[i] = 0ull;
    __syncwarp();
  }
  for (int i = lane; i < k; i += 32) {
    const uint64_t key = list[i];
    ids_out[row * k + i] = key ? key_id(key) : -1;
    if (scores_out) scores_out[row * k + i] = key ? key_score(key) : -INFINITY;
  }
  __syncwarp();
  }
}

// ---- collector.py:178-183 without the [n, I] matrix ---------------------------------------------
__global__ void __launch_bounds__(256) topk_hits_kernel(const int64_t* __restrict__ ids, int64_t n, int k,
                                                        const int64_t* __restrict__ pos_off,
                                                        const int64_t* __restrict__ pos_items,
                                                        int32_t* __restrict__ out) {
  const int64_t total = n * (int64_t)(k + 1);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = idx / (k + 1);
    const int i = (int)(idx - u * (k + 1));
    int64_t lo = pos_off[u], hi = pos_off[u + 1];
    if (i == k) {
      out[idx] = (int32_t)(hi - lo);
      continue;
    }
    const int64_t id = ids[u * k + i];
    int hit = 0;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      const int64_t v = pos_items[mid];
      if (v == id) { hit = 1; break; }
      if (v < id) lo = mid + 1; else hi = mid;
    }
    out[idx] = hit;
  }
}

// ---- metrics.py summed over users ---------------------------------------------------------------
// sums[5][k]: recall, mrr, ndcg, hit, precision (the order of overall.yaml's default metrics).
__device__ __forceinline__ double warp_sum_f64(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

__global__ void __launch_bounds__(256) topk_metric_sums_kernel(const int32_t* __restrict__ rec, int64_t n, int k,
                                                               double* __restrict__ sums) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);  // [warps][5*k]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < warps * 5 * k; i += blockDim.x) acc[i] = 0.0;
  __syncthreads();
  double* mine = acc + warp * 5 * k;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + warp * 32; base < n; base += stride) {
    const int64_t u = base + lane;
    const bool valid = u < n;
    const int32_t* row = rec + (valid ? u : 0) * (int64_t)(k + 1);
    const int pos_len = row[k];
    const int ilen = pos_len < k ? pos_len : k;
    int cum = 0, first = -1;
    double dcg = 0.0, idcg = 0.0;
    for (int j = 0; j < k; ++j) {
      const int hit = row[j] != 0;
      const double disc = 1.0 / log2((double)(j + 2));
      if (hit) {
        if (first < 0) first = j;
        dcg += disc;
      }
      cum += hit;
      if (j < ilen) idcg += disc;  // idcg freezes after min(pos_len, k) terms (metrics.py:191-207)
      double v0 = (double)cum / (double)pos_len;
      double v1 = first >= 0 ? 1.0 / (double)(first + 1) : 0.0;
      double v2 = dcg / idcg;
      double v3 = cum > 0 ? 1.0 : 0.0;
      double v4 = (double)cum / (double)(j + 1);
      if (!valid) v0 = v1 = v2 = v3 = v4 = 0.0;
      v0 = warp_sum_f64(v0);
      v1 = warp_sum_f64(v1);
      v2 = warp_sum_f64(v2);
      v3 = warp_sum_f64(v3);
      v4 = warp_sum_f64(v4);
      if (lane == 0) {
        mine[0 * k + j] += v0;
        mine[1 * k + j] += v1;
        mine[2 * k + j] += v2;
        mine[3 * k + j] += v3;
        mine[4 * k + j] += v4;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 5 * k; i += blockDim.x) {
    double t = 0.0;
    for (int w = 0; w < warps; ++w) t += acc[w * 5 * k + i];
    atomicAdd(&sums[i], t);
  }
}

int check_score_model(const kge_model_t* m) {
  KGE_REQUIRE(m, KGE_E_ARG, "model is NULL");
  KGE_REQUIRE(m->model >= KGE_TRANSE && m->model <= KGE_TRANSH, KGE_E_ARG, "unknown model kind %d", m->model);
  const int ph = (m->model == KGE_ROTATE || m->model == KGE_COMPLEX) ? 2 : 1;
  const int pr = (m->model == KGE_COMPLEX || m->model == KGE_TRANSH) ? 2 : 1;
  for (int p = 0; p < ph; ++p) KGE_REQUIRE(m->user.w[p] && m->entity.w[p], KGE_E_ARG, "NULL weight table");
  for (int p = 0; p < pr; ++p) KGE_REQUIRE(m->relation.w[p], KGE_E_ARG, "NULL relation table");
  KGE_REQUIRE(m->d >= 1, KGE_E_UNSUPPORTED, "embedding_size %d unsupported", m->d);
  return 0;
}

struct TilePlan {
  int kpad, n_splits, tiles_per_split;
  size_t smem;
  int64_t n_blocks;
};

int plan_tiles(const kge_model_t* m, int64_t n, int64_t n_targets, int k, bool topk, TilePlan& pl) {
  const int parts = (m->model == KGE_ROTATE || m->model == KGE_COMPLEX) ? 2 : 1;
  pl.kpad = (m->d + KC - 1) / KC * KC;
  size_t smem = (size_t)BU * parts * pl.kpad * 4 + (size_t)BT * TS * 4;
  if (topk) smem += (size_t)BU * k * 8 + (size_t)BU * BT * 8 + BU * 8 + BU * (BT / 32) * 4 + BU * 4 * 2 + BU * 8;
  smem += BU * 8;   // rowid
  if (m->model == KGE_TRANSH) smem += (size_t)pl.kpad * 4;   // projection factor
  pl.smem = smem;
  KGE_REQUIRE(smem <= 220 * 1024, KGE_E_UNSUPPORTED, "embedding_size %d needs %zu bytes of shared memory", m->d, smem);
  pl.n_blocks = (n + BU - 1) / BU;
  const int64_t n_tiles = (n_targets + BT - 1) / BT;
  int splits = 1;
  if (topk) {
    // enough CTAs for two waves at 2 CTAs/SM, but at least 8 tiles per split
    const int64_t want = (int64_t)kge_num_sms() * 4;
    int64_t s = (want + pl.n_blocks - 1) / pl.n_blocks;
    const int64_t max_s = (n_tiles + 7) / 8;
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    if (s > 65535) s = 65535;
    splits = (int)s;
  }
  pl.tiles_per_split = (int)((n_tiles + splits - 1) / splits);
  pl.n_splits = (int)((n_tiles + pl.tiles_per_split - 1) / pl.tiles_per_split);
  if (pl.n_splits < 1) pl.n_splits = 1;
  return 0;
}

int tile_mode(int model) {
  if (model == KGE_TORUSE) return MODE_TORUS;
  if (model == KGE_TRANSH) return MODE_PROJ;
  return (model == KGE_TRANSE || model == KGE_ROTATE) ? MODE_DIST : MODE_DOT;
}

}  // namespace

extern "C" int kge_predict(const kge_model_t* model, const int64_t* heads, const int64_t* rels, const int64_t* tails,
                           int64_t n, int head_is_user, float* out, kge_stream_t stream) {
  if (int e = check_score_model(model)) return e;
  KGE_REQUIRE(n >= 0, KGE_E_ARG, "negative n");
  if (n == 0) return 0;
  KGE_REQUIRE(heads && tails && out, KGE_E_ARG, "NULL heads / tails / out");
  RowCfg c = {};
  KGE_REQUIRE(kge_pick_rowcfg(model->d, c), KGE_E_UNSUPPORTED, "embedding_size %d unsupported", model->d);
  ScoreArgs a;
  a.m = *model;
  a.heads = heads;
  a.rels = rels;
  a.tails = tails;
  a.n = n;
  a.head_is_user = head_is_user;
  a.rel_row = model->ui_relation;
  const int threads = 256;
  int64_t g = (n + threads / c.g - 1) / (threads / c.g);
  const int64_t cap = (int64_t)kge_num_sms() * 8;
  const int grid = (int)(g < cap ? g : cap);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(V, G, N)                                                                          \
  switch (model->model) {                                                                      \
    case KGE_TRANSE: predict_kernel<KGE_TRANSE, V, G, N><<<grid, threads, 0, st>>>(a, out); break;     \
    case KGE_DISTMULT: predict_kernel<KGE_DISTMULT, V, G, N><<<grid, threads, 0, st>>>(a, out); break; \
    case KGE_ROTATE: predict_kernel<KGE_ROTATE, V, G, N><<<grid, threads, 0, st>>>(a, out); break;     \
    case KGE_TORUSE: predict_kernel<KGE_TORUSE, V, G, N><<<grid, threads, 0, st>>>(a, out); break;     \
    case KGE_TRANSH: predict_kernel<KGE_TRANSH, V, G, N><<<grid, threads, 0, st>>>(a, out); break;     \
    default: predict_kernel<KGE_COMPLEX, V, G, N><<<grid, threads, 0, st>>>(a, out); break;            \
  }
  KGE_DISPATCH_ROWCFG(c, CALL);
#undef CALL
  KGE_LAUNCH_CHECK();
  return 0;
}

extern "C" int kge_transd_project(const float* emb, const float* vec, const int64_t* ids, int64_t n, int32_t d,
                                  const float* rel_vec, const int64_t* rel_ids, int64_t rel_row, float* out,
                                  kge_stream_t stream) {
  KGE_REQUIRE(n >= 0 && d >= 1, KGE_E_ARG, "bad n / d");
  if (n == 0) return 0;
  KGE_REQUIRE(emb && vec && rel_vec && out && (rel_ids || rel_row >= 0), KGE_E_ARG, "NULL argument");
  RowCfg c = {};
  KGE_REQUIRE(kge_pick_rowcfg(d, c), KGE_E_UNSUPPORTED, "embedding_size %d unsupported", d);
  const int threads = 256;
  int64_t g = (n + threads / c.g - 1) / (threads / c.g);
  const int64_t cap = (int64_t)kge_num_sms() * 8;
  const int grid = (int)(g < cap ? g : cap);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(V, G, N) \
  transd_project_kernel<V, G, N><<<grid, threads, 0, st>>>(emb, vec, ids, n, d, rel_vec, rel_ids, rel_row, out)
  KGE_DISPATCH_ROWCFG(c, CALL);
#undef CALL
  KGE_LAUNCH_CHECK();
  return 0;
}

extern "C" int kge_transh_project(const float* emb, const int64_t* ids, int64_t n, int32_t d, const float* norm_vec,
                                  const int64_t* rel_ids, int64_t rel_row, float* out, kge_stream_t stream) {
  KGE_REQUIRE(n >= 0 && d >= 1, KGE_E_ARG, "bad n / d");
  if (n == 0) return 0;
  KGE_REQUIRE(emb && norm_vec && out && (rel_ids || rel_row >= 0), KGE_E_ARG, "NULL argument");
  RowCfg c = {};
  KGE_REQUIRE(kge_pick_rowcfg(d, c), KGE_E_UNSUPPORTED, "embedding_size %d unsupported", d);
  const int threads = 256;
  int64_t g = (n + threads / c.g - 1) / (threads / c.g);
  const int64_t cap = (int64_t)kge_num_sms() * 8;
  const int grid = (int)(g < cap ? g : cap);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(V, G, N) transh_project_kernel<V, G, N><<<grid, threads, 0, st>>>(emb, ids, n, d, norm_vec, rel_ids, rel_row, out)
  KGE_DISPATCH_ROWCFG(c, CALL);
#undef CALL
  KGE_LAUNCH_CHECK();
  return 0;
}

template <int MODE, bool TOPK>
static int launch_tile(const TileArgs& a, const TilePlan& pl, cudaStream_t st) {
  auto kern = fullsort_tile_kernel<MODE, TOPK>;
  KGE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  // blockIdx.x is limited to 2^31-1 rows/BU: fine for any table that fits the device
  dim3 grid((unsigned)pl.n_blocks, (unsigned)pl.n_splits);
  kern<<<grid, TILE_THREADS, pl.smem, st>>>(a);
  KGE_LAUNCH_CHECK();
  return 0;
}

extern "C" int kge_full_sort_scores(const kge_model_t* model, const int64_t* heads, const int64_t* rels, int64_t n,
                                    int head_is_user, int64_t n_targets, float* out, kge_stream_t stream) {
  if (int e = check_score_model(model)) return e;
  KGE_REQUIRE(model->model != KGE_TRANSH || rels == nullptr, KGE_E_UNSUPPORTED,
              "TransH full-sort takes the user->item relation only (transh.py:125-145; no KG scoring in the reference)");
  KGE_REQUIRE(n >= 0 && n_targets >= 1 && n_targets <= model->entity.rows, KGE_E_ARG, "bad n / n_targets");
  if (n == 0) return 0;
  KGE_REQUIRE(heads && out, KGE_E_ARG, "NULL heads / out");
  TilePlan pl = {};
  if (int e = plan_tiles(model, n, n_targets, 0, false, pl)) return e;
  TileArgs a = {};
  a.s.m = *model;
  a.s.heads = heads;
  a.s.rels = rels;
  a.s.tails = nullptr;
  a.s.n = n;
  a.s.head_is_user = head_is_user;
  a.s.rel_row = model->ui_relation_fullsort;
  a.n_targets = n_targets;
  a.kpad = pl.kpad;
  a.tiles_per_split = pl.tiles_per_split;
  a.out = out;
  a.n_splits = 1;
  switch (tile_mode(model->model)) {
    case MODE_DIST: return launch_tile<MODE_DIST, false>(a, pl, (cudaStream_t)stream);
    case MODE_TORUS: return launch_tile<MODE_TORUS, false>(a, pl, (cudaStream_t)stream);
    case MODE_PROJ: return launch_tile<MODE_PROJ, false>(a, pl, (cudaStream_t)stream);
    default: return launch_tile<MODE_DOT, false>(a, pl, (cudaStream_t)stream);
  }
}

extern "C" int64_t kge_full_sort_topk_workspace_bytes(const kge_model_t* model, int64_t n, int64_t n_targets,
                                                      int32_t k) {
  if (!model || n < 0 || n_targets < 1 || k < 1 || k > KMAX) return -1;
  TilePlan pl = {};
  if (plan_tiles(model, n > 0 ? n : 1, n_targets, k, true, pl)) return -1;
  return (int64_t)n * pl.n_splits * k * 8;
}

extern "C" int kge_full_sort_topk(const kge_model_t* model, const int64_t* heads, const int64_t* rels, int64_t n,
                                  int head_is_user, int64_t n_targets, const int64_t* hist_off,
                                  const int64_t* hist_items, int mask_first, int32_t k, int64_t* ids_out,
                                  float* scores_out, void* workspace, int64_t workspace_bytes, kge_stream_t stream) {
  if (int e = check_score_model(model)) return e;
  KGE_REQUIRE(model->model != KGE_TRANSH || rels == nullptr, KGE_E_UNSUPPORTED,
              "TransH full-sort takes the user->item relation only (transh.py:125-145; no KG scoring in the reference)");
  KGE_REQUIRE(n >= 0 && n_targets >= 1 && n_targets <= model->entity.rows, KGE_E_ARG, "bad n / n_targets");
  KGE_REQUIRE(k >= 1 && k <= KMAX, KGE_E_UNSUPPORTED, "k=%d outside [1, %d]", k, KMAX);
  KGE_REQUIRE(k <= n_targets, KGE_E_ARG, "k=%d larger than the number of targets", k);
  KGE_REQUIRE(n_targets < 0xFFFFFFFFll, KGE_E_UNSUPPORTED, "more than 2^32-2 targets");
  if (n == 0) return 0;
  KGE_REQUIRE(heads && ids_out, KGE_E_ARG, "NULL heads / ids_out");
  KGE_REQUIRE((hist_off == nullptr) == (hist_items == nullptr) || hist_off, KGE_E_ARG, "hist_items without hist_off");
  TilePlan pl = {};
  if (int e = plan_tiles(model, n, n_targets, k, true, pl)) return e;
  const int64_t need = n * pl.n_splits * k * 8;
  KGE_REQUIRE(workspace && workspace_bytes >= need, KGE_E_ARG, "workspace too small: need %lld bytes", (long long)need);
  TileArgs a = {};
  a.s.m = *model;
  a.s.heads = heads;
  a.s.rels = rels;
  a.s.tails = nullptr;
  a.s.n = n;
  a.s.head_is_user = head_is_user;
  a.s.rel_row = model->ui_relation_fullsort;
  a.n_targets = n_targets;
  a.kpad = pl.kpad;
  a.tiles_per_split = pl.tiles_per_split;
  a.hist_off = hist_off;
  a.hist_items = hist_items;
  a.mask_first = mask_first;
  a.k = k;
  a.part_keys = reinterpret_cast<uint64_t*>(workspace);
  a.n_splits = pl.n_splits;
  cudaStream_t st = (cudaStream_t)stream;
  const int mode = tile_mode(model->model);
  if (int e = mode == MODE_DIST ? launch_tile<MODE_DIST, true>(a, pl, st)
                                : (mode == MODE_TORUS ? launch_tile<MODE_TORUS, true>(a, pl, st)
                                   : (mode == MODE_PROJ ? launch_tile<MODE_PROJ, true>(a, pl, st)
                                                        : launch_tile<MODE_DOT, true>(a, pl, st))))
    return e;
  const int warps = 8;
  const int64_t mg = (n + warps - 1) / warps;
  topk_merge_kernel<<<(unsigned)mg, warps * 32, (size_t)warps * k * 8, st>>>(a.part_keys, n, pl.n_splits, k, ids_out,
                                                                            scores_out, nullptr, nullptr, 0, 0);
  KGE_LAUNCH_CHECK();
  return 0;
}

// ---- device-gated exact top-k for a row list that only exists on the device ------------------------------
// Used by the tensor-core path (mma_topk.cu) for the rows its filter hands back: `row_map[0 .. *row_count)` are
// their indices.  No host round trip: two fixed-size launches cover every possible count --
//   entries [0, FB_SPLIT_ROWS)      row blocks x target splits, so that a handful of rows still fills the GPU
//   entries [FB_SPLIT_ROWS, n)      one CTA per row block (plenty of rows: no splits needed), CTAs stride
// -- and CTAs beyond the count leave at once (~10 us per call when nothing was flagged).
namespace {
constexpr int64_t FB_SPLIT_ROWS = 1024;
constexpr int FB_SPLITS = 48;
struct FallbackPlan {
  TilePlan tp;
  int splits, tiles_per_split;
  int64_t keys_a, keys_b;   // part_keys elements of the two regions
};
int plan_fallback(const kge_model_t* m, int64_t n, int64_t n_targets, int k, FallbackPlan& fp) {
  if (int e = plan_tiles(m, n > 0 ? n : 1, n_targets, k, true, fp.tp)) return e;
  const int64_t n_tiles = (n_targets + BT - 1) / BT;
  int64_t s = FB_SPLITS;
  if (s > (n_tiles + 7) / 8) s = (n_tiles + 7) / 8;
  if (s < 1) s = 1;
  fp.tiles_per_split = (int)((n_tiles + s - 1) / s);
  fp.splits = (int)((n_tiles + fp.tiles_per_split - 1) / fp.tiles_per_split);
  const int64_t ra = n < FB_SPLIT_ROWS ? n : FB_SPLIT_ROWS;
  fp.keys_a = ra * fp.splits * k;
  fp.keys_b = (n - ra) * k;
  return 0;
}
}  // namespace

int64_t kge_topk_rows_indirect_workspace_bytes(const kge_model_t* model, int64_t n, int64_t n_targets, int32_t k) {
  FallbackPlan fp = {};
  if (!model || k < 1 || k > KMAX || plan_fallback(model, n, n_targets, k, fp)) return -1;
  return (fp.keys_a + fp.keys_b) * 8;
}

int kge_topk_rows_indirect(const kge_model_t* model, const int64_t* heads, const int64_t* rels, int64_t n,
                           int head_is_user, int64_t n_targets, const int64_t* hist_off, const int64_t* hist_items,
                           int mask_first, int32_t k, const int32_t* row_map, const int32_t* row_count,
                           int64_t* ids_out, float* scores_out, void* workspace, cudaStream_t st) {
  if (int e = check_score_model(model)) return e;
  FallbackPlan fp = {};
  if (int e = plan_fallback(model, n, n_targets, k, fp)) return e;
  TileArgs a = {};
  a.s.m = *model;
  a.s.heads = heads;
  a.s.rels = rels;
  a.s.n = n;
  a.s.head_is_user = head_is_user;
  a.s.rel_row = model->ui_relation_fullsort;
  a.n_targets = n_targets;
  a.kpad = fp.tp.kpad;
  a.hist_off = hist_off;
  a.hist_items = hist_items;
  a.mask_first = mask_first;
  a.k = k;
  a.row_map = row_map;
  a.row_count = row_count;
  const int mode = tile_mode(model->model);
  auto kern = mode == MODE_DIST ? fullsort_tile_kernel<MODE_DIST, true>
                                : (mode == MODE_TORUS ? fullsort_tile_kernel<MODE_TORUS, true>
                                   : (mode == MODE_PROJ ? fullsort_tile_kernel<MODE_PROJ, true>
                                                        : fullsort_tile_kernel<MODE_DOT, true>));
  KGE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp.tp.smem));
  const int warps = 8;
  const int64_t ra = n < FB_SPLIT_ROWS ? n : FB_SPLIT_ROWS;
  uint64_t* keys = reinterpret_cast<uint64_t*>(workspace);
  // region A: few rows, many splits
  a.part_keys = keys;
  a.n_splits = fp.splits;
  a.tiles_per_split = fp.tiles_per_split;
  a.row_begin = 0;
  a.row_end = ra;
  kern<<<dim3((unsigned)((ra + BU - 1) / BU), (unsigned)fp.splits), TILE_THREADS, fp.tp.smem, st>>>(a);
  KGE_LAUNCH_CHECK();
  topk_merge_kernel<<<(unsigned)((ra + warps - 1) / warps), warps * 32, (size_t)warps * k * 8, st>>>(
      a.part_keys, n, a.n_splits, k, ids_out, scores_out, row_map, row_count, 0, ra);
  KGE_LAUNCH_CHECK();
  if (n > ra) {   // region B: one CTA per row block, striding
    a.part_keys = keys + fp.keys_a;
    a.n_splits = 1;
    a.tiles_per_split = (int)((n_targets + BT - 1) / BT);
    a.row_begin = ra;
    a.row_end = n;
    int64_t g = (n - ra + BU - 1) / BU;
    const int64_t cap = (int64_t)kge_num_sms() * 2;
    if (g > cap) g = cap;
    kern<<<dim3((unsigned)g, 1), TILE_THREADS, fp.tp.smem, st>>>(a);
    KGE_LAUNCH_CHECK();
    topk_merge_kernel<<<(unsigned)g, warps * 32, (size_t)warps * k * 8, st>>>(a.part_keys, n, 1, k, ids_out, scores_out,
                                                                             row_map, row_count, ra, n);
    KGE_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int kge_topk_hits(const int64_t* ids, int64_t n, int32_t k, const int64_t* pos_off, const int64_t* pos_items,
                             int32_t* out, kge_stream_t stream) {
  KGE_REQUIRE(n >= 0 && k >= 1, KGE_E_ARG, "bad n / k");
  if (n == 0) return 0;
  KGE_REQUIRE(ids && pos_off && pos_items && out, KGE_E_ARG, "NULL argument");
  const int64_t total = n * (int64_t)(k + 1);
  int64_t g = (total + 255) / 256;
  const int64_t cap = (int64_t)kge_num_sms() * 8;
  topk_hits_kernel<<<(unsigned)(g < cap ? g : cap), 256, 0, (cudaStream_t)stream>>>(ids, n, k, pos_off, pos_items, out);
  KGE_LAUNCH_CHECK();
  return 0;
}

extern "C" int kge_topk_metric_sums(const int32_t* rec_topk, int64_t n, int32_t k, double* sums, kge_stream_t stream) {
  KGE_REQUIRE(n >= 0 && k >= 1 && k <= 128, KGE_E_ARG, "bad n / k");
  if (n == 0) return 0;
  KGE_REQUIRE(rec_topk && sums, KGE_E_ARG, "NULL argument");
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)kge_num_sms() * 4;
  topk_metric_sums_kernel<<<(unsigned)(g < cap ? g : cap), 256, (size_t)8 * 5 * k * 8, (cudaStream_t)stream>>>(rec_topk, n, k,
                                                                                                         sums);
  KGE_LAUNCH_CHECK();
  return 0;
}
