// Pieces shared by the CUDA-core scoring kernels (score.cu) and the tcgen05 full-sort path
// (mma_topk.cu): the 64-bit top-k key, the warp-cooperative sorted-list insertion and the query
// transform q(head, relation) of the four scorers.
#pragma once

#include "common.cuh"

namespace {

constexpr int KMAX = 128;  // largest supported k

__device__ __forceinline__ uint64_t make_key(float score, uint32_t id) {
  uint32_t b = __float_as_uint(score);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // order-preserving map of fp32 to uint32
  return ((uint64_t)b << 32) | (uint64_t)(0xFFFFFFFFu - id);  // larger key = better (score desc, id asc)
}
__device__ __forceinline__ float key_score(uint64_t key) {
  uint32_t b = (uint32_t)(key >> 32);
  b = (b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b;
  return __uint_as_float(b);
}
__device__ __forceinline__ int64_t key_id(uint64_t key) { return (int64_t)(0xFFFFFFFFu - (uint32_t)key); }

// Insert `c` into the descending list (capacity k) cooperatively by one warp.  `len` is warp-uniform.
__device__ __forceinline__ void warp_insert(uint64_t* list, int& len, int k, uint64_t c, int lane) {
  if (len == k && c <= list[k - 1]) return;
  int pos = 0;
  for (int base = 0; base < len; base += 32) {
    const int i = base + lane;
    const bool gt = (i < len) && (list[i] > c);
    pos += __popc(__ballot_sync(0xffffffffu, gt));
  }
  const int newlen = len < k ? len + 1 : k;
  uint64_t moved[KMAX / 32];
#pragma unroll
  for (int q = 0; q < KMAX / 32; ++q) {
    const int i = q * 32 + lane;
    moved[q] = (i > pos && i < newlen) ? list[i - 1] : 0ull;
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < KMAX / 32; ++q) {
    const int i = q * 32 + lane;
    if (i > pos && i < newlen) list[i] = moved[q];
  }
  if (lane == 0) list[pos] = c;
  len = newlen;
  __syncwarp();
}

__device__ __forceinline__ float fracf_signed(float x) { return x - truncf(x); }   // torch.frac

struct ScoreArgs {
  kge_model_t m;
  const int64_t* heads;
  const int64_t* rels;   // NULL: the user->item relation row
  const int64_t* tails;  // predict only
  int64_t n;
  int head_is_user;
  int rel_row;  // row used when rels == NULL
};


// q(head, relation) at column c of the embedding (both parts for the complex models):
//   TransE   q = h + r                     (transe.py:55-57)
//   DistMult q = h * r                     (distmult.py:50-51)
//   RotatE   q = rot(h, theta) = (re | im) (rotate.py:61-66)
//   ComplEx  q = (hr*rr | hi*rr + hr*ri - hi*ri)   (complex.py:53-62 regrouped by tail part)
//   TorusE   q = frac(h) + frac(r)         (toruse.py:66-76; torch.frac keeps the sign: x - trunc(x))
//   TransH   q = h * (1 - sum(w) * w) + r  (transh.py:53-58, 73-74; `proj` = the factor vector of the relation, built
//                                           by the caller -- every query row shares the user->item relation)
__device__ __forceinline__ void query_value(const ScoreArgs& a, int64_t qrow, int c, float& q0, float& q1,
                                            const float* proj = nullptr) {
  const int d = a.m.d;
  const int model = a.m.model;
  const kge_table_t& HT = a.head_is_user ? a.m.user : a.m.entity;
  const int64_t h_id = __ldg(a.heads + qrow);
  const int64_t r_id = a.rels ? __ldg(a.rels + qrow) : (int64_t)a.rel_row;
  const float h0 = __ldg(HT.w[0] + h_id * d + c);
  const float r0 = __ldg(a.m.relation.w[0] + r_id * d + c);
  q1 = 0.f;
  if (model == KGE_TRANSE) {
    q0 = h0 + r0;
  } else if (model == KGE_TORUSE) {
    q0 = fracf_signed(h0) + fracf_signed(r0);
  } else if (model == KGE_TRANSH) {
    q0 = h0 * proj[c] + r0;
  } else if (model == KGE_DISTMULT) {
    q0 = h0 * r0;
  } else if (model == KGE_ROTATE) {
    const float h1 = __ldg(HT.w[1] + h_id * d + c);
    float sn, cs;
    sincosf(r0, &sn, &cs);
    q0 = cs * h0 - sn * h1;
    q1 = cs * h1 + sn * h0;
  } else {
    const float h1 = __ldg(HT.w[1] + h_id * d + c);
    const float r1 = __ldg(a.m.relation.w[1] + r_id * d + c);
    q0 = h0 * r0;
    q1 = h1 * r0 + h0 * r1 - h1 * r1;
  }
}

}  // namespace
