// Fused KGE training step for sm_100a: gather -> score -> loss -> analytic gradient scatter,
// then exact row-lazy Adam on the touched rows.
//
// Replaces, per step (paths under /root/reference/hopwise/):
//   model/knowledge_graph_embedding_recommender/{transe.py:75-98, distmult.py:68-95,
//   rotate.py:98-131, complex.py:95-128}   (calculate_loss: 8-16 gathers, cat, scorer, loss)
//   trainer/trainer.py:261                  (loss.backward(): dense [rows, d] gradients)
//   trainer/trainer.py:264 + torch.optim.Adam (dense update of every table)
//
// Data layout in HBM: the tables stay exactly where torch keeps them (fp32 [rows, d],
// one matrix per re/im part), so state_dict() is zero-copy.  Next to every table live m, v
// (Adam moments), g (a gradient accumulator that is all-zero between steps: only touched
// rows are ever written and the update kernel zeroes them again), last_step[rows] and a
// per-step unique-row list.  Nothing of size [rows, d] is traversed per step.
//
// Kernel A (train_fwd_kernel): a group of G lanes owns one positive triple; rows move as
// 128-bit loads; gradients leave as RED.ADD.F32x4.  The user->item relation row is shared by
// every rec triple, so its gradient is accumulated in registers, reduced per CTA in shared
// memory and flushed once per CTA.
// Kernel B (adam_apply_kernel): one group per unique touched row: replay the zero-gradient
// steps the row skipped (dense Adam keeps moving a row after its last gradient), apply the
// real step, zero g.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace {

struct AdamDev {
  float lr, b1, b2, eps, omb1, omb2;
  int step;  // update being produced
  int cap;
  const float2* table;  // {lr/(1-b1^j), 1/sqrt(1-b2^j)}
  int tlen;
};

struct TrainArgs {
  kge_model_t m;
  kge_batch_t b;
  AdamDev adam;
  float w_rec, w_kg;       // weight of one (positive, negative) pair in the scalar loss
  float wpos_rec, wpos_kg; // BCE models: weight of the positive term (= w * k)
  int with_grad;
  float* loss;
  int cap_user, cap_entity, cap_relation;  // per-CTA capacity of the staged unique-row lists
};

__device__ __forceinline__ float2 adam_consts(const AdamDev& A, int j) {
  if (j < A.tlen) return __ldg(A.table + j);
  return make_float2(A.lr, 1.f);
}

// Zero-gradient Adam steps s+1 .. t_end on one row fragment (torch.optim.Adam with g = 0:
// m.lerp_(0, 1-b1); v.mul_(b2); p.addcdiv_(m, sqrt(v)/sqrt(1-b2^j) + eps, -lr/(1-b1^j))).
// After `cap` steps m has decayed by b1^cap (7e-10 at cap 200) and the remaining movement is
// below fp32 resolution of the weights; m and v then decay in closed form.
template <int E>
__device__ __forceinline__ void adam_replay(float (&p)[E], float (&m)[E], float (&v)[E], int s, int t_end,
                                            const AdamDev& A) {
  const int n = t_end - s;
  if (n <= 0) return;
  const int nrep = n < A.cap ? n : A.cap;
  for (int j = s + 1; j <= s + nrep; ++j) {
    const float2 c = adam_consts(A, j);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      m[e] -= A.omb1 * m[e];
      v[e] *= A.b2;
      const float den = sqrt_approx(v[e]) * c.y + A.eps;
      p[e] -= c.x * __fdividef(m[e], den);
    }
  }
  if (n > nrep) {
    const float f1 = powf(A.b1, (float)(n - nrep)), f2 = powf(A.b2, (float)(n - nrep));
#pragma unroll
    for (int e = 0; e < E; ++e) {
      m[e] *= f1;
      v[e] *= f2;
    }
  }
}

// Current value (as of step-1) of one part of a row, replaying lazily-skipped steps on the fly.
template <int VEC, int G, int NCH>
__device__ __forceinline__ void load_current(const kge_table_t& T, int part, int64_t row, int last, int d, int gl,
                                             const AdamDev& A, float (&x)[VEC * NCH]) {
  frag_load<VEC, G, NCH>(T.w[part], row, d, gl, x);
  if (last >= 0 && last < A.step - 1) {
    float m[VEC * NCH], v[VEC * NCH];
    frag_load<VEC, G, NCH>(T.m[part], row, d, gl, m);
    frag_load<VEC, G, NCH>(T.v[part], row, d, gl, v);
    adam_replay<VEC * NCH>(x, m, v, last, A.step - 1, A);
  }
}

__device__ __forceinline__ int row_last(const kge_table_t& T, int64_t row) {
  return T.last_step ? __ldg(T.last_step + row) : -1;
}

template <int G>
__device__ __forceinline__ void touch_row(const kge_table_t& T, int32_t* counter, int64_t row, int step, int gl) {
  if (gl == 0) {
    if (*reinterpret_cast<volatile int32_t*>(T.touch_step + row) != step) {
      const int old = atomicExch(T.touch_step + row, step);
      if (old != step) {
        const int pos = atomicAdd(counter, 1);
        T.uniq[pos] = (int32_t)row;
      }
    }
  }
}

// Same, but the new row goes to a per-CTA list in shared memory (table index t: 0 user, 1 entity,
// 2 relation); the CTA publishes its lists with one global atomicAdd per table at the end, so
// the unique-row counters are not hammered by one returning atomic per touched row.
struct StageLists {
  int* cnt;      // [3]
  int32_t* list[3];
};
template <int G>
__device__ __forceinline__ void touch_row_staged(const kge_table_t& T, const StageLists& S, int t, int64_t row,
                                                 int step, int gl) {
  if (gl == 0) {
    if (*reinterpret_cast<volatile int32_t*>(T.touch_step + row) != step) {
      const int old = atomicExch(T.touch_step + row, step);
      if (old != step) {
        const int pos = atomicAdd(&S.cnt[t], 1);
        S.list[t][pos] = (int32_t)row;
      }
    }
  }
}

__device__ __forceinline__ float softplusf(float z) { return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z))); }
__device__ __forceinline__ float sigmoidf(float z) { return 1.f / (1.f + expf(-z)); }

template <int MODEL, int VEC, int G, int NCH>
__global__ void __launch_bounds__(256) train_fwd_kernel(const TrainArgs a) {
  constexpr int E = VEC * NCH;
  constexpr int PH = (MODEL == KGE_ROTATE || MODEL == KGE_COMPLEX) ? 2 : 1;  // head / tail parts
  constexpr int PR = (MODEL == KGE_COMPLEX) ? 2 : 1;                          // relation parts
  extern __shared__ float s_racc[];  // [PR][d] user->item relation gradient of this CTA, then the staged lists
  __shared__ float s_loss[8];
  __shared__ int s_cnt[3];
  __shared__ int s_base[3];
  StageLists S;
  S.cnt = s_cnt;
  S.list[0] = reinterpret_cast<int32_t*>(s_racc + PR * a.m.d);
  S.list[1] = S.list[0] + a.cap_user;
  S.list[2] = S.list[1] + a.cap_entity;
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;

  const int d = a.m.d;
  const int gl = (threadIdx.x & 31) % G;
  const int groups_per_cta = blockDim.x / G;
  const int64_t n_groups = (int64_t)gridDim.x * groups_per_cta;
  const int64_t n_rec = a.b.n_rec, n_total = a.b.n_rec + a.b.n_kg;
  const int step = a.adam.step;
  int32_t* cnt = a.m.counters + (step & 1) * 4;
  const float margin = a.m.margin;

  for (int i = threadIdx.x; i < PR * d; i += blockDim.x) s_racc[i] = 0.f;
  __syncthreads();

  float lsum = 0.f;
  float racc[PR][E];
#pragma unroll
  for (int p = 0; p < PR; ++p)
#pragma unroll
    for (int e = 0; e < E; ++e) racc[p][e] = 0.f;
  bool rec_seen = false;

  for (int64_t inst = (int64_t)blockIdx.x * groups_per_cta + threadIdx.x / G; inst < n_total; inst += n_groups) {
    const bool is_rec = inst < n_rec;
    const int64_t i = is_rec ? inst : inst - n_rec;
    const int64_t n_seg = is_rec ? a.b.n_rec : a.b.n_kg;
    const int K = is_rec ? a.b.k_rec : a.b.k_kg;
    const kge_table_t& HT = is_rec ? a.m.user : a.m.entity;
    const int64_t h_id = is_rec ? __ldg(a.b.user + i) : __ldg(a.b.head + i);
    const int64_t r_id = is_rec ? (int64_t)a.m.ui_relation : __ldg(a.b.relation + i);
    const int64_t tp_id = is_rec ? __ldg(a.b.item + i) : __ldg(a.b.tail + i);
    const int64_t* negs = is_rec ? a.b.neg_item : a.b.neg_tail;
    const float w = is_rec ? a.w_rec : a.w_kg;
    const float wpos = is_rec ? a.wpos_rec : a.wpos_kg;

    float h[PH][E], r[PR][E], tp[PH][E];
    {
      const int lh = row_last(HT, h_id), lr_ = row_last(a.m.relation, r_id), lt = row_last(a.m.entity, tp_id);
#pragma unroll
      for (int p = 0; p < PH; ++p) load_current<VEC, G, NCH>(HT, p, h_id, lh, d, gl, a.adam, h[p]);
#pragma unroll
      for (int p = 0; p < PR; ++p) load_current<VEC, G, NCH>(a.m.relation, p, r_id, lr_, d, gl, a.adam, r[p]);
#pragma unroll
      for (int p = 0; p < PH; ++p) load_current<VEC, G, NCH>(a.m.entity, p, tp_id, lt, d, gl, a.adam, tp[p]);
    }

    // gradient fragments of the anchor, relation and positive tail
    float gh[PH][E], gr[PR][E], gtp[PH][E];
#pragma unroll
    for (int p = 0; p < PH; ++p)
#pragma unroll
      for (int e = 0; e < E; ++e) { gh[p][e] = 0.f; gtp[p][e] = 0.f; }
#pragma unroll
    for (int p = 0; p < PR; ++p)
#pragma unroll
      for (int e = 0; e < E; ++e) gr[p][e] = 0.f;
    bool any_grad = false;
    float inst_loss = 0.f;

    if (MODEL == KGE_TRANSE) {
      // pair loss: clamp_min(margin + ||x - tp + eps|| - ||x - tn + eps||, 0), x = h + r
      float x[E], up[E];
      float sp = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        x[e] = h[0][e] + r[0][e];
        const float dp = frag_valid<VEC, G, NCH>(d, gl, e) ? (x[e] - tp[0][e] + 1e-6f) : 0.f;
        up[e] = dp;
        sp += dp * dp;
      }
      const float np_ = sqrtf(group_sum<G>(sp));
      const float inv_p = np_ > 0.f ? 1.f / np_ : 0.f;
      // the positive and the negative term are formed by the same operations, so that a pair
      // whose negative equals its positive cancels exactly, as it does under autograd
#pragma unroll
      for (int e = 0; e < E; ++e) up[e] = w * (up[e] * inv_p);
      float nact = 0.f;
      for (int j = 0; j < K; ++j) {
        const int64_t tn_id = __ldg(negs + (int64_t)j * n_seg + i);
        float tn[E];
        load_current<VEC, G, NCH>(a.m.entity, 0, tn_id, row_last(a.m.entity, tn_id), d, gl, a.adam, tn);
        float sn = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const float dn = frag_valid<VEC, G, NCH>(d, gl, e) ? (x[e] - tn[e] + 1e-6f) : 0.f;
          tn[e] = dn;
          sn += dn * dn;
        }
        const float nn_ = sqrtf(group_sum<G>(sn));
        const float z = margin + np_ - nn_;
        if (z >= 0.f) {
          inst_loss += z * w;
          if (a.with_grad) {
            const float inv_n = nn_ > 0.f ? 1.f / nn_ : 0.f;
            nact += 1.f;
#pragma unroll
            for (int e = 0; e < E; ++e) {
              tn[e] = w * (tn[e] * inv_n);  // gradient of the negative tail
              gh[0][e] += up[e] - tn[e];
              gtp[0][e] -= up[e];
            }
            frag_atomic_add<VEC, G, NCH>(a.m.entity.g[0], tn_id, d, gl, tn);
            touch_row_staged<G>(a.m.entity, S, 1, tn_id, step, gl);
          }
        }
      }
      if (nact > 0.f) {
        any_grad = true;
#pragma unroll
        for (int e = 0; e < E; ++e) gr[0][e] = gh[0][e];
      }
    } else if (MODEL == KGE_DISTMULT) {
      // pair loss: clamp_min(margin - s+ + s-, 0), s = sum h*r*t
      float q[E], dacc[E];
      float sp = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        q[e] = h[0][e] * r[0][e];
        sp += q[e] * tp[0][e];
        dacc[e] = 0.f;
      }
      sp = group_sum<G>(sp);
      float nact = 0.f;
      for (int j = 0; j < K; ++j) {
        const int64_t tn_id = __ldg(negs + (int64_t)j * n_seg + i);
        float tn[E];
        load_current<VEC, G, NCH>(a.m.entity, 0, tn_id, row_last(a.m.entity, tn_id), d, gl, a.adam, tn);
        float sn = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) sn += q[e] * tn[e];
        sn = group_sum<G>(sn);
        const float z = margin - sp + sn;
        if (z >= 0.f) {
          inst_loss += z * w;
          if (a.with_grad) {
            nact += 1.f;
            float gtn[E];
#pragma unroll
            for (int e = 0; e < E; ++e) {
              dacc[e] += w * (tn[e] - tp[0][e]);
              gtn[e] = w * q[e];
            }
            frag_atomic_add<VEC, G, NCH>(a.m.entity.g[0], tn_id, d, gl, gtn);
            touch_row_staged<G>(a.m.entity, S, 1, tn_id, step, gl);
          }
        }
      }
      if (nact > 0.f) {
        any_grad = true;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          gh[0][e] = r[0][e] * dacc[e];
          gr[0][e] = h[0][e] * dacc[e];
          gtp[0][e] = -nact * w * q[e];
        }
      }
    } else if (MODEL == KGE_ROTATE) {
      // score = margin - || (rot(h, theta) - t) ||_2 over the stacked (re, im) vector; BCE with logits
      float cs[E], sn_[E], rre[E], rim[E], qre[E], qim[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        sincosf(r[0][e], &sn_[e], &cs[e]);
        rre[e] = cs[e] * h[0][e] - sn_[e] * h[1][e];
        rim[e] = cs[e] * h[1][e] + sn_[e] * h[0][e];
        qre[e] = 0.f;
        qim[e] = 0.f;
      }
      for (int j = -1; j < K; ++j) {
        const bool pos = j < 0;
        const int64_t t_id = pos ? tp_id : __ldg(negs + (int64_t)j * n_seg + i);
        float tre[E], tim[E];
        if (pos) {
#pragma unroll
          for (int e = 0; e < E; ++e) { tre[e] = tp[0][e]; tim[e] = tp[1][e]; }
        } else {
          const int lt = row_last(a.m.entity, t_id);
          load_current<VEC, G, NCH>(a.m.entity, 0, t_id, lt, d, gl, a.adam, tre);
          load_current<VEC, G, NCH>(a.m.entity, 1, t_id, lt, d, gl, a.adam, tim);
        }
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          tre[e] = rre[e] - tre[e];  // residual; padding lanes are 0 - 0
          tim[e] = rim[e] - tim[e];
          ss += tre[e] * tre[e] + tim[e] * tim[e];
        }
        const float nrm = sqrtf(group_sum<G>(ss));
        const float sc = margin - nrm;
        float dl;
        if (pos) { inst_loss += wpos * softplusf(-sc); dl = -wpos * sigmoidf(-sc); }
        else     { inst_loss += w * softplusf(sc);     dl =  w * sigmoidf(sc); }
        if (a.with_grad) {
          any_grad = true;
          const float f = nrm > 0.f ? -dl / nrm : 0.f;  // dL/d(residual) = dl * (-e / nrm)
#pragma unroll
          for (int e = 0; e < E; ++e) {
            tre[e] *= f;
            tim[e] *= f;
            qre[e] += tre[e];
            qim[e] += tim[e];
            tre[e] = -tre[e];  // gradient of the tail
            tim[e] = -tim[e];
          }
          if (pos) {
#pragma unroll
            for (int e = 0; e < E; ++e) { gtp[0][e] = tre[e]; gtp[1][e] = tim[e]; }
          } else {
            frag_atomic_add<VEC, G, NCH>(a.m.entity.g[0], t_id, d, gl, tre);
            frag_atomic_add<VEC, G, NCH>(a.m.entity.g[1], t_id, d, gl, tim);
            touch_row_staged<G>(a.m.entity, S, 1, t_id, step, gl);
          }
        }
      }
      if (any_grad) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          gh[0][e] = cs[e] * qre[e] + sn_[e] * qim[e];
          gh[1][e] = -sn_[e] * qre[e] + cs[e] * qim[e];
          gr[0][e] = -qre[e] * rim[e] + qim[e] * rre[e];
        }
      }
    } else {
      // ComplEx as written in the reference: s = sum tr*A + ti*B,
      // A = hr*rr, B = hi*rr + hr*ri - hi*ri (the 4th term pairs with tail_im)
      float A_[E], B_[E], Tr[E], Ti[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        A_[e] = h[0][e] * r[0][e];
        B_[e] = h[1][e] * r[0][e] + h[0][e] * r[1][e] - h[1][e] * r[1][e];
        Tr[e] = 0.f;
        Ti[e] = 0.f;
      }
      for (int j = -1; j < K; ++j) {
        const bool pos = j < 0;
        const int64_t t_id = pos ? tp_id : __ldg(negs + (int64_t)j * n_seg + i);
        float tre[E], tim[E];
        if (pos) {
#pragma unroll
          for (int e = 0; e < E; ++e) { tre[e] = tp[0][e]; tim[e] = tp[1][e]; }
        } else {
          const int lt = row_last(a.m.entity, t_id);
          load_current<VEC, G, NCH>(a.m.entity, 0, t_id, lt, d, gl, a.adam, tre);
          load_current<VEC, G, NCH>(a.m.entity, 1, t_id, lt, d, gl, a.adam, tim);
        }
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) ss += tre[e] * A_[e] + tim[e] * B_[e];
        const float sc = group_sum<G>(ss);
        float dl;
        if (pos) { inst_loss += wpos * softplusf(-sc); dl = -wpos * sigmoidf(-sc); }
        else     { inst_loss += w * softplusf(sc);     dl =  w * sigmoidf(sc); }
        if (a.with_grad) {
          any_grad = true;
#pragma unroll
          for (int e = 0; e < E; ++e) {
            Tr[e] += dl * tre[e];
            Ti[e] += dl * tim[e];
            tre[e] = dl * A_[e];
            tim[e] = dl * B_[e];
          }
          if (pos) {
#pragma unroll
            for (int e = 0; e < E; ++e) { gtp[0][e] = tre[e]; gtp[1][e] = tim[e]; }
          } else {
            frag_atomic_add<VEC, G, NCH>(a.m.entity.g[0], t_id, d, gl, tre);
            frag_atomic_add<VEC, G, NCH>(a.m.entity.g[1], t_id, d, gl, tim);
            touch_row_staged<G>(a.m.entity, S, 1, t_id, step, gl);
          }
        }
      }
      if (any_grad) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          gh[0][e] = r[0][e] * Tr[e] + r[1][e] * Ti[e];
          gh[1][e] = r[0][e] * Ti[e] - r[1][e] * Ti[e];
          gr[0][e] = h[0][e] * Tr[e] + h[1][e] * Ti[e];
          gr[1][e] = h[0][e] * Ti[e] - h[1][e] * Ti[e];
        }
      }
    }

    if (gl == 0) lsum += inst_loss;

    if (any_grad) {
#pragma unroll
      for (int p = 0; p < PH; ++p) {
        frag_atomic_add<VEC, G, NCH>(HT.g[p], h_id, d, gl, gh[p]);
        frag_atomic_add<VEC, G, NCH>(a.m.entity.g[p], tp_id, d, gl, gtp[p]);
      }
      touch_row_staged<G>(HT, S, is_rec ? 0 : 1, h_id, step, gl);
      touch_row_staged<G>(a.m.entity, S, 1, tp_id, step, gl);
      if (is_rec) {
        rec_seen = true;
#pragma unroll
        for (int p = 0; p < PR; ++p)
#pragma unroll
          for (int e = 0; e < E; ++e) racc[p][e] += gr[p][e];
      } else {
#pragma unroll
        for (int p = 0; p < PR; ++p) frag_atomic_add<VEC, G, NCH>(a.m.relation.g[p], r_id, d, gl, gr[p]);
        touch_row_staged<G>(a.m.relation, S, 2, r_id, step, gl);
      }
    }
  }

  // ---- CTA epilogue: user->item relation gradient and the loss --------------------------------
  if (rec_seen) {
#pragma unroll
    for (int p = 0; p < PR; ++p) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int j = e / VEC;
        const int col = (gl + j * G) * VEC + (e % VEC);
        if (col < d) atomicAdd(&s_racc[p * d + col], racc[p][e]);
      }
    }
  }
  const int any_rec = __syncthreads_or(rec_seen ? 1 : 0);
  if (any_rec) {
    for (int i = threadIdx.x; i < PR * d; i += blockDim.x) {
      const int p = i / d, col = i - p * d;
      atomicAdd(a.m.relation.g[p] + (int64_t)a.m.ui_relation * d + col, s_racc[i]);
    }
    if (threadIdx.x == 0) touch_row_staged<1>(a.m.relation, S, 2, a.m.ui_relation, step, 0);
  }
  // ---- publish the staged unique-row lists: one returning atomic per table per CTA -----------------
  __syncthreads();
  if (threadIdx.x < 3) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(cnt + threadIdx.x, s_cnt[threadIdx.x]) : 0;
  __syncthreads();
  {
    int32_t* dst[3] = {a.m.user.uniq, a.m.entity.uniq, a.m.relation.uniq};
#pragma unroll
    for (int t = 0; t < 3; ++t)
      for (int i = threadIdx.x; i < s_cnt[t]; i += blockDim.x) dst[t][s_base[t] + i] = S.list[t][i];
  }
  lsum = warp_sum(lsum);
  if ((threadIdx.x & 31) == 0) s_loss[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_loss[i];
    if (t != 0.f) atomicAdd(a.loss, t);
  }
}

// ---- kernel B: exact lazy Adam on the touched rows ------------------------------------------
struct ApplyArgs {
  kge_model_t m;
  AdamDev adam;
  float scale;
  const float* scale_dev;  // optional device-side factor (the incoming grad of the loss)
};

template <int VEC, int G, int NCH>
__device__ __forceinline__ void adam_row(const kge_table_t& T, int64_t row, int d, int gl, const AdamDev& A,
                                         float scale) {
  constexpr int E = VEC * NCH;
  const int last = T.last_step[row];
  const float2 c = adam_consts(A, A.step);
  for (int part = 0; part < T.parts; ++part) {
    float p[E], m[E], v[E], g[E];
    frag_load<VEC, G, NCH>(T.w[part], row, d, gl, p);
    frag_load<VEC, G, NCH>(T.m[part], row, d, gl, m);
    frag_load<VEC, G, NCH>(T.v[part], row, d, gl, v);
    frag_load_cg<VEC, G, NCH>(T.g[part], row, d, gl, g);
    if (last >= 0 && last < A.step - 1) adam_replay<E>(p, m, v, last, A.step - 1, A);
    float z[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float ge = g[e] * scale;
      m[e] += A.omb1 * (ge - m[e]);
      v[e] = v[e] * A.b2 + A.omb2 * ge * ge;
      const float den = sqrtf(v[e]) * c.y + A.eps;
      p[e] -= c.x * (m[e] / den);
      z[e] = 0.f;
    }
    frag_store<VEC, G, NCH>(T.w[part], row, d, gl, p);
    frag_store<VEC, G, NCH>(T.m[part], row, d, gl, m);
    frag_store<VEC, G, NCH>(T.v[part], row, d, gl, v);
    frag_store<VEC, G, NCH>(T.g[part], row, d, gl, z);
  }
  __syncwarp(group_mask<G>());  // every lane has read last_step
  if (gl == 0) T.last_step[row] = A.step;
}

template <int VEC, int G, int NCH>
__global__ void __launch_bounds__(256) adam_apply_kernel(const ApplyArgs a) {
  const int gl = (threadIdx.x & 31) % G;
  const int groups_per_cta = blockDim.x / G;
  const int64_t n_groups = (int64_t)gridDim.x * groups_per_cta;
  const int32_t* cnt = a.m.counters + (a.adam.step & 1) * 4;
  const int64_t cu = cnt[0], ce = cnt[1], cr = cnt[2];
  const int64_t total = cu + ce + cr;
  const float scale = a.scale_dev ? a.scale * __ldg(a.scale_dev) : a.scale;
  for (int64_t idx = (int64_t)blockIdx.x * groups_per_cta + threadIdx.x / G; idx < total; idx += n_groups) {
    if (idx < cu) adam_row<VEC, G, NCH>(a.m.user, a.m.user.uniq[idx], a.m.d, gl, a.adam, scale);
    else if (idx < cu + ce) adam_row<VEC, G, NCH>(a.m.entity, a.m.entity.uniq[idx - cu], a.m.d, gl, a.adam, scale);
    else adam_row<VEC, G, NCH>(a.m.relation, a.m.relation.uniq[idx - cu - ce], a.m.d, gl, a.adam, scale);
  }
  // hand the next step a zeroed set of counters (the other parity)
  if (blockIdx.x == 0 && threadIdx.x < 4) a.m.counters[((a.adam.step + 1) & 1) * 4 + threadIdx.x] = 0;
}

// ---- dense catch-up of every row to adam.step (before weights are read by others) ------------
template <int VEC, int G, int NCH>
__global__ void __launch_bounds__(256) adam_flush_kernel(const kge_table_t T, int d, const AdamDev A) {
  constexpr int E = VEC * NCH;
  const int gl = (threadIdx.x & 31) % G;
  const int groups_per_cta = blockDim.x / G;
  const int64_t n_groups = (int64_t)gridDim.x * groups_per_cta;
  for (int64_t row = (int64_t)blockIdx.x * groups_per_cta + threadIdx.x / G; row < T.rows; row += n_groups) {
    const int last = T.last_step[row];
    if (last < 0 || last >= A.step) continue;
    for (int part = 0; part < T.parts; ++part) {
      float p[E], m[E], v[E];
      frag_load<VEC, G, NCH>(T.w[part], row, d, gl, p);
      frag_load<VEC, G, NCH>(T.m[part], row, d, gl, m);
      frag_load<VEC, G, NCH>(T.v[part], row, d, gl, v);
      adam_replay<E>(p, m, v, last, A.step, A);
      frag_store<VEC, G, NCH>(T.w[part], row, d, gl, p);
      frag_store<VEC, G, NCH>(T.m[part], row, d, gl, m);
      frag_store<VEC, G, NCH>(T.v[part], row, d, gl, v);
    }
    __syncwarp(group_mask<G>());
    if (gl == 0) T.last_step[row] = A.step;
  }
}

// ---- discard / pack / add (row-sparse exchange) --------------------------------------------------
template <int VEC, int G, int NCH>
__global__ void __launch_bounds__(256) grad_take_kernel(const kge_table_t T, int d, const int32_t* counter,
                                                        int64_t* ids_out, float* rows_out) {
  constexpr int E = VEC * NCH;
  const int gl = (threadIdx.x & 31) % G;
  const int groups_per_cta = blockDim.x / G;
  const int64_t n_groups = (int64_t)gridDim.x * groups_per_cta;
  const int64_t total = *counter;
  for (int64_t idx = (int64_t)blockIdx.x * groups_per_cta + threadIdx.x / G; idx < total; idx += n_groups) {
    const int64_t row = T.uniq[idx];
    for (int part = 0; part < T.parts; ++part) {
      float g[E], z[E];
      frag_load_cg<VEC, G, NCH>(T.g[part], row, d, gl, g);
#pragma unroll
      for (int e = 0; e < E; ++e) z[e] = 0.f;
      frag_store<VEC, G, NCH>(T.g[part], row, d, gl, z);
      if (rows_out) frag_store<VEC, G, NCH>(rows_out + (int64_t)part * d, idx * T.parts, d, gl, g);
    }
    if (gl == 0) {
      T.touch_step[row] = -1;
      if (ids_out) ids_out[idx] = row;
    }
  }
}

template <int VEC, int G, int NCH>
__global__ void __launch_bounds__(256) grad_add_kernel(const kge_table_t T, int d, int step, int32_t* counter,
                                                       const int64_t* ids, const float* rows, const int32_t* count_dev) {
  constexpr int E = VEC * NCH;
  const int gl = (threadIdx.x & 31) % G;
  const int groups_per_cta = blockDim.x / G;
  const int64_t n_groups = (int64_t)gridDim.x * groups_per_cta;
  const int64_t total = *count_dev;
  for (int64_t idx = (int64_t)blockIdx.x * groups_per_cta + threadIdx.x / G; idx < total; idx += n_groups) {
    const int64_t row = ids[idx];
    for (int part = 0; part < T.parts; ++part) {
      float g[E], x[E];
      frag_load_cg<VEC, G, NCH>(T.g[part], row, d, gl, g);
      frag_load<VEC, G, NCH>(rows + (int64_t)part * d, idx * T.parts, d, gl, x);
#pragma unroll
      for (int e = 0; e < E; ++e) g[e] += x[e];
      frag_store<VEC, G, NCH>(T.g[part], row, d, gl, g);
    }
    touch_row<G>(T, counter, row, step, gl);
  }
}

AdamDev make_adam_dev(const kge_model_t* m, const kge_adam_t* a) {
  AdamDev A;
  A.lr = a->lr;
  A.b1 = a->beta1;
  A.b2 = a->beta2;
  A.eps = a->eps;
  A.omb1 = (float)(1.0 - (double)a->beta1);
  A.omb2 = (float)(1.0 - (double)a->beta2);
  A.step = a->step;
  A.cap = a->replay_cap > 0 ? a->replay_cap : 200;
  A.table = reinterpret_cast<const float2*>(m->adam_table);
  A.tlen = m->adam_table ? m->adam_table_len : 0;
  return A;
}

int grid_for(int64_t n_groups_needed, int groups_per_cta, int ctas_per_sm) {
  int64_t g = (n_groups_needed + groups_per_cta - 1) / groups_per_cta;
  const int64_t cap = (int64_t)kge_num_sms() * ctas_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

bool table_has_state(const kge_table_t& T) {
  for (int p = 0; p < T.parts; ++p)
    if (!T.m[p] || !T.v[p] || !T.g[p]) return false;
  return T.last_step && T.touch_step && T.uniq;
}

int check_model(const kge_model_t* m, bool need_state) {
  KGE_REQUIRE(m, KGE_E_ARG, "model is NULL");
  KGE_REQUIRE(m->model >= KGE_TRANSE && m->model <= KGE_COMPLEX, KGE_E_ARG, "unknown model kind %d", m->model);
  const int ph = (m->model == KGE_ROTATE || m->model == KGE_COMPLEX) ? 2 : 1;
  const int pr = (m->model == KGE_COMPLEX) ? 2 : 1;
  KGE_REQUIRE(m->user.parts == ph && m->entity.parts == ph && m->relation.parts == pr, KGE_E_ARG,
              "table parts do not match the model kind");
  for (int p = 0; p < ph; ++p) KGE_REQUIRE(m->user.w[p] && m->entity.w[p], KGE_E_ARG, "NULL weight table");
  for (int p = 0; p < pr; ++p) KGE_REQUIRE(m->relation.w[p], KGE_E_ARG, "NULL relation table");
  RowCfg c;
  KGE_REQUIRE(kge_pick_rowcfg(m->d, c), KGE_E_UNSUPPORTED, "embedding_size %d unsupported (max 512, or 256 when not a multiple of 4)", m->d);
  if (need_state) {
    KGE_REQUIRE(table_has_state(m->user) && table_has_state(m->entity) && table_has_state(m->relation) && m->counters,
                KGE_E_STATE, "optimiser state buffers missing");
  }
  return 0;
}

}  // namespace

// =================================== C ABI ===================================================

extern "C" int kge_adam_table_fill(float lr, float beta1, float beta2, float* out_host, int32_t len) {
  // entries beyond the table are (lr, 1): both corrections are 1 to fp32 resolution there
  const double b = beta1 > beta2 ? beta1 : beta2;
  int need = 2;
  if (b > 0.0 && b < 1.0) need = (int)ceil(log(1e-10) / log(b)) + 2;
  if (!out_host) return need;
  for (int j = 0; j < len; ++j) {
    if (j == 0) { out_host[0] = lr; out_host[1] = 1.f; continue; }
    const double bc1 = 1.0 - pow((double)beta1, j), bc2 = 1.0 - pow((double)beta2, j);
    out_host[2 * j] = (float)((double)lr / bc1);
    out_host[2 * j + 1] = (float)(1.0 / sqrt(bc2));
  }
  return need;
}

extern "C" int kge_train_forward(const kge_model_t* model, const kge_batch_t* b, const kge_adam_t* adam,
                                 int with_grad, float* loss_out, kge_stream_t stream) {
  if (int e = check_model(model, with_grad != 0)) return e;
  KGE_REQUIRE(b && adam && loss_out, KGE_E_ARG, "NULL batch / adam / loss_out");
  KGE_REQUIRE(b->n_rec >= 0 && b->n_kg >= 0 && b->k_rec >= 1 && b->k_kg >= 1, KGE_E_ARG, "bad batch sizes");
  KGE_REQUIRE(adam->step >= 1, KGE_E_ARG, "adam.step is 1-based");
  const int64_t n_total = b->n_rec + b->n_kg;
  if (n_total == 0) return 0;
  KGE_REQUIRE(b->n_rec == 0 || (b->user && b->item && b->neg_item), KGE_E_ARG, "NULL rec id array");
  KGE_REQUIRE(b->n_kg == 0 || (b->head && b->relation && b->tail && b->neg_tail), KGE_E_ARG, "NULL KG id array");
  const bool lazy = model->user.last_step && model->entity.last_step && model->relation.last_step;
  KGE_REQUIRE(!lazy || (model->user.m[0] && model->entity.m[0]), KGE_E_STATE, "last_step without moments");

  TrainArgs a;
  a.m = *model;
  a.b = *b;
  a.adam = make_adam_dev(model, adam);
  a.with_grad = with_grad;
  a.loss = loss_out;
  const double pr = (double)b->n_rec * b->k_rec, pk = (double)b->n_kg * b->k_kg;
  if (model->model == KGE_TRANSE || model->model == KGE_DISTMULT) {
    a.w_rec = a.w_kg = (float)(1.0 / (pr + pk));
    a.wpos_rec = a.wpos_kg = 0.f;
  } else {
    a.w_rec = pr > 0 ? (float)(0.5 / pr) : 0.f;
    a.w_kg = pk > 0 ? (float)(0.5 / pk) : 0.f;
    a.wpos_rec = pr > 0 ? (float)(0.5 / (double)b->n_rec) : 0.f;
    a.wpos_kg = pk > 0 ? (float)(0.5 / (double)b->n_kg) : 0.f;
  }
  RowCfg c;
  kge_pick_rowcfg(model->d, c);
  const int threads = 256;
  const int gpc = threads / c.g;
  int grid = grid_for(n_total, gpc, 8);
  // per-CTA staged lists must hold every row the CTA can touch; grow the grid until they fit 48 KB
  const int kmax = b->k_rec > b->k_kg ? b->k_rec : b->k_kg;
  int64_t inst_per_cta;
  for (;;) {
    const int64_t iters = (n_total + (int64_t)grid * gpc - 1) / ((int64_t)grid * gpc);
    inst_per_cta = iters * gpc;
    if (inst_per_cta * (3 + kmax) * 4 <= 48 * 1024 || inst_per_cta <= gpc) break;
    grid *= 2;
  }
  a.cap_user = (int)inst_per_cta;
  a.cap_entity = (int)(inst_per_cta * (2 + kmax));
  a.cap_relation = (int)inst_per_cta + 1;
  const size_t smem = (size_t)model->relation.parts * model->d * sizeof(float) +
                      ((size_t)a.cap_user + a.cap_entity + a.cap_relation) * sizeof(int32_t);
  KGE_REQUIRE(smem <= 200 * 1024, KGE_E_UNSUPPORTED, "%d negatives per triple need %zu bytes of shared memory", kmax, smem);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_FWD(M, V, G, N)                                                                                  \
  do {                                                                                                          \
    if (smem > 48 * 1024)                                                                                       \
      KGE_CUDA(cudaFuncSetAttribute(train_fwd_kernel<M, V, G, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)smem));                                                                \
    train_fwd_kernel<M, V, G, N><<<grid, threads, smem, st>>>(a);                                               \
  } while (0)
#define CALL(V, G, N)                                                \
  switch (model->model) {                                            \
    case KGE_TRANSE: LAUNCH_FWD(KGE_TRANSE, V, G, N); break;         \
    case KGE_DISTMULT: LAUNCH_FWD(KGE_DISTMULT, V, G, N); break;     \
    case KGE_ROTATE: LAUNCH_FWD(KGE_ROTATE, V, G, N); break;         \
    default: LAUNCH_FWD(KGE_COMPLEX, V, G, N); break;                \
  }
  KGE_DISPATCH_ROWCFG(c, CALL);
#undef CALL
#undef LAUNCH_FWD
  KGE_LAUNCH_CHECK();
  return 0;
}

extern "C" int kge_adam_apply(const kge_model_t* model, const kge_adam_t* adam, float grad_scale,
                              const float* grad_scale_dev, kge_stream_t stream) {
  if (int e = check_model(model, true)) return e;
  KGE_REQUIRE(adam && adam->step >= 1, KGE_E_ARG, "bad adam");
  ApplyArgs a;
  a.m = *model;
  a.adam = make_adam_dev(model, adam);
  a.scale = grad_scale;
  a.scale_dev = grad_scale_dev;
  RowCfg c;
  kge_pick_rowcfg(model->d, c);
  const int threads = 256;
  const int grid = kge_num_sms() * 4;
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(V, G, N) adam_apply_kernel<V, G, N><<<grid, threads, 0, st>>>(a)
  KGE_DISPATCH_ROWCFG(c, CALL);
#undef CALL
  KGE_LAUNCH_CHECK();
  return 0;
}

extern "C" int kge_adam_flush(const kge_model_t* model, const kge_adam_t* adam, kge_stream_t stream) {
  if (int e = check_model(model, true)) return e;
  KGE_REQUIRE(adam && adam->step >= 0, KGE_E_ARG, "bad adam");
  if (adam->step == 0) return 0;
  AdamDev A = make_adam_dev(model, adam);
  RowCfg c;
  kge_pick_rowcfg(model->d, c);
  const int threads = 256;
  cudaStream_t st = (cudaStream_t)stream;
  const kge_table_t* tabs[3] = {&model->user, &model->entity, &model->relation};
  for (int t = 0; t < 3; ++t) {
    const int grid = grid_for(tabs[t]->rows, threads / c.g, 8);
#define CALL(V, G, N) adam_flush_kernel<V, G, N><<<grid, threads, 0, st>>>(*tabs[t], model->d, A)
    KGE_DISPATCH_ROWCFG(c, CALL);
#undef CALL
    KGE_LAUNCH_CHECK();
  }
  return 0;
}

static int grad_take(const kge_model_t* model, int which, int step, int64_t* ids_out, float* rows_out,
                     int32_t* count_out, cudaStream_t st) {
  const kge_table_t* tabs[3] = {&model->user, &model->entity, &model->relation};
  const kge_table_t& T = *tabs[which];
  int32_t* counter = model->counters + (step & 1) * 4 + which;
  RowCfg c;
  kge_pick_rowcfg(model->d, c);
  const int threads = 256;
  const int grid = kge_num_sms() * 4;
  if (count_out) KGE_CUDA(cudaMemcpyAsync(count_out, counter, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
#define CALL(V, G, N) grad_take_kernel<V, G, N><<<grid, threads, 0, st>>>(T, model->d, counter, ids_out, rows_out)
  KGE_DISPATCH_ROWCFG(c, CALL);
#undef CALL
  KGE_LAUNCH_CHECK();
  KGE_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
  return 0;
}

extern "C" int kge_grad_discard(const kge_model_t* model, int32_t step, kge_stream_t stream) {
  if (int e = check_model(model, true)) return e;
  for (int t = 0; t < 3; ++t)
    if (int e = grad_take(model, t, step, nullptr, nullptr, nullptr, (cudaStream_t)stream)) return e;
  return 0;
}

extern "C" int kge_grad_pack(const kge_model_t* model, int32_t which, int32_t step, int64_t* ids_out,
                             float* rows_out, int32_t* count_out, kge_stream_t stream) {
  if (int e = check_model(model, true)) return e;
  KGE_REQUIRE(which >= 0 && which < 3 && ids_out && rows_out && count_out, KGE_E_ARG, "bad pack arguments");
  return grad_take(model, which, step, ids_out, rows_out, count_out, (cudaStream_t)stream);
}

extern "C" int kge_grad_add(const kge_model_t* model, int32_t which, int32_t step, const int64_t* ids,
                            const float* rows, const int32_t* count_dev, int64_t max_count, kge_stream_t stream) {
  if (int e = check_model(model, true)) return e;
  KGE_REQUIRE(which >= 0 && which < 3 && ids && rows && count_dev && max_count >= 0, KGE_E_ARG, "bad add arguments");
  if (max_count == 0) return 0;
  const kge_table_t* tabs[3] = {&model->user, &model->entity, &model->relation};
  int32_t* counter = model->counters + (step & 1) * 4 + which;
  RowCfg c;
  kge_pick_rowcfg(model->d, c);
  const int threads = 256;
  const int grid = grid_for(max_count, threads / c.g, 4);
#define CALL(V, G, N) \
  grad_add_kernel<V, G, N><<<grid, threads, 0, (cudaStream_t)stream>>>(*tabs[which], model->d, step, counter, ids, rows, count_dev)
  KGE_DISPATCH_ROWCFG(c, CALL);
#undef CALL
  KGE_LAUNCH_CHECK();
  return 0;
}
