// Batch assembly on the device: the row gathers of hopwise's loaders.
//
// Replaces (paths under /root/reference/hopwise/):
//   data/dataloader/general_dataloader.py:66-70     TrainDataLoader.collate_fn: inter_feat[index]
//   data/dataloader/knowledge_dataloader.py:69-75   KGDataLoader.collate_fn: kg_feat[index]
//   data/interaction.py:130-139                     Interaction.__getitem__: every column indexed by the same index
// One launch gathers every id column of a batch by the batch's index vector (the reference indexes the columns one
// after another on the host): out[c][i] = column[c][index[i]].  Pure HBM/L2-bound int64 traffic: each index is read
// once and reused for all columns, writes are coalesced.
#include "common.cuh"

namespace {

constexpr int MAX_COLS = 8;

struct GatherArgs {
  const int64_t* cols[MAX_COLS];
  int64_t* outs[MAX_COLS];
  const int64_t* index;
  int64_t n, rows;
  int n_cols;
  int32_t* status;  // set to 1 when an index falls outside [0, rows)
};

__global__ void __launch_bounds__(256) gather_columns_kernel(const GatherArgs a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = __ldg(a.index + i);
    if (r < 0 || r >= a.rows) {
      if (a.status) *a.status = 1;
      continue;
    }
#pragma unroll
    for (int c = 0; c < MAX_COLS; ++c)
      if (c < a.n_cols) a.outs[c][i] = __ldg(a.cols[c] + r);
  }
}

// ids staged as int32 over PCIe (every id is a row index below 2^31), widened to the int64 the kernels read
__global__ void __launch_bounds__(256) widen_ids_kernel(const int32_t* __restrict__ src, int64_t* __restrict__ dst,
                                                        int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t quads = n / 4;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += stride) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(src) + q);
    reinterpret_cast<longlong2*>(dst)[2 * q] = make_longlong2(v.x, v.y);
    reinterpret_cast<longlong2*>(dst)[2 * q + 1] = make_longlong2(v.z, v.w);
  }
  for (int64_t i = quads * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

}  // namespace

extern "C" int kge_widen_ids_i32(const int32_t* src, int64_t* dst, int64_t n, kge_stream_t stream) {
  KGE_REQUIRE(n >= 0, KGE_E_ARG, "negative n");
  if (n == 0) return 0;
  KGE_REQUIRE(src && dst, KGE_E_ARG, "NULL argument");
  KGE_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, KGE_E_ARG,
              "buffers must be 16-byte aligned");
  int64_t grid = (n / 4 + 255) / 256;
  const int64_t cap = (int64_t)kge_num_sms() * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  widen_ids_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  KGE_LAUNCH_CHECK();
  return 0;
}

extern "C" int kge_gather_columns(const int64_t* const* columns, int32_t n_columns, int64_t rows, const int64_t* index,
                                  int64_t n, int64_t* const* outs, int32_t* status, kge_stream_t stream) {
  KGE_REQUIRE(n >= 0 && rows >= 0 && n_columns >= 1 && n_columns <= MAX_COLS, KGE_E_ARG, "bad n / rows / n_columns");
  if (n == 0) return 0;
  KGE_REQUIRE(columns && outs && index, KGE_E_ARG, "NULL argument");
  GatherArgs a = {};
  for (int c = 0; c < n_columns; ++c) {
    KGE_REQUIRE(columns[c] && outs[c], KGE_E_ARG, "NULL column %d", c);
    a.cols[c] = columns[c];
    a.outs[c] = outs[c];
  }
  a.index = index;
  a.n = n;
  a.rows = rows;
  a.n_cols = n_columns;
  a.status = status;
  const int threads = 256;
  int64_t grid = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)kge_num_sms() * 8;
  if (grid > cap) grid = cap;
  gather_columns_kernel<<<(int)grid, threads, 0, (cudaStream_t)stream>>>(a);
  KGE_LAUNCH_CHECK();
  return 0;
}

// One training batch of the reference's RSKG loader in one library call: the KG half is drawn first (gather the
// triples by the batch's index vector, corrupt the tails), then the recommendation half, both samplers on the one
// MT19937 stream (knowledge_dataloader.py:131-145, general_dataloader.py:66-70, abstract_dataloader.py:185-198).
// Four launches queued back to back; what it saves over calling the pieces is host time, which bounds the loop at the
// reference's batch size.  out: head | relation | tail | neg_tail | user | item | neg_item (neg_item: neg_num per row,
// j-major), uniform candidates.
extern "C" int64_t kge_assemble_batch_workspace_bytes(int64_t n_kg, int64_t n_rec, int32_t neg_num) {
  if (n_kg < 0 || n_rec < 0 || neg_num < 1) return -1;
  const int64_t a = kge_sample_workspace_bytes(n_kg), b = kge_sample_workspace_bytes(n_rec * neg_num);
  return a > b ? a : b;
}

extern "C" int kge_assemble_batch(uint32_t* mt_state, const int64_t* kg_head, const int64_t* kg_rel,
                                  const int64_t* kg_tail, int64_t kg_rows, const int64_t* kg_index, int64_t n_kg,
                                  const int64_t* kg_used_off, const int64_t* kg_used_vals, int64_t entity_num,
                                  const int64_t* inter_user, const int64_t* inter_item, int64_t inter_rows,
                                  const int64_t* rec_index, int64_t n_rec, int32_t neg_num,
                                  const int64_t* rec_used_off, const int64_t* rec_used_vals, int64_t item_num,
                                  int64_t* out, void* workspace, kge_stream_t stream) {
  KGE_REQUIRE(n_kg >= 0 && n_rec >= 0 && neg_num >= 1 && out, KGE_E_ARG, "bad sizes / NULL out");
  int64_t* head = out;
  int64_t* neg_tail = out + 3 * n_kg;
  int64_t* user = neg_tail + n_kg;
  int64_t* neg_item = user + 2 * n_rec;
  if (n_kg > 0) {
    const int64_t* cols[3] = {kg_head, kg_rel, kg_tail};
    int64_t* outs[3] = {head, head + n_kg, head + 2 * n_kg};
    if (int e = kge_gather_columns(cols, 3, kg_rows, kg_index, n_kg, outs, nullptr, stream)) return e;
    if (int e = kge_sample_negatives(mt_state, head, n_kg, 1, kg_used_off, kg_used_vals, 1, entity_num, neg_tail, workspace,
                                     stream))
      return e;
  }
  if (n_rec > 0) {
    const int64_t* cols[2] = {inter_user, inter_item};
    int64_t* outs[2] = {user, user + n_rec};
    if (int e = kge_gather_columns(cols, 2, inter_rows, rec_index, n_rec, outs, nullptr, stream)) return e;
    if (int e = kge_sample_negatives(mt_state, user, n_rec, neg_num, rec_used_off, rec_used_vals, 1, item_num, neg_item,
                                     workspace, stream))
      return e;
  }
  return 0;
}
