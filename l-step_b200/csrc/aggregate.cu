// K2 (gather side) — neighbourhood PE aggregate of LSTEP.compute_neighborhood_pe
// (/root/reference/models/LSTEP.py:228-238):
//   S[i] = sum_k [ pe[nbr[i,k]] || (nbr[i,k] != 0) * cos(fp32(q_time[i] - nbr_t[i,k]) * w) ]
// One CTA per query row, two warp-aligned roles working concurrently: `t` threads own one time
// frequency each (K accurate cosines, never materialising the [K, d+t] concatenation), d/4
// threads own one 128-bit column group of the PE row and add the K gathered rows in order
// k = 0..K-1 (the PE table is L2 resident: 7.6 MB at Reddit size). Padded slots (id 0) read
// pe[0] like any other row — it is non-zero after an update.
#include <algorithm>

#include "common.cuh"
#include "gather_bodies.cuh"

namespace lstep {

template <int VEC, bool kLookup>
__global__ void __launch_bounds__(512) nbr_aggregate_kernel(const float* __restrict__ pe, const double* __restrict__ q_time,
                                                            const int32_t* __restrict__ nbr, const float* __restrict__ nbr_t,
                                                            int64_t n_rows, int K, const float* __restrict__ tw, int d, int t, int t_pad,
                                                            float* __restrict__ S, int64_t ldS, int64_t period, LookupArgs lk) {
  pdl_launch_dependents();
  pdl_wait();
  nbr_aggregate_rows<VEC, kLookup>(blockIdx.x, gridDim.x, false, pe, q_time, nbr, nbr_t, n_rows, K, tw, d, t, t_pad, S, ldS, period, lk, t);
}

__global__ void __launch_bounds__(256) nbr_aggregate_bwd_kernel(const float* __restrict__ dS,
                                                                const int32_t* __restrict__ nbr, int64_t n_rows, int K,
                                                                int d, int t, float* __restrict__ dpe) {
  const int in1 = d + t;
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      const float g = dS[row * in1 + c];
      for (int k = 0; k < K; ++k) atomicAdd(dpe + (int64_t)nbr[row * K + k] * d + c, g);
    }
  }
}

// a5 stand-alone: out[i][j] = time_feature(dt[i], w[j]) (the function every fused kernel calls)
__global__ void __launch_bounds__(256) time_features_kernel(const float* __restrict__ dt, int64_t n, const float* __restrict__ w, int t,
                                                            float* __restrict__ out) {
  const int64_t total = n * t;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = time_feature(dt[i / t], w[i % t]);
}

int launch_pe_mlp(const float* A, int64_t lda, const float* pe, RowIds base_ids, int64_t n_rows, int64_t expected_rows,
                  const int32_t* n_rows_dev, const lstep_pe_mlp* m, float* out, int64_t out_stride, float* pe_inplace,
                  cudaStream_t st, bool late_trigger = false);

static bool aligned16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

template <bool kLookup>
static int launch_nbr_aggregate_t(const float* pe, const double* q_time, const int32_t* nbr, const float* nbr_t, int64_t n_rows, int K,
                                  const float* tw, int d, int t, float* S, int64_t ldS, int64_t period, LookupArgs lk, cudaStream_t st) {
  const bool v4 = d % 4 == 0 && ldS % 4 == 0 && aligned16(pe) && aligned16(S);
  const int dvec = v4 ? d / 4 : d;
  const int t_pad = (int)align_up((size_t)t, 32);
  const int threads = (int)align_up((size_t)t_pad + dvec, 32);
  if (threads > 512) return LSTEP_ERR_UNSUPPORTED;
  const size_t smem = (size_t)K * 8;
  if (smem > 48 * 1024) return LSTEP_ERR_UNSUPPORTED;
  const int64_t grid = n_rows < (int64_t)num_sms() * 16 ? n_rows : (int64_t)num_sms() * 16;
  if (v4)
    launch_k(nbr_aggregate_kernel<4, kLookup>, dim3((unsigned)grid), dim3(threads), smem, st, pe, q_time, nbr, nbr_t, n_rows, K, tw, d, t,
             t_pad, S, ldS, period, lk);
  else
    launch_k(nbr_aggregate_kernel<1, kLookup>, dim3((unsigned)grid), dim3(threads), smem, st, pe, q_time, nbr, nbr_t, n_rows, K, tw, d, t,
             t_pad, S, ldS, period, lk);
  return check_launch("nbr_aggregate");
}

int launch_nbr_aggregate(const float* pe, const double* q_time, const int32_t* nbr, const float* nbr_t,
                         int64_t n_rows, int K, const float* tw, int d, int t, float* S, int64_t ldS, int64_t period,
                         cudaStream_t st) {
  return launch_nbr_aggregate_t<false>(pe, q_time, nbr, nbr_t, n_rows, K, tw, d, t, S, ldS, period, LookupArgs{}, st);
}

// lookup + aggregate in one launch (row r: node q_node.at(r), time q_time[q_node.time_index(r)])
int launch_nbr_lookup_aggregate(const lstep_csr* csr, RowIds q_node, const float* pe, const double* q_time, int64_t n_rows, int K,
                                const float* tw, int d, int t, float* S, int64_t ldS, uint32_t* err_flag, cudaStream_t st) {
  if (!csr || K <= 0 || !q_node.p[0] || !q_time) return LSTEP_ERR_INVALID_ARG;
  LookupArgs lk{csr->indptr, csr->nbr, csr->t, csr->num_rows, q_node, err_flag};
  return launch_nbr_aggregate_t<true>(pe, q_time, nullptr, nullptr, n_rows, K, tw, d, t, S, ldS, q_node.period, lk, st);
}

}  // namespace lstep

using namespace lstep;

extern "C" int lstep_nbr_aggregate(const float* pe, int64_t pe_rows, const double* q_time, const int32_t* nbr,
                                   const float* nbr_t, int64_t n_rows, int K, const float* tw, int d, int t, float* S,
                                   void* stream) {
  if (n_rows < 0 || K <= 0 || d <= 0 || t < 0 || pe_rows <= 0) return LSTEP_ERR_INVALID_ARG;
  if (n_rows == 0) return LSTEP_OK;
  if (!pe || !q_time || !nbr || !nbr_t || !S || (t > 0 && !tw)) return LSTEP_ERR_INVALID_ARG;
  return launch_nbr_aggregate(pe, q_time, nbr, nbr_t, n_rows, K, tw, d, t, S, d + t, 0, as_stream(stream));
}

extern "C" int lstep_nbr_aggregate_bwd(const float* dS, const int32_t* nbr, int64_t n_rows, int K, int d, int t,
                                       float* dpe, int64_t pe_rows, void* stream) {
  if (n_rows < 0 || K <= 0 || d <= 0 || pe_rows <= 0) return LSTEP_ERR_INVALID_ARG;
  if (n_rows == 0) return LSTEP_OK;
  if (!dS || !nbr || !dpe) return LSTEP_ERR_INVALID_ARG;
  const int64_t grid = n_rows < (int64_t)num_sms() * 8 ? n_rows : (int64_t)num_sms() * 8;
  nbr_aggregate_bwd_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(dS, nbr, n_rows, K, d, t, dpe);
  return check_launch("nbr_aggregate_bwd");
}

extern "C" int lstep_neighborhood_pe(const float* pe, int64_t pe_rows, const int64_t* q_node, const double* q_time,
                                     const int32_t* nbr, const float* nbr_t, int64_t n_rows, int K,
                                     const lstep_pe_mlp* mlp, float* out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  if (n_rows < 0 || K <= 0 || !mlp || pe_rows <= 0) return LSTEP_ERR_INVALID_ARG;
  if (n_rows == 0) return LSTEP_OK;
  if (!pe || !q_node || !q_time || !nbr || !nbr_t || !out || !workspace) return LSTEP_ERR_INVALID_ARG;
  const int64_t ldS = (int64_t)align_up((size_t)(mlp->d + mlp->t), 4);  // 16-byte aligned rows
  const size_t need = (size_t)n_rows * ldS * sizeof(float);
  if (workspace_bytes < need) return LSTEP_ERR_WORKSPACE;
  float* S = reinterpret_cast<float*>(workspace);
  int rc = launch_nbr_aggregate(pe, q_time, nbr, nbr_t, n_rows, K, mlp->tw, mlp->d, mlp->t, S, ldS, 0, as_stream(stream));
  if (rc != LSTEP_OK) return rc;
  return launch_pe_mlp(S, ldS, pe, single_ids(q_node), n_rows, n_rows, nullptr, mlp, out, mlp->d, nullptr, as_stream(stream));
}

extern "C" int lstep_nbr_lookup_aggregate(const lstep_csr* csr, const int64_t* q_node, const double* q_time, int64_t n_rows, int K,
                                          const float* pe, int64_t pe_rows, const float* tw, int d, int t, float* S,
                                          uint32_t* err_flag, void* stream) {
  if (n_rows < 0 || K <= 0 || d <= 0 || t < 0 || pe_rows <= 0) return LSTEP_ERR_INVALID_ARG;
  if (n_rows == 0) return LSTEP_OK;
  if (!csr || !pe || !q_node || !q_time || !S || (t > 0 && !tw)) return LSTEP_ERR_INVALID_ARG;
  return launch_nbr_lookup_aggregate(csr, single_ids(q_node), pe, q_time, n_rows, K, tw, d, t, S, d + t, err_flag, as_stream(stream));
}

extern "C" int lstep_time_features(const float* dt, int64_t n, const float* w, int t, float* out, void* stream) {
  if (n < 0 || t <= 0 || (n > 0 && (!dt || !w || !out))) return LSTEP_ERR_INVALID_ARG;
  if (n == 0) return LSTEP_OK;
  const int64_t blocks = std::min<int64_t>(ceil_div(n * t, 256), (int64_t)num_sms() * 8);
  time_features_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(dt, n, w, t, out);
  return check_launch("time_features");
}
