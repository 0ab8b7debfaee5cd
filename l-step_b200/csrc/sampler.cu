// K1 — most-recent-K temporal neighbour lookup over a time-sorted CSR.
// Replaces NeighborSampler.get_historical_neighbors('recent') + find_neighbors_before
// (/root/reference/utils/utils.py:129-146, 148-213).
//
// One warp per query row. The strictly-earlier count c = |{j : t[j] < tq}| (np.searchsorted
// side='left', compared in fp64) is found by a warp-cooperative 32-ary search: each step the 32
// lanes probe 32 interior pivots of the remaining range, a ballot gives the number of pivots
// below tq, and the range shrinks 33x (a hub with 10^5 entries needs 3 steps + a final 32-wide
// probe). Lanes then copy the last min(K,c) entries right-aligned and zero-fill the left.
#include "common.cuh"

namespace lstep {

template <typename IdT, bool kWithEid, bool kHook>
__global__ void __launch_bounds__(256) sample_recent_kernel(const int64_t* __restrict__ indptr,
                                                            const int32_t* __restrict__ c_nbr,
                                                            const int32_t* __restrict__ c_eid,
                                                            const double* __restrict__ c_t, int64_t num_rows, RowIds q_node,
                                                            const double* __restrict__ q_time, int64_t n_rows,
                                                            int64_t n_valid, int K, IdT* __restrict__ out_nbr,
                                                            IdT* __restrict__ out_eid, float* __restrict__ out_t,
                                                            uint32_t* err_flag, PhaseBHook hook) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  if (kHook && blockIdx.x == 0)
    for (int c = threadIdx.x; c < hook.d; c += blockDim.x) hook.pe0[c] = 0.f;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;

  int64_t end = 0;  // one past the last strictly-earlier entry
  int64_t cnt = 0;
  if (row < n_valid) {
    const int64_t node = q_node.at(row);
    if (node < 0 || node >= num_rows) {
      if (lane == 0 && err_flag) atomicOr(err_flag, LSTEP_FLAG_NODE_OUT_OF_RANGE);
    } else {
      const int64_t lo = indptr[node];
      int64_t a = lo, b = indptr[node + 1];
      const double tq = q_time[q_node.time_index(row)];
      // invariant: entries < a are earlier than tq, entries >= b are not
      while (b - a > 32) {
        const int64_t len = b - a;
        const int64_t p = a + (len * (lane + 1)) / 33;
        const int j = __popc(__ballot_sync(kFull, c_t[p] < tq));
        const int64_t pa = a + (len * j) / 33, pb = a + (len * (j + 1)) / 33;
        if (j < 32) b = pb;
        if (j > 0) a = pa + 1;
      }
      const int64_t p = a + lane;
      const bool less = (p < b) && (c_t[p] < tq);
      end = a + __popc(__ballot_sync(kFull, less));
      cnt = end - lo;
    }
  }
  const int take = (int)(cnt < K ? cnt : K);
  const int pad = K - take;
  const int64_t obase = row * (int64_t)K;
  const int64_t first = end - take;
  bool has_zero = false;
  for (int k = lane; k < K; k += 32) {
    IdT n = 0, e = 0;
    float tt = 0.f;
    if (k >= pad) {
      const int64_t j = first + (k - pad);
      n = (IdT)c_nbr[j];
      if (kWithEid) e = (IdT)c_eid[j];
      tt = (float)c_t[j];  // f64 -> f32 round-to-nearest, as the numpy store does (utils.py:166,208)
    }
    out_nbr[obase + k] = n;
    if (kWithEid) out_eid[obase + k] = e;
    out_t[obase + k] = tt;
    if (kHook) {
      if (n > 0) {
        const int c = atomicAdd(hook.cnt_of + n, 1);
        hook.rank[obase + k] = c;
        if (c == 0) hook.U[atomicAdd(hook.counters + 0, 1)] = (int64_t)n;
      } else {
        has_zero = true;
      }
    }
  }
  if (kHook) {
    if (__any_sync(kFull, has_zero) && lane == 0) hook.counters[1] = 1;
  }
}

template <typename IdT, bool kWithEid>
int launch_sample(const lstep_csr* csr, RowIds q_node, const double* q_time, int64_t n_rows,
                         int64_t n_valid, int K, IdT* out_nbr, IdT* out_eid, float* out_t, uint32_t* err_flag,
                         void* stream) {
  if (!csr || K <= 0 || n_rows < 0 || n_valid < 0) return LSTEP_ERR_INVALID_ARG;
  if (n_rows == 0) return LSTEP_OK;
  if (!out_nbr || !out_t || (n_valid > 0 && (!q_node.p[0] || !q_time))) return LSTEP_ERR_INVALID_ARG;
  if (n_valid > n_rows) n_valid = n_rows;
  const int warps = 8;
  const int64_t blocks = ceil_div(n_rows, warps);
  launch_k(sample_recent_kernel<IdT, kWithEid, false>, dim3((unsigned)blocks), dim3(warps * 32), 0, as_stream(stream), 
      csr->indptr, csr->nbr, csr->eid, csr->t, csr->num_rows, q_node, q_time, n_rows, n_valid, K, out_nbr, out_eid,
      out_t, err_flag, PhaseBHook{});
  return check_launch("sample_recent");
}

// lookup + per-destination counting for update_pe phase B (one launch instead of two)
int launch_sample_count(const lstep_csr* csr, const int64_t* q_node, const double* q_time, int64_t n_rows, int64_t n_valid,
                        int K, int32_t* out_nbr, float* out_t, uint32_t* err_flag, PhaseBHook hook, void* stream) {
  if (!csr || K <= 0 || n_rows <= 0 || !out_nbr || !out_t || !q_node || !q_time) return LSTEP_ERR_INVALID_ARG;
  if (n_valid > n_rows) n_valid = n_rows;
  const int warps = 8;
  const int64_t blocks = ceil_div(n_rows, warps);
  launch_k(sample_recent_kernel<int32_t, false, true>, dim3((unsigned)blocks), dim3(warps * 32), 0, as_stream(stream), 
      csr->indptr, csr->nbr, csr->eid, csr->t, csr->num_rows, single_ids(q_node), q_time, n_rows, n_valid, K, out_nbr,
      nullptr, out_t, err_flag, hook);
  return check_launch("sample_recent+count");
}

template int launch_sample<int32_t, false>(const lstep_csr*, RowIds, const double*, int64_t, int64_t, int, int32_t*, int32_t*, float*,
                                           uint32_t*, void*);

}  // namespace lstep

extern "C" int lstep_sample_recent(const lstep_csr* csr, const int64_t* q_node, const double* q_time, int64_t n_rows,
                                   int64_t n_valid, int K, int64_t* out_nbr, int64_t* out_eid, float* out_t,
                                   uint32_t* err_flag, void* stream) {
  if (out_eid)
    return lstep::launch_sample<int64_t, true>(csr, lstep::single_ids(q_node), q_time, n_rows, n_valid, K, out_nbr, out_eid, out_t,
                                               err_flag, stream);
  return lstep::launch_sample<int64_t, false>(csr, lstep::single_ids(q_node), q_time, n_rows, n_valid, K, out_nbr, nullptr, out_t,
                                              err_flag, stream);
}

extern "C" int lstep_sample_recent_compact(const lstep_csr* csr, const int64_t* q_node, const double* q_time,
                                           int64_t n_rows, int64_t n_valid, int K, int32_t* out_nbr, float* out_t,
                                           uint32_t* err_flag, void* stream) {
  return lstep::launch_sample<int32_t, false>(csr, lstep::single_ids(q_node), q_time, n_rows, n_valid, K, out_nbr, nullptr, out_t,
                                              err_flag, stream);
}
