// Host-fed streaming step: the per-batch call a loop holding HOST (numpy) batches makes.
//
// The reference's loops hand every batch to the model as host arrays and read the predictions back
// (train_LSTEP_link_prediction.py:204-313, evaluate_model_utils.py:38-142). lstep_pe_step_host is the
// native form of that hand-over for the PE path: one call packs the batch (endpoints, times, the sorted
// unique batch nodes, the C query id sets) into a pinned slot, moves it with ONE async copy, enqueues the
// whole step (pe_step_core, csrc/step.cu), reduces the [C][n][d] neighbourhood PEs to the per-query row
// sums a caller reads back, and copies those to a pinned result slot, all on the caller's stream, without
// synchronising. lstep_host_step_result waits for a ticket's event and hands out the pinned result, so a
// loop can keep `slots - 1` steps in flight (results read one step behind hide the host's own time).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"

namespace lstep {

int pe_step_core(const lstep_pe_stream* s, const lstep_csr* csr, const int64_t* src, const int64_t* dst, const double* tq,
                 int64_t n_edges, const int64_t* ids, int64_t n_ids, double current_time, int head, int len, int append_slot,
                 const float* G, const int64_t* const* query_ids_host, int n_queries, float* nbr_out, int K,
                 const lstep_pe_mlp* mlp_nbr, const lstep_pe_mlp* mlp_upd, void* workspace, size_t workspace_bytes,
                 uint32_t* err_flag, void* stream);

// The batch moves in and the result moves out through KERNELS that read / write the pinned slots directly
// (pinned host memory is device-addressable under unified virtual addressing): a copy-engine transfer between
// two kernels costs a stream hand-over in each direction (~6-8 us each for these 14 KB / 3 KB payloads) and
// is slower than a 4-CTA copy kernel. The copy-in runs on a side stream (it overlaps the previous step; the
// compute stream waits on its event), the result kernel on another (the next step does not queue behind it).
// LSTEP_HOST_MEMCPY=1 selects cudaMemcpyAsync instead (A/B).
__global__ void __launch_bounds__(256) stage_in_kernel(const uint4* __restrict__ host_slot, uint4* __restrict__ dev_slot, int64_t n16) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x)
    dev_slot[i] = __ldcv(host_slot + i);  // the host rewrites the pinned slot between launches: never cached
}

// res[row] = sum_c x[row][c]: one warp per row, lanes stride the columns, fixed-order shuffle tree
__global__ void __launch_bounds__(256) row_sum_kernel(const float* __restrict__ x, int64_t n_rows, int d, float* __restrict__ res) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const float* p = x + row * d;
  float acc = 0.f;
  for (int c = lane; c < d; c += 32) acc += ld_dep(p + c);
  acc = warp_sum(acc);
  if (lane == 0) res[row] = acc;
}

constexpr int kMaxSlots = 8;

}  // namespace lstep

using namespace lstep;

struct lstep_host_stepper {
  int slots;
  int64_t max_edges;
  int max_queries;
  int d;
  size_t in_bytes;   // capacity of one input slot
  size_t res_floats; // capacity of one result slot
  char* h_in[kMaxSlots];
  char* d_in[kMaxSlots];
  float* h_res[kMaxSlots];
  float* d_res[kMaxSlots];
  cudaEvent_t done[kMaxSlots];       // result of the slot's step is in the pinned result slot
  cudaEvent_t in_ready[kMaxSlots];   // the slot's batch is on the device
  cudaEvent_t step_done[kMaxSlots];  // the slot's step has written nbr_out
  float* d_nbr[kMaxSlots];           // [max_queries][max_edges][d] per slot: the result kernel of step i reads it while step i+1 runs
  cudaStream_t in_stream, out_stream;  // side streams: copy-in overlaps the previous step, the result path leaves the critical path
  int64_t res_n[kMaxSlots];
  int64_t ticket_of[kMaxSlots];  // ticket currently held by the slot, -1 = none
  float* d_nbr_out;              // [max_queries][max_edges][d] when the caller keeps no copy of its own
  int64_t next_ticket;
  uint64_t h2d_bytes, d2h_bytes;
};

// sorted unique ids of src[0..n) U dst[0..n) -> out[<= 2n]. Small id spaces: a bitmap over [0, V1) scanned in
// order (1 us for 11 k nodes); otherwise (or when an id is out of range, which the device lookup reports) sort.
static int64_t unique_sorted_ids(const int64_t* src, const int64_t* dst, size_t n, int64_t V1, int64_t* out) {
  static thread_local std::vector<uint64_t> bits;
  if (V1 > 0 && V1 <= (1 << 18)) {
    const size_t words = (size_t)(V1 + 63) / 64;
    if (bits.size() < words) bits.assign(words, 0);
    bool ok = true;
    for (size_t i = 0; i < n && ok; ++i) {
      const int64_t a = src[i], b = dst[i];
      if (a < 0 || a >= V1 || b < 0 || b >= V1) {
        ok = false;
        break;
      }
      bits[(size_t)a >> 6] |= 1ull << (a & 63);
      bits[(size_t)b >> 6] |= 1ull << (b & 63);
    }
    if (ok) {
      int64_t m = 0;
      for (size_t w = 0; w < words; ++w) {
        uint64_t x = bits[w];
        if (!x) continue;
        bits[w] = 0;
        while (x) {
          out[m++] = (int64_t)(w * 64 + (size_t)__builtin_ctzll(x));
          x &= x - 1;
        }
      }
      return m;
    }
    std::fill(bits.begin(), bits.begin() + words, 0);
  }
  memcpy(out, src, 8 * n);
  memcpy(out + n, dst, 8 * n);
  std::sort(out, out + 2 * n);
  return std::unique(out, out + 2 * n) - out;
}

static int cuda_fail(cudaError_t e, const char* where) {
  set_cuda_error(e, where);
  return LSTEP_ERR_CUDA;
}

extern "C" void lstep_host_stepper_destroy(lstep_host_stepper* h) {
  if (!h) return;
  for (int i = 0; i < h->slots; ++i) {
    if (h->done[i]) cudaEventDestroy(h->done[i]);
    if (h->in_ready[i]) cudaEventDestroy(h->in_ready[i]);
    if (h->step_done[i]) cudaEventDestroy(h->step_done[i]);
    if (h->d_nbr[i]) cudaFree(h->d_nbr[i]);
    if (h->h_in[i]) cudaFreeHost(h->h_in[i]);
    if (h->h_res[i]) cudaFreeHost(h->h_res[i]);
    if (h->d_in[i]) cudaFree(h->d_in[i]);
    if (h->d_res[i]) cudaFree(h->d_res[i]);
  }
  if (h->d_nbr_out) cudaFree(h->d_nbr_out);
  if (h->in_stream) cudaStreamDestroy(h->in_stream);
  if (h->out_stream) cudaStreamDestroy(h->out_stream);
  delete h;
}

extern "C" int lstep_host_stepper_create(int slots, int64_t max_edges, int max_queries, int d, lstep_host_stepper** out) {
  if (!out || slots < 1 || slots > kMaxSlots || max_edges <= 0 || max_queries < 0 || max_queries > 8 || d <= 0)
    return LSTEP_ERR_INVALID_ARG;
  lstep_host_stepper* h = new (std::nothrow) lstep_host_stepper();
  if (!h) return LSTEP_ERR_INVALID_ARG;
  h->slots = slots;
  h->max_edges = max_edges;
  h->max_queries = max_queries;
  h->d = d;
  // src | dst | t | ids (<= 2n) | C query sets, 8 bytes each
  h->in_bytes = align_up(sizeof(int64_t) * (size_t)max_edges * (size_t)(5 + max_queries), 16);
  h->res_floats = (size_t)std::max(1, max_queries) * (size_t)max_edges;
  h->next_ticket = 0;
  h->h2d_bytes = h->d2h_bytes = 0;
  h->d_nbr_out = nullptr;
  h->in_stream = h->out_stream = nullptr;
  for (int i = 0; i < kMaxSlots; ++i) {
    h->h_in[i] = h->d_in[i] = nullptr;
    h->h_res[i] = h->d_res[i] = nullptr;
    h->done[i] = h->in_ready[i] = h->step_done[i] = nullptr;
    h->d_nbr[i] = nullptr;
    h->ticket_of[i] = -1;
    h->res_n[i] = 0;
  }
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < slots && e == cudaSuccess; ++i) {
    if ((e = cudaHostAlloc((void**)&h->h_in[i], h->in_bytes, cudaHostAllocDefault)) != cudaSuccess) break;
    if ((e = cudaHostAlloc((void**)&h->h_res[i], sizeof(float) * h->res_floats, cudaHostAllocDefault)) != cudaSuccess) break;
    if ((e = cudaMalloc((void**)&h->d_in[i], h->in_bytes)) != cudaSuccess) break;
    if ((e = cudaMalloc((void**)&h->d_res[i], sizeof(float) * h->res_floats)) != cudaSuccess) break;
    if ((e = cudaEventCreateWithFlags(&h->done[i], cudaEventDisableTiming)) != cudaSuccess) break;
    if ((e = cudaEventCreateWithFlags(&h->in_ready[i], cudaEventDisableTiming)) != cudaSuccess) break;
    if ((e = cudaEventCreateWithFlags(&h->step_done[i], cudaEventDisableTiming)) != cudaSuccess) break;
    if ((e = cudaMalloc((void**)&h->d_nbr[i], sizeof(float) * h->res_floats * (size_t)d)) != cudaSuccess) break;
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->in_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    lstep_host_stepper_destroy(h);
    return cuda_fail(e, "host_stepper_create");
  }
  *out = h;
  return LSTEP_OK;
}

extern "C" int lstep_pe_step_host(lstep_host_stepper* h, const lstep_pe_stream* s, const lstep_csr* csr, int64_t n_edges,
                                  const int64_t* src_host, const int64_t* dst_host, const double* t_host,
                                  const int64_t* ids_host, int64_t n_ids, int head, int len, int append_slot, const float* G,
                                  const int64_t* const* query_ids_host_arrays, int n_queries, float* nbr_out, int K,
                                  const lstep_pe_mlp* mlp_nbr, const lstep_pe_mlp* mlp_upd, void* workspace,
                                  size_t workspace_bytes, uint32_t* err_flag, void* stream, int64_t* ticket) {
  if (!h || !s || !ticket || n_edges <= 0 || n_edges > h->max_edges || n_queries < 0 || n_queries > h->max_queries)
    return LSTEP_ERR_INVALID_ARG;
  if (!src_host || !dst_host || !t_host || (n_queries > 0 && !query_ids_host_arrays)) return LSTEP_ERR_INVALID_ARG;
  if (ids_host && (n_ids < 0 || n_ids > 2 * n_edges)) return LSTEP_ERR_INVALID_ARG;
  for (int c = 0; c < n_queries; ++c)
    if (!query_ids_host_arrays[c]) return LSTEP_ERR_INVALID_ARG;
  if (s->d != h->d) return LSTEP_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  const int slot = (int)(h->next_ticket % h->slots);
  cudaError_t e;
  if (h->ticket_of[slot] >= 0) {  // the slot's previous step (copy in, kernels, copy out) must have drained
    if ((e = cudaEventSynchronize(h->done[slot])) != cudaSuccess) return cuda_fail(e, "pe_step_host slot wait");
  }
  const size_t n = (size_t)n_edges;
  int64_t* hp = reinterpret_cast<int64_t*>(h->h_in[slot]);
  int64_t* h_src = hp;
  int64_t* h_dst = hp + n;
  double* h_t = reinterpret_cast<double*>(hp + 2 * n);
  int64_t* h_ids = hp + 3 * n;  // room for 2n
  int64_t* h_q = hp + 5 * n;
  memcpy(h_src, src_host, 8 * n);
  memcpy(h_dst, dst_host, 8 * n);
  memcpy(h_t, t_host, 8 * n);
  double tmax = t_host[0];
  for (size_t i = 1; i < n; ++i) tmax = t_host[i] > tmax ? t_host[i] : tmax;
  if (ids_host) {
    memcpy(h_ids, ids_host, 8 * (size_t)n_ids);
  } else {  // sorted unique endpoints (evaluate_model_utils.py:54-55 does this with torch.unique on the host)
    n_ids = unique_sorted_ids(src_host, dst_host, n, s->V1, h_ids);
  }
  for (int c = 0; c < n_queries; ++c) memcpy(h_q + (size_t)c * n, query_ids_host_arrays[c], 8 * n);
  {
    // every id indexes the [V1, d] table and the ring on the device: reject out-of-range ids here (the reference
    // raises IndexError on them, utils.py:140 / LSTEP.py:303; the Python host maps this status to IndexError)
    const int64_t V1 = s->V1;
    bool ok = n_ids == 0 || (h_ids[0] >= 0 && h_ids[n_ids - 1] < V1);  // sorted
    if (ids_host)
      for (int64_t i = 0; i < n_ids && ok; ++i) ok = h_ids[i] >= 0 && h_ids[i] < V1;
    for (size_t i = 0; i < n && ok; ++i) ok = src_host[i] >= 0 && src_host[i] < V1 && dst_host[i] >= 0 && dst_host[i] < V1;
    for (size_t i = 0; i < (size_t)n_queries * n && ok; ++i) ok = h_q[i] >= 0 && h_q[i] < V1;
    if (!ok) return LSTEP_ERR_ID_RANGE;
  }
  const size_t bytes = 8 * n * (size_t)(5 + n_queries);
  const bool use_memcpy = tuning().host_memcpy != 0;
  // copy-in on its own stream: it only has to wait for the slot's previous use (synchronised above), so it runs
  // while the previous step is still computing; the compute stream picks it up through an event
  if (use_memcpy) {
    if ((e = cudaMemcpyAsync(h->d_in[slot], h->h_in[slot], bytes, cudaMemcpyHostToDevice, h->in_stream)) != cudaSuccess)
      return cuda_fail(e, "pe_step_host h2d");
  } else {
    stage_in_kernel<<<4, 256, 0, h->in_stream>>>(reinterpret_cast<const uint4*>(h->h_in[slot]), reinterpret_cast<uint4*>(h->d_in[slot]),
                                                 (int64_t)((bytes + 15) / 16));
    int rc0 = check_launch("stage_in");
    if (rc0 != LSTEP_OK) return rc0;
  }
  if ((e = cudaEventRecord(h->in_ready[slot], h->in_stream)) != cudaSuccess) return cuda_fail(e, "pe_step_host event");
  if ((e = cudaStreamWaitEvent(st, h->in_ready[slot], 0)) != cudaSuccess) return cuda_fail(e, "pe_step_host wait");
  h->h2d_bytes += bytes;
  const int64_t* dp = reinterpret_cast<const int64_t*>(h->d_in[slot]);
  const int64_t* qdev[8];
  for (int c = 0; c < n_queries; ++c) {
    // the same host array passed twice (negative sources == sources under random negative sampling) becomes the same
    // device pointer, which is what pe_step_core recognises as an identical query set
    int u = 0;
    while (u < c && query_ids_host_arrays[u] != query_ids_host_arrays[c]) ++u;
    qdev[c] = dp + 5 * n + (size_t)u * n;
  }
  float* outp = nbr_out ? nbr_out : h->d_nbr[slot];
  int rc = pe_step_core(s, csr, dp, dp + n, reinterpret_cast<const double*>(dp + 2 * n), n_edges, dp + 3 * n, n_ids, tmax, head,
                        len, append_slot, G, qdev, n_queries, outp, K, mlp_nbr, mlp_upd, workspace, workspace_bytes, err_flag,
                        stream);
  if (rc != LSTEP_OK) return rc;
  const int64_t rows = (int64_t)n_queries * n_edges;
  // result path on its own stream, behind the step: the next step does not queue behind it. (A caller-owned
  // nbr_out is reused by the caller's next step, so in that case the result kernel stays on the compute stream.)
  cudaStream_t rs = nbr_out ? st : h->out_stream;
  if (!nbr_out) {
    if ((e = cudaEventRecord(h->step_done[slot], st)) != cudaSuccess) return cuda_fail(e, "pe_step_host event");
    if ((e = cudaStreamWaitEvent(rs, h->step_done[slot], 0)) != cudaSuccess) return cuda_fail(e, "pe_step_host wait");
  }
  if (rows > 0) {
    // the row sums go straight into the pinned result slot (visible to the host once the event below has fired)
    row_sum_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, rs>>>(outp, rows, h->d, use_memcpy ? h->d_res[slot] : h->h_res[slot]);
    if ((rc = check_launch("row_sum")) != LSTEP_OK) return rc;
    if (use_memcpy &&
        (e = cudaMemcpyAsync(h->h_res[slot], h->d_res[slot], sizeof(float) * (size_t)rows, cudaMemcpyDeviceToHost, rs)) != cudaSuccess)
      return cuda_fail(e, "pe_step_host d2h");
    h->d2h_bytes += sizeof(float) * (size_t)rows;
  }
  if ((e = cudaEventRecord(h->done[slot], rs)) != cudaSuccess) return cuda_fail(e, "pe_step_host event");
  h->res_n[slot] = rows;
  h->ticket_of[slot] = h->next_ticket;
  *ticket = h->next_ticket++;
  return LSTEP_OK;
}

extern "C" int lstep_host_step_result(lstep_host_stepper* h, int64_t ticket, const float** result_host, int64_t* n_floats) {
  if (!h || ticket < 0 || ticket >= h->next_ticket) return LSTEP_ERR_INVALID_ARG;
  const int slot = (int)(ticket % h->slots);
  if (h->ticket_of[slot] != ticket) return LSTEP_ERR_INVALID_ARG;  // the slot has been reused: result gone
  cudaError_t e = cudaEventSynchronize(h->done[slot]);
  if (e != cudaSuccess) return cuda_fail(e, "host_step_result");
  if (result_host) *result_host = h->h_res[slot];
  if (n_floats) *n_floats = h->res_n[slot];
  return LSTEP_OK;
}

extern "C" void lstep_host_stepper_bytes(const lstep_host_stepper* h, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
  if (!h) return;
  if (h2d_bytes) *h2d_bytes = h->h2d_bytes;
  if (d2h_bytes) *d2h_bytes = h->d2h_bytes;
}

// A whole run of consecutive batches in one call (an evaluation split: evaluate_model_utils.py:38-142 loops over
// batches of the split): the loop of lstep_pe_step_host + lstep_host_step_result lives here, results read one step
// behind, so the host cost per batch is the CUDA launches alone. Steady-state regime only (the history ring is full:
// *len_io == T, one collapsed filter G for every step); batch b is edges [b*batch_size, min((b+1)*batch_size, n_total)).
// results_host: [n_batches][n_queries][batch_size] floats (the tail of a ragged last batch is left untouched).
// Synchronous: every result has been read when the call returns.
extern "C" int lstep_pe_steps_host(lstep_host_stepper* h, const lstep_pe_stream* s, const lstep_csr* csr, int64_t n_total,
                                   int64_t batch_size, const int64_t* src_host, const int64_t* dst_host, const double* t_host,
                                   const int64_t* const* query_ids_host_arrays, int n_queries, int* head_io, int* len_io,
                                   const float* G, int K, const lstep_pe_mlp* mlp_nbr, const lstep_pe_mlp* mlp_upd,
                                   void* workspace, size_t workspace_bytes, uint32_t* err_flag, void* stream,
                                   float* results_host, int64_t* n_done_out) {
  if (n_done_out) *n_done_out = 0;
  if (!h || !s || !head_io || !len_io || !results_host || n_total < 0 || batch_size <= 0 || batch_size > h->max_edges)
    return LSTEP_ERR_INVALID_ARG;
  if (n_queries < 0 || n_queries > 8 || (n_queries > 0 && !query_ids_host_arrays)) return LSTEP_ERR_INVALID_ARG;
  if (*len_io != s->T) return LSTEP_ERR_UNSUPPORTED;  // masked regime: the filter changes every step, use the per-step call
  int head = *head_io;
  int64_t pending = -1, pending_b = -1, pending_n = 0;
  auto collect = [&](int64_t ticket, int64_t b, int64_t n) -> int {
    const float* r = nullptr;
    int64_t nf = 0;
    const int rc = lstep_host_step_result(h, ticket, &r, &nf);
    if (rc != LSTEP_OK) return rc;
    float* dst = results_host + (size_t)b * n_queries * batch_size;
    for (int c = 0; c < n_queries; ++c) memcpy(dst + (size_t)c * batch_size, r + (size_t)c * n, sizeof(float) * (size_t)n);
    return LSTEP_OK;
  };
  int64_t b = 0;
  for (int64_t lo = 0; lo < n_total; lo += batch_size, ++b) {
    const int64_t n = n_total - lo < batch_size ? n_total - lo : batch_size;
    const int64_t* q[8];
    for (int c = 0; c < n_queries; ++c) q[c] = query_ids_host_arrays[c] + lo;
    int64_t ticket = -1;
    // full ring: the oldest slot (head) is overwritten, then becomes the newest
    int rc = lstep_pe_step_host(h, s, csr, n, src_host + lo, dst_host + lo, t_host + lo, nullptr, 0, head, s->T, head, G, q, n_queries,
                                nullptr, K, mlp_nbr, mlp_upd, workspace, workspace_bytes, err_flag, stream, &ticket);
    if (rc != LSTEP_OK) {
      // batches [0, b) HAVE been applied to the ring: report the position reached (head, and the count in *len_io's place is
      // not possible — len stays T — so the caller derives it from the returned head and n_done below) and drain the pending result
      if (pending >= 0) (void)collect(pending, pending_b, pending_n);
      *head_io = head;
      if (n_done_out) *n_done_out = b;
      return rc;
    }
    head = (head + 1) % s->T;
    if (pending >= 0 && (rc = collect(pending, pending_b, pending_n)) != LSTEP_OK) {
      *head_io = head;
      if (n_done_out) *n_done_out = b + 1;
      return rc;
    }
    pending = ticket;
    pending_b = b;
    pending_n = n;
  }
  *head_io = head;
  if (n_done_out) *n_done_out = b;
  if (pending >= 0) {
    const int rc = collect(pending, pending_b, pending_n);
    if (rc != LSTEP_OK) return rc;
  }
  return LSTEP_OK;
}
