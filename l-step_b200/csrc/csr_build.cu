// a1 — time-sorted CSR of the undirected temporal adjacency, built on the device.
// Replaces get_neighbor_sampler's adjacency-list loop and NeighborSampler.__init__'s per-node
// stable sort (/root/reference/utils/utils.py:292-299, 95-102).
//
// The reference's order for node v is: entries in insertion order (edge order, the source-side
// entry of an edge before its destination-side entry), stably sorted by time. That is the order
// a stable sort by (owner, time) of the flat insertion sequence produces, so the build is a
// least-significant-digit radix sort of a permutation: 8 passes over the order-preserving
// 64-bit image of the fp64 time, then ceil(bits(num_rows)/8) passes over the owner id. Each
// pass is histogram -> exclusive scan -> stable scatter (warp match + cross-warp prefix, rounds
// in index order). One-off setup work: O(E) per pass, not on the per-batch path.
#include "common.cuh"

namespace lstep {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;

__device__ __forceinline__ uint64_t time_key(double t) {
  if (t == 0.0) t = 0.0;  // -0.0 and +0.0 compare equal in the reference's sort
  uint64_t u = (uint64_t)__double_as_longlong(t);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__global__ void entries_from_edges_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                          const int64_t* __restrict__ eid, const double* __restrict__ t, int64_t E,
                                          int64_t num_rows, int32_t* __restrict__ owner, int32_t* __restrict__ nbr,
                                          int32_t* __restrict__ eids, uint64_t* __restrict__ keys,
                                          uint32_t* __restrict__ idx, uint32_t* err_flag) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = src[e], d = dst[e], id = eid[e];
  if (s < 0 || s >= num_rows || d < 0 || d >= num_rows || id < 0 || id > 0x7fffffffLL) {
    if (err_flag) atomicOr(err_flag, LSTEP_FLAG_NODE_OUT_OF_RANGE);
    s = d = 0;
    id = 0;
  }
  const uint64_t k = time_key(t[e]);
  owner[2 * e] = (int32_t)s;  // adj[src].append((dst, eid, t))   utils.py:298
  nbr[2 * e] = (int32_t)d;
  owner[2 * e + 1] = (int32_t)d;  // adj[dst].append((src, eid, t))   utils.py:299
  nbr[2 * e + 1] = (int32_t)s;
  eids[2 * e] = eids[2 * e + 1] = (int32_t)id;
  keys[2 * e] = keys[2 * e + 1] = k;
  idx[2 * e] = (uint32_t)(2 * e);
  idx[2 * e + 1] = (uint32_t)(2 * e + 1);
}

__global__ void entries_from_lists_kernel(const int64_t* __restrict__ owner_in, const int64_t* __restrict__ nbr_in,
                                          const int64_t* __restrict__ eid_in, const double* __restrict__ t, int64_t n,
                                          int64_t num_rows, int32_t* __restrict__ owner, int32_t* __restrict__ nbr,
                                          int32_t* __restrict__ eids, uint64_t* __restrict__ keys,
                                          uint32_t* __restrict__ idx, uint32_t* err_flag) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t o = owner_in[i], v = nbr_in[i], id = eid_in[i];
  if (o < 0 || o >= num_rows || v < 0 || v > 0x7fffffffLL || id < 0 || id > 0x7fffffffLL) {
    if (err_flag) atomicOr(err_flag, LSTEP_FLAG_NODE_OUT_OF_RANGE);
    o = v = id = 0;
  }
  owner[i] = (int32_t)o;
  nbr[i] = (int32_t)v;
  eids[i] = (int32_t)id;
  keys[i] = time_key(t[i]);
  idx[i] = (uint32_t)i;
}

__global__ void owner_keys_kernel(const int32_t* __restrict__ owner, const uint32_t* __restrict__ idx, int64_t n,
                                  uint64_t* __restrict__ keys) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) keys[i] = (uint64_t)(uint32_t)owner[idx[i]];
}

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n,
                                                                  int shift, uint32_t* __restrict__ hist,
                                                                  int64_t num_blocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  for (int r = 0; r < kSortItems; ++r) {
    const int64_t i = base + r * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255], 1u);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * num_blocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of a uint32 array by one CTA (setup path; n <= a few million)
__global__ void __launch_bounds__(1024) scan_u32_kernel(uint32_t* __restrict__ a, int64_t n) {
  __shared__ uint32_t s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t chunk = (n + blockDim.x - 1) / blockDim.x;
  const int64_t lo = tid * chunk, hi = lo + chunk < n ? lo + chunk : n;
  uint32_t sum = 0;
  for (int64_t i = lo; i < hi; ++i) sum += a[i];
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t v = s_warp[lane], iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(kFull, iv, o);
      if (lane >= o) iv += x;
    }
    s_warp[lane] = iv - v;
  }
  __syncthreads();
  uint32_t run = s_warp[wid] + incl - sum;
  for (int64_t i = lo; i < hi; ++i) {
    const uint32_t v = a[i];
    a[i] = run;
    run += v;
  }
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint64_t* __restrict__ keys_in,
                                                                     const uint32_t* __restrict__ idx_in, int64_t n,
                                                                     int shift, const uint32_t* __restrict__ hist,
                                                                     int64_t num_blocks,
                                                                     uint64_t* __restrict__ keys_out,
                                                                     uint32_t* __restrict__ idx_out) {
  __shared__ uint32_t base[256];
  __shared__ uint32_t wc[kSortThreads / 32][256];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  base[tid] = hist[(int64_t)tid * num_blocks + blockIdx.x];
  const int64_t tile = (int64_t)blockIdx.x * kSortTile;
  for (int r = 0; r < kSortItems; ++r) {
    for (int w = 0; w < kSortThreads / 32; ++w) wc[w][tid] = 0;
    __syncthreads();
    const int64_t i = tile + r * kSortThreads + tid;
    const bool valid = i < n;
    uint64_t key = 0;
    uint32_t id = 0;
    int digit = 256 + lane;  // inactive lanes never match anyone
    if (valid) {
      key = keys_in[i];
      id = idx_in[i];
      digit = (int)((key >> shift) & 255);
    }
    const unsigned peers = __match_any_sync(kFull, digit);
    const int rank_in_warp = __popc(peers & ((1u << lane) - 1));
    if (valid && rank_in_warp == 0) wc[wid][digit] = __popc(peers);
    __syncthreads();
    if (valid) {
      uint32_t pre = 0;
      for (int w = 0; w < wid; ++w) pre += wc[w][digit];
      const uint32_t pos = base[digit] + pre + rank_in_warp;
      keys_out[pos] = key;
      idx_out[pos] = id;
    }
    __syncthreads();
    uint32_t tot = 0;
    for (int w = 0; w < kSortThreads / 32; ++w) tot += wc[w][tid];
    base[tid] += tot;
    __syncthreads();
  }
}

__global__ void csr_emit_kernel(const uint32_t* __restrict__ idx, const int32_t* __restrict__ owner,
                                const int32_t* __restrict__ nbr, const int32_t* __restrict__ eids,
                                const uint64_t* __restrict__ tkeys_unused, const double* __restrict__ t_entries,
                                int entries_per_time, int64_t n, int64_t num_rows, int64_t* __restrict__ indptr,
                                int32_t* __restrict__ out_nbr, int32_t* __restrict__ out_eid,
                                double* __restrict__ out_t) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > n) return;
  int64_t prev = -1, cur = num_rows;
  if (i > 0) prev = owner[idx[i - 1]];
  if (i < n) {
    const uint32_t j = idx[i];
    cur = owner[j];
    out_nbr[i] = nbr[j];
    out_eid[i] = eids[j];
    out_t[i] = t_entries[j / entries_per_time];
  }
  for (int64_t v = prev + 1; v <= cur; ++v) indptr[v] = i;  // rows (prev, cur] start at i
}

struct BuildWs {
  int32_t *owner, *nbr, *eid;
  uint64_t *keys0, *keys1;
  uint32_t *idx0, *idx1;
  uint32_t* hist;
  int64_t num_blocks;
  size_t bytes;
};

static BuildWs carve_build(void* base, int64_t n) {
  BuildWs w;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + o : nullptr;
    o = align_up(o + bytes, 256);
    return p;
  };
  const size_t nn = (size_t)(n > 0 ? n : 1);
  w.num_blocks = ceil_div((int64_t)nn, kSortTile);
  w.owner = (int32_t*)take(4 * nn);
  w.nbr = (int32_t*)take(4 * nn);
  w.eid = (int32_t*)take(4 * nn);
  w.keys0 = (uint64_t*)take(8 * nn);
  w.keys1 = (uint64_t*)take(8 * nn);
  w.idx0 = (uint32_t*)take(4 * nn);
  w.idx1 = (uint32_t*)take(4 * nn);
  w.hist = (uint32_t*)take(4 * 256 * (size_t)w.num_blocks);
  w.bytes = o;
  return w;
}

static int radix_passes(BuildWs& w, int64_t n, int first_bit, int n_bits, uint64_t*& kin, uint64_t*& kout,
                        uint32_t*& iin, uint32_t*& iout, cudaStream_t st) {
  for (int shift = first_bit; shift < first_bit + n_bits; shift += 8) {
    radix_hist_kernel<<<(unsigned)w.num_blocks, kSortThreads, 0, st>>>(kin, n, shift, w.hist, w.num_blocks);
    scan_u32_kernel<<<1, 1024, 0, st>>>(w.hist, 256 * w.num_blocks);
    radix_scatter_kernel<<<(unsigned)w.num_blocks, kSortThreads, 0, st>>>(kin, iin, n, shift, w.hist, w.num_blocks,
                                                                         kout, iout);
    int rc = check_launch("radix pass");
    if (rc != LSTEP_OK) return rc;
    uint64_t* tk = kin;
    kin = kout;
    kout = tk;
    uint32_t* ti = iin;
    iin = iout;
    iout = ti;
  }
  return LSTEP_OK;
}

static int sort_and_emit(BuildWs& w, int64_t n, int64_t num_rows, const double* t_entries, int entries_per_time,
                         int64_t* out_indptr, int32_t* out_nbr, int32_t* out_eid, double* out_t, cudaStream_t st) {
  uint64_t *kin = w.keys0, *kout = w.keys1;
  uint32_t *iin = w.idx0, *iout = w.idx1;
  int rc = radix_passes(w, n, 0, 64, kin, kout, iin, iout, st);  // by time
  if (rc != LSTEP_OK) return rc;
  const int64_t blocks = ceil_div(n > 0 ? n : 1, 256);
  owner_keys_kernel<<<(unsigned)blocks, 256, 0, st>>>(w.owner, iin, n, kin);
  int bits = 1;
  while (bits < 32 && (1LL << bits) < num_rows) ++bits;
  rc = radix_passes(w, n, 0, (bits + 7) / 8 * 8, kin, kout, iin, iout, st);  // then by owner (stable)
  if (rc != LSTEP_OK) return rc;
  csr_emit_kernel<<<(unsigned)ceil_div(n + 1, 256), 256, 0, st>>>(iin, w.owner, w.nbr, w.eid, nullptr, t_entries,
                                                                  entries_per_time, n, num_rows, out_indptr, out_nbr,
                                                                  out_eid, out_t);
  return check_launch("csr_emit");
}

}  // namespace lstep

using namespace lstep;

extern "C" size_t lstep_csr_build_workspace_bytes(int64_t n_entries, int64_t num_rows) {
  (void)num_rows;
  if (n_entries < 0) return 0;
  return carve_build(nullptr, n_entries).bytes;
}

extern "C" int lstep_csr_build_from_edges(const int64_t* src, const int64_t* dst, const int64_t* eid, const double* t,
                                          int64_t num_edges, int64_t num_rows, int64_t* out_indptr, int32_t* out_nbr,
                                          int32_t* out_eid, double* out_t, void* workspace, size_t workspace_bytes,
                                          uint32_t* err_flag, void* stream) {
  if (num_edges < 0 || num_rows <= 0 || !out_indptr) return LSTEP_ERR_INVALID_ARG;
  if (num_rows > 0x7fffffffLL || 2 * num_edges > 0xffffffffLL) return LSTEP_ERR_ID_RANGE;
  const int64_t n = 2 * num_edges;
  if (n > 0 && (!src || !dst || !eid || !t || !out_nbr || !out_eid || !out_t)) return LSTEP_ERR_INVALID_ARG;
  if (!workspace || workspace_bytes < carve_build(nullptr, n).bytes) return LSTEP_ERR_WORKSPACE;
  BuildWs w = carve_build(workspace, n);
  cudaStream_t st = as_stream(stream);
  if (num_edges > 0) {
    entries_from_edges_kernel<<<(unsigned)ceil_div(num_edges, 256), 256, 0, st>>>(
        src, dst, eid, t, num_edges, num_rows, w.owner, w.nbr, w.eid, w.keys0, w.idx0, err_flag);
    int rc = check_launch("entries_from_edges");
    if (rc != LSTEP_OK) return rc;
  }
  return sort_and_emit(w, n, num_rows, t, 2, out_indptr, out_nbr, out_eid, out_t, st);
}

extern "C" int lstep_csr_build_from_entries(const int64_t* owner, const int64_t* nbr, const int64_t* eid,
                                            const double* t, int64_t n_entries, int64_t num_rows, int64_t* out_indptr,
                                            int32_t* out_nbr, int32_t* out_eid, double* out_t, void* workspace,
                                            size_t workspace_bytes, uint32_t* err_flag, void* stream) {
  if (n_entries < 0 || num_rows <= 0 || !out_indptr) return LSTEP_ERR_INVALID_ARG;
  if (num_rows > 0x7fffffffLL || n_entries > 0xffffffffLL) return LSTEP_ERR_ID_RANGE;
  if (n_entries > 0 && (!owner || !nbr || !eid || !t || !out_nbr || !out_eid || !out_t)) return LSTEP_ERR_INVALID_ARG;
  if (!workspace || workspace_bytes < carve_build(nullptr, n_entries).bytes) return LSTEP_ERR_WORKSPACE;
  BuildWs w = carve_build(workspace, n_entries);
  cudaStream_t st = as_stream(stream);
  if (n_entries > 0) {
    entries_from_lists_kernel<<<(unsigned)ceil_div(n_entries, 256), 256, 0, st>>>(
        owner, nbr, eid, t, n_entries, num_rows, w.owner, w.nbr, w.eid, w.keys0, w.idx0, err_flag);
    int rc = check_launch("entries_from_lists");
    if (rc != LSTEP_OK) return rc;
  }
  return sort_and_emit(w, n_entries, num_rows, t, 1, out_indptr, out_nbr, out_eid, out_t, st);
}
