// Change-log PE history (SURVEY §7 / §8(e) capacity caveat: "a versioned (change-log) history — piecewise-constant rows
// make the collapsed filter a [span sum] per version").
//
// The reference keeps the last T snapshots of the WHOLE table ([V1, T, d]: 688 GB at 10 M nodes, T = 100) although a step
// changes only the batch nodes and their sampled neighbours (~13 k of 10 M rows at B = 2000). The dense ring of the
// streaming API inherits that: every step writes one row per node (the H term, 2 * V1 * d * 4 B = 13.8 GB per step at
// 10 M nodes — 4.6 ms of pure copy against 0.3 ms of kernels). Here the same history is stored as
//
//     base[v]                 value of node v at the OLDEST step of the window (window index 0)
//     events[slot][e]         (node, row) for every row a step changed, one list per window step (ring over T slots),
//                             with a per-slot hash table node -> e and a per-node bit mask of the slots holding an event
//
// i.e. V1 + T * (rows changed per step) rows instead of V1 * T: 6.9 GB + 0.9 GB for the 10 M-node graph. It is exactly the
// same function of time: x[v, f] = row of the latest event of v at a window index <= f, else base[v].
//
//   filter   out[n, c] = sum_f G[f, c] x[n, f, c] = w_0[c] base[n, c] + sum_j w_j[c] row_j[c],  w_j[c] = sum of G[f, c] over the
//            window indices the version j covers (plain fp32 sums of G over the span: no prefix-sum differences, no
//            cancellation). A node costs (events in window + 1) row reads instead of T.
//   append   the rows a step changed are {sorted batch nodes} u {distinct phase-B destinations U}; U comes from the push
//            kernel's list, batch nodes already in U are recognised by the push kernel's stamp map. Each gets a slot in
//            the step's event list (atomic counter; order inside a step is irrelevant: lookups go through the hash).
//   retire   when the window slides, the events of the step that leaves become the nodes' new base rows.
//
// Node-id sharded groups: a rank keeps base rows / masks / events of the nodes it owns only (local row = (v - row_add) /
// row_mul for v % row_mul == row_add).
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "step_core.cuh"

namespace lstep {
namespace {

constexpr unsigned long long kEmpty = ~0ull;

__device__ __forceinline__ uint32_t hash_node(uint32_t v) {
  v ^= v >> 16;
  v *= 0x7feb352du;
  v ^= v >> 15;
  v *= 0x846ca68bu;
  v ^= v >> 16;
  return v;
}

__device__ __forceinline__ int hash_find(const unsigned long long* __restrict__ tab, int H, uint32_t node) {
  uint32_t h = hash_node(node) & (uint32_t)(H - 1);
  for (int probes = 0; probes < H; ++probes) {
    const unsigned long long e = __ldcg(tab + h);
    if (e == kEmpty) return -1;
    if ((uint32_t)(e >> 32) == node) return (int)(uint32_t)e;
    h = (h + 1) & (uint32_t)(H - 1);
  }
  return -1;
}

// Table replicas of the other ranks of a peer group (csrc/peer.cu): a row written to the local table is stored to the same
// row of every replica as well (NVLink peer stores; published by the group's next barrier).
struct PeerTables {
  float* p[LSTEP_MAX_PEERS];
  int n;  // number of OTHER replicas
};

// One CTA (128 threads) per batch node.
__global__ void __launch_bounds__(128) changelog_filter_kernel(lstep_changelog cl, int head, int len, const int64_t* __restrict__ ids,
                                                               int64_t n_ids, const float* __restrict__ G, float* __restrict__ out,
                                                               int64_t out_stride, const int64_t* __restrict__ out_ids, PeerTables peers) {
  __shared__ int s_idx[128];    // event index at window position f, or -1
  __shared__ int s_evf[129];    // window positions that carry an event (ascending), then `len`
  __shared__ int s_evi[128];
  __shared__ int s_n;
  const int tid = threadIdx.x;
  const int d = cl.d, dvec = d >> 2, T = cl.T;
  for (int64_t n = blockIdx.x; n < n_ids; n += gridDim.x) {
    const int64_t v = ids[n];
    const int64_t lrow = (v - cl.row_add) / cl.row_mul;
    if (tid < len) {
      const int slot = (head + tid) % T;
      int idx = -1;
      if (cl.ev_mask[lrow * 4 + (slot >> 5)] & (1u << (slot & 31))) idx = hash_find(cl.ev_hash + (size_t)slot * cl.H, cl.H, (uint32_t)v);
      s_idx[tid] = idx;
    }
    __syncthreads();
    if (tid == 0) {
      int c = 0;
      for (int f = 0; f < len; ++f)
        if (s_idx[f] >= 0) {
          s_evf[c] = f;
          s_evi[c] = s_idx[f];
          ++c;
        }
      s_evf[c] = len;
      s_n = c;
    }
    __syncthreads();
    const int ne = s_n;
    if (tid < dvec) {
      auto span = [&](int f0, int f1) {  // sum of G[f, 4 columns] over window positions [f0, f1)
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int f = f0; f < f1; ++f) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(G + (size_t)f * d) + tid);
          w.x += g.x;
          w.y += g.y;
          w.z += g.z;
          w.w += g.w;
        }
        return w;
      };
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      {
        const float4 b = __ldcg(reinterpret_cast<const float4*>(cl.base + lrow * (int64_t)d) + tid);
        const float4 w = span(0, s_evf[0]);
        acc.x = w.x * b.x;
        acc.y = w.y * b.y;
        acc.z = w.z * b.z;
        acc.w = w.w * b.w;
      }
      for (int j0 = 0; j0 < ne; j0 += 4) {  // four row loads in flight
        float4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u;
          if (j < ne) {
            const int slot = (head + s_evf[j]) % T;
            r[u] = __ldcg(reinterpret_cast<const float4*>(cl.ev_row + ((size_t)slot * cl.cap + s_evi[j]) * d) + tid);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u;
          if (j < ne) {
            const float4 w = span(s_evf[j], s_evf[j + 1]);
            acc.x = fmaf(w.x, r[u].x, acc.x);
            acc.y = fmaf(w.y, r[u].y, acc.y);
            acc.z = fmaf(w.z, r[u].z, acc.z);
            acc.w = fmaf(w.w, r[u].w, acc.w);
          }
        }
      }
      const int64_t orow = out_ids ? out_ids[n] : n;
      *reinterpret_cast<float4*>(out + orow * out_stride + 4 * tid) = acc;
      for (int g = 0; g < peers.n; ++g) *reinterpret_cast<float4*>(peers.p[g] + orow * out_stride + 4 * tid) = acc;
    }
    __syncthreads();
  }
  if (peers.n) __threadfence_system();
}

// The events of `slot` (the step leaving the window) become base rows; their mask bits are cleared.
__global__ void __launch_bounds__(256) changelog_retire_kernel(lstep_changelog cl, int slot) {
  const int cnt = min(cl.ev_cnt[slot], cl.cap);
  const int dvec = cl.d >> 2;
  const int64_t total = (int64_t)cnt * dvec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i / dvec), c = (int)(i % dvec);
    const int64_t v = cl.ev_node[(size_t)slot * cl.cap + e];
    const int64_t lrow = (v - cl.row_add) / cl.row_mul;
    reinterpret_cast<float4*>(cl.base + lrow * (int64_t)cl.d)[c] =
        reinterpret_cast<const float4*>(cl.ev_row + ((size_t)slot * cl.cap + e) * cl.d)[c];
    if (c == 0) atomicAnd(cl.ev_mask + lrow * 4 + (slot >> 5), ~(1u << (slot & 31)));
  }
}

// Items [0, nU) = the distinct phase-B destinations U (nU read from the device when n_u_dev != NULL), items [nU, nU + n_ids)
// = the batch nodes, skipped when the stamp map says U already holds them. One warp per item.
__global__ void __launch_bounds__(256) changelog_append_kernel(lstep_changelog cl, int slot, const float* __restrict__ table,
                                                               const int64_t* __restrict__ U, const int32_t* __restrict__ n_u_dev, int64_t n_u,
                                                               const int64_t* __restrict__ ids, int64_t n_ids,
                                                               const int32_t* __restrict__ stamp_map, int stamp, int with_row0,
                                                               uint32_t* err_flag, PeerTables peers) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  if (n_u_dev) {
    const int64_t nd = __ldcg(n_u_dev);
    n_u = nd < n_u ? nd : n_u;
  }
  const int dvec = cl.d >> 2;
  for (int64_t it = warp; it < n_u + n_ids + (with_row0 ? 1 : 0); it += n_warps) {
    int64_t v;
    if (it < n_u) {
      v = __ldcg(reinterpret_cast<const long long*>(U) + it);
    } else {
      // a batch node, or (last item) the padding row 0, which update_pe zeroes in every call (LSTEP.py:317)
      v = it < n_u + n_ids ? ids[it - n_u] : 0;
      if (it == n_u + n_ids && n_ids > 0 && ids[0] == 0) continue;
      if (U && stamp_map && __ldcg(stamp_map + v) == stamp) continue;  // phase B changed this row too: U holds it
    }
    if (v % cl.row_mul != cl.row_add) continue;  // not owned by this rank
    const int64_t lrow = (v - cl.row_add) / cl.row_mul;
    int idx = 0;
    if (lane == 0) idx = atomicAdd(cl.ev_cnt + slot, 1);
    idx = __shfl_sync(kFull, idx, 0);
    const float4* srow = reinterpret_cast<const float4*>(table + v * (int64_t)cl.d);
    if (idx >= cl.cap) {
      if (lane == 0 && err_flag) atomicOr(err_flag, LSTEP_FLAG_CHANGELOG_FULL);
      for (int g = 0; g < peers.n; ++g)  // the replicas must not diverge even when the event is lost
        for (int c = lane; c < dvec; c += 32) reinterpret_cast<float4*>(peers.p[g] + v * (int64_t)cl.d)[c] = __ldcg(srow + c);
      continue;
    }
    float4* drow = reinterpret_cast<float4*>(cl.ev_row + ((size_t)slot * cl.cap + idx) * cl.d);
    for (int c = lane; c < dvec; c += 32) {
      const float4 x = __ldcg(srow + c);
      drow[c] = x;
      for (int g = 0; g < peers.n; ++g) reinterpret_cast<float4*>(peers.p[g] + v * (int64_t)cl.d)[c] = x;  // publish the owned row
    }
    if (lane == 0) {
      cl.ev_node[(size_t)slot * cl.cap + idx] = (int32_t)v;
      unsigned long long* tab = cl.ev_hash + (size_t)slot * cl.H;
      const unsigned long long entry = ((unsigned long long)(uint32_t)v << 32) | (uint32_t)idx;
      uint32_t h = hash_node((uint32_t)v) & (uint32_t)(cl.H - 1);
      while (atomicCAS(tab + h, kEmpty, entry) != kEmpty) h = (h + 1) & (uint32_t)(cl.H - 1);
      atomicOr(cl.ev_mask + lrow * 4 + (slot >> 5), 1u << (slot & 31));
    }
  }
  if (peers.n) __threadfence_system();
}

bool valid(const lstep_changelog* cl) {
  return cl && cl->base && cl->ev_node && cl->ev_row && cl->ev_cnt && cl->ev_hash && cl->ev_mask && cl->rows > 0 && cl->T > 0 && cl->T <= 128 &&
         cl->cap > 0 && cl->H >= 2 * cl->cap && (cl->H & (cl->H - 1)) == 0 && cl->d > 0 && cl->d % 4 == 0 && cl->d <= 512 && cl->row_mul > 0 &&
         cl->row_add >= 0 && cl->row_add < cl->row_mul;
}

}  // namespace

}  // namespace lstep

using namespace lstep;

namespace lstep {
static PeerTables other_tables(const lstep_peer_group* grp) {
  PeerTables pt{};
  if (grp)
    for (int g = 0; g < grp->world; ++g)
      if (g != grp->rank) pt.p[pt.n++] = grp->table[g];
  return pt;
}
int changelog_filter_peer(const lstep_changelog* cl, int head, int len, const int64_t* ids, int64_t n_ids, const float* G, float* out,
                          int64_t out_stride, const int64_t* out_ids, const lstep_peer_group* grp, void* stream);
int changelog_append_peer(const lstep_changelog* cl, int slot, int retire, const float* table, const int64_t* U, const int32_t* n_u_dev,
                          int64_t n_u, const int64_t* ids, int64_t n_ids, const int32_t* stamp_map, int stamp, int with_row0,
                          uint32_t* err_flag, const lstep_peer_group* grp, void* stream);
}  // namespace lstep

extern "C" int lstep_changelog_filter(const lstep_changelog* cl, int head, int len, const int64_t* ids, int64_t n_ids, const float* G,
                                      float* out, int64_t out_stride, const int64_t* out_ids, void* stream) {
  return changelog_filter_peer(cl, head, len, ids, n_ids, G, out, out_stride, out_ids, nullptr, stream);
}
int lstep::changelog_filter_peer(const lstep_changelog* cl, int head, int len, const int64_t* ids, int64_t n_ids, const float* G, float* out,
                                 int64_t out_stride, const int64_t* out_ids, const lstep_peer_group* grp, void* stream) {
  if (!valid(cl) || head < 0 || head >= cl->T || len < 0 || len > cl->T || n_ids < 0) return LSTEP_ERR_INVALID_ARG;
  if (n_ids == 0) return LSTEP_OK;
  if (!ids || !G || !out || out_stride % 4 != 0 || (reinterpret_cast<uintptr_t>(out) & 15)) return LSTEP_ERR_INVALID_ARG;
  const int64_t grid = std::min<int64_t>(n_ids, (int64_t)num_sms() * 16);
  changelog_filter_kernel<<<(unsigned)grid, 128, 0, as_stream(stream)>>>(*cl, head, len, ids, n_ids, G, out, out_stride, out_ids, other_tables(grp));
  return check_launch("changelog_filter");
}

/* Close a step: when `retire` is set the events of `slot` (the oldest step, leaving the window) are folded into the base rows
 * first; then the rows of `table` for U[0 .. n_u) (n_u = min(n_u, *n_u_dev) when n_u_dev != NULL) and for the ids whose
 * stamp_map entry differs from `stamp` become the slot's new events. */
extern "C" int lstep_changelog_append(const lstep_changelog* cl, int slot, int retire, const float* table, const int64_t* U,
                                      const int32_t* n_u_dev, int64_t n_u, const int64_t* ids, int64_t n_ids, const int32_t* stamp_map,
                                      int stamp, int with_row0, uint32_t* err_flag, void* stream) {
  return changelog_append_peer(cl, slot, retire, table, U, n_u_dev, n_u, ids, n_ids, stamp_map, stamp, with_row0, err_flag, nullptr, stream);
}
int lstep::changelog_append_peer(const lstep_changelog* cl, int slot, int retire, const float* table, const int64_t* U, const int32_t* n_u_dev,
                                 int64_t n_u, const int64_t* ids, int64_t n_ids, const int32_t* stamp_map, int stamp, int with_row0,
                                 uint32_t* err_flag, const lstep_peer_group* grp, void* stream) {
  if (!valid(cl) || slot < 0 || slot >= cl->T || n_u < 0 || n_ids < 0 || !table) return LSTEP_ERR_INVALID_ARG;
  if ((n_u > 0 && !U) || (n_ids > 0 && !ids)) return LSTEP_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  cudaError_t e;
  if (retire) {
    changelog_retire_kernel<<<num_sms() * 2, 256, 0, st>>>(*cl, slot);
    int rc = check_launch("changelog_retire");
    if (rc != LSTEP_OK) return rc;
  }
  if ((e = cudaMemsetAsync(cl->ev_hash + (size_t)slot * cl->H, 0xff, sizeof(unsigned long long) * cl->H, st)) != cudaSuccess ||
      (e = cudaMemsetAsync(cl->ev_cnt + slot, 0, sizeof(int32_t), st)) != cudaSuccess) {
    set_cuda_error(e, "changelog memset");
    return LSTEP_ERR_CUDA;
  }
  if (n_u + n_ids == 0 && !with_row0) return LSTEP_OK;
  const int64_t warps = n_u + n_ids + 1;
  const int64_t grid = std::min<int64_t>(ceil_div(warps * 32, 256), (int64_t)num_sms() * 8);
  changelog_append_kernel<<<(unsigned)grid, 256, 0, st>>>(*cl, slot, table, U, n_u_dev, n_u, ids, n_ids, stamp_map, stamp, with_row0, err_flag,
                                                          other_tables(grp));
  return check_launch("changelog_append");
}

/* lstep_pe_step on a change-log history: filter of the batch nodes into the table, the step's kernels, retire + append.
 * head / len: the window (slot of its oldest step, number of valid steps <= T); the step's events go to slot
 * (head + len) % T while the history is filling, to `head` (after retiring it) once len == T. q_off / q_rows as in
 * lstep_pe_step_sharded (q_rows < 0: all edges). filter_ids / n_filter: the batch nodes whose history THIS rank holds
 * (= ids for a single GPU); their filtered rows are written to filter_out (row i, pitch d) when filter_out != NULL — the
 * caller all-gathers and scatters them — and straight into the table otherwise. */
extern "C" int lstep_pe_step_changelog(const lstep_pe_stream* s, const lstep_changelog* cl, const lstep_csr* csr, int64_t lo, int64_t n_edges,
                                       const int64_t* ids, int64_t n_ids, double current_time, int head, int len, const float* G,
                                       const int64_t* const* query_ids_host, int n_queries, int64_t q_off, int64_t q_rows, float* nbr_out,
                                       int K, const lstep_pe_mlp* mlp_nbr, const lstep_pe_mlp* mlp_upd, void* workspace,
                                       size_t workspace_bytes, uint32_t* err_flag, void* stream, int phase) {
  // phase 0: everything; 1: filter only (into filter destination = table); 2: everything after the filter
  if (!s || !s->src || !s->dst || !s->t || !s->cur || !valid(cl) || lo < 0 || n_edges < 0 || n_ids < 0 || !mlp_upd) return LSTEP_ERR_INVALID_ARG;
  if (cl->d != s->d || head < 0 || head >= cl->T || len < 0 || len > cl->T) return LSTEP_ERR_INVALID_ARG;
  if (!update_push_available(mlp_upd)) return LSTEP_ERR_UNSUPPORTED;  // the append needs the push kernel's U list and stamp map
  int rc;
  if (phase != 2 && n_ids > 0) {
    if (cl->row_mul != 1) return LSTEP_ERR_INVALID_ARG;  // sharded groups filter through lstep_changelog_filter themselves
    if ((rc = lstep_changelog_filter(cl, head, len, ids, n_ids, G, s->cur, s->d, ids, stream)) != LSTEP_OK) return rc;
  }
  if (phase == 1) return LSTEP_OK;
  StepOpts opt;
  opt.skip_dft = opt.skip_append = true;
  opt.q_off = q_off;
  opt.q_rows = q_rows;
  int stamp = 0;
  opt.stamp_out = &stamp;
  lstep_pe_stream s2 = *s;
  if (!s2.ring) s2.ring = s2.cur;  // (never touched: the core neither filters nor appends)
  const int64_t qr = q_rows < 0 ? n_edges : q_rows;
  rc = pe_step_core_ex(&s2, csr, s->src + lo, s->dst + lo, s->t + lo, n_edges, ids, n_ids, current_time, 0, 0, 0, nullptr, query_ids_host,
                       qr > 0 ? n_queries : 0, nbr_out, K, mlp_nbr, mlp_upd, workspace, workspace_bytes, err_flag, stream, opt);
  if (rc != LSTEP_OK) return rc;
  const int64_t* U = nullptr;
  const int32_t *n_dest = nullptr, *stamp_map = nullptr;
  int64_t n_u_max = 0;
  if (n_ids > 0) {
    update_ws_phase_b_lists(workspace, n_ids, n_edges, K, s->d, mlp_upd->t, s->V1, &U, &n_dest, &stamp_map);
    const int64_t total = n_ids * (int64_t)K;
    n_u_max = (total < s->V1 - 1 ? total : s->V1 - 1) + 1;
  }
  const bool full = len == cl->T;
  const int slot = full ? head : (head + len) % cl->T;
  return lstep_changelog_append(cl, slot, full ? 1 : 0, s->cur, U, n_dest, n_u_max, ids, n_ids, stamp_map, stamp, n_ids > 0 ? 1 : 0, err_flag,
                                stream);
}
