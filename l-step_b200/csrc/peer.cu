// Peer group: the scale-out PE step over NVLink peer memory (BASELINE config 5: 10 M nodes, 8 x B200).
//
// Layout. Every rank holds a REPLICA of the current PE table [V1, d] (6.9 GB at 10 M nodes) and of the temporal CSR, and
// OWNS the nodes v with v % world == rank: their PE history (change log, csrc/changelog.cu) and every piece of work whose
// result is a row of an owned node. Replicas are kept equal by the owners STORING the rows they change straight into the
// other replicas (plain 16-byte stores to peer pointers: the NVSwitch carries them; no NCCL call, no staging buffer, no
// host involvement, no data-dependent message sizes), and two flag barriers per step order those stores against the
// readers:
//
//     filter     owned batch nodes: history -> filtered row -> table row in EVERY replica          (1/world of the rows)
//     ---------- barrier 1: everyone's filtered rows AND everyone's final rows of the previous step have landed ----------
//     gather     a6 lookup + aggregate of this rank's 1/world share of the query rows  ||  a7 edge aggregate of the owned
//                batch nodes                                                                     (reads the local replica)
//     MLP pair   neighbourhood MLP of the share -> outputs  ||  phase-A MLP of the owned batch nodes -> local new_rows
//     bcast      new_rows[i] -> every rank's new_rows buffer at row pos_mine[i] (the node's index in the batch's id list)
//     ---------- barrier 2: phase A's rows of ALL batch nodes are in every rank's new_rows buffer --------------------------
//     push       lookup of ALL batch nodes (replicated: 4 k warp searches), accumulation for the OWNED destinations only
//                (exact fixed point, csrc/update_push.cu); the owned batch nodes' phase-A rows go into the local table
//     MLP (B)    owned destinations: table row <- row + tanh(mlp(aggregate))                      (local table)
//     append     owned changed rows (owned batch nodes, owned destinations, row 0 on rank 0): event of the change log
//                AND the same row of every other replica
//
// Why two barriers are enough (what a rank reads between two barriers is never written by another rank in between):
//   * between barrier 1 and barrier 2 a rank reads arbitrary rows of its replica (gather, base rows of the neighbourhood
//     MLP). Remote stores into a replica come from the owners' filter and append kernels only. A rank passes barrier 2 of
//     step s only after EVERY rank has announced it, i.e. after every rank's gather / MLP pair of step s have completed; its
//     append(s) and filter(s+1) — the only kernels that store into other replicas — come later in its stream. And nobody
//     passes barrier 1 of step s+1 before everyone's append(s) / filter(s+1) have completed and their stores are fenced.
//   * after barrier 2 a rank touches only rows it owns (push applies owned phase-A rows, the phase-B MLP reads and writes
//     owned destinations, the append reads owned rows); remote stores from other ranks' append(s) / filter(s+1) go to rows
//     THEY own. The new_rows buffer is rewritten by bcast(s+1), which follows barrier 1 of step s+1, i.e. every rank's
//     push(s) has completed.
// A non-owner's replica is stale for the rows a step changes between that step's barrier 2 and the next barrier 1 — an
// interval in which it does not read them.
//
// Results equal the single-GPU step's bit for bit: every row is computed by one rank from the same inputs with the same
// kernels, and phase B's sums are exact 32.32 fixed point (independent of who adds what in which order).
//
// Barrier = flag exchange in peer memory: rank r announces epoch e by storing e into flags[g][r] of every rank g
// (st.release.sys after a system-scope fence; the storing kernels fence their own peer stores too); a rank waits by polling
// its OWN flag block (ld.acquire.sys) until all `world` entries are >= e. The wait is bounded (timeout_ms on the global
// timer): a missing peer raises LSTEP_FLAG_PEER_TIMEOUT instead of hanging the device. The inserted kernels are plain
// launches (no programmatic overlap across a barrier).
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "step_core.cuh"

namespace lstep {

int changelog_filter_peer(const lstep_changelog* cl, int head, int len, const int64_t* ids, int64_t n_ids, const float* G, float* out,
                          int64_t out_stride, const int64_t* out_ids, const lstep_peer_group* grp, void* stream);
int changelog_append_peer(const lstep_changelog* cl, int slot, int retire, const float* table, const int64_t* U, const int32_t* n_u_dev,
                          int64_t n_u, const int64_t* ids, int64_t n_ids, const int32_t* stamp_map, int stamp, int with_row0,
                          uint32_t* err_flag, const lstep_peer_group* grp, void* stream);

namespace {

struct PeerPtrs {
  float* p[LSTEP_MAX_PEERS];
  int n;
};
struct FlagPtrs {
  uint32_t* p[LSTEP_MAX_PEERS];
  int n, rank;
};

bool valid_group(const lstep_peer_group* g) {
  if (!g || g->world < 1 || g->world > LSTEP_MAX_PEERS || g->rank < 0 || g->rank >= g->world) return false;
  for (int i = 0; i < g->world; ++i)
    if (!g->table[i] || !g->new_rows[i] || !g->flags[i]) return false;
  return true;
}

// dst[g][dst_rows[i]][:] = src[i][:] for every rank g (the local one included): one warp per row
__global__ void __launch_bounds__(256) peer_rows_bcast_kernel(const float* __restrict__ src, int64_t n_rows, int d,
                                                              const int64_t* __restrict__ dst_rows, PeerPtrs dst) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int dvec = d >> 2;
  for (int64_t i = warp; i < n_rows; i += n_warps) {
    const int64_t r = dst_rows ? dst_rows[i] : i;
    const float4* s = reinterpret_cast<const float4*>(src + i * (int64_t)d);
    for (int c = lane; c < dvec; c += 32) {
      const float4 x = __ldcg(s + c);
      for (int g = 0; g < dst.n; ++g) reinterpret_cast<float4*>(dst.p[g] + r * (int64_t)d)[c] = x;
    }
  }
  __threadfence_system();
}

__global__ void peer_signal_kernel(FlagPtrs f, uint32_t epoch) {
  // every earlier kernel of the stream has completed (plain launch) and fenced its peer stores; fence again, then announce
  __threadfence_system();
  const int g = threadIdx.x;
  if (g < f.n) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[g] + f.rank), "r"(epoch) : "memory");
}

__global__ void peer_wait_kernel(const uint32_t* flags, int world, uint32_t epoch, unsigned long long timeout_ns, uint32_t* err_flag) {
  const int g = threadIdx.x;
  bool ok = true;
  if (g < world) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + g) : "memory");
      if ((int32_t)(v - epoch) >= 0) break;
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        ok = false;
        break;
      }
      __nanosleep(100);
    }
  }
  if (!ok && err_flag) atomicOr(err_flag, LSTEP_FLAG_PEER_TIMEOUT);
  __threadfence_system();
}

}  // namespace

int peer_rows_bcast(const float* src, int64_t n_rows, int d, const int64_t* dst_rows, const lstep_peer_group* grp, int which, cudaStream_t st) {
  if (!valid_group(grp) || !src || n_rows < 0 || d <= 0 || d % 4 != 0) return LSTEP_ERR_INVALID_ARG;
  if (n_rows == 0) return LSTEP_OK;
  PeerPtrs dst{};
  for (int g = 0; g < grp->world; ++g) dst.p[dst.n++] = which == 0 ? grp->table[g] : grp->new_rows[g];
  const int64_t grid = std::min<int64_t>(ceil_div(n_rows * 32, 256), (int64_t)num_sms() * 8);
  peer_rows_bcast_kernel<<<(unsigned)grid, 256, 0, st>>>(src, n_rows, d, dst_rows, dst);
  return check_launch("peer_rows_bcast");
}

int peer_signal(const lstep_peer_group* g, uint32_t epoch, cudaStream_t st) {
  if (!valid_group(g)) return LSTEP_ERR_INVALID_ARG;
  FlagPtrs f{};
  for (int i = 0; i < g->world; ++i) f.p[i] = g->flags[i];
  f.n = g->world;
  f.rank = g->rank;
  peer_signal_kernel<<<1, 32, 0, st>>>(f, epoch);
  return check_launch("peer_signal");
}

int peer_wait(const lstep_peer_group* g, uint32_t epoch, int timeout_ms, uint32_t* err_flag, cudaStream_t st) {
  if (!valid_group(g)) return LSTEP_ERR_INVALID_ARG;
  const unsigned long long ns = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 2000) * 1000000ull;
  peer_wait_kernel<<<1, 32, 0, st>>>(g->flags[g->rank], g->world, epoch, ns, err_flag);
  return check_launch("peer_wait");
}

}  // namespace lstep

using namespace lstep;

extern "C" int lstep_ipc_alloc(size_t bytes, void** ptr) {
  if (!ptr || bytes == 0) return LSTEP_ERR_INVALID_ARG;
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, bytes);
  if (e != cudaSuccess) {
    set_cuda_error(e, "ipc_alloc");
    return LSTEP_ERR_CUDA;
  }
  return LSTEP_OK;
}
extern "C" int lstep_ipc_free(void* ptr) {
  if (!ptr) return LSTEP_OK;
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) {
    set_cuda_error(e, "ipc_free");
    return LSTEP_ERR_CUDA;
  }
  return LSTEP_OK;
}
extern "C" int lstep_ipc_export(void* ptr, unsigned char handle_out[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!ptr || !handle_out) return LSTEP_ERR_INVALID_ARG;
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) {
    set_cuda_error(e, "ipc_export");
    return LSTEP_ERR_CUDA;
  }
  memcpy(handle_out, &h, 64);
  return LSTEP_OK;
}
extern "C" int lstep_ipc_open(const unsigned char handle[64], void** ptr) {
  if (!handle || !ptr) return LSTEP_ERR_INVALID_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_cuda_error(e, "ipc_open");
    return LSTEP_ERR_CUDA;
  }
  return LSTEP_OK;
}
extern "C" int lstep_ipc_close(void* ptr) {
  if (!ptr) return LSTEP_OK;
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) {
    set_cuda_error(e, "ipc_close");
    return LSTEP_ERR_CUDA;
  }
  return LSTEP_OK;
}

extern "C" int lstep_peer_signal(const lstep_peer_group* g, uint32_t epoch, void* stream) { return peer_signal(g, epoch, as_stream(stream)); }
extern "C" int lstep_peer_wait(const lstep_peer_group* g, uint32_t epoch, int timeout_ms, uint32_t* err_flag, void* stream) {
  return peer_wait(g, epoch, timeout_ms, err_flag, as_stream(stream));
}

extern "C" int lstep_pe_step_peer(const lstep_pe_stream* s, const lstep_changelog* cl, const lstep_csr* csr, const lstep_peer_group* grp,
                                  int64_t lo, int64_t n_edges, const int64_t* ids, int64_t n_ids, const int64_t* ids_mine,
                                  const int64_t* pos_mine, int64_t n_mine, double current_time, int head, int len, const float* G,
                                  const int64_t* const* query_ids_host, int n_queries, int64_t q_off, int64_t q_rows, float* nbr_out,
                                  int K, const lstep_pe_mlp* mlp_nbr, const lstep_pe_mlp* mlp_upd, void* workspace,
                                  size_t workspace_bytes, uint32_t* err_flag, uint32_t epoch_base, int timeout_ms, int phases,
                                  void* stream) {
  if (!s || !s->src || !s->dst || !s->t || !s->cur || !cl || !csr || !valid_group(grp) || lo < 0 || n_edges < 0 || n_ids < 0 || n_mine < 0 ||
      n_mine > n_ids || !mlp_nbr || !mlp_upd || (phases & ~7) || q_rows < 0)
    return LSTEP_ERR_INVALID_ARG;
  if (grp->table[grp->rank] != s->cur || cl->row_mul != grp->world || cl->row_add != grp->rank || cl->d != s->d) return LSTEP_ERR_INVALID_ARG;
  if (head < 0 || head >= cl->T || len < 0 || len > cl->T) return LSTEP_ERR_INVALID_ARG;
  if (!update_push_available(mlp_upd)) return LSTEP_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  int rc;
  if (phases & 1) {
    if (n_mine > 0 && (rc = changelog_filter_peer(cl, head, len, ids_mine, n_mine, G, s->cur, s->d, ids_mine, grp, stream)) != LSTEP_OK) return rc;
    if ((rc = peer_signal(grp, epoch_base + 1, st)) != LSTEP_OK) return rc;
  }
  if (!(phases & 6)) return LSTEP_OK;
  if ((phases & 2) && (rc = peer_wait(grp, epoch_base + 1, timeout_ms, err_flag, st)) != LSTEP_OK) return rc;
  PeerPlan plan;
  plan.grp = grp;
  plan.ids_mine = ids_mine;
  plan.pos_mine = pos_mine;
  plan.n_mine = n_mine;
  plan.epoch2 = epoch_base + 2;
  plan.timeout_ms = timeout_ms;
  plan.phases = phases & 6;
  StepOpts opt;
  opt.skip_dft = opt.skip_append = true;
  opt.q_off = q_off;
  opt.q_rows = q_rows;
  opt.peer = &plan;
  int stamp = 0;
  opt.stamp_out = &stamp;
  lstep_pe_stream s2 = *s;
  if (!s2.ring) s2.ring = s2.cur;  // (never touched: the core neither filters nor appends)
  rc = pe_step_core_ex(&s2, csr, s->src + lo, s->dst + lo, s->t + lo, n_edges, ids, n_ids, current_time, 0, 0, 0, nullptr, query_ids_host,
                       q_rows > 0 ? n_queries : 0, nbr_out, K, mlp_nbr, mlp_upd, workspace, workspace_bytes, err_flag, stream, opt);
  if (rc != LSTEP_OK || !(phases & 4)) return rc;
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
  return changelog_append_peer(cl, slot, full ? 1 : 0, s->cur, U, n_dest, n_u_max, ids, n_ids, stamp_map, stamp, n_ids > 0 ? 1 : 0, err_flag, grp,
                               stream);
}

extern "C" int lstep_pe_steps_peer(const lstep_pe_stream* s, const lstep_changelog* cl, const lstep_csr* csr, const lstep_peer_group* grp,
                                   int64_t n_steps, const int64_t* lo_host, const int64_t* n_edges_host, const double* tmax_host,
                                   const int64_t* ids, const int64_t* ids_off_host, const int64_t* ids_mine, const int64_t* pos_mine,
                                   const int64_t* mine_off_host, int* head_io, const float* G, const int64_t* const* query_ids_host,
                                   const int64_t* q_base_host, int n_queries, float* nbr_out, int64_t out_step_stride, int K,
                                   const lstep_pe_mlp* mlp_nbr, const lstep_pe_mlp* mlp_upd, void* workspace, size_t workspace_bytes,
                                   uint32_t* err_flag, uint32_t* epoch_io, int timeout_ms, void* stream) {
  if (!cl || !valid_group(grp) || n_steps < 0 || !lo_host || !n_edges_host || !tmax_host || !ids || !ids_off_host || !mine_off_host || !head_io ||
      !epoch_io || n_queries < 0 || n_queries > 8 || (n_queries > 0 && (!query_ids_host || !q_base_host)))
    return LSTEP_ERR_INVALID_ARG;
  const int T = cl->T;
  int head = *head_io;
  uint32_t epoch = *epoch_io;
  int rc = LSTEP_OK;
  for (int64_t i = 0; i < n_steps; ++i) {
    const int64_t n = n_edges_host[i];
    const int64_t q_off = grp->rank * n / grp->world, q_rows = (grp->rank + 1) * n / grp->world - q_off;
    const int64_t* q[8] = {};
    for (int c = 0; c < n_queries; ++c) q[c] = query_ids_host[c] + q_base_host[i] + q_off;
    const int64_t m0 = mine_off_host[i], m1 = mine_off_host[i + 1];
    rc = lstep_pe_step_peer(s, cl, csr, grp, lo_host[i], n, ids + ids_off_host[i], ids_off_host[i + 1] - ids_off_host[i],
                            ids_mine ? ids_mine + m0 : nullptr, pos_mine ? pos_mine + m0 : nullptr, m1 - m0, tmax_host[i], head, T, G, q, n_queries,
                            q_off, q_rows, nbr_out ? nbr_out + i * out_step_stride : nullptr, K, mlp_nbr, mlp_upd, workspace, workspace_bytes,
                            err_flag, epoch, timeout_ms, 7, stream);
    if (rc != LSTEP_OK) break;
    epoch += 2;
    head = (head + 1) % T;
  }
  *head_io = head;
  *epoch_io = epoch;
  return rc;
}
