// Internal interface of the streaming step's core (csrc/step.cu), shared with the host-fed stepper and the change-log history.
#pragma once
#include "common.cuh"

namespace lstep {

// Owner-computes step of a peer group (csrc/peer.cu): phase A runs for the OWNED batch nodes only and its rows are stored
// into every rank's new_rows buffer; phase B runs for the owned destinations only. `phases` bit 1: everything up to the
// announcement of barrier 2; bit 2: from the wait on barrier 2 on (a single-process group runs bit 1 of every rank first).
struct PeerPlan {
  const lstep_peer_group* grp = nullptr;
  const int64_t* ids_mine = nullptr;  // owned batch nodes (ascending)
  const int64_t* pos_mine = nullptr;  // their positions in the batch's id list
  int64_t n_mine = 0;
  uint32_t epoch2 = 0;
  int timeout_ms = 2000;
  int phases = 6;
};
// Variations of the step for callers that own the history themselves (node-id sharded groups, the change-log history): the DFT
// filter and the ring append are then done by the caller, and the a6 queries may cover only a share [q_off, q_off + q_rows)
// of the batch's edges. stamp_out receives the step's stamp (the value the push kernel wrote into the per-node stamp map
// for every row phase B changed).
struct StepOpts {
  bool skip_dft = false, skip_append = false;
  int64_t q_off = 0, q_rows = -1;  // -1: all n_edges
  int* stamp_out = nullptr;
  const PeerPlan* peer = nullptr;
};
// csrc/peer.cu
int peer_rows_bcast(const float* src, int64_t n_rows, int d, const int64_t* dst_rows, const lstep_peer_group* grp, int which, cudaStream_t st);
int peer_signal(const lstep_peer_group* g, uint32_t epoch, cudaStream_t st);
int peer_wait(const lstep_peer_group* g, uint32_t epoch, int timeout_ms, uint32_t* err_flag, cudaStream_t st, bool announce = false);

int pe_step_core_ex(const lstep_pe_stream* s, const lstep_csr* csr, const int64_t* src, const int64_t* dst, const double* tq,
                    int64_t n_edges, const int64_t* ids, int64_t n_ids, double current_time, int head, int len, int append_slot,
                    const float* G, const int64_t* const* query_ids_host, int n_queries, float* nbr_out, int K,
                    const lstep_pe_mlp* mlp_nbr, const lstep_pe_mlp* mlp_upd, void* workspace, size_t workspace_bytes,
                    uint32_t* err_flag, void* stream, const StepOpts& opt);
int pe_step_core(const lstep_pe_stream* s, const lstep_csr* csr, const int64_t* src, const int64_t* dst, const double* tq,
                 int64_t n_edges, const int64_t* ids, int64_t n_ids, double current_time, int head, int len, int append_slot,
                 const float* G, const int64_t* const* query_ids_host, int n_queries, float* nbr_out, int K,
                 const lstep_pe_mlp* mlp_nbr, const lstep_pe_mlp* mlp_upd, void* workspace, size_t workspace_bytes,
                 uint32_t* err_flag, void* stream);
// lists phase B leaves in an update workspace: U (distinct destinations), their count (device), the per-node stamp map
void update_ws_phase_b_lists(void* workspace, int64_t n_ids, int64_t n_edges, int K, int d, int t, int64_t pe_rows, const int64_t** U,
                             const int32_t** n_dest_dev, const int32_t** stamp_map);
bool update_push_available(const lstep_pe_mlp* mlp);

}  // namespace lstep
