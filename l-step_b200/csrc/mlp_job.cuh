// Shared by the two PE-MLP kernels (csrc/mlp_cluster.cu: fp32 SIMT split-K clusters; csrc/mlp_umma.cu: tcgen05 3xTF32):
// the job / option structs of a launch and the mbarrier + bulk-copy primitives.
#pragma once
#include "common.cuh"

namespace lstep {

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    if (++spins > (1u << 24)) __trap();  // a lost bulk copy must fault, not hang the device
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(s_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}

// Optional fixed-point input (update_pe phase B, csrc/update_push.cu): row r of the aggregate is
// acc[r][0 .. d+t) in 32.32 fixed point instead of A[r][:]; reset_map[node of row r] is cleared once the
// row has been consumed (the claim map of the push kernel returns to all-zero).
struct FixedRows {
  const unsigned long long* acc;
  int32_t* reset_map;
  // late_trigger: let the NEXT kernel of the chain become resident only once this kernel's CTAs are past their
  // dependency wait, i.e. once everything BEFORE this kernel is complete. update_pe's phase-B push kernel relies on
  // it: its lookup / claim phase runs before its own wait, concurrently with the phase-A MLP and with nothing else.
  int late_trigger;
  // ring_slot != NULL (streaming step): every result row is also written into the history ring's new slot,
  // ring_slot + node * ring_stride — the step's ring append then only copies the rows this kernel does not write and
  // has nothing left to do after its dependency wait.
  float* ring_slot;
  int64_t ring_stride;
  // pe_rows > 0: base ids outside [0, pe_rows) read row 0 instead of memory beyond the table (the lookup kernels have
  // already raised LSTEP_FLAG_NODE_OUT_OF_RANGE for them: the caller sees IndexError, this kernel just must not fault)
  int64_t pe_rows;
  // ids_stable: the base-id arrays of the launch are not written by any kernel of the stream (the streaming step's
  // query / batch node lists): the first link of the kernel's dependent load chain, id -> base row, is then taken
  // before the dependency wait.
  int ids_stable;
};


// One launch can carry TWO independent MLP jobs (different rows, weights and outputs) on disjoint sets of clusters:
// clusters [0, split) run job 0, the rest job 1. The streaming step uses it for the neighbourhood MLP of the C query
// sets and the phase-A MLP of update_pe, which both only read the table (phase A then writes its rows to a side
// buffer that the push kernel applies): one launch, one set of fixed costs, instead of two links of the chain.
// Output fan-out of a job whose rows are the rows of n_uniq DISTINCT query sets of `period` rows each while the caller
// asked for n_out sets, some of them identical (the eval loop passes the batch's sources twice: positive and negative
// source, evaluate_model_utils.py:51-52): computed row u * period + i is stored to every output set c with
// src_of[c] == u, at out + (c * period + i) * out_stride. n_out == 0: plain out + row * out_stride.
struct OutFan {
  int64_t period;
  int n_out;
  signed char src_of[8];
};

struct MlpJob {
  const float* A;
  int64_t lda;
  RowIds base_ids;
  int64_t n_rows;
  const int32_t* n_rows_dev;
  lstep_pe_mlp m;
  float* out;
  int64_t out_stride;
  float* pe_inplace;
  OutFan fan;
};


// csrc/mlp_cluster.cu. expected_rows: the typical row count (n_rows is only an upper bound when n_rows_dev carries the real one);
// acc_fixed / reset_map / ring_slot: see FixedRows; fan: see OutFan; pe_rows > 0 clamps base ids to the table.
int launch_pe_mlp_cluster(const float* A, int64_t lda, const float* pe, RowIds base_ids, int64_t n_rows, int64_t expected_rows,
                          const int32_t* n_rows_dev, const lstep_pe_mlp* m, float* out, int64_t out_stride, float* pe_inplace,
                          const unsigned long long* acc_fixed, int32_t* reset_map, cudaStream_t st, bool late_trigger = false,
                          float* ring_slot = nullptr, int64_t ring_stride = 0, const OutFan* fan = nullptr, int64_t pe_rows = 0);
// two float-input jobs in one launch; LSTEP_ERR_UNSUPPORTED when they do not fit one round of clusters
int launch_pe_mlp_cluster_pair(const float* pe, const float* A0, int64_t lda0, RowIds ids0, int64_t rows0, const lstep_pe_mlp* m0,
                               float* out0, int64_t out_stride0, const float* A1, int64_t lda1, RowIds ids1, int64_t rows1,
                               const lstep_pe_mlp* m1, float* out1, int64_t out_stride1, cudaStream_t st, bool late_trigger,
                               int64_t pe_rows, const OutFan* fan0);
bool pe_mlp_cluster_supports(const lstep_pe_mlp* m);

// csrc/mlp_umma.cu: tensor-core form of the same launch (LSTEP_ERR_UNSUPPORTED when the shape / packing does not fit)
int launch_pe_mlp_umma(const MlpJob& j0, const MlpJob* j1, const float* pe, FixedRows fx, cudaStream_t st);
bool pe_mlp_umma_wanted(const lstep_pe_mlp* m, int64_t expected_rows);

}  // namespace lstep
