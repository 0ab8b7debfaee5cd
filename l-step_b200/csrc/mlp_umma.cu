// PE MLP on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) with fp32-grade accuracy ("BF16x3"):
//
//   out = base + tanh( [Ws base + bs] + W2 relu(W1 A + b1) + b2 )        (same function as csrc/mlp_cluster.cu;
//   models/LSTEP.py:240-247, :294-301, :329-336)
//
// Why: the MLP is the one contraction of the path. The fp32 SIMT kernel runs it at 13-17 % of the FMA peak and is 40 % of
// the step at B = 200; at B = 2000 (Flights shape: ~20 000 MLP rows, 4 GFLOP per step) the step is MLP-compute-bound.
// A 64-row tile costs 234 tensor instructions of 88 cycles here (20.6 k cycles) against ~53 k cycles of perfectly
// issued FFMA (and ~5x that at the SIMT kernel's measured efficiency).
//
// Accuracy: the 1e-5 parity bar rules out single-pass BF16 / TF32. Every fp32 operand is split EXACTLY into three
// bf16 values x = x1 + x2 + x3 (8 + 8 + 8 significand bits; x1 = bf16(x), x2 = bf16(x - x1), x3 = x - x1 - x2) and a
// product a*w is evaluated as the six partial products whose weight is >= 2^-16:
//     a3 w1 + a1 w3 + a2 w2 + a2 w1 + a1 w2 + a1 w1          (dropped: a2 w3, a3 w2, a3 w3 <= 2^-24 relative)
// each exact in the fp32 accumulator (8 x 8 significand bits). That is the per-product accuracy of an fp32 FMA; a bf16
// instruction covers K = 16, so the six passes cost what three TF32 passes (K = 8, 11-bit operands, 2^-22 per operand)
// would. What then limits the accuracy is the accumulator: every tcgen05.mma adds its K = 16 partial sum into the
// fp32 TMEM accumulator with TRUNCATION (measured: 234 adds into one accumulator leave a biased error of ~40 ulp,
// 7x the SIMT kernel's). The five small products are therefore accumulated in a SECOND accumulator (their
// truncation errors are 2^-8 smaller in absolute terms) and added to the large one once, in fp32 registers, when the
// tile is read back: 17 / 22 truncating adds remain on the large accumulator. Two accumulators per layer fit TMEM
// because the tile is M = 64: a 64-row accumulator occupies the lower 16 lanes of each 32-lane sub-partition, the
// second one the upper 16 lanes of the same columns (lane offset 16), so a 32x32b tcgen05.ld returns the large sums
// in lanes 0-15 and the small sums in lanes 16-31 of a warp and one shuffle adds them.
// Weights are split once at pack time (lstep_pack_linear_tc), activations while they are staged.
//
// Tile / pipeline (one CTA = one 64-row tile at a time, persistent over tiles; cta_group::1, M = 64, N = d padded to 16):
//   layer 1   D1[64 x N] = A[64 x K1] W1^T                          K1 = d+t padded to 16   (TMEM columns [0, N))
//   layer 2   D2[64 x N] = Base[64 x K2] Ws^T + H[64 x K2] W2^T     K2 = d padded to 16     (TMEM columns [256, 256+N))
//   with H = relu(D1 + b1) read back from TMEM (tcgen05.ld), split and re-staged as the A operand; the Base chunks are
//   issued first, so the tensor pipe stays busy while D1 is being drained.
// The K dimension is streamed in chunks of 16 (one instruction's K) through an 8-stage shared-memory ring; a stage holds
// the A chunk [3 parts][2 k-groups][64 rows][8 bf16] (6 KB) and the weight chunk [3 parts][2 k-groups][N rows][8 bf16]
// (16.5 KB at N = 176) in the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core matrices: SBO = 128 B between
// 8-row groups, LBO = rows * 16 B between the two k-groups), so one chunk = 6 tcgen05.mma instructions.
// Warp roles: warp 0 = TMEM allocation + MMA issue (one elected thread); warp 1 = weight loader (one cp.async.bulk per
// chunk: the packed weights are already in stage layout); warps 2-9 = two groups of four activation-producer warps that
// take alternate chunks (global -> split -> st.shared, or TMEM -> relu -> split -> st.shared; a warp reads the TMEM
// sub-partition warp % 4 = 16 rows of the tile) and share the epilogue (TMEM -> + bias, tanh, + base -> global) by
// alternate 16-column blocks.
// full / empty mbarriers per stage; tcgen05.commit frees a stage and publishes D1 / D2.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "mlp_job.cuh"

namespace lstep {
namespace {

constexpr int kM = 64;         // rows per tile (UMMA M)
constexpr int kKC = 16;        // K elements per chunk = K of one bf16 instruction (2 k-groups of 8 elements)
constexpr int kStages = 8;
constexpr int kGroups = 2;               // producer groups (alternate chunks)
constexpr int kGroupThreads = 128;       // 4 warps: one per TMEM sub-partition
constexpr int kProducers = kGroups * kGroupThreads;  // warps 2..9
constexpr int kThreads = 64 + kProducers;
constexpr int kD2Col = 256;    // TMEM column of the layer-2 accumulator
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSmallLane = 16u << 16;  // TMEM address offset of the second (small-term) accumulator: upper half sub-partitions

__device__ __forceinline__ void mb_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
// exact three-way bf16 split of two fp32 values, packed as bf16x2 words (low half = first element)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {  // one F2FP: {hi, lo} -> bf16x2, round to nearest even
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void split2(float x, float y, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
  p1 = pack_bf16x2(x, y);
  const float rx = x - __uint_as_float(p1 << 16), ry = y - __uint_as_float(p1 & 0xffff0000u);  // exact
  p2 = pack_bf16x2(rx, ry);
  p3 = pack_bf16x2(rx - __uint_as_float(p2 << 16), ry - __uint_as_float(p2 & 0xffff0000u));   // exact: <= 8 significant bits left
}
// 8 consecutive k of one row -> one 16-byte k-group in each of the three part planes
__device__ __forceinline__ void split8_store(const float (&v)[8], unsigned char* plane0, uint32_t part_stride, uint32_t off) {
  uint4 a, b, c;
  split2(v[0], v[1], a.x, b.x, c.x);
  split2(v[2], v[3], a.y, b.y, c.y);
  split2(v[4], v[5], a.z, b.z, c.z);
  split2(v[6], v[7], a.w, b.w, c.w);
  *reinterpret_cast<uint4*>(plane0 + off) = a;
  *reinterpret_cast<uint4*>(plane0 + part_stride + off) = b;
  *reinterpret_cast<uint4*>(plane0 + 2 * part_stride + off) = c;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, no swizzle: start address, LBO (between the two 16-byte k-groups of one instruction), SBO (between 8-row
// groups), all in 16-byte units; bits [46,48) = 1 (sm_100 descriptor version); layout type 0.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void producer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kProducers) : "memory"); }

#ifdef LSTEP_UMMA_TIMING
__device__ long long g_umma_clk[32];
#define UT_SET(i, v) do { if (blockIdx.x == 0) g_umma_clk[i] = (v); } while (0)
#define UT_ADD(i, v) do { if (blockIdx.x == 0) g_umma_clk[i] += (v); } while (0)
#else
#define UT_SET(i, v)
#define UT_ADD(i, v)
#endif

struct UmmaShape {
  int Np, K1p, K2p, Q1, Q2;
  uint32_t a_part, w_part, a_bytes, w_bytes, stage_bytes;
  size_t smem;
};
__host__ __device__ inline UmmaShape umma_shape(int d, int t) {
  UmmaShape s;
  s.Np = (int)align_up((size_t)d, 16);
  s.K1p = (int)align_up((size_t)(d + t), kKC);
  s.K2p = (int)align_up((size_t)d, kKC);
  s.Q1 = s.K1p / kKC;
  s.Q2 = s.K2p / kKC;
  s.a_part = 2u * kM * 16u;               // one part plane of the A chunk: [2 k-groups][64 rows][16 B]
  s.w_part = 2u * (uint32_t)s.Np * 16u;   // one part plane of the weight chunk: [2 k-groups][Np rows][16 B]
  s.a_bytes = 3u * s.a_part;
  s.w_bytes = 3u * s.w_part;
  s.stage_bytes = s.a_bytes + s.w_bytes;
  s.smem = (size_t)kStages * s.stage_bytes + 2 * (size_t)s.Np * sizeof(float) + 128;
  return s;
}

// packed (bf16): [chunk q][part][k-group][n][8]: element (n, k) of W[out, in] with q = k / 16, kg = (k % 16) / 8, e = k % 8
__global__ void pack_linear_umma_kernel(const float* __restrict__ w, int out_f, int in_f, int Np, int Kp, __nv_bfloat16* __restrict__ packed) {
  const int64_t total = (int64_t)Kp * Np;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / Np), n = (int)(i % Np);
    const float x = (n < out_f && k < in_f) ? w[(size_t)n * in_f + k] : 0.f;
    const __nv_bfloat16 x1 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(x1);
    const __nv_bfloat16 x2 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 x3 = __float2bfloat16_rn(r1 - __bfloat162float(x2));
    const int q = k / kKC, kg = (k % kKC) / 8, e = k % 8;
    const size_t part = (size_t)2 * Np * 8;  // elements per part plane
    const size_t base = (size_t)q * 3 * part + ((size_t)kg * Np + n) * 8 + e;
    packed[base] = x1;
    packed[base + part] = x2;
    packed[base + 2 * part] = x3;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
    pe_mlp_umma_kernel(const __grid_constant__ MlpJob job0, const __grid_constant__ MlpJob job1, int split, const float* pe, FixedRows fx) {
  const bool second = (int)blockIdx.x >= split;
  const MlpJob& jb = second ? job1 : job0;
  const lstep_pe_mlp& m = jb.m;
  const int d = m.d, in1 = m.d + m.t;
  const bool has_self = m.ws_tc != nullptr;
  const UmmaShape sh = umma_shape(d, m.t);
  const int Np = sh.Np;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t d1_full, d2_full, tmem_free;
  __shared__ uint32_t s_tmem;
  __shared__ int64_t s_node[kM];
  unsigned char* stages = smem_raw;
  float* bias1 = reinterpret_cast<float*>(smem_raw + (size_t)kStages * sh.stage_bytes);
  float* bias2 = bias1 + Np;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  TL_ENTRY(fx.acc ? 4 : 2);
  if (!fx.late_trigger) pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mb_init(&full_bar[s], 1 + kGroupThreads / 32);  // weight loader (arrive.expect_tx) + one arrive per warp of the owning group
      mb_init(&empty_bar[s], 1);                      // tcgen05.commit
    }
    mb_init(&d1_full, 1);
    mb_init(&d2_full, 1);
    mb_init(&tmem_free, kProducers / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&s_tmem)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int c = tid; c < Np; c += kThreads) {  // packed biases are zero padded to lstep_packed_ld(d) >= Np
    bias1[c] = m.b1[c];
    bias2[c] = m.b2[c] + (has_self ? m.bs[c] : 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  pdl_wait();  // from here on: data written by the preceding kernels (aggregate rows, table rows, the device row count)
  TL_WAITED(fx.acc ? 4 : 2);
  int64_t n_rows = jb.n_rows;
  if (jb.n_rows_dev) {
    const int64_t nd = ld_dep(jb.n_rows_dev);
    n_rows = nd < n_rows ? nd : n_rows;
  }
  if (fx.late_trigger) pdl_launch_dependents();
  const int64_t n_tiles = (n_rows + kM - 1) / kM;
  const int64_t cta = second ? (int64_t)blockIdx.x - split : (int64_t)blockIdx.x;
  const int64_t n_ctas = second ? (int64_t)gridDim.x - split : (int64_t)split;
  const int Q1 = sh.Q1, QB = has_self ? sh.Q2 : 0, QH = sh.Q2, QG = Q1 + QB, Q = QG + QH;
  const uint32_t a_lbo = kM * 16, w_lbo = (uint32_t)Np * 16, sbo = 128;
  // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Np >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
  const int64_t pe_rows = fx.pe_rows;
  auto clamp_row = [pe_rows](int64_t node) { return (pe_rows > 0 && (node < 0 || node >= pe_rows)) ? (int64_t)0 : node; };

  if (warp == 0) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      uint32_t g = 0, it = 0;
      for (int64_t tile = cta; tile < n_tiles; tile += n_ctas, ++it) {
        mb_wait(&tmem_free, (it & 1) ^ 1);  // the previous tile's epilogue has drained D1 / D2 (first tile: passes at once)
        tc_fence_after();
#ifdef LSTEP_UMMA_TIMING
        const long long t_tile = clock64();
        if (it == 0) { UT_SET(0, t_tile); UT_SET(1, 0); UT_SET(2, 0); UT_SET(3, 0); }
#endif
        for (int q = 0; q < Q; ++q, ++g) {
          const uint32_t s = g % kStages, ph = (g / kStages) & 1;
#ifdef LSTEP_UMMA_TIMING
          const long long tw0 = clock64();
#endif
          mb_wait(&full_bar[s], ph);
#ifdef LSTEP_UMMA_TIMING
          if (it == 0) UT_ADD(q < Q1 ? 1 : (q < QG ? 2 : 3), clock64() - tw0);  // waiting for operands: layer 1 / base / H chunks
#endif
          tc_fence_after();
          const uint32_t a0 = s_u32(stages + (size_t)s * sh.stage_bytes), w0 = a0 + sh.a_bytes;
          const uint32_t dbig = tmem + (q < Q1 ? 0u : (uint32_t)kD2Col), dsmall = dbig + kSmallLane;
          const uint32_t first = (q == 0 || q == Q1) ? 0u : 1u;
          uint64_t da[3], dw[3];
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            da[p] = umma_desc(a0 + (uint32_t)p * sh.a_part, a_lbo, sbo);
            dw[p] = umma_desc(w0 + (uint32_t)p * sh.w_part, w_lbo, sbo);
          }
          // the six partial products of weight >= 2^-16: the large one into the large accumulator, the five small ones
          // (smallest first) into the small one
          umma_bf16(dbig, da[0], dw[0], idesc, first);
          umma_bf16(dsmall, da[2], dw[0], idesc, first);
          umma_bf16(dsmall, da[0], dw[2], idesc, 1u);
          umma_bf16(dsmall, da[1], dw[1], idesc, 1u);
          umma_bf16(dsmall, da[1], dw[0], idesc, 1u);
          umma_bf16(dsmall, da[0], dw[1], idesc, 1u);
          umma_commit(&empty_bar[s]);              // stage free once these MMAs have read it
          if (q == Q1 - 1) umma_commit(&d1_full);  // layer 1 complete
        }
        umma_commit(&d2_full);
#ifdef LSTEP_UMMA_TIMING
        if (it == 0) UT_SET(4, clock64() - t_tile);  // issue loop of the tile
#endif
      }
    }
  } else if (warp == 1) {
    // ===================================================================== weight loader
    if (lane == 0) {
      uint32_t g = 0;
      const size_t chunk_floats = sh.w_bytes / 4;
      for (int64_t tile = cta; tile < n_tiles; tile += n_ctas) {
        for (int q = 0; q < Q; ++q, ++g) {
          const uint32_t s = g % kStages, ph = (g / kStages) & 1;
          mb_wait(&empty_bar[s], ph ^ 1);
          const float* src = q < Q1 ? m.w1_tc + (size_t)q * chunk_floats
                                    : (q < QG ? m.ws_tc + (size_t)(q - Q1) * chunk_floats : m.w2_tc + (size_t)(q - QG) * chunk_floats);
          mb_expect_tx(&full_bar[s], sh.w_bytes);
          bulk_load(stages + (size_t)s * sh.stage_bytes + sh.a_bytes, src, sh.w_bytes, &full_bar[s]);
        }
      }
    }
  } else {
    // ===================================================================== activation producers + epilogue (2 x 128 threads)
    const int pt = tid - 64;                 // 0..255
    const int grp = pt / kGroupThreads;      // producer group: takes the chunks q with q % 2 == grp
    const int gw = (pt % kGroupThreads) >> 5;  // warp inside the group 0..3
    const int sub = warp & 3;                // TMEM sub-partition this warp may read: lanes [32*sub, 32*sub + 32)
    // M = 64 accumulator layout: row r lives in TMEM lane (r % 16) + 32 * (r / 16); the small-term accumulator in lane + 16.
    // A 32x32b load of sub-partition `sub` therefore gives lanes 0-15 the large sums and lanes 16-31 the small sums of
    // rows 16 * sub + (lane % 16); after the shuffle-add, lanes 0-15 keep columns 0-7 of a 16-column block (k-group 0)
    // and lanes 16-31 columns 8-15 (k-group 1).
    const int my_row = 16 * sub + (lane & 15);
    const int my_half = lane >> 4;
    const uint32_t t_lane = tmem + ((uint32_t)(32 * sub) << 16);
    // gather mapping: 16 consecutive rows x 2 k-groups per warp (64-byte row segments; 8 lanes of a store wavefront
    // write 8 consecutive rows of one k-group: conflict free); the 4 warps of a group cover the 64 rows of the tile
    const int g_r = lane & 15, g_kg = lane >> 4;
    const bool a_vec = (jb.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(jb.A) & 15) == 0);
    const bool p_vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(pe) & 15) == 0);
    const float* __restrict__ A = jb.A;
    const int64_t lda = jb.lda;
    uint32_t it = 0;
    int64_t g_tile = 0;  // global chunk counter at the start of the current tile
    for (int64_t tile = cta; tile < n_tiles; tile += n_ctas, ++it, g_tile += Q) {
      const int64_t row0 = tile * kM;
      if (pt < kM) {
        const int64_t row = row0 + pt;
        s_node[pt] = row < n_rows ? clamp_row(jb.base_ids.at_dep(row)) : 0;
      }
      producer_sync();
#ifdef LSTEP_UMMA_TIMING
      const long long tp0 = clock64();
      if (it == 0 && pt == 0) UT_SET(8, tp0);
#endif
      // this thread's 8 consecutive k of one row of global chunk q
      auto load_item = [&](int q, float (&v)[8]) {
        const int r = gw * 16 + g_r;
        const int64_t row = row0 + r;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (row >= n_rows) return;
        const bool is_a = q < Q1;
        const int kg0 = (is_a ? q : q - Q1) * kKC + g_kg * 8;
        const int lim = is_a ? in1 : d;
        if (is_a && fx.acc) {
          constexpr float kInv = 2.3283064365386963e-10f;  // 2^-32
          const unsigned long long* p64 = fx.acc + row * (int64_t)in1 + kg0;
          if (kg0 + 7 < lim && (reinterpret_cast<uintptr_t>(p64) & 15) == 0) {
            ulonglong2 x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = __ldcg(reinterpret_cast<const ulonglong2*>(p64) + i);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              v[2 * i] = (float)(long long)x[i].x * kInv;
              v[2 * i + 1] = (float)(long long)x[i].y * kInv;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (kg0 + i < lim) v[i] = (float)(long long)__ldcg(p64 + i) * kInv;
          }
          return;
        }
        const float* p32 = is_a ? A + row * lda + kg0 : pe + s_node[r] * (int64_t)d + kg0;
        if ((is_a ? a_vec : p_vec) && kg0 + 7 < lim) {
          const float4 x = ld_dep(reinterpret_cast<const float4*>(p32)), y = ld_dep(reinterpret_cast<const float4*>(p32) + 1);
          v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
          v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (kg0 + i < lim) v[i] = ld_dep(p32 + i);
        }
      };
      auto publish = [&](int q, const float (&v0)[8]) {
        const uint32_t g = (uint32_t)(g_tile + q);
        const uint32_t s = g % kStages, ph = (g / kStages) & 1;
        mb_wait(&empty_bar[s], ph ^ 1);
        unsigned char* st_a = stages + (size_t)s * sh.stage_bytes;
        split8_store(v0, st_a, sh.a_part, (uint32_t)(g_kg * kM + gw * 16 + g_r) * 16u);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mb_arrive(&full_bar[s]);
      };
      // ---- global chunks of this group (q = grp, grp + 2, ...), one chunk of register prefetch (two static buffers)
      {
        float b0[8], b1[8], b2[8];  // two chunks of this group in flight beyond the one being published
        int q = grp;
        if (q < QG) load_item(q, b0);
        if (q + 2 < QG) load_item(q + 2, b1);
        while (q < QG) {
          if (q + 4 < QG) load_item(q + 4, b2);
          publish(q, b0);
          q += 2;
          if (q >= QG) break;
          if (q + 4 < QG) load_item(q + 4, b0);
          publish(q, b1);
          q += 2;
          if (q >= QG) break;
          if (q + 4 < QG) load_item(q + 4, b1);
          publish(q, b2);
          q += 2;
        }
      }
      // ---- H chunks: relu(D1 + b1) from TMEM, one row per thread
#ifdef LSTEP_UMMA_TIMING
      if (it == 0 && pt == 0) UT_SET(9, clock64() - tp0);   // global chunks published
#endif
      mb_wait(&d1_full, it & 1);
#ifdef LSTEP_UMMA_TIMING
      if (it == 0 && pt == 0) UT_SET(10, clock64() - tp0);  // D1 complete
#endif
      tc_fence_after();
      for (int qh = ((QG & 1) == grp ? 0 : 1); qh < QH; qh += 2) {  // tile chunk QG + qh belongs to group (QG + qh) % 2
        const uint32_t g = (uint32_t)(g_tile + QG + qh);
        const uint32_t s = g % kStages, ph = (g / kStages) & 1;
        float v[16];
        tmem_ld16(t_lane + (uint32_t)(qh * kKC), v);
        float h8[8];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float sum = v[i] + __shfl_xor_sync(kFull, v[i], 16);  // large + small accumulator (commutative: both lanes agree)
          if ((i >> 3) == my_half) h8[i & 7] = fmaxf(sum + bias1[qh * kKC + i], 0.f);
        }
        mb_wait(&empty_bar[s], ph ^ 1);
        unsigned char* st_a = stages + (size_t)s * sh.stage_bytes;
        split8_store(h8, st_a, sh.a_part, (uint32_t)(my_half * kM + my_row) * 16u);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mb_arrive(&full_bar[s]);
      }
      // ---- epilogue: out = base + tanh(D2 + b2 (+ bs)), one row per thread, 16 columns per TMEM load; the two groups
      // take alternate 16-column blocks
#ifdef LSTEP_UMMA_TIMING
      if (it == 0 && pt == 0) UT_SET(11, clock64() - tp0);  // H chunks published
#endif
      mb_wait(&d2_full, it & 1);
#ifdef LSTEP_UMMA_TIMING
      if (it == 0 && pt == 0) UT_SET(12, clock64() - tp0);  // D2 complete
#endif
      tc_fence_after();
      const int64_t row = row0 + my_row;
      const bool live = row < n_rows;
      const int64_t node = s_node[my_row];
      const float* brow = pe + node * (int64_t)d;
      float* orow = jb.out ? jb.out + row * jb.out_stride : jb.pe_inplace + node * (int64_t)d;
      float* rrow = fx.ring_slot ? fx.ring_slot + node * fx.ring_stride : nullptr;
      const bool o_vec = p_vec && (jb.out ? (jb.out_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(jb.out) & 15) == 0)
                                          : (reinterpret_cast<uintptr_t>(jb.pe_inplace) & 15) == 0);
      for (int cb = 16 * grp; cb < Np; cb += 16 * kGroups) {
        float v[16];
        tmem_ld16(t_lane + (uint32_t)(kD2Col + cb), v);  // (warp-collective: every lane takes part, live or not)
        float z8[8];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float sum = v[i] + __shfl_xor_sync(kFull, v[i], 16);
          if ((i >> 3) == my_half) z8[i & 7] = sum;
        }
        if (!live) continue;
#pragma unroll
        for (int kg = 0; kg < 2; ++kg) {
          const int c = cb + my_half * 8 + kg * 4;
          if (c >= d) break;
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (o_vec && c + 3 < d) {
            b = ld_dep(reinterpret_cast<const float4*>(brow + c));
          } else {
            b.x = ld_dep(brow + c);
            if (c + 1 < d) b.y = ld_dep(brow + c + 1);
            if (c + 2 < d) b.z = ld_dep(brow + c + 2);
            if (c + 3 < d) b.w = ld_dep(brow + c + 3);
          }
          float4 o;
          o.x = b.x + tanhf(z8[kg * 4 + 0] + bias2[c + 0]);
          o.y = b.y + tanhf(z8[kg * 4 + 1] + bias2[c + 1]);
          o.z = b.z + tanhf(z8[kg * 4 + 2] + bias2[c + 2]);
          o.w = b.w + tanhf(z8[kg * 4 + 3] + bias2[c + 3]);
          auto store = [&](float* dst) {
            if (o_vec && c + 3 < d) {
              *reinterpret_cast<float4*>(dst) = o;
            } else {
              dst[0] = o.x;
              if (c + 1 < d) dst[1] = o.y;
              if (c + 2 < d) dst[2] = o.z;
              if (c + 3 < d) dst[3] = o.w;
            }
          };
          if (jb.out && jb.fan.n_out > 0) {  // identical query sets were computed once: fan the row out
            const int64_t u = row / jb.fan.period, i = row % jb.fan.period;
            for (int cc = 0; cc < jb.fan.n_out; ++cc)
              if (jb.fan.src_of[cc] == u) store(jb.out + (cc * jb.fan.period + i) * jb.out_stride + c);
          } else {
            store(orow + c);
          }
          if (rrow) store(rrow + c);
        }
      }
      if (fx.reset_map && live && grp == 0 && my_half == 0) fx.reset_map[node] = 0;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mb_arrive(&tmem_free);
#ifdef LSTEP_UMMA_TIMING
      if (it == 0 && pt == 0) UT_SET(13, clock64() - tp0);  // epilogue done
#endif
      producer_sync();  // s_node is rewritten by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
  TL_EXIT(fx.acc ? 4 : 2);
}

}  // namespace

bool pe_mlp_umma_wanted(const lstep_pe_mlp* m, int64_t expected_rows) {
  const Tuning& tn = tuning();
  if (!tn.mlp_umma || !m || !m->w1_tc || !m->w2_tc) return false;
  const UmmaShape sh = umma_shape(m->d, m->t);
  if (sh.Np > 256 || sh.smem > 224 * 1024) return false;
  return expected_rows >= tn.mlp_umma_min_rows;
}

int launch_pe_mlp_umma(const MlpJob& j0, const MlpJob* j1, const float* pe, FixedRows fx, cudaStream_t st) {
  const lstep_pe_mlp* m = &j0.m;
  if (!m->w1_tc || !m->w2_tc) return LSTEP_ERR_UNSUPPORTED;
  if (j1 && (j1->m.d != m->d || j1->m.t != m->t || (j1->m.ws_tc != nullptr) != (m->ws_tc != nullptr) || !j1->m.w1_tc || !j1->m.w2_tc))
    return LSTEP_ERR_UNSUPPORTED;
  if ((m->ws != nullptr) != (m->ws_tc != nullptr)) return LSTEP_ERR_UNSUPPORTED;
  const UmmaShape sh = umma_shape(m->d, m->t);
  if (sh.Np > 256 || sh.smem > 224 * 1024) return LSTEP_ERR_UNSUPPORTED;
  static size_t attr_done = 0;  // dynamic shared memory the kernel has been allowed so far
  if (attr_done < sh.smem) {
    cudaError_t e = cudaFuncSetAttribute(pe_mlp_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
    if (e != cudaSuccess) {
      set_cuda_error(e, "pe_mlp_umma attr");
      return LSTEP_ERR_CUDA;
    }
    attr_done = sh.smem;
  }
  const int sms = num_sms();
  int64_t c0 = ceil_div(j0.n_rows, kM), c1 = j1 ? ceil_div(j1->n_rows, kM) : 0;
  if (c0 + c1 > sms) {  // persistent: CTAs walk the tiles of their job
    if (j1) {
      const int64_t a = std::max<int64_t>(1, c0 * sms / (c0 + c1));
      c1 = std::max<int64_t>(1, sms - a);
      c0 = a;
    } else {
      c0 = sms;
    }
  }
  launch_k(pe_mlp_umma_kernel, dim3((unsigned)(c0 + c1)), dim3(kThreads), sh.smem, st, j0, j1 ? *j1 : j0, (int)c0, pe, fx);
  return check_launch("pe_mlp_umma");
}

#ifdef LSTEP_UMMA_TIMING
extern "C" int lstep_debug_umma_clocks(long long* out32) {
  return cudaMemcpyFromSymbol(out32, g_umma_clk, sizeof(long long) * 32) == cudaSuccess ? 0 : 4;
}
#endif

}  // namespace lstep

using namespace lstep;

extern "C" size_t lstep_packed_tc_floats(int out_features, int in_features) {
  if (out_features <= 0 || in_features <= 0) return 0;
  // three bf16 planes per element = 6 bytes = 1.5 floats
  return ((size_t)3 * align_up((size_t)in_features, kKC) * align_up((size_t)out_features, 16) + 1) / 2;
}

extern "C" int lstep_pack_linear_tc(const float* weight, int out_features, int in_features, float* packed, void* stream) {
  if (!weight || !packed || out_features <= 0 || in_features <= 0) return LSTEP_ERR_INVALID_ARG;
  const int Np = (int)align_up((size_t)out_features, 16), Kp = (int)align_up((size_t)in_features, kKC);
  pack_linear_umma_kernel<<<128, 256, 0, as_stream(stream)>>>(weight, out_features, in_features, Np, Kp, reinterpret_cast<__nv_bfloat16*>(packed));
  return check_launch("pack_linear_tc");
}
