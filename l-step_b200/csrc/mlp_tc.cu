// PE MLP on the tensor cores with fp32-grade accuracy ("3xTF32").
//
//   out = base + tanh( [Ws base + bs] + W2 relu(W1 A + b1) + b2 )        (same function as csrc/mlp.cu)
//
// The parity bar (1e-5) rules out single-pass TF32: a TF32 operand keeps 11 significand bits. Every fp32
// operand x is therefore split as x = hi + lo with hi = tf32(x), lo = tf32(x - hi) (22 of the 24 bits), and
// each product is evaluated as hi*hi + hi*lo + lo*hi in three tensor-core passes that accumulate in fp32 —
// the dropped lo*lo term is 2^-22 relative, the same order as fp32 rounding of a 272-term dot product.
//
// Why mma.sync (m16n8k8) and not tcgen05: the headline workload has 300..1200 MLP rows per launch. A
// tcgen05 tile is 128 (or 64) rows — 7..13 CTAs would carry the whole launch — while 16-row warp-level
// tiles spread the same work over 50..75 SMs; at these sizes the launch is bound by issue and by streaming
// the 0.9 MB of weights into each SM, not by tensor throughput. (A tcgen05 variant only pays off at the
// 10^4-row launches of the B=2000 configs; see DESIGN.md.)
//
// Layout / schedule
//   * weights are pre-packed (lstep_pack_linear_tc) in B-fragment order, hi and lo interleaved: for every
//     k-step of 8 input rows and n-tile of 8 outputs, lane l holds {b0_hi, b1_hi, b0_lo, b1_lo} with
//     b0 = W[n = 8nt + l/4][k = 8ks + l%4], b1 = W[n][k + 4] — one conflict-free 128-bit shared load per
//     (k-step, n-tile) feeds three MMAs;
//   * a CTA owns R = 16*MT rows and all ldo columns; 8 consumer warps split the n-tiles, one producer warp
//     streams the flat weight-tile sequence (16 input rows per tile) with cp.async.bulk into a 4-stage ring
//     (full / empty mbarriers), persistent over row tiles exactly like the SIMT kernel;
//   * activations are staged once per row tile as raw fp32, row-major with a +4 pitch so the A-fragment
//     loads (row = lane/4 (+8), col = lane%4 (+4)) hit 32 distinct banks; the hi / lo split of an A
//     fragment is three ALU instructions per element, done in registers.
#include "common.cuh"

namespace lstep {

constexpr int kTcKTile = 16;   // input rows per streamed weight tile (2 k-steps of 8)
constexpr int kTcStages = 4;
constexpr int kTcWarps = 8;    // consumer warps

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(to_tf32(x));
  lo = __uint_as_float(to_tf32(x - hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// packed[(ks * NT + nt) * 32 + lane] = {b0_hi, b1_hi, b0_lo, b1_lo}
__global__ void pack_linear_tc_kernel(const float* __restrict__ w, int out_f, int in_f, int in_pad, int ldo,
                                      float4* __restrict__ packed) {
  const int NT = ldo >> 3;
  const int64_t total = (int64_t)(in_pad >> 3) * NT * 32;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int lane = (int)(i & 31);
    const int64_t f = i >> 5;
    const int nt = (int)(f % NT), ks = (int)(f / NT);
    const int n = nt * 8 + (lane >> 2), k0 = ks * 8 + (lane & 3), k1 = k0 + 4;
    const float b0 = (n < out_f && k0 < in_f) ? w[(size_t)n * in_f + k0] : 0.f;
    const float b1 = (n < out_f && k1 < in_f) ? w[(size_t)n * in_f + k1] : 0.f;
    float4 o;
    split_tf32(b0, o.x, o.z);
    split_tf32(b1, o.y, o.w);
    packed[i] = o;
  }
}

// mbarrier / bulk-copy primitives (same as csrc/mlp.cu)
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    if (++spins > (1u << 24)) __trap();  // a lost bulk copy must fault, not hang the device
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(tc_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tc_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kTcWarps * 32) : "memory"); }

// dynamic smem: weight ring [kTcStages][tile_f4 float4] | A1 [R][p1] | A2 [R][p2]
//   A1 = aggregate rows (in1_pad columns), A2 = [h ; base] rows (2*d_pad columns); p = columns + 4
template <int MT, int NTW>
__global__ void __launch_bounds__((kTcWarps + 1) * 32) pe_mlp_tc_kernel(const float* __restrict__ A, int64_t lda, const float* pe,
                                                                        RowIds base_ids, int64_t n_rows,
                                                                        const int32_t* __restrict__ n_rows_dev, lstep_pe_mlp m,
                                                                        int ldo, float* __restrict__ out, int64_t out_stride,
                                                                        float* pe_inplace) {
  constexpr int R = 16 * MT;
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t full_bar[kTcStages];
  __shared__ __align__(8) uint64_t empty_bar[kTcStages];
  const int d = m.d, in1 = m.d + m.t;
  const int in1_pad = (in1 + kTcKTile - 1) / kTcKTile * kTcKTile;
  const int d_pad = (d + kTcKTile - 1) / kTcKTile * kTcKTile;
  if (n_rows_dev) {
    const int64_t nd = *n_rows_dev;
    n_rows = nd < n_rows ? nd : n_rows;
  }
  const int64_t n_row_tiles = (n_rows + R - 1) / R;
  if ((int64_t)blockIdx.x >= n_row_tiles) return;
  const int NT = ldo >> 3;
  const int tile_f4 = 2 * NT * 32;  // float4 per streamed weight tile (2 k-steps)
  const uint32_t tile_bytes = (uint32_t)tile_f4 * 16u;
  const int p1 = in1_pad + 4, p2 = 2 * d_pad + 4;
  float4* Wst = reinterpret_cast<float4*>(smem);
  float* A1 = smem + (size_t)kTcStages * tile_f4 * 4;
  float* A2 = A1 + (size_t)R * p1;
  const int tid = threadIdx.x;
  const bool has_self = m.ws_tc != nullptr;
  const float4* w1p = reinterpret_cast<const float4*>(m.w1_tc);
  const float4* w2p = reinterpret_cast<const float4*>(m.w2_tc);
  const float4* wsp = reinterpret_cast<const float4*>(m.ws_tc);
  const int t1 = in1_pad / kTcKTile, t2 = d_pad / kTcKTile;
  const int ntiles = t1 + t2 + (has_self ? t2 : 0);
  const int64_t my_tiles = (n_row_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      tc_mbar_init(&full_bar[s], 1);
      tc_mbar_init(&empty_bar[s], kTcWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= kTcWarps * 32) {  // ---- producer warp: stream the weight tiles of every row tile
    if (tid == kTcWarps * 32) {
      int64_t it = 0;
      for (int64_t rt = 0; rt < my_tiles; ++rt) {
        for (int i = 0; i < ntiles; ++i, ++it) {
          const int s = (int)(it % kTcStages);
          if (it >= kTcStages) tc_mbar_wait(&empty_bar[s], (uint32_t)(((it / kTcStages) - 1) & 1));
          const float4* src = i < t1 ? w1p + (size_t)i * tile_f4
                                     : (i < t1 + t2 ? w2p + (size_t)(i - t1) * tile_f4 : wsp + (size_t)(i - t1 - t2) * tile_f4);
          tc_mbar_expect_tx(&full_bar[s], tile_bytes);
          tc_bulk_g2s(Wst + (size_t)s * tile_f4, src, tile_bytes, &full_bar[s]);
        }
      }
    }
    return;
  }

  const int warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int ncons = kTcWarps * 32;
  // n-tiles of this warp: nt = warp + 8*j, j < NTW (strided so every warp gets work when NT < 8*NTW)
  float acc[MT][NTW][4];

  auto zero_acc = [&]() {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int j = 0; j < NTW; ++j) acc[mt][j][0] = acc[mt][j][1] = acc[mt][j][2] = acc[mt][j][3] = 0.f;
  };
  // one streamed weight tile (2 k-steps) against the activation plane `Ap` with pitch p, columns k0..k0+15
  auto mma_tile = [&](const float4* wt, const float* Ap, int p, int k0) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t ah[MT][4], al[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int r0 = (mt * 16 + gid) * p + k0 + ks * 8 + tig;
        const float a0 = Ap[r0], a1 = Ap[r0 + 8 * p], a2 = Ap[r0 + 4], a3 = Ap[r0 + 8 * p + 4];
        float h, l;
        split_tf32(a0, h, l);
        ah[mt][0] = __float_as_uint(h);
        al[mt][0] = __float_as_uint(l);
        split_tf32(a1, h, l);
        ah[mt][1] = __float_as_uint(h);
        al[mt][1] = __float_as_uint(l);
        split_tf32(a2, h, l);
        ah[mt][2] = __float_as_uint(h);
        al[mt][2] = __float_as_uint(l);
        split_tf32(a3, h, l);
        ah[mt][3] = __float_as_uint(h);
        al[mt][3] = __float_as_uint(l);
      }
#pragma unroll
      for (int j = 0; j < NTW; ++j) {
        const int nt = warp + kTcWarps * j;
        if (nt < NT) {
          const float4 b = wt[(size_t)(ks * NT + nt) * 32 + lane];
          const uint32_t bh0 = __float_as_uint(b.x), bh1 = __float_as_uint(b.y), bl0 = __float_as_uint(b.z), bl1 = __float_as_uint(b.w);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_tf32(acc[mt][j], al[mt], bh0, bh1);  // small terms first
            mma_tf32(acc[mt][j], ah[mt], bl0, bl1);
            mma_tf32(acc[mt][j], ah[mt], bh0, bh1);
          }
        }
      }
    }
  };

  for (int64_t rt = 0; rt < my_tiles; ++rt) {
    const int64_t row0 = ((int64_t)blockIdx.x + rt * gridDim.x) * R;
    const int64_t it0 = rt * ntiles;
    // ---- stage the aggregate rows (A1) and the base rows (second half of A2)
    for (int idx = tid; idx < R * in1_pad; idx += ncons) {
      const int r = idx / in1_pad, k = idx % in1_pad;
      const float v = (row0 + r < n_rows && k < in1) ? A[(row0 + r) * lda + k] : 0.f;
      A1[r * p1 + k] = v;
    }
    if (has_self) {
      for (int idx = tid; idx < R * d_pad; idx += ncons) {
        const int r = idx / d_pad, k = idx % d_pad;
        A2[r * p2 + d_pad + k] = (row0 + r < n_rows && k < d) ? pe[base_ids.at(row0 + r) * (int64_t)d + k] : 0.f;
      }
    }
    tc_consumer_sync();

    // ---- layer 1: h = relu(W1 a + b1) -> first half of A2 (hi / lo), padded columns zero
    zero_acc();
    for (int i = 0; i < t1; ++i) {
      const int64_t it = it0 + i;
      const int s = (int)(it % kTcStages);
      tc_mbar_wait(&full_bar[s], (uint32_t)((it / kTcStages) & 1));
      mma_tile(Wst + (size_t)s * tile_f4, A1, p1, i * kTcKTile);
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(&empty_bar[s]);
    }
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      const int nt = warp + kTcWarps * j;
      if (nt < NT) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = mt * 16 + gid + (e >> 1) * 8, c = nt * 8 + 2 * tig + (e & 1);
            if (c < d_pad) {
              A2[r * p2 + c] = c < d ? fmaxf(acc[mt][j][e] + m.b1[c], 0.f) : 0.f;
            }
          }
      }
    }
    tc_consumer_sync();

    // ---- layer 2 (+ self term): z = W2 h [+ Ws base] over the concatenated k-run [h ; base]
    zero_acc();
    for (int i = t1; i < ntiles; ++i) {
      const int64_t it = it0 + i;
      const int s = (int)(it % kTcStages);
      tc_mbar_wait(&full_bar[s], (uint32_t)((it / kTcStages) & 1));
      mma_tile(Wst + (size_t)s * tile_f4, A2, p2, (i - t1) * kTcKTile);
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(&empty_bar[s]);
    }
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      const int nt = warp + kTcWarps * j;
      if (nt < NT) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = mt * 16 + gid + (e >> 1) * 8, c = nt * 8 + 2 * tig + (e & 1);
            const int64_t row = row0 + r;
            if (row < n_rows && c < d) {
              const int64_t bid = base_ids.at(row);
              const float base = pe[bid * (int64_t)d + c];  // exact fp32 base for the residual (hi+lo keeps 22 bits)
              const float z = acc[mt][j][e] + m.b2[c] + (has_self ? m.bs[c] : 0.f);
              const float o = base + tanhf(z);
              if (out)
                out[row * out_stride + c] = o;
              else
                pe_inplace[bid * (int64_t)d + c] = o;
            }
          }
      }
    }
    tc_consumer_sync();  // the activation planes are reused by the next row tile
  }
}

template <int MT, int NTW>
static int launch_tc(const float* A, int64_t lda, const float* pe, RowIds base_ids, int64_t n_rows, const int32_t* n_rows_dev,
                     const lstep_pe_mlp* m, float* out, int64_t out_stride, float* pe_inplace, cudaStream_t st) {
  constexpr int R = 16 * MT;
  const int ldo = lstep_packed_ld(m->d);
  const int in1_pad = lstep_packed_rows(m->d + m->t), d_pad = lstep_packed_rows(m->d);
  const int NT = ldo >> 3;
  const size_t smem = (size_t)kTcStages * 2 * NT * 32 * 16 + sizeof(float) * (size_t)R * ((in1_pad + 4) + (2 * d_pad + 4));
  if (smem > 220 * 1024) return LSTEP_ERR_UNSUPPORTED;
  auto kern = pe_mlp_tc_kernel<MT, NTW>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) {
      set_cuda_error(e, "pe_mlp_tc attr");
      return LSTEP_ERR_CUDA;
    }
    attr_set = true;
  }
  int64_t blocks = ceil_div(n_rows, R);
  if (blocks > num_sms()) blocks = num_sms();  // persistent: one CTA per SM walks the row tiles
  kern<<<(unsigned)blocks, (kTcWarps + 1) * 32, smem, st>>>(A, lda, pe, base_ids, n_rows, n_rows_dev, *m, ldo, out, out_stride, pe_inplace);
  return check_launch("pe_mlp_tc");
}

// returns LSTEP_ERR_UNSUPPORTED when the shape does not fit (the caller then uses the SIMT kernel)
int launch_pe_mlp_tc(const float* A, int64_t lda, const float* pe, RowIds base_ids, int64_t n_rows, int64_t expected_rows,
                     const int32_t* n_rows_dev, const lstep_pe_mlp* m, float* out, int64_t out_stride, float* pe_inplace,
                     cudaStream_t st) {
  if (!m->w1_tc || !m->w2_tc || (m->ws && !m->ws_tc)) return LSTEP_ERR_UNSUPPORTED;
  const int NT = lstep_packed_ld(m->d) >> 3;
  if (NT > 3 * kTcWarps) return LSTEP_ERR_UNSUPPORTED;
  const bool two = expected_rows >= (int64_t)num_sms() * 24;  // 32-row tiles once 16-row tiles would wrap the SMs anyway
  if (NT <= kTcWarps) return two ? launch_tc<2, 1>(A, lda, pe, base_ids, n_rows, n_rows_dev, m, out, out_stride, pe_inplace, st)
                                 : launch_tc<1, 1>(A, lda, pe, base_ids, n_rows, n_rows_dev, m, out, out_stride, pe_inplace, st);
  if (NT <= 2 * kTcWarps) return two ? launch_tc<2, 2>(A, lda, pe, base_ids, n_rows, n_rows_dev, m, out, out_stride, pe_inplace, st)
                                     : launch_tc<1, 2>(A, lda, pe, base_ids, n_rows, n_rows_dev, m, out, out_stride, pe_inplace, st);
  return two ? launch_tc<2, 3>(A, lda, pe, base_ids, n_rows, n_rows_dev, m, out, out_stride, pe_inplace, st)
             : launch_tc<1, 3>(A, lda, pe, base_ids, n_rows, n_rows_dev, m, out, out_stride, pe_inplace, st);
}

}  // namespace lstep

using namespace lstep;

extern "C" size_t lstep_packed_tc_floats(int out_features, int in_features) {
  if (out_features <= 0 || in_features <= 0) return 0;
  return (size_t)(lstep_packed_rows(in_features) >> 3) * (lstep_packed_ld(out_features) >> 3) * 32 * 4;
}

extern "C" int lstep_pack_linear_tc(const float* weight, int out_features, int in_features, float* packed, void* stream) {
  if (!weight || !packed || out_features <= 0 || in_features <= 0) return LSTEP_ERR_INVALID_ARG;
  pack_linear_tc_kernel<<<64, 256, 0, as_stream(stream)>>>(weight, out_features, in_features, lstep_packed_rows(in_features),
                                                           lstep_packed_ld(out_features), reinterpret_cast<float4*>(packed));
  return check_launch("pack_linear_tc");
}
