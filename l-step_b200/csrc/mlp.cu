// PE MLP:  out = base + tanh( [Ws base + bs] + W2 relu(W1 A + b1) + b2 )
// The arithmetic of models/LSTEP.py:240-247 (neighbourhood), :294-301 (update phase A, with the
// self term) and :329-336 (update phase B, where the self term is computed and discarded — Q3).
//
// fp32 FMA throughout: the parity bar (1e-5) rules out single-pass TF32/BF16 tensor-core math.
// Weights are pre-transposed to [in][ldo] (lstep_pack_linear) so that thread c reads column c
// with unit stride across the warp; they total 0.85 MB and stay L2 resident. One CTA owns R
// rows and all output columns: the R x (d+t) aggregate tile and the R x d base tile sit in shared
// memory (k-major, so one 128-bit broadcast load feeds 4 rows), each thread keeps R
// accumulators for its column and streams its weight column through registers with 8 loads in
// flight. R is picked per launch so the grid still covers the 148 SMs.
#include "common.cuh"

namespace lstep {

__global__ void pack_linear_kernel(const float* __restrict__ w, const float* __restrict__ b, int out_f, int in_f,
                                   int ldo, float* __restrict__ pw, float* __restrict__ pb) {
  const int64_t total = (int64_t)in_f * ldo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / ldo), c = (int)(i % ldo);
    pw[i] = c < out_f ? w[(size_t)c * in_f + k] : 0.f;
  }
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < ldo; c += blockDim.x) pb[c] = (b && c < out_f) ? b[c] : 0.f;
}

// dynamic smem: As[in1][R] | Bs[d][R] | Hs[d][R]
template <int R>
__global__ void __launch_bounds__(256) pe_mlp_kernel(const float* __restrict__ A, const float* pe,
                                                     const int64_t* __restrict__ base_ids, int64_t n_rows,
                                                     const int32_t* __restrict__ n_rows_dev, lstep_pe_mlp m, int ldo,
                                                     float* __restrict__ out, int64_t out_stride,
                                                     float* pe_inplace) {
  static_assert(R % 4 == 0, "R must be a multiple of 4");
  extern __shared__ __align__(16) float smem[];
  const int d = m.d, in1 = m.d + m.t;
  if (n_rows_dev) {
    const int64_t nd = *n_rows_dev;
    n_rows = nd < n_rows ? nd : n_rows;
  }
  const int64_t row0 = (int64_t)blockIdx.x * R;
  if (row0 >= n_rows) return;
  float* As = smem;
  float* Bs = As + (size_t)in1 * R;
  float* Hs = Bs + (size_t)d * R;
  const int c = threadIdx.x;
  const int nthr = blockDim.x;

  // ---- stage A tile and base rows, k-major
  for (int r = 0; r < R; ++r) {
    const int64_t row = row0 + r;
    const bool ok = row < n_rows;
    const float* arow = A + row * (int64_t)in1;
    for (int k = c; k < in1; k += nthr) As[k * R + r] = ok ? arow[k] : 0.f;
    const float* brow = ok ? pe + base_ids[row] * (int64_t)d : nullptr;
    for (int k = c; k < d; k += nthr) Bs[k * R + r] = ok ? brow[k] : 0.f;
  }
  __syncthreads();

  float acc[R];
  // ---- layer 1: h = relu(W1 a + b1)
  if (c < ldo) {
    const float b1 = m.b1[c];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = b1;
    const float* wcol = m.w1 + c;
    int k = 0;
    for (; k + 8 <= in1; k += 8) {
      float w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = __ldg(wcol + (size_t)(k + u) * ldo);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4* a4 = reinterpret_cast<const float4*>(As + (k + u) * R);
#pragma unroll
        for (int q = 0; q < R / 4; ++q) {
          const float4 a = a4[q];
          acc[4 * q + 0] = fmaf(a.x, w[u], acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(a.y, w[u], acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(a.z, w[u], acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(a.w, w[u], acc[4 * q + 3]);
        }
      }
    }
    for (; k < in1; ++k) {
      const float w = __ldg(wcol + (size_t)k * ldo);
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fmaf(As[k * R + r], w, acc[r]);
    }
    if (c < d) {
#pragma unroll
      for (int r = 0; r < R; ++r) Hs[c * R + r] = fmaxf(acc[r], 0.f);
    }
  }
  __syncthreads();

  // ---- layer 2 (+ self term): z = W2 h + b2 [+ Ws base + bs]; out = base + tanh(z)
  if (c < d) {
    const float b2 = m.b2[c];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = b2;
    {
      const float* wcol = m.w2 + c;
      int k = 0;
      for (; k + 8 <= d; k += 8) {
        float w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = __ldg(wcol + (size_t)(k + u) * ldo);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4* h4 = reinterpret_cast<const float4*>(Hs + (k + u) * R);
#pragma unroll
          for (int q = 0; q < R / 4; ++q) {
            const float4 a = h4[q];
            acc[4 * q + 0] = fmaf(a.x, w[u], acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(a.y, w[u], acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(a.z, w[u], acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(a.w, w[u], acc[4 * q + 3]);
          }
        }
      }
      for (; k < d; ++k) {
        const float w = __ldg(wcol + (size_t)k * ldo);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fmaf(Hs[k * R + r], w, acc[r]);
      }
    }
    if (m.ws) {
      // the reference adds the two Linear outputs: (Ws base + bs) + (W2 h + b2)
      float sacc[R];
      const float bs = m.bs[c];
#pragma unroll
      for (int r = 0; r < R; ++r) sacc[r] = bs;
      const float* wcol = m.ws + c;
      int k = 0;
      for (; k + 8 <= d; k += 8) {
        float w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = __ldg(wcol + (size_t)(k + u) * ldo);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4* b4 = reinterpret_cast<const float4*>(Bs + (k + u) * R);
#pragma unroll
          for (int q = 0; q < R / 4; ++q) {
            const float4 a = b4[q];
            sacc[4 * q + 0] = fmaf(a.x, w[u], sacc[4 * q + 0]);
            sacc[4 * q + 1] = fmaf(a.y, w[u], sacc[4 * q + 1]);
            sacc[4 * q + 2] = fmaf(a.z, w[u], sacc[4 * q + 2]);
            sacc[4 * q + 3] = fmaf(a.w, w[u], sacc[4 * q + 3]);
          }
        }
      }
      for (; k < d; ++k) {
        const float w = __ldg(wcol + (size_t)k * ldo);
#pragma unroll
        for (int r = 0; r < R; ++r) sacc[r] = fmaf(Bs[k * R + r], w, sacc[r]);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] += sacc[r];
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = row0 + r;
      if (row < n_rows) {
        const float o = Bs[c * R + r] + tanhf(acc[r]);
        if (out)
          out[row * out_stride + c] = o;
        else
          pe_inplace[base_ids[row] * (int64_t)d + c] = o;
      }
    }
  }
}

template <int R>
static int launch_mlp_r(const float* A, const float* pe, const int64_t* base_ids, int64_t n_rows,
                        const int32_t* n_rows_dev, const lstep_pe_mlp* m, float* out, int64_t out_stride,
                        float* pe_inplace, cudaStream_t st) {
  const int ldo = lstep_packed_ld(m->d);
  const size_t smem = sizeof(float) * (size_t)R * ((size_t)m->d + m->t + 2 * (size_t)m->d);
  if (smem > 200 * 1024 || ldo > 256) return LSTEP_ERR_UNSUPPORTED;
  auto kern = pe_mlp_kernel<R>;
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      set_cuda_error(e, "pe_mlp attr");
      return LSTEP_ERR_CUDA;
    }
    attr_set = true;
  }
  const int64_t blocks = ceil_div(n_rows, R);
  kern<<<(unsigned)blocks, ldo, smem, st>>>(A, pe, base_ids, n_rows, n_rows_dev, *m, ldo, out, out_stride, pe_inplace);
  return check_launch("pe_mlp");
}

// n_rows is the host-side upper bound of rows; *n_rows_dev (optional) the device-side count.
int launch_pe_mlp(const float* A, const float* pe, const int64_t* base_ids, int64_t n_rows, int64_t expected_rows,
                  const int32_t* n_rows_dev, const lstep_pe_mlp* m, float* out, int64_t out_stride, float* pe_inplace,
                  cudaStream_t st) {
  if (n_rows <= 0) return LSTEP_OK;
  if (!A || !pe || !base_ids || !m || (!out && !pe_inplace)) return LSTEP_ERR_INVALID_ARG;
  // rows per CTA: keep >= ~1 CTA per SM when the problem allows it
  if (expected_rows >= (int64_t)kNumSMs * 16) return launch_mlp_r<16>(A, pe, base_ids, n_rows, n_rows_dev, m, out, out_stride, pe_inplace, st);
  if (expected_rows >= (int64_t)kNumSMs * 6) return launch_mlp_r<8>(A, pe, base_ids, n_rows, n_rows_dev, m, out, out_stride, pe_inplace, st);
  return launch_mlp_r<4>(A, pe, base_ids, n_rows, n_rows_dev, m, out, out_stride, pe_inplace, st);
}

}  // namespace lstep

using namespace lstep;

extern "C" int lstep_packed_ld(int out_features) { return (int)align_up((size_t)out_features, 32); }

extern "C" int lstep_pack_linear(const float* weight, const float* bias, int out_features, int in_features,
                                 float* packed_w, float* packed_b, void* stream) {
  if (!weight || !packed_w || !packed_b || out_features <= 0 || in_features <= 0) return LSTEP_ERR_INVALID_ARG;
  const int ldo = lstep_packed_ld(out_features);
  pack_linear_kernel<<<64, 256, 0, as_stream(stream)>>>(weight, bias, out_features, in_features, ldo, packed_w,
                                                        packed_b);
  return check_launch("pack_linear");
}

extern "C" int lstep_pe_mlp_apply(const float* A, const float* pe, const int64_t* base_ids, int64_t n_rows,
                                  const lstep_pe_mlp* mlp, float* out, int64_t out_stride, float* pe_inplace,
                                  void* stream) {
  if (n_rows < 0) return LSTEP_ERR_INVALID_ARG;
  return launch_pe_mlp(A, pe, base_ids, n_rows, n_rows, nullptr, mlp, out, out_stride, pe_inplace, as_stream(stream));
}
