// Compatibility mode for consumers of the materialised [N x K] distance matrix.
// The hot path never builds this matrix; `memo['encode']['distance']` of the reference
// (vq/algorithms/vq/quantizers.py:97-98) is produced ON DEMAND by this kernel for the components that read it:
// EntropyLoss (vq/algorithms/vq/losses.py:130-153), MultinomialAnchor (vq/algorithms/cvqvae/anchors.py:88-104),
// user callbacks, and the `materialize_distance` debug switch of the quantizer modules.
// Plain fp32 CUDA-core tile kernel (64 x 64 outputs per block, 4 x 4 per thread): exact fp32 products, no tensor
// cores, no operand planes — a debugging / compatibility aid, not a throughput path.
#include <math.h>

#include "common.cuh"

namespace vqb {

constexpr int kDT = 64;   // tile edge
constexpr int kDK = 16;   // contraction slab

template <typename TX>
__global__ void __launch_bounds__(256) distance_matrix_kernel(const TX* __restrict__ x, int64_t N, int D,
                                                              const float* __restrict__ W, int64_t K, int cosine,
                                                              float* __restrict__ out) {
  __shared__ float xs[kDK][kDT + 1], ws[kDK][kDT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t row0 = (int64_t)blockIdx.y * kDT, col0 = (int64_t)blockIdx.x * kDT;
  float acc[4][4] = {}, xx[4] = {}, ee[4] = {};
  for (int d0 = 0; d0 < D; d0 += kDK) {
    // 64 rows x 16 contraction elements per operand: 1024 elements, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + i * 256;
      const int r = e / kDK, dd = e % kDK;
      const int64_t gr = row0 + r, gc = col0 + r;
      xs[dd][r] = (gr < N && d0 + dd < D) ? to_f32<TX>(x[gr * D + d0 + dd]) : 0.f;
      ws[dd][r] = (gc < K && d0 + dd < D) ? W[gc * D + d0 + dd] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int dd = 0; dd < kDK; ++dd) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = xs[dd][ty * 4 + i];
        b[i] = ws[dd][tx * 4 + i];
        xx[i] = fmaf(a[i], a[i], xx[i]);
        ee[i] = fmaf(b[i], b[i], ee[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = row0 + ty * 4 + i;
    if (r >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = col0 + tx * 4 + j;
      if (c >= K) continue;
      float v;
      if (cosine) {   // 1 - <x/|x|, e/|e|>, F.normalize eps                      distances.py:41-45
        v = 1.f - acc[i][j] / (fmaxf(sqrtf(xx[i]), kNormEps) * fmaxf(sqrtf(ee[j]), kNormEps));
      } else {        // torch.cdist (mm path): sqrt(clamp_min(|x|^2 - 2 x.e + |e|^2, 0))   distances.py:32
        v = sqrtf(fmaxf(xx[i] - 2.f * acc[i][j] + ee[j], 0.f));
      }
      out[r * K + c] = v;
    }
  }
}

}  // namespace vqb

using namespace vqb;

extern "C" int vqb_distance_matrix(const void* x, int x_dtype, int64_t N, int D, const float* W, int64_t K, int cosine,
                                   float* out, void* stream) {
  VQB_REQUIRE(x && W && out, "vqb_distance_matrix: null pointer");
  VQB_REQUIRE(N >= 1 && K >= 1 && D >= 1, "vqb_distance_matrix: bad shape");
  const dim3 grid((unsigned)((K + kDT - 1) / kDT), (unsigned)((N + kDT - 1) / kDT));
  VQB_REQUIRE(grid.y <= 65535, "vqb_distance_matrix: N <= 4 194 240 rows per call (materialise in row chunks)");
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == VQB_F32)
    distance_matrix_kernel<float><<<grid, 256, 0, st>>>((const float*)x, N, D, W, K, cosine, out);
  else if (x_dtype == VQB_BF16)
    distance_matrix_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, N, D, W, K, cosine, out);
  else
    VQB_REQUIRE(false, "vqb_distance_matrix: bad dtype");
  VQB_LAUNCH_OK();
  return VQB_OK;
}
