// Finite scalar quantisation: tanh bound, round (half-to-even), straight-through, mixed-radix index
// pack — one pass, HBM-bound.  Rows are D<=16 scalars (5 or 6 in the shipped configs), so a block
// stages a contiguous [256 tokens x D] slab through shared memory to keep global traffic coalesced.
#include <math.h>

#include "common.cuh"

namespace vqb {

constexpr int kFsqTokens = 256;  // tokens per block iteration == threads per block

template <typename TX, typename TO>
__global__ void __launch_bounds__(kFsqTokens) fsq_forward_kernel(const TX* __restrict__ x, int64_t N,
                                                                 const vqb_fsq_params p, TO* __restrict__ zq,
                                                                 int32_t* __restrict__ index) {
  __shared__ float slab[kFsqTokens * 16];
  const int D = p.D;
  for (int64_t base = (int64_t)blockIdx.x * kFsqTokens; base < N; base += (int64_t)gridDim.x * kFsqTokens) {
    const int64_t ntok = min((int64_t)kFsqTokens, N - base);
    const int nelem = (int)ntok * D;
    for (int i = threadIdx.x; i < nelem; i += kFsqTokens) slab[i] = to_f32<TX>(x[base * D + i]);
    __syncthreads();
    if (threadIdx.x < ntok) {
      int code = 0;
#pragma unroll 1
      for (int d = 0; d < D; ++d) {
        const float v = slab[threadIdx.x * D + d];
        // z = tanh(x + atanh(odd/max_)) * max_ - odd ; z /= 2      fsq/quantizers.py:118-119
        float z = __fsub_rn(__fmul_rn(tanhf(__fadd_rn(v, p.shift[d])), p.max_[d]), p.odd[d]);
        z = z * 0.5f;
        const float r = rintf(z);  // torch.round: half to even; ste value z + (r - z) == r exactly
        slab[threadIdx.x * D + d] = __fdiv_rn(r, p.half[d]);                     // :123
        code += ((int)r + (int)p.half[d]) * p.cumprod[d];                        // :124-125, :65-68
      }
      if (index) index[base + threadIdx.x] = code;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nelem; i += kFsqTokens) zq[base * D + i] = from_f32<TO>(slab[i]);
    __syncthreads();
  }
}

template <typename TG, typename TX>
__global__ void __launch_bounds__(256) fsq_backward_kernel(const TG* __restrict__ gz, const TX* __restrict__ x,
                                                           int64_t total, const vqb_fsq_params p,
                                                           TX* __restrict__ gx) {
  const int D = p.D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const float t = tanhf(__fadd_rn(to_f32<TX>(x[i]), p.shift[d]));
    // d zq / d x = max_/(2*half) * (1 - tanh^2)   (round is straight-through)
    const float g = to_f32<TG>(gz[i]) / p.half[d] * 0.5f * p.max_[d] * (1.f - t * t);
    gx[i] = from_f32<TX>(g);
  }
}

__global__ void fsq_decode_kernel(const int32_t* __restrict__ index, int64_t N, const vqb_fsq_params p,
                                  float* __restrict__ z) {
  const int D = p.D;
  const int64_t total = N * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / D;
    const int d = (int)(i - n * D);
    const int digit = (index[n] / p.cumprod[d]) % p.levels[d];          // fsq/quantizers.py:59-63
    z[i] = __fsub_rn(__fdiv_rn((float)digit, p.half[d]), 1.f);           // :137
  }
}

static int check_params(const vqb_fsq_params* p) {
  VQB_REQUIRE(p, "fsq: null params");
  VQB_REQUIRE(p->D >= 1 && p->D <= 16, "fsq: D must be in 1..16 (got %d)", p->D);
  return VQB_OK;
}

}  // namespace vqb

using namespace vqb;

extern "C" {

int vqb_fsq_forward(const void* x, int x_dtype, int64_t N, const vqb_fsq_params* p, void* zq, int out_dtype,
                    int32_t* index, void* stream) {
  if (int e = check_params(p)) return e;
  VQB_REQUIRE(x && zq, "vqb_fsq_forward: null pointer");
  if (N <= 0) return VQB_OK;
  int64_t blocks64 = (N + kFsqTokens - 1) / kFsqTokens;
  const int blocks = (int)(blocks64 < (int64_t)sm_count() * 8 ? blocks64 : (int64_t)sm_count() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == VQB_F32 && out_dtype == VQB_F32)
    fsq_forward_kernel<float, float><<<blocks, kFsqTokens, 0, st>>>((const float*)x, N, *p, (float*)zq, index);
  else if (x_dtype == VQB_BF16 && out_dtype == VQB_BF16)
    fsq_forward_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, kFsqTokens, 0, st>>>((const __nv_bfloat16*)x, N, *p, (__nv_bfloat16*)zq, index);
  else if (x_dtype == VQB_BF16 && out_dtype == VQB_F32)
    fsq_forward_kernel<__nv_bfloat16, float><<<blocks, kFsqTokens, 0, st>>>((const __nv_bfloat16*)x, N, *p, (float*)zq, index);
  else
    VQB_REQUIRE(false, "vqb_fsq_forward: unsupported dtype combination");
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_fsq_backward(const void* gz, int g_dtype, const void* x, int x_dtype, int64_t N, const vqb_fsq_params* p,
                     void* gx, int gx_dtype, void* stream) {
  if (int e = check_params(p)) return e;
  VQB_REQUIRE(gz && x && gx, "vqb_fsq_backward: null pointer");
  VQB_REQUIRE(gx_dtype == x_dtype, "vqb_fsq_backward: gx dtype must equal x dtype");
  if (N <= 0) return VQB_OK;
  const int64_t total = N * p->D;
  int64_t blocks64 = (total + 255) / 256;
  const int blocks = (int)(blocks64 < (int64_t)sm_count() * 8 ? blocks64 : (int64_t)sm_count() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (g_dtype == VQB_F32 && x_dtype == VQB_F32)
    fsq_backward_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)gz, (const float*)x, total, *p, (float*)gx);
  else if (g_dtype == VQB_BF16 && x_dtype == VQB_BF16)
    fsq_backward_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)gz, (const __nv_bfloat16*)x, total, *p, (__nv_bfloat16*)gx);
  else if (g_dtype == VQB_F32 && x_dtype == VQB_BF16)
    fsq_backward_kernel<float, __nv_bfloat16><<<blocks, 256, 0, st>>>((const float*)gz, (const __nv_bfloat16*)x, total, *p, (__nv_bfloat16*)gx);
  else
    VQB_REQUIRE(false, "vqb_fsq_backward: unsupported dtype combination");
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_fsq_decode(const int32_t* index, int64_t N, const vqb_fsq_params* p, float* z, void* stream) {
  if (int e = check_params(p)) return e;
  VQB_REQUIRE(index && z, "vqb_fsq_decode: null pointer");
  if (N <= 0) return VQB_OK;
  const int64_t total = N * p->D;
  int64_t blocks64 = (total + 255) / 256;
  const int blocks = (int)(blocks64 < (int64_t)sm_count() * 8 ? blocks64 : (int64_t)sm_count() * 8);
  fsq_decode_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(index, N, *p, z);
  VQB_LAUNCH_OK();
  return VQB_OK;
}

}  // extern "C"
