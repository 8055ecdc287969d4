// Finite scalar quantisation: tanh bound, round (half-to-even), straight-through, mixed-radix index
// pack — one pass, HBM-bound.  Rows are D<=16 scalars (5 or 6 in the shipped configs), so a block
// stages a contiguous [256 tokens x D] slab through shared memory to keep global traffic coalesced.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace vqb {

constexpr int kFsqTokens = 256;  // tokens per block iteration == threads per block

// tanh with ABSOLUTE error <= ~5e-7 from two MUFU ops: 1 - 2 / (1 + exp(2v)).  The quantizer rounds
// z = (tanh(.) * max_ - odd) / 2 to an integer, so only the absolute error matters: |dz| <= max_/2 * 5e-7 < 2e-6,
// inside the documented 1e-5 round-half boundary band (libdevice tanhf costs ~4x more instructions and made
// the kernel issue-bound at 30 % of HBM bandwidth).
__device__ __forceinline__ float fsq_tanh(float v) {
  const float e = __expf(2.f * v);              // +inf for large v -> 1 ; 0 for very negative v -> -1
  return 1.f - __fdividef(2.f, e + 1.f);
}
// r / half with the exact fp32 quotient: half is a small integer, a power of two for the usual level counts
__device__ __forceinline__ float fsq_div(float r, float half) {
  const int h = (int)half;
  return (h & (h - 1)) == 0 ? r * (1.f / half) : __fdiv_rn(r, half);
}

// One thread per token, D known at compile time: the D independent tanhf chains of a token are fully unrolled
// (ILP) with every per-channel constant in registers.  A block stages a contiguous [256 tokens x D] slab through
// shared memory so that global loads and stores stay coalesced although a token row is only 20-24 bytes.
template <typename TX, typename TO, int DT>
__global__ void __launch_bounds__(kFsqTokens) fsq_forward_kernel(const TX* __restrict__ x, int64_t N,
                                                                 const vqb_fsq_params p, TO* __restrict__ zq,
                                                                 int32_t* __restrict__ index) {
  __shared__ __align__(16) float slab[kFsqTokens * (DT > 0 ? DT : 16)];
  const int D = DT > 0 ? DT : p.D;
  for (int64_t base = (int64_t)blockIdx.x * kFsqTokens; base < N; base += (int64_t)gridDim.x * kFsqTokens) {
    const int64_t ntok = min((int64_t)kFsqTokens, N - base);
    const int nelem = (int)ntok * D;
    // 128-bit staging copies for full slabs (256*D*sizeof is a multiple of 16 and the slab base is 16 B aligned)
    constexpr int VEC = 16 / sizeof(TX);
    const bool vec_in = ntok == kFsqTokens && ((uintptr_t)x % 16 == 0);
    if (vec_in) {
      const uint4* src = reinterpret_cast<const uint4*>(x + base * D);
      for (int i = threadIdx.x; i < nelem / VEC; i += kFsqTokens) {
        const uint4 raw = src[i];
        if constexpr (sizeof(TX) == 4) {
          *reinterpret_cast<uint4*>(slab + i * 4) = raw;
        } else {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h[j]);
            slab[i * 8 + 2 * j] = f.x;
            slab[i * 8 + 2 * j + 1] = f.y;
          }
        }
      }
    } else {
      for (int i = threadIdx.x; i < nelem; i += kFsqTokens) slab[i] = to_f32<TX>(x[base * D + i]);
    }
    __syncthreads();
    if (threadIdx.x < ntok) {
      int code = 0;
      float out[DT > 0 ? DT : 16];
#pragma unroll
      for (int d = 0; d < (DT > 0 ? DT : 16); ++d) {
        if (DT > 0 || d < D) {
          const float v = slab[threadIdx.x * D + d];
          // z = tanh(x + atanh(odd/max_)) * max_ - odd ; z /= 2      fsq/quantizers.py:118-119
          float z = __fsub_rn(__fmul_rn(fsq_tanh(__fadd_rn(v, p.shift[d])), p.max_[d]), p.odd[d]);
          z = z * 0.5f;
          const float r = rintf(z);  // torch.round: half to even; ste value z + (r - z) == r exactly
          out[d] = fsq_div(r, p.half[d]);                                      // :123
          code += ((int)r + (int)p.half[d]) * p.cumprod[d];                      // :124-125, :65-68
        }
      }
#pragma unroll
      for (int d = 0; d < (DT > 0 ? DT : 16); ++d)
        if (DT > 0 || d < D) slab[threadIdx.x * D + d] = out[d];
      if (index) index[base + threadIdx.x] = code;
    }
    __syncthreads();
    constexpr int VECO = 16 / sizeof(TO);
    if (ntok == kFsqTokens && ((uintptr_t)zq % 16 == 0)) {
      uint4* dst = reinterpret_cast<uint4*>(zq + base * D);
      for (int i = threadIdx.x; i < nelem / VECO; i += kFsqTokens) {
        if constexpr (sizeof(TO) == 4) {
          dst[i] = *reinterpret_cast<const uint4*>(slab + i * 4);
        } else {
          uint4 raw;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(slab[i * 8 + 2 * j], slab[i * 8 + 2 * j + 1]);
          dst[i] = raw;
        }
      }
    } else {
      for (int i = threadIdx.x; i < nelem; i += kFsqTokens) zq[base * D + i] = from_f32<TO>(slab[i]);
    }
    __syncthreads();
  }
}

// Token-per-thread variant WITHOUT the shared-memory slab: a thread reads its own 10-32 byte token with the widest
// loads the token size allows (the neighbouring lanes' pieces of the same 32-byte sectors are served by L1), so there
// are no block barriers and every thread keeps several independent loads in flight.
template <typename T, int DT>
__device__ __forceinline__ void fsq_load_token(const T* __restrict__ p, float (&v)[DT]) {
  constexpr int BYTES = DT * (int)sizeof(T);
  if constexpr (BYTES % 8 == 0) {
    T tmp[DT];
#pragma unroll
    for (int i = 0; i < BYTES / 8; ++i) reinterpret_cast<uint2*>(tmp)[i] = reinterpret_cast<const uint2*>(p)[i];
#pragma unroll
    for (int d = 0; d < DT; ++d) v[d] = to_f32<T>(tmp[d]);
  } else if constexpr (BYTES % 4 == 0) {
    T tmp[DT];
#pragma unroll
    for (int i = 0; i < BYTES / 4; ++i) reinterpret_cast<uint32_t*>(tmp)[i] = reinterpret_cast<const uint32_t*>(p)[i];
#pragma unroll
    for (int d = 0; d < DT; ++d) v[d] = to_f32<T>(tmp[d]);
  } else {
#pragma unroll
    for (int d = 0; d < DT; ++d) v[d] = to_f32<T>(p[d]);
  }
}
template <typename T, int DT>
__device__ __forceinline__ void fsq_store_token(T* __restrict__ p, const float (&v)[DT]) {
  constexpr int BYTES = DT * (int)sizeof(T);
  T tmp[DT];
#pragma unroll
  for (int d = 0; d < DT; ++d) tmp[d] = from_f32<T>(v[d]);
  if constexpr (BYTES % 8 == 0) {
#pragma unroll
    for (int i = 0; i < BYTES / 8; ++i) reinterpret_cast<uint2*>(p)[i] = reinterpret_cast<const uint2*>(tmp)[i];
  } else if constexpr (BYTES % 4 == 0) {
#pragma unroll
    for (int i = 0; i < BYTES / 4; ++i) reinterpret_cast<uint32_t*>(p)[i] = reinterpret_cast<const uint32_t*>(tmp)[i];
  } else {
#pragma unroll
    for (int d = 0; d < DT; ++d) p[d] = tmp[d];
  }
}

template <typename TX, typename TO, int DT>
__global__ void __launch_bounds__(256) fsq_forward_direct_kernel(const TX* __restrict__ x, int64_t N, const vqb_fsq_params p,
                                                                 TO* __restrict__ zq, int32_t* __restrict__ index) {
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    float v[DT], out[DT];
    fsq_load_token<TX, DT>(x + n * DT, v);
    int code = 0;
#pragma unroll
    for (int d = 0; d < DT; ++d) {
      float z = __fsub_rn(__fmul_rn(fsq_tanh(__fadd_rn(v[d], p.shift[d])), p.max_[d]), p.odd[d]);
      z = z * 0.5f;
      const float r = rintf(z);
      out[d] = fsq_div(r, p.half[d]);
      code += ((int)r + (int)p.half[d]) * p.cumprod[d];
    }
    fsq_store_token<TO, DT>(zq + n * DT, out);
    if (index) index[n] = code;
  }
}

template <typename TG, typename TX>
__global__ void __launch_bounds__(256) fsq_backward_kernel(const TG* __restrict__ gz, const TX* __restrict__ x,
                                                           int64_t total, const vqb_fsq_params p,
                                                           TX* __restrict__ gx) {
  const int D = p.D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const float t = fsq_tanh(__fadd_rn(to_f32<TX>(x[i]), p.shift[d]));
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
  static int direct = -1;     // VQB_FSQ_DIRECT=0: the shared-memory slab kernel (developer A/B switch)
  if (direct < 0) {
    const char* e = getenv("VQB_FSQ_DIRECT");
    direct = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
#define VQB_FSQ_SLAB(TX, TO, DT)                                                                            \
  fsq_forward_kernel<TX, TO, DT><<<blocks, kFsqTokens, 0, st>>>((const TX*)x, N, *p, (TO*)zq, index)
#define VQB_FSQ_LAUNCH(TX, TO, DT)                                                                          \
  do {                                                                                                      \
    if (direct)                                                                                             \
      fsq_forward_direct_kernel<TX, TO, DT><<<blocks, 256, 0, st>>>((const TX*)x, N, *p, (TO*)zq, index);   \
    else                                                                                                    \
      VQB_FSQ_SLAB(TX, TO, DT);                                                                             \
  } while (0)
#define VQB_FSQ_D(TX, TO)                                        \
  switch (p->D) {                                                \
    case 3: VQB_FSQ_LAUNCH(TX, TO, 3); break;                    \
    case 4: VQB_FSQ_LAUNCH(TX, TO, 4); break;                    \
    case 5: VQB_FSQ_LAUNCH(TX, TO, 5); break;                    \
    case 6: VQB_FSQ_LAUNCH(TX, TO, 6); break;                    \
    case 7: VQB_FSQ_LAUNCH(TX, TO, 7); break;                    \
    case 8: VQB_FSQ_LAUNCH(TX, TO, 8); break;                    \
    default: VQB_FSQ_SLAB(TX, TO, 0); break;                     \
  }
  if (x_dtype == VQB_F32 && out_dtype == VQB_F32) { VQB_FSQ_D(float, float) }
  else if (x_dtype == VQB_BF16 && out_dtype == VQB_BF16) { VQB_FSQ_D(__nv_bfloat16, __nv_bfloat16) }
  else if (x_dtype == VQB_BF16 && out_dtype == VQB_F32) { VQB_FSQ_D(__nv_bfloat16, float) }
  else { VQB_REQUIRE(false, "vqb_fsq_forward: unsupported dtype combination"); }
#undef VQB_FSQ_D
#undef VQB_FSQ_LAUNCH
#undef VQB_FSQ_SLAB
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
