// Shared helpers for the vqb200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vqb200.h"

namespace vqb {

void set_error(const char* fmt, ...);

#define VQB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::vqb::set_error(__VA_ARGS__);      \
      return VQB_ERR_ARG;                 \
    }                                     \
  } while (0)

#define VQB_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::vqb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VQB_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define VQB_LAUNCH_OK() VQB_CUDA_OK(cudaGetLastError())

constexpr float kNormEps = 1e-12f;  // F.normalize default eps

// plane formats (include/vqb200.h): 1..3 bf16 planes, one fp16 plane, or the fp16 (hi, lo * 2^11) pair
constexpr int kPairShift = 11;       // lo' = (v - hi) * 2^11; vqb_assign folds 2^-11 back with scale-input-d
__host__ __device__ inline bool is_f16(int planes) { return (planes & 0x10) != 0; }       // fp16 family
__host__ __device__ inline bool is_f16x2(int planes) { return planes == VQB_PLANES_F16X2; }
__host__ __device__ inline int plane_count(int planes) { return VQB_PLANE_COUNT(planes); }
__host__ __device__ inline bool planes_valid(int planes) {
  return (planes >= 1 && planes <= 3) || planes == VQB_PLANES_F16 || planes == VQB_PLANES_F16X2;
}

__host__ __device__ inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ---- dtype-generic row element access -------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- packed (score, index) keys ----------------------------------------------------------
// orderable(): monotone map fp32 -> uint32 (larger float => larger uint).
__host__ __device__ __forceinline__ uint32_t orderable(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float from_orderable(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
// Smaller key == better (higher score, then lower index).
__host__ __device__ __forceinline__ unsigned long long make_key(float score, uint32_t index) {
  return (static_cast<unsigned long long>(~orderable(score)) << 32) | index;
}
// Initial value of a key slot ("no candidate yet").  A row whose scores are all NaN keeps it.
constexpr unsigned long long kNoKey = ~0ull;
__host__ __device__ __forceinline__ uint32_t key_index(unsigned long long k) { return static_cast<uint32_t>(k); }
__host__ __device__ __forceinline__ float key_score(unsigned long long k) {
  return from_orderable(~static_cast<uint32_t>(k >> 32));
}

// ---- warp helpers ------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int W>
__device__ __forceinline__ float group_sum(float v) {  // sum over aligned groups of W lanes
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int W>
__device__ __forceinline__ float group_max(float v) {  // max over aligned groups of W lanes
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// One-plane fp16 rows keep |v| < 2^15: a row whose largest component is >= 2^15 is scaled by an exact power of two
// (the row arg-max of <row, b_j> does not change; such rows do not occur in latent spaces)
__device__ __forceinline__ float f16_row_scale(float row_absmax) {
  int e;
  frexpf(row_absmax, &e);  // row_absmax = m * 2^e, 0.5 <= m < 1
  return (e > 15 && row_absmax < INFINITY) ? ldexpf(1.f, 15 - e) : 1.f;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// The hot step is a chain of short kernels (pack, pack, assign, gather, backward); with PDL the next grid is
// scheduled while the previous one drains, so the ~2 us launch + ramp of each link overlaps its predecessor.
// Every PDL-launched kernel calls pdl_wait() before touching global memory (it returns once the preceding
// grid has completed and its writes are visible; a no-op for a normal launch).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// SM count of the CURRENT device (cached per device ordinal: one process may drive several GPUs)
inline int sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  const bool cached = dev >= 0 && dev < 64;
  if (cached && cache[dev] > 0) return cache[dev];
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (cached) cache[dev] = n;
  return n;
}

}  // namespace vqb
