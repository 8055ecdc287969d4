// Operand packing (fp32/bf16 rows -> exact bf16 planes for the tcgen05 contraction) and the
// row l2-normalisation forward/backward.  All HBM-bound, one pass, sub-warp group per row.
#include <math.h>
#include <stdarg.h>

#include "common.cuh"

namespace vqb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// lanes per row: each lane handles ~4 elements, rows of a warp stay sector-aligned.
static inline int lanes_per_row(int D) {
  int g = 1;
  while (g < 32 && g * 4 < D) g <<= 1;
  return g;
}

template <typename T, int G>
__global__ void __launch_bounds__(256) pack_rows_kernel(
    const T* __restrict__ src, int64_t rows, int64_t rows_pad, int D, int Dp, int normalize, int planes,
    __nv_bfloat16* __restrict__ dst, float* __restrict__ half_sqnorm, float* __restrict__ writeback,
    unsigned long long* __restrict__ keys, int64_t n_keys, uint4* __restrict__ zero_fill, int64_t n_zero16,
    float* __restrict__ lo_norm_max) {
  pdl_wait();               // PDL: the source rows / key buffer may still be in use by the preceding launch
  pdl_launch_dependents();
  // fused memset of the assignment keys
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_keys;
       i += (int64_t)gridDim.x * blockDim.x)
    keys[i] = ~0ull;
  // fused zero-fill of the step's statistics buffer
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_zero16; i += (int64_t)gridDim.x * blockDim.x)
    zero_fill[i] = make_uint4(0u, 0u, 0u, 0u);

  const int lane = threadIdx.x % G;
  const int64_t rows_per_block = blockDim.x / G;
  for (int64_t r = blockIdx.x * rows_per_block + threadIdx.x / G; r < rows_pad;
       r += (int64_t)gridDim.x * rows_per_block) {
    const bool real = r < rows;
    float ss = 0.f, amax = 0.f;
    if (real)
      for (int d = lane; d < D; d += G) {
        float v = to_f32<T>(src[r * D + d]);
        ss = fmaf(v, v, ss);
        amax = fmaxf(amax, fabsf(v));
      }
    ss = group_sum<G>(ss);
    const float denom = normalize ? fmaxf(sqrtf(ss), kNormEps) : 1.f;
    const float f16_scale = planes == VQB_PLANES_F16 ? f16_row_scale(group_max<G>(amax)) : 1.f;
    float ss2 = 0.f, lo2 = 0.f;
    for (int d = lane; d < Dp; d += G) {
      float v = 0.f;
      if (real && d < D) {
        v = to_f32<T>(src[r * D + d]);
        if (normalize) v = __fdiv_rn(v, denom);
        if (writeback) writeback[r * D + d] = v;
      }
      ss2 = fmaf(v, v, ss2);
      if (planes == VQB_PLANES_F16) {
        dst[r * Dp + d] = __ushort_as_bfloat16(__half_as_ushort(__float2half_rn(v * f16_scale)));
      } else if (is_f16x2(planes)) {
        // fp16 pair: hi = fp16(v), lo' = fp16((v - hi) * 2^11); the subtraction is exact
        const __half hi = __float2half_rn(v);
        lo2 = fmaf(v - __half2float(hi), v - __half2float(hi), lo2);
        const __half lo = __float2half_rn((v - __half2float(hi)) * (float)(1 << kPairShift));
        dst[r * Dp + d] = __ushort_as_bfloat16(__half_as_ushort(hi));
        dst[(rows_pad + r) * Dp + d] = __ushort_as_bfloat16(__half_as_ushort(lo));
      } else {
        // exact 3-way split: v == hi + mid + lo
        float rem = v;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          if (p < planes) {
            __nv_bfloat16 h = __float2bfloat16_rn(rem);
            dst[((int64_t)p * rows_pad + r) * Dp + d] = h;
            rem = rem - __bfloat162float(h);
          }
        }
      }
    }
    if (half_sqnorm) {
      ss2 = group_sum<G>(ss2);
      if (lane == 0) half_sqnorm[r] = real ? 0.5f * ss2 : INFINITY;
    }
    if (lo_norm_max) {   // max_j |row_j - hi_j|_2: the error bound of a one-term (hi plane only) contraction
      lo2 = group_sum<G>(lo2);
      if (lane == 0 && real) atomicMax(reinterpret_cast<int*>(lo_norm_max), __float_as_int(sqrtf(lo2)));
    }
  }
}

// Vectorised variant for D % 8 == 0 (Dp == D or zero-padded tail): every lane owns 8 contiguous elements per
// step (128-bit loads/stores), keeps them in registers across the norm reduction, and writes each plane once.
template <typename T, int G, int NV>
__global__ void __launch_bounds__(256) pack_rows_vec_kernel(
    const T* __restrict__ src, int64_t rows, int64_t rows_pad, int D, int Dp, int normalize, int planes,
    __nv_bfloat16* __restrict__ dst, float* __restrict__ half_sqnorm, float* __restrict__ writeback,
    unsigned long long* __restrict__ keys, int64_t n_keys, uint4* __restrict__ zero_fill, int64_t n_zero16,
    float* __restrict__ lo_norm_max, int fold_role) {
  pdl_wait();               // the source rows / the key buffer may still be in use by the preceding launch
  pdl_launch_dependents();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_keys; i += (int64_t)gridDim.x * blockDim.x)
    keys[i] = ~0ull;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_zero16; i += (int64_t)gridDim.x * blockDim.x)
    zero_fill[i] = make_uint4(0u, 0u, 0u, 0u);
  const int lane = threadIdx.x % G;
  const int64_t rows_per_block = blockDim.x / G;
  for (int64_t r = blockIdx.x * rows_per_block + threadIdx.x / G; r < rows_pad;
       r += (int64_t)gridDim.x * rows_per_block) {  // rows_pad is a multiple of 256: warp-uniform trip count
    const bool real = r < rows;
    float v[NV][8];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      const int d0 = (it * G + lane) * 8;
      if (real && d0 < D) {
        if constexpr (sizeof(T) == 4) {
          const float4 a = *reinterpret_cast<const float4*>(src + r * D + d0);
          const float4 b = *reinterpret_cast<const float4*>(src + r * D + d0 + 4);
          v[it][0] = a.x; v[it][1] = a.y; v[it][2] = a.z; v[it][3] = a.w;
          v[it][4] = b.x; v[it][5] = b.y; v[it][6] = b.z; v[it][7] = b.w;
        } else {
          const uint4 raw = *reinterpret_cast<const uint4*>(src + r * D + d0);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h[i]);
            v[it][2 * i] = f.x;
            v[it][2 * i + 1] = f.y;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[it][i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = fmaf(v[it][i], v[it][i], ss);
    }
    ss = group_sum<G>(ss);
    const float denom = normalize ? fmaxf(sqrtf(ss), kNormEps) : 1.f;
    float f16_scale = 1.f;
    if (planes == VQB_PLANES_F16) {
      float amax = 0.f;
#pragma unroll
      for (int it = 0; it < NV; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) amax = fmaxf(amax, fabsf(v[it][i]));
      f16_scale = f16_row_scale(group_max<G>(amax));
    }
    float ss2 = 0.f, lo2 = 0.f;
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      if (normalize) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[it][i] = __fdiv_rn(v[it][i], denom);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) ss2 = fmaf(v[it][i], v[it][i], ss2);
    }
    if (half_sqnorm || fold_role >= 0) ss2 = group_sum<G>(ss2);   // warp-uniform condition; the row's |.|^2
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      const int d0 = (it * G + lane) * 8;
      if (d0 < Dp) {
        if (fold_role >= 0 && d0 == D && real) {
          // vqb_fold_l2_side fused into the pack: the lane that owns the first padding chunk writes the folded L2 side
          // terms instead of zeros (columns D..D+2: the codes' term / the tokens' ones, D+3..D+5 the other way round)
          const int own = fold_role == 1 ? 0 : 3, partner = 3 - own;
          const float hv = fold_role == 0 ? 0.f : -0.5f * ss2;
          const __nv_bfloat16 hi = __float2bfloat16_rn(hv);
          const float r1 = hv - __bfloat162float(hi);
          const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
          const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
          const __nv_bfloat16 piece[3] = {hi, mid, lo};
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            if (p < planes) {
              uint4 raw = make_uint4(0u, 0u, 0u, 0u);
              __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(&raw);
              if (p == 0)
                for (int c = 0; c < 3; ++c) e[partner + c] = __float2bfloat16(1.f);
              if (planes >= 3) e[own] = piece[p];
              else if (p == 0)
                for (int c = 0; c < 3; ++c) e[own + c] = piece[c];
              *reinterpret_cast<uint4*>(dst + ((int64_t)p * rows_pad + r) * Dp + d0) = raw;
            }
          }
          continue;
        }
        if (writeback && real && d0 < D) {
          *reinterpret_cast<float4*>(writeback + r * D + d0) = make_float4(v[it][0], v[it][1], v[it][2], v[it][3]);
          *reinterpret_cast<float4*>(writeback + r * D + d0 + 4) = make_float4(v[it][4], v[it][5], v[it][6], v[it][7]);
        }
        float rem[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rem[i] = v[it][i];
        if (planes == VQB_PLANES_F16) {
          uint4 raw;
          __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(rem[2 * i] * f16_scale, rem[2 * i + 1] * f16_scale);
          *reinterpret_cast<uint4*>(dst + r * Dp + d0) = raw;
          continue;
        }
        if (is_f16x2(planes)) {
          uint4 raw_hi, raw_lo;
          __half2* hh = reinterpret_cast<__half2*>(&raw_hi);
          __half2* hl = reinterpret_cast<__half2*>(&raw_lo);
          constexpr float kUp = (float)(1 << kPairShift);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            hh[i] = __floats2half2_rn(rem[2 * i], rem[2 * i + 1]);
            const float2 back = __half22float2(hh[i]);
            lo2 = fmaf(rem[2 * i] - back.x, rem[2 * i] - back.x, lo2);
            lo2 = fmaf(rem[2 * i + 1] - back.y, rem[2 * i + 1] - back.y, lo2);
            hl[i] = __floats2half2_rn((rem[2 * i] - back.x) * kUp, (rem[2 * i + 1] - back.y) * kUp);
          }
          *reinterpret_cast<uint4*>(dst + r * Dp + d0) = raw_hi;
          *reinterpret_cast<uint4*>(dst + (rows_pad + r) * Dp + d0) = raw_lo;
          continue;
        }
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          if (p < planes) {
            uint4 raw;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              h[i] = __floats2bfloat162_rn(rem[2 * i], rem[2 * i + 1]);
              const float2 back = __bfloat1622float2(h[i]);
              rem[2 * i] -= back.x;       // exact: v == hi + mid + lo
              rem[2 * i + 1] -= back.y;
            }
            *reinterpret_cast<uint4*>(dst + ((int64_t)p * rows_pad + r) * Dp + d0) = raw;
          }
        }
      }
    }
    if (half_sqnorm) {
      if (lane == 0) half_sqnorm[r] = real ? 0.5f * ss2 : INFINITY;
    }
    if (lo_norm_max) {
      lo2 = group_sum<G>(lo2);
      if (lane == 0 && real) atomicMax(reinterpret_cast<int*>(lo_norm_max), __float_as_int(sqrtf(lo2)));
    }
  }
}

template <typename TI, typename TO, int G>
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const TI* __restrict__ x, int64_t rows, int D,
                                                         TO* __restrict__ y) {
  const int lane = threadIdx.x % G;
  const int64_t rows_per_block = blockDim.x / G;
  for (int64_t base = blockIdx.x * rows_per_block; base < rows; base += (int64_t)gridDim.x * rows_per_block) {
    const int64_t r = base + threadIdx.x / G;  // warp-uniform trip count: shuffles below need every lane
    const bool valid = r < rows;
    float ss = 0.f;
    if (valid)
      for (int d = lane; d < D; d += G) {
        float v = to_f32<TI>(x[r * D + d]);
        ss = fmaf(v, v, ss);
      }
    ss = group_sum<G>(ss);
    const float denom = fmaxf(sqrtf(ss), kNormEps);
    if (valid)
      for (int d = lane; d < D; d += G) y[r * D + d] = from_f32<TO>(__fdiv_rn(to_f32<TI>(x[r * D + d]), denom));
  }
}

// gx = (gy - (gy . y) y) / denom with y = x / denom   (exact Jacobian of F.normalize away from the eps clamp;
// when ||x|| < eps the forward is x/eps and the Jacobian is I/eps)
template <typename TG, typename TI, typename TO, int G>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const TG* __restrict__ gy, const TI* __restrict__ x,
                                                         int64_t rows, int D, TO* __restrict__ gx) {
  const int lane = threadIdx.x % G;
  const int64_t rows_per_block = blockDim.x / G;
  for (int64_t base = blockIdx.x * rows_per_block; base < rows; base += (int64_t)gridDim.x * rows_per_block) {
    const int64_t r = base + threadIdx.x / G;
    const bool valid = r < rows;
    float ss = 0.f, gd = 0.f;
    for (int d = lane; valid && d < D; d += G) {
      float v = to_f32<TI>(x[r * D + d]);
      float g = to_f32<TG>(gy[r * D + d]);
      ss = fmaf(v, v, ss);
      gd = fmaf(g, v, gd);
    }
    ss = group_sum<G>(ss);
    gd = group_sum<G>(gd);
    const float nrm = sqrtf(ss);
    const bool clamped = nrm < kNormEps;
    const float denom = fmaxf(nrm, kNormEps);
    const float inv = 1.f / denom;
    const float proj = clamped ? 0.f : gd * inv * inv;  // (g . y) y == (g . x) x / denom^2
    for (int d = lane; valid && d < D; d += G) {
      float v = to_f32<TI>(x[r * D + d]);
      float g = to_f32<TG>(gy[r * D + d]);
      gx[r * D + d] = from_f32<TO>((g - proj * v) * inv);
    }
  }
}

// out[r] = 1 / max(||x_r||, eps) for r < rows, 0 in the padding up to rows_pad (per-column scale of vqb_assign).
// f16_rows: the rows are fed as a VQB_PLANES_F16 plane, i.e. possibly scaled by f16_row_scale -> fold it in.
template <typename T, int G>
__global__ void __launch_bounds__(256) row_inv_norm_kernel(const T* __restrict__ x, int64_t rows, int64_t rows_pad, int D,
                                                           int f16_rows, float* __restrict__ out) {
  pdl_wait();               // PDL: inputs come from the preceding launches
  pdl_launch_dependents();
  const int lane = threadIdx.x % G;
  const int64_t rows_per_block = blockDim.x / G;
  for (int64_t r = blockIdx.x * rows_per_block + threadIdx.x / G; r < rows_pad;
       r += (int64_t)gridDim.x * rows_per_block) {
    float ss = 0.f, amax = 0.f;
    if (r < rows)
      for (int d = lane; d < D; d += G) {
        const float v = to_f32<T>(x[r * D + d]);
        ss = fmaf(v, v, ss);
        amax = fmaxf(amax, fabsf(v));
      }
    ss = group_sum<G>(ss);
    const float scale = f16_rows ? f16_row_scale(group_max<G>(amax)) : 1.f;
    if (lane == 0) out[r] = r < rows ? __fdiv_rn(1.f, fmaxf(sqrtf(ss), kNormEps) * scale) : 0.f;
  }
}

// L2 distance with spare K padding: the side terms ride in the contraction instead of the epilogue (see
// vqb_fold_l2_side in vqb200.h).  One thread per row; -h is split EXACTLY into three bf16 pieces (an fp32 value is
// hi + mid + lo), kept in one column across three planes or, for fewer planes, in three columns of plane 0.
__global__ void __launch_bounds__(256) fold_l2_side_kernel(__nv_bfloat16* __restrict__ planes, int n_planes, int64_t rows,
                                                           int64_t rows_pad, int D, int Dp,
                                                           const float* __restrict__ half_sqnorm, int role) {
  pdl_wait();               // the planes and the side vector come from the preceding pack launch
  pdl_launch_dependents();
  const int own = role == 1 ? D : D + 3, partner = role == 1 ? D + 3 : D;
  const __nv_bfloat16 one = __float2bfloat16(1.f);
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    __nv_bfloat16* row0 = planes + r * Dp;
    for (int c = 0; c < 3; ++c) row0[partner + c] = one;
    if (half_sqnorm == nullptr) continue;
    const float v = -half_sqnorm[r];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hi);            // exact
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));   // exact: 24 = 3 x 8 significant bits
    if (n_planes >= 3) {
      row0[own] = hi;
      row0[rows_pad * Dp + own] = mid;
      row0[2 * rows_pad * Dp + own] = lo;
    } else {
      row0[own] = hi;
      row0[own + 1] = mid;
      row0[own + 2] = lo;
    }
  }
}

template <int G, typename F>
static inline void launch_rows(int64_t rows, F&& f) {
  const int rows_per_block = 256 / G;
  int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  f((int)blocks);
}

#define VQB_DISPATCH_G(G_, ...)                  \
  switch (G_) {                                  \
    case 1: { constexpr int G = 1; __VA_ARGS__; } break;   \
    case 2: { constexpr int G = 2; __VA_ARGS__; } break;   \
    case 4: { constexpr int G = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int G = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int G = 16; __VA_ARGS__; } break; \
    default: { constexpr int G = 32; __VA_ARGS__; } break; \
  }

}  // namespace vqb

using namespace vqb;

extern "C" {

int vqb_abi_version(void) { return VQB200_ABI_VERSION; }
const char* vqb_last_error(void) { return g_err; }

int vqb_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  VQB_CUDA_OK(cudaGetDevice(&dev));
  if (sm) VQB_CUDA_OK(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
  if (major) VQB_CUDA_OK(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
  if (minor) VQB_CUDA_OK(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
  return VQB_OK;
}

int64_t vqb_operand_dp(int D) {
  if (D <= 16) return 16;
  if (D <= 32) return 32;
  return round_up(D, 64);
}
int64_t vqb_operand_rows_pad(int64_t rows) { return round_up(rows < 1 ? 1 : rows, 256); }
size_t vqb_operand_bytes(int64_t rows, int D, int planes) {
  return (size_t)plane_count(planes) * (size_t)vqb_operand_rows_pad(rows) * (size_t)vqb_operand_dp(D) * 2;
}

static int pack_rows_impl(const void* src, int src_dtype, int64_t rows, int D, int normalize, int planes,
                  void* dst_planes, float* half_sqnorm, float* writeback, unsigned long long* keys,
                  int64_t n_keys, void* zero_fill, int64_t zero_bytes, float* lo_norm_max, int fold_role, void* stream) {
  VQB_REQUIRE(src && dst_planes, "vqb_pack_rows: null pointer");
  VQB_REQUIRE(zero_fill == nullptr || ((uintptr_t)zero_fill % 16 == 0 && zero_bytes % 16 == 0 && zero_bytes >= 0),
              "vqb_pack_rows: zero_fill must be 16-byte aligned and a multiple of 16 bytes");
  uint4* zf = (uint4*)zero_fill;
  const int64_t nz = zero_fill ? zero_bytes / 16 : 0;
  VQB_REQUIRE(rows >= 0 && D >= 1 && D <= 8192, "vqb_pack_rows: bad shape rows=%lld D=%d", (long long)rows, D);
  VQB_REQUIRE(planes_valid(planes), "vqb_pack_rows: planes must be 1..3, VQB_PLANES_F16 or VQB_PLANES_F16X2 (got %d)", planes);
  VQB_REQUIRE(!is_f16x2(planes) || normalize, "vqb_pack_rows: VQB_PLANES_F16X2 needs normalize = 1 (|v| <= 1)");
  VQB_REQUIRE(planes != VQB_PLANES_F16 || (!normalize && src_dtype == VQB_BF16),
              "vqb_pack_rows: VQB_PLANES_F16 takes un-normalised bf16 rows (8 significant bits are exact in fp16)");
  VQB_REQUIRE(src_dtype == VQB_F32 || src_dtype == VQB_BF16, "vqb_pack_rows: bad dtype %d", src_dtype);
  const int64_t rows_pad = vqb_operand_rows_pad(rows);
  const int Dp = (int)vqb_operand_dp(D);
  cudaStream_t st = (cudaStream_t)stream;
  const bool aligned = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst_planes % 16 == 0) &&
                       (writeback == nullptr || (uintptr_t)writeback % 16 == 0);
  if (D % 8 == 0 && Dp <= 2048 && aligned) {  // vector path
    const int slices = Dp / 8;
    int gv = 1;
    while (gv < 32 && gv < slices) gv <<= 1;
    int nv = 1;
    while (nv * gv < slices) nv <<= 1;
    const int rpb = 256 / gv;
    int64_t blocks64 = (rows_pad + rpb - 1) / rpb;
    const int blocks = (int)(blocks64 < (int64_t)sm_count() * 8 ? blocks64 : (int64_t)sm_count() * 8);
#define VQB_PACK_CASE(G_, NV_)                                                                                       \
    if (gv == G_ && nv == NV_) {                                                                                     \
      if (src_dtype == VQB_F32)                                                                                      \
        launch_pdl(pack_rows_vec_kernel<float, G_, NV_>, blocks, 256, 0, st, (const float*)src, rows, rows_pad, D, Dp, \
            normalize, planes, (__nv_bfloat16*)dst_planes, half_sqnorm, writeback, keys, keys ? n_keys : 0, zf, nz, lo_norm_max, fold_role);        \
      else                                                                                                           \
        launch_pdl(pack_rows_vec_kernel<__nv_bfloat16, G_, NV_>, blocks, 256, 0, st, (const __nv_bfloat16*)src, rows, \
            rows_pad, D, Dp, normalize, planes, (__nv_bfloat16*)dst_planes, half_sqnorm, writeback, keys,            \
            keys ? n_keys : 0, zf, nz, lo_norm_max, fold_role);                                                                           \
      VQB_LAUNCH_OK();                                                                                               \
      return VQB_OK;                                                                                                 \
    }
    VQB_PACK_CASE(1, 1) VQB_PACK_CASE(2, 1) VQB_PACK_CASE(4, 1) VQB_PACK_CASE(8, 1) VQB_PACK_CASE(16, 1)
    VQB_PACK_CASE(32, 1) VQB_PACK_CASE(32, 2) VQB_PACK_CASE(32, 4) VQB_PACK_CASE(32, 8)
#undef VQB_PACK_CASE
  }
  const int g = lanes_per_row(Dp);
  VQB_DISPATCH_G(g, launch_rows<G>(rows_pad, [&](int blocks) {
    if (src_dtype == VQB_F32)
      launch_pdl(pack_rows_kernel<float, G>, blocks, 256, 0, st, (const float*)src, rows, rows_pad, D, Dp, normalize,
                                                          planes, (__nv_bfloat16*)dst_planes, half_sqnorm,
                                                          writeback, keys, keys ? n_keys : 0, zf, nz, lo_norm_max);
    else
      launch_pdl(pack_rows_kernel<__nv_bfloat16, G>, blocks, 256, 0, st, 
          (const __nv_bfloat16*)src, rows, rows_pad, D, Dp, normalize, planes, (__nv_bfloat16*)dst_planes,
          half_sqnorm, writeback, keys, keys ? n_keys : 0, zf, nz, lo_norm_max);
  }));
  VQB_LAUNCH_OK();
  if (fold_role >= 0) {   // rows that do not take the vector path: the fold runs as its own launch
    VQB_REQUIRE(fold_role == 0 || half_sqnorm, "vqb_pack_rows_fold: this row width needs half_sqnorm for roles 1 and 2");
    int64_t blocks = (rows + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    VQB_CUDA_OK(launch_pdl(fold_l2_side_kernel, (int)blocks, 256, 0, st, (__nv_bfloat16*)dst_planes, planes, rows, rows_pad, D,
                           Dp, fold_role == 0 ? (const float*)nullptr : (const float*)half_sqnorm, fold_role == 1 ? 1 : 0));
    VQB_LAUNCH_OK();
  }
  return VQB_OK;
}

int vqb_pack_rows(const void* src, int src_dtype, int64_t rows, int D, int normalize, int planes,
                  void* dst_planes, float* half_sqnorm, float* writeback, unsigned long long* keys,
                  int64_t n_keys, void* zero_fill, int64_t zero_bytes, float* lo_norm_max, void* stream) {
  return pack_rows_impl(src, src_dtype, rows, D, normalize, planes, dst_planes, half_sqnorm, writeback, keys, n_keys,
                        zero_fill, zero_bytes, lo_norm_max, -1, stream);
}

int vqb_pack_rows_fold(const void* src, int src_dtype, int64_t rows, int D, int normalize, int planes,
                       void* dst_planes, float* half_sqnorm, float* writeback, unsigned long long* keys,
                       int64_t n_keys, void* zero_fill, int64_t zero_bytes, int fold_role, void* stream) {
  VQB_REQUIRE(planes >= 1 && planes <= 3, "vqb_pack_rows_fold: exact bf16 planes only (1..3), got %d", planes);
  VQB_REQUIRE(fold_role >= 0 && fold_role <= 2, "vqb_pack_rows_fold: fold_role must be 0, 1 or 2");
  VQB_REQUIRE(D >= 1 && vqb_operand_dp(D) - D >= VQB_L2_FOLD_COLUMNS, "vqb_pack_rows_fold: needs %d spare columns (D=%d)",
              VQB_L2_FOLD_COLUMNS, D);
  return pack_rows_impl(src, src_dtype, rows, D, normalize, planes, dst_planes, half_sqnorm, writeback, keys, n_keys,
                        zero_fill, zero_bytes, nullptr, fold_role, stream);
}

int vqb_row_inv_norm(const void* x, int x_dtype, int64_t rows, int D, int f16_rows, float* out, void* stream) {
  VQB_REQUIRE(x && out, "vqb_row_inv_norm: null pointer");
  VQB_REQUIRE(rows >= 1 && D >= 1, "vqb_row_inv_norm: bad shape");
  const int64_t rows_pad = vqb_operand_rows_pad(rows);
  cudaStream_t st = (cudaStream_t)stream;
  const int g = lanes_per_row(D);
  VQB_DISPATCH_G(g, launch_rows<G>(rows_pad, [&](int blocks) {
    if (x_dtype == VQB_F32)
      launch_pdl(row_inv_norm_kernel<float, G>, blocks, 256, 0, st, (const float*)x, rows, rows_pad, D, f16_rows, out);
    else
      launch_pdl(row_inv_norm_kernel<__nv_bfloat16, G>, blocks, 256, 0, st, (const __nv_bfloat16*)x, rows, rows_pad, D, f16_rows, out);
  }));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_l2norm_forward(const void* x, int x_dtype, int64_t rows, int D, void* y, int y_dtype, void* stream) {
  VQB_REQUIRE(x && y, "vqb_l2norm_forward: null pointer");
  VQB_REQUIRE(rows >= 0 && D >= 1, "vqb_l2norm_forward: bad shape");
  if (rows == 0) return VQB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = lanes_per_row(D);
  VQB_DISPATCH_G(g, launch_rows<G>(rows, [&](int blocks) {
    if (x_dtype == VQB_F32 && y_dtype == VQB_F32)
      l2norm_fwd_kernel<float, float, G><<<blocks, 256, 0, st>>>((const float*)x, rows, D, (float*)y);
    else if (x_dtype == VQB_BF16 && y_dtype == VQB_F32)
      l2norm_fwd_kernel<__nv_bfloat16, float, G><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, rows, D, (float*)y);
    else if (x_dtype == VQB_BF16 && y_dtype == VQB_BF16)
      l2norm_fwd_kernel<__nv_bfloat16, __nv_bfloat16, G><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, rows, D, (__nv_bfloat16*)y);
    else
      l2norm_fwd_kernel<float, __nv_bfloat16, G><<<blocks, 256, 0, st>>>((const float*)x, rows, D, (__nv_bfloat16*)y);
  }));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_fold_l2_side(void* planes, int n_planes, int64_t rows, int D, const float* half_sqnorm, int role, void* stream) {
  VQB_REQUIRE(planes, "vqb_fold_l2_side: null pointer");
  VQB_REQUIRE(n_planes >= 1 && n_planes <= 3, "vqb_fold_l2_side: exact bf16 planes only (1..3), got %d", n_planes);
  VQB_REQUIRE(role == 0 || role == 1, "vqb_fold_l2_side: role must be 0 (tokens) or 1 (codes)");
  VQB_REQUIRE(rows >= 0 && D >= 1, "vqb_fold_l2_side: bad shape rows=%lld D=%d", (long long)rows, D);
  const int Dp = (int)vqb_operand_dp(D);
  VQB_REQUIRE(Dp - D >= VQB_L2_FOLD_COLUMNS, "vqb_fold_l2_side: needs %d spare columns (D=%d, Dp=%d)", VQB_L2_FOLD_COLUMNS, D, Dp);
  VQB_REQUIRE(role == 0 || half_sqnorm, "vqb_fold_l2_side: the codes role needs half_sqnorm");
  if (rows == 0) return VQB_OK;
  int64_t blocks = (rows + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  VQB_CUDA_OK(launch_pdl(fold_l2_side_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, (__nv_bfloat16*)planes, n_planes,
                         rows, vqb_operand_rows_pad(rows), D, Dp, half_sqnorm, role));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_l2norm_backward(const void* gy, int g_dtype, const void* x, int x_dtype, int64_t rows, int D, void* gx,
                        int gx_dtype, void* stream) {
  VQB_REQUIRE(gy && x && gx, "vqb_l2norm_backward: null pointer");
  VQB_REQUIRE(g_dtype == VQB_F32, "vqb_l2norm_backward: upstream gradient must be fp32");
  if (rows == 0) return VQB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = lanes_per_row(D);
  VQB_DISPATCH_G(g, launch_rows<G>(rows, [&](int blocks) {
    if (x_dtype == VQB_F32 && gx_dtype == VQB_F32)
      l2norm_bwd_kernel<float, float, float, G><<<blocks, 256, 0, st>>>((const float*)gy, (const float*)x, rows, D, (float*)gx);
    else if (x_dtype == VQB_BF16 && gx_dtype == VQB_BF16)
      l2norm_bwd_kernel<float, __nv_bfloat16, __nv_bfloat16, G><<<blocks, 256, 0, st>>>((const float*)gy, (const __nv_bfloat16*)x, rows, D, (__nv_bfloat16*)gx);
    else if (x_dtype == VQB_BF16 && gx_dtype == VQB_F32)
      l2norm_bwd_kernel<float, __nv_bfloat16, float, G><<<blocks, 256, 0, st>>>((const float*)gy, (const __nv_bfloat16*)x, rows, D, (float*)gx);
    else
      l2norm_bwd_kernel<float, float, __nv_bfloat16, G><<<blocks, 256, 0, st>>>((const float*)gy, (const float*)x, rows, D, (__nv_bfloat16*)gx);
  }));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

}  // extern "C"
