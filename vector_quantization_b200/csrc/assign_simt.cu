// vqb_assign dispatcher + the CUDA-core (fp32 FMA) assignment kernel.  The SIMT kernel has the same
// contract as the tcgen05 kernel (assign_tc.cu) and exists as its on-device cross-check: it rebuilds the
// exact fp32 operands from the bf16 planes and accumulates in fp32 in a fixed order.
#include <math.h>

#include "common.cuh"

namespace vqb {

int assign_tc_launch(const void* a_planes, int pa, int64_t a_rows, int64_t a_plane_rows, const void* b_planes, int pb,
                     int64_t b_rows, int64_t b_plane_rows, int D, const float* b_side, int side_mode,
                     int64_t b_index_offset, unsigned long long* keys, unsigned long long* second_keys,
                     const int* a_rows_dev, cudaStream_t st);

constexpr int SA = 128;  // A rows per block (one per thread)
constexpr int SB = 32;   // B rows per inner tile
constexpr int SD = 32;   // depth chunk

__device__ __forceinline__ float load_planes(const __nv_bfloat16* __restrict__ base, int planes,
                                             int64_t plane_stride, int64_t off) {
  if (planes == VQB_PLANES_F16) return __half2float(__ushort_as_half(__bfloat16_as_ushort(base[off])));
  if (is_f16x2(planes)) {  // fp16 pair: v = hi + lo' * 2^-11
    const __half hi = __ushort_as_half(__bfloat16_as_ushort(base[off]));
    const __half lo = __ushort_as_half(__bfloat16_as_ushort(base[plane_stride + off]));
    return __half2float(hi) + __half2float(lo) * (1.f / (float)(1 << kPairShift));
  }
  float v = 0.f;
  // lo -> hi so that the exact sum is reproduced bit for bit (|lo| << |mid| << |hi|)
  for (int p = planes - 1; p >= 0; --p) v += __bfloat162float(base[p * plane_stride + off]);
  return v;
}

__global__ void __launch_bounds__(SA) assign_simt_kernel(const __nv_bfloat16* __restrict__ A, int pa,
                                                         int64_t a_rows, int64_t a_rows_pad,
                                                         const __nv_bfloat16* __restrict__ B, int pb,
                                                         int64_t b_rows, int64_t b_rows_pad, int Dp,
                                                         const float* __restrict__ h, int side_mode, int64_t b_off,
                                                         unsigned long long* __restrict__ keys) {
  __shared__ float As[SA][SD + 1];
  __shared__ float Bs[SB][SD];
  const int t = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * SA;
  const int64_t row = row0 + t;
  // this block's slice of B tiles
  const int64_t n_btiles = (b_rows + SB - 1) / SB;
  const int64_t per = (n_btiles + gridDim.y - 1) / gridDim.y;
  const int64_t bt0 = blockIdx.y * per, bt1 = min(n_btiles, bt0 + per);

  float best = -INFINITY;
  uint32_t best_j = 0xffffffffu;
  for (int64_t bt = bt0; bt < bt1; ++bt) {
    const int64_t j0 = bt * SB;
    float acc[SB];
#pragma unroll
    for (int j = 0; j < SB; ++j) acc[j] = 0.f;
    for (int d0 = 0; d0 < Dp; d0 += SD) {
      __syncthreads();
      for (int i = t; i < SA * SD; i += SA) {
        const int r = i / SD, d = i % SD;
        As[r][d] = (d0 + d < Dp && row0 + r < a_rows_pad) ? load_planes(A, pa, a_rows_pad * Dp, (row0 + r) * Dp + d0 + d) : 0.f;
      }
      for (int i = t; i < SB * SD; i += SA) {
        const int r = i / SD, d = i % SD;
        Bs[r][d] = (d0 + d < Dp && j0 + r < b_rows_pad) ? load_planes(B, pb, b_rows_pad * Dp, (j0 + r) * Dp + d0 + d) : 0.f;
      }
      __syncthreads();
      float a[SD];
#pragma unroll
      for (int d = 0; d < SD; ++d) a[d] = As[t][d];
#pragma unroll
      for (int j = 0; j < SB; ++j)
#pragma unroll
        for (int d = 0; d < SD; ++d) acc[j] = fmaf(a[d], Bs[j][d], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < SB; ++j) {
      if (j0 + j < b_rows) {
        const float s = (h == nullptr || side_mode == 0) ? acc[j] : (side_mode == 1 ? acc[j] - h[j0 + j] : acc[j] * h[j0 + j]);
        if (s > best) {
          best = s;
          best_j = (uint32_t)(j0 + j + b_off);
        }
      }
    }
  }
  if (row < a_rows && best_j != 0xffffffffu) atomicMin(keys + row, make_key(best, best_j));
}

}  // namespace vqb

using namespace vqb;

extern "C" int vqb_assign_ex(const void* a_planes, int pa, int64_t a_rows, int64_t a_plane_rows, const void* b_planes,
                             int pb, int64_t b_rows, int64_t b_plane_rows, int D, const float* b_half_sqnorm,
                             int side_mode, int64_t b_index_offset, unsigned long long* keys,
                             unsigned long long* second_keys, const int* a_rows_dev, int backend, void* stream);

extern "C" int vqb_assign(const void* a_planes, int pa, int64_t a_rows, int64_t a_plane_rows, const void* b_planes,
                          int pb, int64_t b_rows, int64_t b_plane_rows, int D, const float* b_half_sqnorm,
                          int side_mode, int64_t b_index_offset, unsigned long long* keys, int backend,
                          void* stream) {
  return vqb_assign_ex(a_planes, pa, a_rows, a_plane_rows, b_planes, pb, b_rows, b_plane_rows, D, b_half_sqnorm, side_mode,
                       b_index_offset, keys, nullptr, nullptr, backend, stream);
}

extern "C" int vqb_assign_ex(const void* a_planes, int pa, int64_t a_rows, int64_t a_plane_rows, const void* b_planes,
                             int pb, int64_t b_rows, int64_t b_plane_rows, int D, const float* b_half_sqnorm,
                             int side_mode, int64_t b_index_offset, unsigned long long* keys,
                             unsigned long long* second_keys, const int* a_rows_dev, int backend, void* stream) {
  VQB_REQUIRE(a_planes && b_planes && keys, "vqb_assign: null pointer");
  VQB_REQUIRE(backend == VQB_BACKEND_TCGEN05 || (second_keys == nullptr && a_rows_dev == nullptr),
              "vqb_assign_ex: runner-up keys / device row count need the tcgen05 backend");
  VQB_REQUIRE(a_rows >= 1 && b_rows >= 1 && D >= 1, "vqb_assign: bad shape a_rows=%lld b_rows=%lld D=%d",
              (long long)a_rows, (long long)b_rows, D);
  VQB_REQUIRE(planes_valid(pa) && planes_valid(pb), "vqb_assign: planes must be 1..3, VQB_PLANES_F16 or VQB_PLANES_F16X2");
  // fp16 and bf16 operands cannot be mixed on the tensor core; the one exception is a ONE-plane bf16 A operand
  // (zero-copy tokens) against fp16 B planes with D <= 64, converted inside the kernel
  VQB_REQUIRE(is_f16(pa) == is_f16(pb) || (pa == 1 && is_f16(pb) && vqb_operand_dp(D) <= 64),
              "vqb_assign: the tensor core cannot mix fp16 and bf16 operands (a_nplanes=0x%x b_nplanes=0x%x)", pa, pb);
  VQB_REQUIRE(b_rows + b_index_offset < 0xffffffffll && b_index_offset >= 0,
              "vqb_assign: column index does not fit 32 bits");
  cudaStream_t st = (cudaStream_t)stream;
  if (a_plane_rows <= 0) a_plane_rows = vqb_operand_rows_pad(a_rows);
  if (b_plane_rows <= 0) b_plane_rows = vqb_operand_rows_pad(b_rows);
  VQB_REQUIRE(a_plane_rows >= a_rows && b_plane_rows >= b_rows, "vqb_assign: plane stride smaller than the row count");
  VQB_REQUIRE(side_mode >= 0 && side_mode <= 2, "vqb_assign: side_mode must be 0 (none), 1 (subtract) or 2 (scale)");
  if (backend == VQB_BACKEND_TCGEN05)
    return assign_tc_launch(a_planes, pa, a_rows, a_plane_rows, b_planes, pb, b_rows, b_plane_rows, D, b_half_sqnorm,
                            side_mode, b_index_offset, keys, second_keys, a_rows_dev, st);
  VQB_REQUIRE(backend == VQB_BACKEND_SIMT, "vqb_assign: unknown backend %d", backend);
  const int Dp = (int)vqb_operand_dp(D);
  const int64_t a_tiles = (a_rows + SA - 1) / SA;
  int64_t splits = ((int64_t)sm_count() * 4 + a_tiles - 1) / a_tiles;
  const int64_t n_btiles = (b_rows + SB - 1) / SB;
  if (splits > n_btiles) splits = n_btiles;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  dim3 grid((unsigned)a_tiles, (unsigned)splits);
  assign_simt_kernel<<<grid, SA, 0, st>>>((const __nv_bfloat16*)a_planes, pa, a_rows, a_plane_rows,
                                          (const __nv_bfloat16*)b_planes, pb, b_rows, b_plane_rows, Dp, b_half_sqnorm,
                                          side_mode, b_index_offset, keys);
  VQB_LAUNCH_OK();
  return VQB_OK;
}
