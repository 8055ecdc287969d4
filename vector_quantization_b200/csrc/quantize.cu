// Codebook gather + straight-through estimator + MSE-loss reduction (forward, one pass) and the
// closed-form backward.  HBM-bound: per token reads x row + codebook row + index, writes z row.
#include <math.h>

#include "common.cuh"

namespace vqb {

constexpr int kMaxPartials = 4096;

static inline int lanes_per_row_q(int D) {
  int g = 1;
  while (g < 32 && g * 4 < D) g <<= 1;
  return g;
}

// block-wide deterministic sum of two values; result valid in thread 0
__device__ __forceinline__ void block_sum2(float& a, float& b, float* sh /* >= 2*8 floats */) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sh[w] = a;
    sh[8 + w] = b;
  }
  __syncthreads();
  if (w == 0) {
    a = l < (blockDim.x >> 5) ? sh[l] : 0.f;
    b = l < (blockDim.x >> 5) ? sh[8 + l] : 0.f;
    a = warp_sum(a);
    b = warp_sum(b);
  }
}

template <typename TX, typename TO, int G>
__global__ void __launch_bounds__(256) gather_ste_loss_kernel(
    const TX* __restrict__ x, int64_t N, int D, const float* __restrict__ W, int64_t K,
    const int64_t* __restrict__ quant, TO* __restrict__ z_out, int want_norm, float* __restrict__ mse4,
    float* __restrict__ partials, unsigned int* __restrict__ ticket) {
  __shared__ float sh[16];
  __shared__ bool is_last;
  const int lane = threadIdx.x % G;
  const int64_t rows_per_block = blockDim.x / G;
  float sse = 0.f, sse_n = 0.f;
  for (int64_t base = blockIdx.x * rows_per_block; base < N; base += (int64_t)gridDim.x * rows_per_block) {
    const int64_t n_raw = base + threadIdx.x / G;  // warp-uniform trip count: shuffles need every lane
    const bool valid = n_raw < N;
    const int64_t n = valid ? n_raw : N - 1;
    int64_t q = quant[n];
    q = q < 0 ? 0 : (q >= K ? K - 1 : q);  // never read out of bounds on a corrupt index
    const float* __restrict__ wrow = W + q * D;
    const TX* __restrict__ xrow = x + n * D;
    float sxx = 0.f, sww = 0.f;
    for (int d = lane; d < D; d += G) {
      const float xv = to_f32<TX>(xrow[d]);
      const float wv = __ldg(wrow + d);
      const float diff = __fsub_rn(wv, xv);
      if (valid) {
        z_out[n * D + d] = from_f32<TO>(__fadd_rn(xv, diff));  // ste value: x + (z - x)
        sse = fmaf(diff, diff, sse);
      }
      sxx = fmaf(xv, xv, sxx);
      sww = fmaf(wv, wv, sww);
    }
    if (want_norm) {
      sxx = group_sum<G>(sxx);
      sww = group_sum<G>(sww);
      const float dx = fmaxf(sqrtf(sxx), kNormEps), dw = fmaxf(sqrtf(sww), kNormEps);
      for (int d = lane; d < D; d += G) {
        const float u = __fdiv_rn(to_f32<TX>(xrow[d]), dx);
        const float v = __fdiv_rn(__ldg(wrow + d), dw);
        const float diff = v - u;
        if (valid) sse_n = fmaf(diff, diff, sse_n);
      }
    }
  }
  block_sum2(sse, sse_n, sh);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = sse;
    partials[kMaxPartials + blockIdx.x] = sse_n;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) {  // fixed order -> run-to-run deterministic
      a += __ldcg(partials + i);
      b += __ldcg(partials + kMaxPartials + i);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (threadIdx.x == 0) {
      const float inv = 1.f / ((float)N * (float)D);
      mse4[0] = a * inv;
      mse4[1] = a * inv;
      mse4[2] = b * inv;
      mse4[3] = b * inv;
      *ticket = 0u;  // self-reset for the next launch
    }
  }
}

template <typename TG, typename TX, typename TO, int G>
__global__ void __launch_bounds__(256) quantize_backward_kernel(
    const TG* __restrict__ gz, const TX* __restrict__ x, const float* __restrict__ W, int64_t K,
    const int64_t* __restrict__ quant, int64_t N, int D, const float* __restrict__ g4, int want_norm,
    TO* __restrict__ gx, float* __restrict__ gW) {
  const int lane = threadIdx.x % G;
  const int64_t rows_per_block = blockDim.x / G;
  const float scale = 2.f / ((float)N * (float)D);
  const float c_cb = g4[0] * scale, c_cm = g4[1] * scale;
  const float c_cbn = want_norm ? g4[2] * scale : 0.f, c_cmn = want_norm ? g4[3] * scale : 0.f;
  for (int64_t base = blockIdx.x * rows_per_block; base < N; base += (int64_t)gridDim.x * rows_per_block) {
    const int64_t n_raw = base + threadIdx.x / G;  // warp-uniform trip count: shuffles need every lane
    const bool valid = n_raw < N;
    const int64_t n = valid ? n_raw : N - 1;
    int64_t q = quant[n];
    q = q < 0 ? 0 : (q >= K ? K - 1 : q);
    const float* __restrict__ wrow = W + q * D;
    const TX* __restrict__ xrow = x + n * D;
    float dx = 1.f, dw = 1.f, uu = 0.f, vv = 0.f, uv = 0.f;
    bool cx = false, cw = false;
    if (want_norm) {
      float sxx = 0.f, sww = 0.f, sxw = 0.f;
      for (int d = lane; d < D; d += G) {
        const float xv = to_f32<TX>(xrow[d]);
        const float wv = __ldg(wrow + d);
        sxx = fmaf(xv, xv, sxx);
        sww = fmaf(wv, wv, sww);
        sxw = fmaf(xv, wv, sxw);
      }
      sxx = group_sum<G>(sxx);
      sww = group_sum<G>(sww);
      sxw = group_sum<G>(sxw);
      const float nx = sqrtf(sxx), nw = sqrtf(sww);
      cx = nx < kNormEps;
      cw = nw < kNormEps;
      dx = fmaxf(nx, kNormEps);
      dw = fmaxf(nw, kNormEps);
      uu = sxx / (dx * dx);
      vv = sww / (dw * dw);
      uv = sxw / (dx * dw);
    }
    for (int d = lane; d < D; d += G) {
      const float xv = to_f32<TX>(xrow[d]);
      const float wv = __ldg(wrow + d);
      float g_x = to_f32<TG>(gz[n * D + d]) + c_cm * (xv - wv);
      float g_w = c_cb * (wv - xv);
      if (want_norm) {
        const float u = xv / dx, v = wv / dw;
        // J_n(x)^T g_u with g_u = c (u - v):  (g_u - (g_u.u) u) / dx   (projection dropped when clamped)
        const float gu = c_cmn * (u - v);
        const float gu_dot_u = c_cmn * (uu - uv);
        g_x += (gu - (cx ? 0.f : gu_dot_u * u)) / dx;
        const float gv = c_cbn * (v - u);
        const float gv_dot_v = c_cbn * (vv - uv);
        g_w += (gv - (cw ? 0.f : gv_dot_v * v)) / dw;
      }
      if (valid) {
        gx[n * D + d] = from_f32<TO>(g_x);
        if (gW) atomicAdd(gW + q * D + d, g_w);
      }
    }
  }
}

__global__ void unpack_keys_kernel(const unsigned long long* __restrict__ keys, int64_t n, int64_t offset,
                                   int64_t* __restrict__ idx, float* __restrict__ score) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    if (idx) idx[i] = (int64_t)key_index(k) - offset;
    if (score) score[i] = key_score(k);
  }
}

__global__ void embedding_gather_kernel(const float* __restrict__ W, int64_t K, int D,
                                        const int64_t* __restrict__ quant, int64_t n, float* __restrict__ out) {
  const int64_t total = n * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D;
    const int d = (int)(i - r * D);
    int64_t q = quant[r];
    q = q < 0 ? 0 : (q >= K ? K - 1 : q);
    out[i] = __ldg(W + q * D + d);
  }
}

__global__ void keys_flip_kernel(unsigned long long* __restrict__ keys, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    keys[i] ^= 0x8000000000000000ull;
}

static inline int grid_for(int64_t rows, int rows_per_block) {
  int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
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

int64_t vqb_loss_partials_count(void) { return 2 * kMaxPartials; }

int vqb_gather_ste_loss(const void* x, int x_dtype, int64_t N, int D, const float* W, int64_t K,
                        const int64_t* quant, void* z_out, int out_dtype, int want_norm, float* mse4,
                        float* partials, unsigned int* ticket, void* stream) {
  VQB_REQUIRE(x && W && quant && z_out && mse4 && partials && ticket, "vqb_gather_ste_loss: null pointer");
  VQB_REQUIRE(N >= 1 && D >= 1 && K >= 1, "vqb_gather_ste_loss: bad shape N=%lld D=%d K=%lld", (long long)N, D,
              (long long)K);
  cudaStream_t st = (cudaStream_t)stream;
  const int g = lanes_per_row_q(D);
  const int blocks = grid_for(N, 256 / g);
  VQB_REQUIRE(blocks <= kMaxPartials, "vqb_gather_ste_loss: grid too large");
#define LAUNCH(TX, TO)                                                                                          \
  VQB_DISPATCH_G(g, (gather_ste_loss_kernel<TX, TO, G><<<blocks, 256, 0, st>>>(                                 \
                        (const TX*)x, N, D, W, K, quant, (TO*)z_out, want_norm, mse4, partials, ticket)))
  if (x_dtype == VQB_F32 && out_dtype == VQB_F32) { LAUNCH(float, float); }
  else if (x_dtype == VQB_BF16 && out_dtype == VQB_F32) { LAUNCH(__nv_bfloat16, float); }
  else if (x_dtype == VQB_BF16 && out_dtype == VQB_BF16) { LAUNCH(__nv_bfloat16, __nv_bfloat16); }
  else if (x_dtype == VQB_F32 && out_dtype == VQB_BF16) { LAUNCH(float, __nv_bfloat16); }
  else { VQB_REQUIRE(false, "vqb_gather_ste_loss: bad dtype"); }
#undef LAUNCH
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_quantize_backward(const void* gz, int g_dtype, const void* x, int x_dtype, const float* W, int64_t K,
                          const int64_t* quant, int64_t N, int D, const float* g4, int want_norm, void* gx,
                          int gx_dtype, float* gW, void* stream) {
  VQB_REQUIRE(gz && x && W && quant && g4 && gx, "vqb_quantize_backward: null pointer");
  VQB_REQUIRE(N >= 1 && D >= 1 && K >= 1, "vqb_quantize_backward: bad shape");
  VQB_REQUIRE(x_dtype == gx_dtype, "vqb_quantize_backward: gx dtype must equal x dtype");
  cudaStream_t st = (cudaStream_t)stream;
  const int g = lanes_per_row_q(D);
  const int blocks = grid_for(N, 256 / g);
#define LAUNCH(TG, TX)                                                                                   \
  VQB_DISPATCH_G(g, (quantize_backward_kernel<TG, TX, TX, G><<<blocks, 256, 0, st>>>(                    \
                        (const TG*)gz, (const TX*)x, W, K, quant, N, D, g4, want_norm, (TX*)gx, gW)))
  if (g_dtype == VQB_F32 && x_dtype == VQB_F32) { LAUNCH(float, float); }
  else if (g_dtype == VQB_F32 && x_dtype == VQB_BF16) { LAUNCH(float, __nv_bfloat16); }
  else if (g_dtype == VQB_BF16 && x_dtype == VQB_BF16) { LAUNCH(__nv_bfloat16, __nv_bfloat16); }
  else if (g_dtype == VQB_BF16 && x_dtype == VQB_F32) { LAUNCH(__nv_bfloat16, float); }
  else { VQB_REQUIRE(false, "vqb_quantize_backward: bad dtype"); }
#undef LAUNCH
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_unpack_keys(const unsigned long long* keys, int64_t n, int64_t offset, int64_t* idx, float* score,
                    void* stream) {
  VQB_REQUIRE(keys && (idx || score), "vqb_unpack_keys: null pointer");
  if (n <= 0) return VQB_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  unpack_keys_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(keys, n, offset, idx, score);
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_embedding_gather(const float* W, int64_t K, int D, const int64_t* quant, int64_t n, float* out,
                         void* stream) {
  VQB_REQUIRE(W && quant && out, "vqb_embedding_gather: null pointer");
  VQB_REQUIRE(K >= 1 && D >= 1, "vqb_embedding_gather: bad shape");
  if (n <= 0) return VQB_OK;
  int64_t blocks64 = (n * D + 255) / 256;
  const int blocks = (int)(blocks64 < (int64_t)sm_count() * 8 ? blocks64 : (int64_t)sm_count() * 8);
  embedding_gather_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, K, D, quant, n, out);
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_keys_flip_sign(unsigned long long* keys, int64_t n, void* stream) {
  VQB_REQUIRE(keys, "vqb_keys_flip_sign: null pointer");
  if (n <= 0) return VQB_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  keys_flip_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(keys, n);
  VQB_LAUNCH_OK();
  return VQB_OK;
}

}  // extern "C"
