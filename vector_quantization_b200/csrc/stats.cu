// Usage / EMA statistics: warp-aggregated scatter-add of per-code counts and feature sums, the
// k-means + EMA codebook update (VQ-KD) and the CVQ-VAE anchor blend.  HBM / L2-atomic bound.
#include <math.h>

#include "common.cuh"

namespace vqb {

// ---- scatter: stats[q*D + d] += x[n,d] ; stats[K*D + q] += 1 ---------------------------------
// G lanes cooperate on one token row (V contiguous elements per lane per step).  Rows of the same warp
// that hit the same code are first combined with shuffles ("warp-aggregated"), so a burst of equal
// indices (low codebook usage) costs one vector atomic instead of 32/G.
template <typename TX, int G, int V>
__global__ void __launch_bounds__(256) scatter_stats_kernel(const TX* __restrict__ x, int64_t N, int D,
                                                            int normalize_x, const int64_t* __restrict__ quant,
                                                            const unsigned long long* __restrict__ keys, int64_t key_offset,
                                                            float* __restrict__ stats, int64_t K) {
  pdl_wait();               // PDL: inputs come from the preceding launches
  pdl_launch_dependents();
  const int lane_in_warp = threadIdx.x & 31;
  const int lane = threadIdx.x % G;
  const int64_t rows_per_block = blockDim.x / G;
  float* __restrict__ counts = stats + K * (int64_t)D;
  constexpr int GROUPS = 32 / G;
  // lanes of my warp that serve the same element slice (same lane-in-group)
  unsigned sublane_mask = 0;
#pragma unroll
  for (int j = 0; j < GROUPS; ++j) sublane_mask |= 1u << (j * G + lane);

  for (int64_t base = blockIdx.x * rows_per_block; base < N; base += (int64_t)gridDim.x * rows_per_block) {
    const int64_t n_raw = base + threadIdx.x / G;
    const bool valid = n_raw < N;
    const int64_t n = valid ? n_raw : N - 1;
    int64_t q64;
    if (quant) {
      q64 = quant[n];
    } else {   // packed keys of vqb_assign: no finite score (NaN token) -> index 0, like torch.argmin
      const unsigned long long kk = keys[n];
      q64 = kk == kNoKey ? 0 : (int64_t)key_index(kk) - key_offset;
    }
    const bool inrange = valid && q64 >= 0 && q64 < K;
    const int q = inrange ? (int)q64 : -1 - (lane_in_warp / G);  // unique negative id: never aggregated
    const TX* __restrict__ xrow = x + n * D;

    float inv = 1.f;
    if (normalize_x) {
      float ss = 0.f;
      for (int d = lane * V; d < D; d += G * V)
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float t = to_f32<TX>(xrow[d + v]);
          ss = fmaf(t, t, ss);
        }
      ss = group_sum<G>(ss);
      inv = 1.f / fmaxf(sqrtf(ss), kNormEps);
    }

    const unsigned same = __match_any_sync(0xffffffffu, q);      // lanes whose row hits my code
    const unsigned peers = same & sublane_mask;                    // ... and serve my element slice
    const int leader_lane = __ffs(peers) - 1;
    const bool leader = leader_lane == lane_in_warp;
    const bool any_dup = __any_sync(0xffffffffu, __popc(same) > G);

    if (inrange && leader && lane == 0) atomicAdd(counts + q, (float)(__popc(same) / G));

    for (int d0 = lane * V; d0 < round_up(D, G * V); d0 += G * V) {  // warp-uniform trip count
      float v[V];
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = (d0 + i < D) ? to_f32<TX>(xrow[d0 + i]) * inv : 0.f;
      if (any_dup) {
#pragma unroll
        for (int j = 0; j < GROUPS; ++j) {
          const int src = j * G + lane;
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float o = __shfl_sync(0xffffffffu, v[i], src);
            if (leader && src != lane_in_warp && ((peers >> src) & 1u)) v[i] += o;
          }
        }
      }
      if (inrange && leader && d0 < D) {
        float* dst = stats + (int64_t)q * D + d0;
        if constexpr (V == 4) {
          atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));  // red.global.add.v4.f32
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) atomicAdd(dst + i, v[i]);
        }
      }
    }
  }
}

__global__ void bincount_kernel(const int64_t* __restrict__ quant, int64_t n, unsigned long long* __restrict__ counts,
                                int64_t K, int add_total) {
  pdl_wait();               // PDL: inputs come from the preceding launches
  pdl_launch_dependents();
  if (add_total && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(counts + K, (unsigned long long)n);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < round_up(n, 32);
       i += (int64_t)gridDim.x * blockDim.x) {
    const bool ok = i < n;
    const int64_t q = ok ? quant[i] : -1;
    const bool inrange = ok && q >= 0 && q < K;
    const int key = inrange ? (int)q : -1 - (int)(threadIdx.x & 31);
    const unsigned same = __match_any_sync(0xffffffffu, key);
    if (inrange && (__ffs(same) - 1) == (int)(threadIdx.x & 31)) atomicAdd(counts + q, (unsigned long long)__popc(same));
  }
}

// ---- VQ-KD: centroid / where / normalise / EMA / normalise ------------------------------------
template <int G>
__global__ void __launch_bounds__(256) kmeans_ema_kernel(const float* __restrict__ stats, float* __restrict__ W,
                                                         int64_t K, int D, float decay, float omd) {
  pdl_wait();               // PDL: inputs come from the preceding launches
  pdl_launch_dependents();
  const int lane = threadIdx.x % G;
  const int64_t rows_per_block = blockDim.x / G;
  const float* __restrict__ counts = stats + K * (int64_t)D;
  for (int64_t base = blockIdx.x * rows_per_block; base < K; base += (int64_t)gridDim.x * rows_per_block) {
    const int64_t k_raw = base + threadIdx.x / G;
    const bool valid = k_raw < K;
    const int64_t k = valid ? k_raw : K - 1;
    const float cnt = counts[k];
    const bool occurred = cnt > 0.f;
    const float den = fmaxf(cnt, 1.f);
    const float* __restrict__ srow = stats + k * D;
    float* __restrict__ wrow = W + k * D;
    float ss = 0.f;
    for (int d = lane; d < D; d += G) {
      const float c = occurred ? __fdiv_rn(srow[d], den) : wrow[d];
      ss = fmaf(c, c, ss);
    }
    ss = group_sum<G>(ss);
    const float dn = fmaxf(sqrtf(ss), kNormEps);
    float ss2 = 0.f;
    for (int d = lane; d < D; d += G) {
      const float c = occurred ? __fdiv_rn(srow[d], den) : wrow[d];
      const float e = __fadd_rn(__fmul_rn(wrow[d], decay), __fmul_rn(__fdiv_rn(c, dn), omd));
      ss2 = fmaf(e, e, ss2);
    }
    ss2 = group_sum<G>(ss2);
    const float dn2 = fmaxf(sqrtf(ss2), kNormEps);
    for (int d = lane; valid && d < D; d += G) {
      const float c = occurred ? __fdiv_rn(srow[d], den) : wrow[d];
      const float e = __fadd_rn(__fmul_rn(wrow[d], decay), __fmul_rn(__fdiv_rn(c, dn), omd));
      wrow[d] = __fdiv_rn(e, dn2);
    }
  }
}

template <typename TX>
__global__ void gather_rows_by_key_kernel(const TX* __restrict__ x, int64_t N, int D,
                                          const unsigned long long* __restrict__ keys, int64_t K, int64_t offset,
                                          float* __restrict__ out) {
  pdl_wait();               // PDL: inputs come from the preceding launches
  pdl_launch_dependents();
  const int64_t total = K * (int64_t)D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / D;
    const int d = (int)(i - k * D);
    const int64_t n = (int64_t)key_index(keys[k]) - offset;
    out[i] = (n >= 0 && n < N) ? to_f32<TX>(x[n * D + d]) : 0.f;
  }
}

__global__ void cvq_update_kernel(float* __restrict__ W, const float* __restrict__ anchors, float anchor_scale,
                                  float* __restrict__ prob, const int64_t* __restrict__ counts,
                                  const int64_t* __restrict__ total_ptr, int64_t K, int D, float decay, float omd,
                                  float eps) {
  pdl_wait();               // PDL: inputs come from the preceding launches
  pdl_launch_dependents();
  const float total = (float)*total_ptr;  // int64 -> fp32, then a true division like `bin_count / numel`
  // one warp per code row; lane 0 owns the probability update
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp; k < K; k += nwarps) {
    const float freq = __fdiv_rn((float)counts[k], total);
    const float p = __fadd_rn(__fmul_rn(prob[k], decay), __fmul_rn(freq, omd));
    // decay_k = 1 - exp(-p*K*10/(1-decay) - eps)      cvqvae/quantizer_callback.py:98-101
    float t = __fmul_rn(__fmul_rn(-p, (float)K), 10.f);
    t = __fsub_rn(__fdiv_rn(t, omd), eps);
    const float dec = __fsub_rn(1.f, expf(t));
    const float omdec = __fsub_rn(1.f, dec);
    __syncwarp();
    if (lane == 0) prob[k] = p;
    if (omdec == 0.f) continue;   // dec == 1 exactly: W*1 + anchor*0 leaves the row unchanged bit for bit (see comm.cu)
    for (int d = lane; d < D; d += 32) {
      const float a = anchors[k * D + d] * anchor_scale;
      W[k * D + d] = __fadd_rn(__fmul_rn(W[k * D + d], dec), __fmul_rn(a, omdec));
    }
  }
}

static inline int pow2_lanes(int n) {
  int g = 1;
  while (g < 32 && g < n) g <<= 1;
  return g;
}
static inline int grid_rows(int64_t rows, int rows_per_block) {
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

int vqb_scatter_stats(const void* x, int x_dtype, int64_t N, int D, int normalize_x, const int64_t* quant,
                      const unsigned long long* keys, int64_t key_index_offset, float* stats, int64_t K, void* stream) {
  VQB_REQUIRE(x && stats, "vqb_scatter_stats: null pointer");
  VQB_REQUIRE((quant != nullptr) != (keys != nullptr), "vqb_scatter_stats: pass exactly one of quant / keys");
  VQB_REQUIRE(N >= 1 && D >= 1 && K >= 1, "vqb_scatter_stats: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (D % 4 == 0);
  const int g = pow2_lanes(vec ? D / 4 : (D + 3) / 4);
  const int blocks = grid_rows(N, 256 / g);
#define LAUNCH(TX)                                                                                              \
  if (vec) {                                                                                                    \
    VQB_DISPATCH_G(g, (launch_pdl(scatter_stats_kernel<TX, G, 4>, blocks, 256, 0, st, (const TX*)x, N, D, normalize_x,  \
                                                                               quant, keys, key_index_offset, stats, K)));              \
  } else {                                                                                                      \
    VQB_DISPATCH_G(g, (launch_pdl(scatter_stats_kernel<TX, G, 1>, blocks, 256, 0, st, (const TX*)x, N, D, normalize_x,  \
                                                                               quant, keys, key_index_offset, stats, K)));              \
  }
  if (x_dtype == VQB_F32) { LAUNCH(float) }
  else if (x_dtype == VQB_BF16) { LAUNCH(__nv_bfloat16) }
  else VQB_REQUIRE(false, "vqb_scatter_stats: bad dtype");
#undef LAUNCH
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_bincount_accumulate(const int64_t* quant, int64_t n, int64_t* counts, int64_t K, int add_total,
                            void* stream) {
  VQB_REQUIRE(quant && counts, "vqb_bincount_accumulate: null pointer");
  if (n <= 0) return VQB_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  launch_pdl(bincount_kernel, blocks, 256, 0, (cudaStream_t)stream, quant, n, (unsigned long long*)counts, K, add_total);
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_kmeans_ema_update(const float* stats, float* W, int64_t K, int D, float decay, float one_minus_decay,
                          void* stream) {
  VQB_REQUIRE(stats && W, "vqb_kmeans_ema_update: null pointer");
  VQB_REQUIRE(K >= 1 && D >= 1, "vqb_kmeans_ema_update: bad shape");
  const int g = pow2_lanes((D + 3) / 4);
  const int blocks = grid_rows(K, 256 / g);
  VQB_DISPATCH_G(g, (launch_pdl(kmeans_ema_kernel<G>, blocks, 256, 0, (cudaStream_t)stream, stats, W, K, D, decay,
                                                                                    one_minus_decay)));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_gather_rows_by_key(const void* x, int x_dtype, int64_t N, int D, const unsigned long long* keys, int64_t K,
                           int64_t offset, float* out, void* stream) {
  VQB_REQUIRE(x && keys && out, "vqb_gather_rows_by_key: null pointer");
  VQB_REQUIRE(K >= 1 && D >= 1, "vqb_gather_rows_by_key: bad shape");
  int64_t total = K * (int64_t)D;
  int blocks = (int)((total + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  if (x_dtype == VQB_F32)
    launch_pdl(gather_rows_by_key_kernel<float>, blocks, 256, 0, (cudaStream_t)stream, (const float*)x, N, D, keys, K, offset, out);
  else if (x_dtype == VQB_BF16)
    launch_pdl(gather_rows_by_key_kernel<__nv_bfloat16>, blocks, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, N, D, keys, K, offset, out);
  else
    VQB_REQUIRE(false, "vqb_gather_rows_by_key: bad dtype");
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_cvq_update(float* W, const float* anchors, float anchor_scale, float* prob, const int64_t* counts,
                   const int64_t* total, int64_t K, int D, float decay, float one_minus_decay, float eps,
                   void* stream) {
  VQB_REQUIRE(W && anchors && prob && counts && total, "vqb_cvq_update: null pointer");
  VQB_REQUIRE(K >= 1 && D >= 1, "vqb_cvq_update: bad shape");
  int blocks = (int)((K * 32 + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  launch_pdl(cvq_update_kernel, blocks, 256, 0, (cudaStream_t)stream, W, anchors, anchor_scale, prob, counts, total, K, D,
                                                              decay, one_minus_decay, eps);
  VQB_LAUNCH_OK();
  return VQB_OK;
}

}  // extern "C"
