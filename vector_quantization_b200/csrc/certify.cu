// Certified one-term assignment: the bookkeeping kernels around vqb_assign_ex.
// For D >= 128 the assignment is bound by the tensor pipe and an fp32 codebook costs TWO MMA terms as an fp16
// (hi, lo * 2^11) pair.  The hi plane alone gives every score to within eps = |x| * max_j |e_j - hi_j|; a row whose
// best and runner-up scores are more than 2 eps apart has provably the same arg-max as the two-term contraction.
// Rows that fail the test are collected here, re-run exactly (two terms) as a compact operand, and scattered back.
#include "common.cuh"

namespace vqb {

// Flag pass: one thread per row; a 1024-row block writes its 32 ballot words and its number of flagged rows.
constexpr int kCertBlock = 1024;

__global__ void __launch_bounds__(kCertBlock) certify_flag_kernel(
    const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ second, int64_t rows,
    const float* __restrict__ row_inv_norm, const float* __restrict__ delta, float noise, uint32_t* __restrict__ ballots,
    int* __restrict__ block_counts) {
  __shared__ int s_count;
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  const float d = *delta + noise;
  const int64_t r = blockIdx.x * (int64_t)kCertBlock + threadIdx.x;
  bool flag = false;
  if (r < rows) {
    const unsigned long long k1 = keys[r], k2 = second[r];
    if (k1 != kNoKey && k2 != kNoKey) {                // no finite score (NaN row) / a single candidate: nothing to prove
      const float s = row_inv_norm ? 1.f / row_inv_norm[r] : 1.f;
      flag = !(key_score(k1) - key_score(k2) > 2.f * d * s);   // NaN margins are flagged as well
    }
  }
  const uint32_t word = __ballot_sync(0xffffffffu, flag);
  if ((threadIdx.x & 31) == 0) {
    ballots[blockIdx.x * (kCertBlock / 32) + (threadIdx.x >> 5)] = word;
    atomicAdd(&s_count, __popc(word));
  }
  __syncthreads();
  if (threadIdx.x == 0) block_counts[blockIdx.x] = s_count;
}

// CVQ-VAE: which codes can receive a non-zero anchor weight this step?  (cvqvae/quantizer_callback.py:94-102)
//   p = ema(p_old, count / total);  decay_k = 1 - exp(-p*K*10/(1-gamma) - eps);  W_k <- W_k*decay_k + anchor_k*(1 - decay_k)
// In fp32 decay_k rounds to exactly 1 once p*K*10/(1-gamma) exceeds ~17.3, i.e. for every code used at more than ~2 % of
// the uniform rate; the anchor of such a code is multiplied by exactly 0.  The flag is evaluated on a LOWER bound of
// p (this rank's count over the global token total, times 0.999): p can only be larger once the counts of the other
// ranks are in, and 1 - decay_k is non-increasing in p, so every code whose true weight is non-zero is flagged on every
// rank (a superset, never a subset).  Only the flagged codes need the column arg-min and an anchor row.
__global__ void __launch_bounds__(kCertBlock) cvq_needy_flag_kernel(
    const float* __restrict__ prob, const int64_t* __restrict__ counts_local, float total_global, int64_t K, float decay,
    float omd, float eps, uint32_t* __restrict__ ballots, int* __restrict__ block_counts) {
  __shared__ int s_count;
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  const int64_t k = blockIdx.x * (int64_t)kCertBlock + threadIdx.x;
  bool flag = false;
  if (k < K) {
    const float freq_lb = __fdiv_rn((float)counts_local[k], total_global);
    const float p = (__fadd_rn(__fmul_rn(prob[k], decay), __fmul_rn(freq_lb, omd))) * 0.999f;
    float t = __fmul_rn(__fmul_rn(-p, (float)K), 10.f);
    t = __fsub_rn(__fdiv_rn(t, omd), eps);
    const float dec = __fsub_rn(1.f, expf(t));
    flag = !(__fsub_rn(1.f, dec) == 0.f);          // NaN probabilities are flagged as well
  }
  const uint32_t word = __ballot_sync(0xffffffffu, flag);
  if ((threadIdx.x & 31) == 0) {
    ballots[blockIdx.x * (kCertBlock / 32) + (threadIdx.x >> 5)] = word;
    atomicAdd(&s_count, __popc(word));
  }
  __syncthreads();
  if (threadIdx.x == 0) block_counts[blockIdx.x] = s_count;
}

// Compaction pass: ascending row order (deterministic, identical on every rank of a sharded run).
__global__ void __launch_bounds__(kCertBlock) certify_compact_kernel(
    const uint32_t* __restrict__ ballots, const int* __restrict__ block_counts, int nblocks, int* __restrict__ row_list,
    int* __restrict__ count, unsigned long long* __restrict__ compact_keys) {
  __shared__ int s_base, s_warp[kCertBlock / 32];
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {   // rows flagged in the blocks before mine (and the total, published by the last block)
    int before = 0, total = 0;
    for (int b = lane; b < nblocks; b += 32) {
      const int c = block_counts[b];
      total += c;
      if (b < (int)blockIdx.x) before += c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      before += __shfl_xor_sync(0xffffffffu, before, o);
      total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if (lane == 0) {
      s_base = before;
      if (blockIdx.x == gridDim.x - 1) *count = total;
    }
  }
  const uint32_t word = ballots[blockIdx.x * (kCertBlock / 32) + warp];
  if (lane == 0) s_warp[warp] = __popc(word);
  __syncthreads();
  int off = s_base;
  for (int w = 0; w < warp; ++w) off += s_warp[w];
  if ((word >> lane) & 1u) {
    const int slot = off + __popc(word & ((1u << lane) - 1u));
    row_list[slot] = (int)(blockIdx.x * kCertBlock + threadIdx.x);
    compact_keys[slot] = kNoKey;
  }
}

__global__ void gather_plane_rows_kernel(const uint4* __restrict__ src, int nplanes, int64_t src_plane_rows, int vecs,
                                         const int* __restrict__ row_list, const int* __restrict__ count, int64_t cap,
                                         uint4* __restrict__ dst, int64_t dst_plane_rows) {
  pdl_wait();
  pdl_launch_dependents();
  int64_t n = *count;
  if (n > cap) n = cap;
  const int64_t total = n * nplanes * vecs;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    const int64_t t = i / vecs;
    const int p = (int)(t % nplanes);
    const int64_t slot = t / nplanes;
    dst[((int64_t)p * dst_plane_rows + slot) * vecs + v] = src[((int64_t)p * src_plane_rows + row_list[slot]) * vecs + v];
  }
}

__global__ void gather_f32_kernel(const float* __restrict__ src, const int* __restrict__ row_list,
                                  const int* __restrict__ count, int64_t cap, float* __restrict__ dst) {
  pdl_wait();
  pdl_launch_dependents();
  int64_t n = *count;
  if (n > cap) n = cap;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[row_list[i]];
}

__global__ void scatter_keys_kernel(const unsigned long long* __restrict__ compact, const int* __restrict__ row_list,
                                    const int* __restrict__ count, int64_t cap, unsigned long long* __restrict__ keys) {
  pdl_wait();
  pdl_launch_dependents();
  int64_t n = *count;
  if (n > cap) n = cap;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    keys[row_list[i]] = compact[i];
}

static inline int blocks_1d(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace vqb

using namespace vqb;

extern "C" {

int64_t vqb_certify_workspace_bytes(int64_t rows) {
  const int64_t nblocks = (rows + kCertBlock - 1) / kCertBlock;
  return nblocks * (kCertBlock / 32) * 4 + nblocks * 4;
}

int vqb_certify(const unsigned long long* keys, const unsigned long long* second_keys, int64_t rows,
                const float* row_inv_norm, const float* delta, float noise, int* row_list, int* count,
                unsigned long long* compact_keys, void* workspace, void* stream) {
  VQB_REQUIRE(keys && second_keys && delta && row_list && count && compact_keys && workspace, "vqb_certify: null pointer");
  VQB_REQUIRE(rows >= 1 && rows < (1ll << 31), "vqb_certify: bad row count");
  const int nblocks = (int)((rows + kCertBlock - 1) / kCertBlock);
  uint32_t* ballots = static_cast<uint32_t*>(workspace);
  int* block_counts = reinterpret_cast<int*>(ballots + (size_t)nblocks * (kCertBlock / 32));
  cudaStream_t st = (cudaStream_t)stream;
  VQB_CUDA_OK(launch_pdl(certify_flag_kernel, nblocks, kCertBlock, 0, st, keys, second_keys, rows, row_inv_norm, delta, noise,
                         ballots, block_counts));
  VQB_CUDA_OK(launch_pdl(certify_compact_kernel, nblocks, kCertBlock, 0, st, (const uint32_t*)ballots,
                         (const int*)block_counts, nblocks, row_list, count, compact_keys));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_cvq_needy_codes(const float* prob, const int64_t* counts_local, float total_global, int64_t K, float decay,
                        float one_minus_decay, float eps, int* code_list, int* count, unsigned long long* compact_keys,
                        void* workspace, void* stream) {
  VQB_REQUIRE(prob && counts_local && code_list && count && compact_keys && workspace, "vqb_cvq_needy_codes: null pointer");
  VQB_REQUIRE(K >= 1 && K < (1ll << 31) && total_global > 0.f, "vqb_cvq_needy_codes: bad arguments");
  const int nblocks = (int)((K + kCertBlock - 1) / kCertBlock);
  uint32_t* ballots = static_cast<uint32_t*>(workspace);
  int* block_counts = reinterpret_cast<int*>(ballots + (size_t)nblocks * (kCertBlock / 32));
  cudaStream_t st = (cudaStream_t)stream;
  VQB_CUDA_OK(launch_pdl(cvq_needy_flag_kernel, nblocks, kCertBlock, 0, st, prob, counts_local, total_global, K, decay,
                         one_minus_decay, eps, ballots, block_counts));
  VQB_CUDA_OK(launch_pdl(certify_compact_kernel, nblocks, kCertBlock, 0, st, (const uint32_t*)ballots,
                         (const int*)block_counts, nblocks, code_list, count, compact_keys));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_gather_plane_rows(const void* src, int nplanes, int64_t src_plane_rows, int Dp, const int* row_list,
                          const int* count, int64_t cap, void* dst, int64_t dst_plane_rows, void* stream) {
  VQB_REQUIRE(src && row_list && count && dst, "vqb_gather_plane_rows: null pointer");
  VQB_REQUIRE(nplanes >= 1 && nplanes <= 3 && Dp >= 8 && Dp % 8 == 0 && cap >= 1, "vqb_gather_plane_rows: bad shape");
  // the grid is sized for a modest fraction of the rows; the grid-stride loop covers the rest
  VQB_CUDA_OK(launch_pdl(gather_plane_rows_kernel, blocks_1d(cap * nplanes * (Dp / 8) / 8 + 1), 256, 0, (cudaStream_t)stream,
                         (const uint4*)src, nplanes, src_plane_rows, Dp / 8, row_list, count, cap, (uint4*)dst, dst_plane_rows));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_gather_f32(const float* src, const int* row_list, const int* count, int64_t cap, float* dst, void* stream) {
  VQB_REQUIRE(src && row_list && count && dst && cap >= 1, "vqb_gather_f32: bad arguments");
  VQB_CUDA_OK(launch_pdl(gather_f32_kernel, blocks_1d(cap / 8 + 1), 256, 0, (cudaStream_t)stream, src, row_list, count, cap, dst));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_scatter_keys(const unsigned long long* compact_keys, const int* row_list, const int* count, int64_t cap,
                     unsigned long long* keys, void* stream) {
  VQB_REQUIRE(compact_keys && row_list && count && keys && cap >= 1, "vqb_scatter_keys: bad arguments");
  VQB_CUDA_OK(launch_pdl(scatter_keys_kernel, blocks_1d(cap / 8 + 1), 256, 0, (cudaStream_t)stream, compact_keys, row_list, count,
                         cap, keys));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

}  // extern "C"
