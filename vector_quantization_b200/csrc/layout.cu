// Data formats either side of the quantizer:
//   * the tokenizer model's `b c h w -> (b h w) c` / `(b h w) c -> b c h w` rearranges
//     (vq/tasks/image_tokenization/models/base.py:124,126-127) as one tiled, coalesced batched transpose;
//   * the tokenise-only output: packed assignment keys -> compact token ids (uint16 for K <= 65536, else
//     int32) for the token dumps (runners/callbacks.py:40-53, tools/tokenize_llamagen.py:93-103).
// Both are HBM-bound single passes.
#include "common.cuh"

namespace vqb {

constexpr int kTile = 32;

// src [batch][rows][cols] -> dst [batch][cols][rows]; 32 x 32 tile through padded shared memory, both the
// global read (along cols) and the global write (along rows) are coalesced.
template <typename T>
__global__ void __launch_bounds__(kTile * 8) transpose_last2_kernel(const T* __restrict__ src, int64_t rows,
                                                                    int64_t cols, T* __restrict__ dst) {
  __shared__ T tile[kTile][kTile + 1];
  const int64_t b = blockIdx.z;
  const T* __restrict__ s = src + b * rows * cols;
  T* __restrict__ d = dst + b * rows * cols;
  const int64_t c0 = (int64_t)blockIdx.x * kTile, r0 = (int64_t)blockIdx.y * kTile;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int i = ty; i < kTile; i += 8) {
    const int64_t r = r0 + i, c = c0 + tx;
    if (r < rows && c < cols) tile[i][tx] = s[r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < kTile; i += 8) {
    const int64_t c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) d[c * rows + r] = tile[tx][i];
  }
}

// Vectorised variant (cols and rows multiples of the 16-byte vector): a TR (rows) x TC (cols) tile is read
// with 128-bit loads along cols, transposed through shared memory, and written with 128-bit stores along rows.
template <typename T, int TR, int TC>
__global__ void __launch_bounds__(256) transpose_last2_vec_kernel(const T* __restrict__ src, int64_t rows, int64_t cols,
                                                                  T* __restrict__ dst) {
  constexpr int VE = 16 / sizeof(T);          // elements per 128-bit vector
  constexpr int PITCH = TC + 4 / sizeof(T) + (sizeof(T) == 8 ? 1 : 0);   // +1 bank word: conflict-free column reads
  __shared__ T tile[TR][PITCH];
  const int64_t b = blockIdx.z;
  const T* __restrict__ s = src + b * rows * cols;
  T* __restrict__ d = dst + b * rows * cols;
  const int64_t r0 = (int64_t)blockIdx.y * TR;
  constexpr int VPR = TC / VE;
  constexpr int VPC = TR / VE;
  // a block walks the column tiles of its row band (gridDim.x < number of column tiles when there are plenty of
  // images): a 4-8 KB tile per block left the kernel bound by block scheduling, not by bandwidth
  for (int64_t c0 = (int64_t)blockIdx.x * TC; c0 < cols; c0 += (int64_t)gridDim.x * TC) {
    // load: (TC / VE) vectors per row
    for (int v = threadIdx.x; v < TR * VPR; v += 256) {
      const int r = v / VPR, cv = (v % VPR) * VE;
      if (r0 + r < rows && c0 + cv < cols) {
        const uint4 raw = *reinterpret_cast<const uint4*>(s + (r0 + r) * cols + c0 + cv);
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int i = 0; i < VE; ++i) tile[r][cv + i] = e[i];
      }
    }
    __syncthreads();
    // store: output row = source column c, TR contiguous elements along r -> TR / VE vectors per output row
    for (int v = threadIdx.x; v < TC * VPC; v += 256) {
      // consecutive lanes -> consecutive vectors of ONE output row (then the next row): full 64-128 B segments per
      // row instead of 16-byte pieces strided by the row pitch; the padded pitch keeps the shared reads conflict-free
      // (long output rows, VPC > 8, keep the column-fastest order: their shared reads would conflict 4-way)
      const int rv = (VPC <= 8 ? v % VPC : v / TC) * VE, c = VPC <= 8 ? v / VPC : v % TC;
      if (c0 + c < cols && r0 + rv < rows) {
        uint4 raw;
        T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
        for (int i = 0; i < VE; ++i) e[i] = tile[rv + i][c];
        *reinterpret_cast<uint4*>(d + (c0 + c) * rows + r0 + rv) = raw;
      }
    }
    __syncthreads();
  }
}

template <typename TO>
__global__ void compact_tokens_kernel(const unsigned long long* __restrict__ keys, int64_t n, int64_t offset,
                                      TO* __restrict__ out) {
  // 8 keys per thread: four 128-bit loads, 16 (uint16) or 32 (int32) bytes stored contiguously
  const int64_t n8 = n / 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    ulonglong2 k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = reinterpret_cast<const ulonglong2*>(keys)[i * 4 + j];
    TO v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] = (TO)(k[j].x == kNoKey ? 0 : (int64_t)key_index(k[j].x) - offset);
      v[2 * j + 1] = (TO)(k[j].y == kNoKey ? 0 : (int64_t)key_index(k[j].y) - offset);
    }
    if constexpr (sizeof(TO) == 2) {
      reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<const uint4*>(v);
    } else {
      reinterpret_cast<uint4*>(out)[2 * i] = *reinterpret_cast<const uint4*>(v);
      reinterpret_cast<uint4*>(out)[2 * i + 1] = *reinterpret_cast<const uint4*>(v + 4);
    }
  }
  for (int64_t i = n8 * 8 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    out[i] = (TO)(k == kNoKey ? 0 : (int64_t)key_index(k) - offset);
  }
}

}  // namespace vqb

using namespace vqb;

extern "C" {

int vqb_transpose_last2(const void* src, int elem_bytes, int64_t batch, int64_t rows, int64_t cols, void* dst,
                        void* stream) {
  VQB_REQUIRE(src && dst, "vqb_transpose_last2: null pointer");
  VQB_REQUIRE(src != dst, "vqb_transpose_last2: in-place transpose is not supported");
  VQB_REQUIRE(elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8, "vqb_transpose_last2: elem_bytes must be 2, 4 or 8");
  VQB_REQUIRE(batch >= 0 && rows >= 0 && cols >= 0, "vqb_transpose_last2: bad shape");
  if (batch == 0 || rows == 0 || cols == 0) return VQB_OK;
  VQB_REQUIRE(batch <= 65535 && (rows + 15) / 16 <= 65535, "vqb_transpose_last2: grid limit (batch, rows/16 <= 65535)");
  cudaStream_t st = (cudaStream_t)stream;
  const int ve = 16 / elem_bytes;
  if (rows % ve == 0 && cols % ve == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0) {
    // tile shape follows the matrix: a narrow side (channels = 8 or 16) gets a narrow tile
#define VQB_TR_CASE(T_, TR_, TC_)                                                                              \
    do {                                                                                                       \
      const int64_t ct = (cols + TC_ - 1) / TC_, rt = (rows + TR_ - 1) / TR_;                                   \
      int64_t gx = ((int64_t)sm_count() * 16 + rt * batch - 1) / (rt * batch); /* enough blocks for ~2 waves */ \
      if (gx > ct) gx = ct;                                                                                    \
      if (gx < 1) gx = 1;                                                                                      \
      const dim3 vgrid((unsigned)gx, (unsigned)rt, (unsigned)batch);                                           \
      transpose_last2_vec_kernel<T_, TR_, TC_><<<vgrid, 256, 0, st>>>((const T_*)src, rows, cols, (T_*)dst);   \
    } while (0)
#define VQB_TR_SHAPE(T_)                                      \
    do {                                                      \
      if (cols <= 16) VQB_TR_CASE(T_, 128, 16);               \
      else if (rows <= 16) VQB_TR_CASE(T_, 16, 128);          \
      else if (cols <= 32) VQB_TR_CASE(T_, 64, 32);           \
      else VQB_TR_CASE(T_, 32, 64);                           \
    } while (0)
    if (elem_bytes == 2) VQB_TR_SHAPE(uint16_t);
    else if (elem_bytes == 4) VQB_TR_SHAPE(uint32_t);
    else VQB_TR_SHAPE(unsigned long long);
#undef VQB_TR_SHAPE
#undef VQB_TR_CASE
    VQB_LAUNCH_OK();
    return VQB_OK;
  }
  const dim3 grid((unsigned)((cols + kTile - 1) / kTile), (unsigned)((rows + kTile - 1) / kTile), (unsigned)batch);
  const dim3 block(kTile, 8);
  if (elem_bytes == 2)
    transpose_last2_kernel<uint16_t><<<grid, block, 0, st>>>((const uint16_t*)src, rows, cols, (uint16_t*)dst);
  else if (elem_bytes == 4)
    transpose_last2_kernel<uint32_t><<<grid, block, 0, st>>>((const uint32_t*)src, rows, cols, (uint32_t*)dst);
  else
    transpose_last2_kernel<unsigned long long><<<grid, block, 0, st>>>((const unsigned long long*)src, rows, cols,
                                                                        (unsigned long long*)dst);
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_compact_tokens(const unsigned long long* keys, int64_t n, int64_t index_offset, void* out, int out_bytes,
                       void* stream) {
  VQB_REQUIRE(keys && out, "vqb_compact_tokens: null pointer");
  VQB_REQUIRE(out_bytes == 2 || out_bytes == 4, "vqb_compact_tokens: out_bytes must be 2 (uint16) or 4 (int32)");
  if (n <= 0) return VQB_OK;
  VQB_REQUIRE((uintptr_t)keys % 16 == 0 && (uintptr_t)out % 16 == 0, "vqb_compact_tokens: 16-byte aligned buffers");
  int blocks = (int)((n / 8 + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_bytes == 2) compact_tokens_kernel<uint16_t><<<blocks, 256, 0, st>>>(keys, n, index_offset, (uint16_t*)out);
  else compact_tokens_kernel<int32_t><<<blocks, 256, 0, st>>>(keys, n, index_offset, (int32_t*)out);
  VQB_LAUNCH_OK();
  return VQB_OK;
}

}  // extern "C"
