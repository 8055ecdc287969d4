// Nearest-code assignment on the 5th-gen tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, hardware swizzle) -> shared-memory ring -> tcgen05.mma kind::f16
//   (bf16 x bf16 -> fp32 accumulators in TMEM, 128 x 256 tile) -> fused arg-max epilogue read with
//   tcgen05.ld (TMEM lane == operand row, so the row arg-max is thread-local) -> 64-bit atomicMin
//   of packed (score, index) keys.  The [rows x codes] score matrix never exists in memory.
//
// The contraction dimension is "virtual": an fp32 operand is an exact sum of bf16 planes, and the
// kernel accumulates the selected plane-pair terms (a_p . b_q) into the same TMEM accumulator, i.e.
// K_virtual = n_terms * Dp.  Persistent CTAs own a contiguous range of (row-tile, code-tile) work
// items so the grid is exactly one CTA per SM with balanced work.
//
// Warp roles (320 threads):  warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
//                            warps 2..9 = epilogue (warp%4 selects the TMEM lane quarter,
//                            (warp-2)/4 the 128-column half of the 256-wide accumulator).
#include <cuda.h>
#include <math.h>

#include "common.cuh"

namespace vqb {

constexpr int BM = 128;           // rows per tile (UMMA M)
constexpr int BN = 256;           // codes per tile (UMMA N)
constexpr int UMMA_K = 16;        // bf16
constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr int kMaxTerms = 6;
constexpr uint32_t kTmemCols = 512;  // two 256-column accumulators

struct TermTable {
  int n;
  int a[kMaxTerms];
  int b[kMaxTerms];
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (an error code on the host), never
// as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("vqb assign_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all tcgen05.mma issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// wait for the loads AND pin the register uses after the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i += 8)
    asm volatile("" : "+r"(r[i]), "+r"(r[i + 1]), "+r"(r[i + 2]), "+r"(r[i + 3]), "+r"(r[i + 4]), "+r"(r[i + 5]),
                 "+r"(r[i + 6]), "+r"(r[i + 7]));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// K-major, hardware-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout).
template <int BK>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  constexpr uint32_t swizzle_bytes = BK * 2;                       // 32 / 64 / 128
  constexpr uint64_t layout = swizzle_bytes == 128 ? 2 : (swizzle_bytes == 64 ? 4 : 6);
  constexpr uint64_t sbo = (8 * swizzle_bytes) >> 4;               // 8-row core-matrix group stride
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);                         // [0,14)  start address
  d |= (uint64_t)1 << 16;                                          // [16,30) leading byte offset (unused, swizzled K-major)
  d |= sbo << 32;                                                  // [32,46) stride byte offset
  d |= (uint64_t)1 << 46;                                          // [46,48) descriptor version (sm_100)
  d |= layout << 61;                                               // [61,64) swizzle mode
  return d;
}

// ---------------------------------------------------------------------------------------------
// epilogue helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

// arg-max over a 32-column chunk held in registers.  USE_SIDE: score = acc - side[col] (L2);
// MASK: columns >= n_valid (zero-padded operand rows of the last code tile) are excluded.
template <bool USE_SIDE, bool MASK>
__device__ __forceinline__ void chunk_argmax(const uint32_t (&r)[32], uint32_t side_saddr, uint32_t col_base,
                                             int n_valid, float& best, uint32_t& best_idx) {
  float s[32];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    if constexpr (USE_SIDE) {
      const float4 h = lds128(side_saddr + j * 4);  // warp-wide broadcast
      s[j] = __uint_as_float(r[j]) - h.x;
      s[j + 1] = __uint_as_float(r[j + 1]) - h.y;
      s[j + 2] = __uint_as_float(r[j + 2]) - h.z;
      s[j + 3] = __uint_as_float(r[j + 3]) - h.w;
    } else {
      s[j] = __uint_as_float(r[j]);
      s[j + 1] = __uint_as_float(r[j + 1]);
      s[j + 2] = __uint_as_float(r[j + 2]);
      s[j + 3] = __uint_as_float(r[j + 3]);
    }
  }
  if constexpr (MASK) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j >= n_valid) s[j] = -INFINITY;
  }
  float m[11];
#pragma unroll
  for (int j = 0; j < 10; ++j) m[j] = fmax3(s[3 * j], s[3 * j + 1], s[3 * j + 2]);
  m[10] = fmaxf(s[30], s[31]);
  const float m0 = fmax3(m[0], m[1], m[2]), m1 = fmax3(m[3], m[4], m[5]), m2 = fmax3(m[6], m[7], m[8]);
  const float mx = fmax3(fmax3(m0, m1, m2), m[9], m[10]);
  if (mx > best) {  // strict: an equal score later in the scan never displaces a lower index
    int j = 31;
#pragma unroll
    for (int jj = 30; jj >= 0; --jj)
      if (s[jj] == mx) j = jj;
    best = mx;
    best_idx = col_base + (uint32_t)j;
  }
}

// One 128-column half of a 128 x 256 accumulator: four 32-column chunks, TMEM loads double-buffered so
// that the load of chunk c+1 is in flight while chunk c is reduced.  The accumulator is released to the
// MMA warp as soon as the last load has landed in registers.
template <bool USE_SIDE, bool MASK>
__device__ __forceinline__ void tile_argmax(uint32_t taddr, uint32_t side_saddr, uint32_t gcol0, int64_t b_rows,
                                            uint64_t* tmem_empty_bar, int lane, float& best, uint32_t& best_idx) {
  uint32_t ra[32], rb[32];
  auto nv = [&](int c) -> int {
    if constexpr (!MASK) return 32;
    const int64_t left = b_rows - (int64_t)(gcol0 + 32 * c);
    return left >= 32 ? 32 : (left < 0 ? 0 : (int)left);
  };
  tmem_ld32(taddr, ra);
  tmem_ld_wait(ra);
  tmem_ld32(taddr + 32, rb);
  chunk_argmax<USE_SIDE, MASK>(ra, side_saddr, gcol0, nv(0), best, best_idx);
  tmem_ld_wait(rb);
  tmem_ld32(taddr + 64, ra);
  chunk_argmax<USE_SIDE, MASK>(rb, side_saddr + 128, gcol0 + 32, nv(1), best, best_idx);
  tmem_ld_wait(ra);
  tmem_ld32(taddr + 96, rb);
  chunk_argmax<USE_SIDE, MASK>(ra, side_saddr + 256, gcol0 + 64, nv(2), best, best_idx);
  tmem_ld_wait(rb);
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(tmem_empty_bar);  // all four loads are in registers: TMEM buffer is free
  chunk_argmax<USE_SIDE, MASK>(rb, side_saddr + 384, gcol0 + 96, nv(3), best, best_idx);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// WHOLE = true : one pipeline stage holds every operand plane of one (row tile, code tile) work item
//                (Dp == BK <= 64): one barrier wait, all term MMAs back to back, two commits per tile.
// WHOLE = false: classic k-blocked ring, one stage = one BK-wide slab of one plane pair (large D).
template <int BK, bool WHOLE>
__global__ void __launch_bounds__(kThreads, 1)
assign_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ TermTable terms, int pa, int pb, int kblocks, int nstages, int64_t a_rows,
                 int64_t a_rows_pad, int64_t b_rows, int64_t b_rows_pad, const float* __restrict__ b_half_sqnorm,
                 int64_t b_index_offset, unsigned long long* __restrict__ keys) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t kABytes = BM * BK * 2, kBBytes = BN * BK * 2;
  const uint32_t stage_bytes = WHOLE ? (uint32_t)pa * kABytes + (uint32_t)pb * kBBytes : kABytes + kBBytes;
  // carve: [stages] | side[2][256] | barriers | tmem ptr        (base re-aligned to 1024 B)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* side_smem = reinterpret_cast<float*>(smem + (size_t)nstages * stage_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(side_smem + 2 * BN);
  uint64_t* empty_bar = full_bar + nstages;
  uint64_t* tmem_full = empty_bar + nstages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t a_tiles = (a_rows + BM - 1) / BM, b_tiles = (b_rows + BN - 1) / BN;
  const int64_t total = a_tiles * b_tiles;
  const int64_t t0 = total * blockIdx.x / gridDim.x, t1 = total * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    for (int s = 0; s < nstages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + i, 1);
      mbar_init(tmem_empty + i, kEpiThreads / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int64_t at = t0 / b_tiles, bt = t0 - at * b_tiles;
      for (int64_t t = t0; t < t1; ++t) {
        const int a_row = (int)(at * BM), b_row = (int)(bt * BN);
        if constexpr (WHOLE) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          mbar_arrive_expect_tx(full_bar + stage, stage_bytes);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          for (int p = 0; p < pa; ++p)
            tma_load_2d(sa + p * kABytes, &tmap_a, full_bar + stage, 0, (int)(p * a_rows_pad) + a_row);
          uint8_t* sb = sa + pa * kABytes;
          for (int p = 0; p < pb; ++p)
            tma_load_2d(sb + p * kBBytes, &tmap_b, full_bar + stage, 0, (int)(p * b_rows_pad) + b_row);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        } else {
#pragma unroll 1
          for (int term = 0; term < terms.n; ++term) {
            const int arow = (int)(terms.a[term] * a_rows_pad) + a_row, brow = (int)(terms.b[term] * b_rows_pad) + b_row;
#pragma unroll 1
            for (int kb = 0; kb < kblocks; ++kb) {
              mbar_wait(empty_bar + stage, phase ^ 1);
              mbar_arrive_expect_tx(full_bar + stage, stage_bytes);
              uint8_t* sa = smem + (size_t)stage * stage_bytes;
              tma_load_2d(sa, &tmap_a, full_bar + stage, kb * BK, arow);
              tma_load_2d(sa + kABytes, &tmap_b, full_bar + stage, kb * BK, brow);
              if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
          }
        }
        if (++bt == b_tiles) { bt = 0; ++at; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      // kind::f16 instruction descriptor: D=f32, A=B=bf16, K-major both, N=256, M=128
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int64_t local = 0;
      for (int64_t t = t0; t < t1; ++t, ++local) {
        const int buf = (int)(local & 1);
        const uint32_t use = (uint32_t)(local >> 1);
        mbar_wait(tmem_empty + buf, (use & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * BN;
        if constexpr (WHOLE) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
          const uint32_t sb = sa + (uint32_t)pa * kABytes;
#pragma unroll
          for (int term = 0; term < kMaxTerms; ++term) {
            if (term < terms.n) {
              const uint64_t adesc = make_smem_desc<BK>(sa + (uint32_t)terms.a[term] * kABytes);
              const uint64_t bdesc = make_smem_desc<BK>(sb + (uint32_t)terms.b[term] * kBBytes);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (term | k) != 0);
            }
          }
          umma_commit(empty_bar + stage);  // smem slot reusable once these MMAs retire
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        } else {
          const int nv = terms.n * kblocks;
#pragma unroll 1
          for (int v = 0; v < nv; ++v) {
            mbar_wait(full_bar + stage, phase);
            tc_fence_after();
            const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
            const uint64_t adesc = make_smem_desc<BK>(sa), bdesc = make_smem_desc<BK>(sa + kABytes);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)  // +32 bytes along K inside the swizzle atom: +2 in (addr >> 4)
              umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (v | k) != 0);
            umma_commit(empty_bar + stage);
            if (++stage == nstages) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(tmem_full + buf);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue: fused arg-max =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;   // TMEM lanes [32*quarter, 32*quarter+32) are the only ones this warp may read
    const int half = ew >> 2;       // 128-column half of the accumulator
    const int etid = threadIdx.x - 64;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const bool use_side = b_half_sqnorm != nullptr;  // uniform for the whole launch
    const uint32_t side_base = smem_u32(side_smem);
    float best = -INFINITY;
    uint32_t best_idx = 0xffffffffu;
    int64_t cur_at = -1;
    int64_t local = 0;
    auto flush = [&]() {
      const int64_t row = cur_at * BM + quarter * 32 + lane;
      if (cur_at >= 0 && row < a_rows && best_idx != 0xffffffffu)
        atomicMin(keys + row, make_key(best, best_idx + (uint32_t)b_index_offset));
    };
    int64_t at = t0 / b_tiles, bt = t0 - at * b_tiles;
    for (int64_t t = t0; t < t1; ++t, ++local) {
      if (at != cur_at) {
        flush();
        cur_at = at;
        best = -INFINITY;
        best_idx = 0xffffffffu;
      }
      const int buf = (int)(local & 1);
      const uint32_t use = (uint32_t)(local >> 1);
      if (use_side) {
        // 256 epilogue threads <-> 256 codes of this tile (the side vector is padded to rows_pad with +inf).
        // The barrier also orders "everyone finished the tile that used this buffer two tiles ago".
        side_smem[buf * BN + etid] = __ldg(b_half_sqnorm + bt * BN + etid);
        named_bar_sync(1, kEpiThreads);
      }
      mbar_wait(tmem_full + buf, use & 1);
      tc_fence_after();
      const uint32_t col0 = (uint32_t)(half * 128);
      const uint32_t taddr = tmem_base + lane_addr + (uint32_t)buf * BN + col0;
      const uint32_t side_saddr = side_base + (uint32_t)(buf * BN + col0) * 4;
      const uint32_t gcol0 = (uint32_t)(bt * BN) + col0;
      const bool partial = (bt + 1) * BN > b_rows;
      if (!partial) {
        if (use_side) tile_argmax<true, false>(taddr, side_saddr, gcol0, b_rows, tmem_empty + buf, lane, best, best_idx);
        else tile_argmax<false, false>(taddr, side_saddr, gcol0, b_rows, tmem_empty + buf, lane, best, best_idx);
      } else {
        if (use_side) tile_argmax<true, true>(taddr, side_saddr, gcol0, b_rows, tmem_empty + buf, lane, best, best_idx);
        else tile_argmax<false, true>(taddr, side_saddr, gcol0, b_rows, tmem_empty + buf, lane, best, best_idx);
      }
      if (++bt == b_tiles) { bt = 0; ++at; }
    }
    flush();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int make_operand_map(CUtensorMap* map, const void* base, int64_t total_rows, int Dp, int BK, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return VQB_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)Dp, (cuuint64_t)total_rows};
  cuuint64_t gstride[1] = {(cuuint64_t)Dp * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw =
      BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (BK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld Dp=%d BK=%d box_rows=%d)", (int)r,
              (long long)total_rows, Dp, BK, box_rows);
    return VQB_ERR_CUDA;
  }
  return VQB_OK;
}

// plane-pair terms kept: i + j <= max(pa, pb) - 1 (every dropped term is below 2^-24 relative for 3 planes);
// smallest terms first so that they are not absorbed by the large ones.
static TermTable make_terms(int pa, int pb) {
  TermTable t{};
  const int order = (pa > pb ? pa : pb) - 1;
  for (int s = order; s >= 0; --s)
    for (int i = 0; i < pa; ++i) {
      const int j = s - i;
      if (j < 0 || j >= pb) continue;
      if (t.n < kMaxTerms) {
        t.a[t.n] = i;
        t.b[t.n] = j;
        ++t.n;
      }
    }
  return t;
}

template <int BK, bool WHOLE>
static int launch(const void* a_planes, int pa, int64_t a_rows, const void* b_planes, int pb, int64_t b_rows, int Dp,
                  const float* h, int64_t off, unsigned long long* keys, cudaStream_t st) {
  const int64_t a_pad = vqb_operand_rows_pad(a_rows), b_pad = vqb_operand_rows_pad(b_rows);
  CUtensorMap ma, mb;
  if (int e = make_operand_map(&ma, a_planes, pa * a_pad, Dp, BK, BM)) return e;
  if (int e = make_operand_map(&mb, b_planes, pb * b_pad, Dp, BK, BN)) return e;
  const uint32_t stage_bytes = WHOLE ? (uint32_t)(pa * BM + pb * BN) * BK * 2 : (uint32_t)(BM + BN) * BK * 2;
  int nstages = (int)(196608 / stage_bytes);
  if (nstages > 8) nstages = 8;
  const size_t smem_bytes = 1024 + (size_t)nstages * stage_bytes + 2 * BN * sizeof(float) + (2 * nstages + 4) * 8 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    VQB_CUDA_OK(cudaFuncSetAttribute(assign_tc_kernel<BK, WHOLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const TermTable terms = make_terms(pa, pb);
  const int64_t total = ((a_rows + BM - 1) / BM) * ((b_rows + BN - 1) / BN);
  int grid = sm_count();
  if (total < grid) grid = (int)total;
  assign_tc_kernel<BK, WHOLE><<<grid, kThreads, smem_bytes, st>>>(ma, mb, terms, pa, pb, Dp / BK, nstages, a_rows, a_pad,
                                                                  b_rows, b_pad, h, off, keys);
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int assign_tc_launch(const void* a_planes, int pa, int64_t a_rows, const void* b_planes, int pb, int64_t b_rows,
                     int D, const float* h, int64_t off, unsigned long long* keys, cudaStream_t st) {
  const int Dp = (int)vqb_operand_dp(D);
  // whole-tile stages need at least a double buffer of all planes of one work item in shared memory
  const bool whole = Dp <= 64 && (size_t)(pa * BM + pb * BN) * Dp * 2 * 2 <= 196608;
#define VQB_LAUNCH(BK_, W_) return launch<BK_, W_>(a_planes, pa, a_rows, b_planes, pb, b_rows, Dp, h, off, keys, st)
  if (Dp == 16) { if (whole) VQB_LAUNCH(16, true); VQB_LAUNCH(16, false); }
  if (Dp == 32) { if (whole) VQB_LAUNCH(32, true); VQB_LAUNCH(32, false); }
  if (whole) VQB_LAUNCH(64, true);
  VQB_LAUNCH(64, false);
#undef VQB_LAUNCH
}

}  // namespace vqb
