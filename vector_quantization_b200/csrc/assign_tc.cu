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
// Warp roles (352 threads):  warp 0 = TMA producer (whole-tile mode: also fetches the next row tile's A planes a
//                            row tile ahead and converts zero-copy bf16 tokens to fp16 in shared memory),
//                            warp 1 = TMEM allocator + MMA issuer,
//                            warps 2..9 = epilogue (warp%4 selects the TMEM lane quarter,
//                            (warp-2)/4 the 128-column half of the 256-wide accumulator),
//                            warp 10 = second MMA issuer (whole-tile mode: the issuers alternate tiles and
//                            each owns one TMEM accumulator; idle in the k-blocked mode).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace vqb {

constexpr int BM = 128;           // rows per tile (UMMA M)
constexpr int BN = 256;           // codes per tile (UMMA N)
constexpr int UMMA_K = 16;        // bf16
constexpr int kMaxTerms = 6;
constexpr uint32_t kTmemCols = 512;  // two 256-column accumulators
#ifndef VQB_ONE_COMMIT   // whole-tile mode: one tcgen05.commit per tile (developer A/B switch)
#define VQB_ONE_COMMIT 1
#endif
constexpr int kEpilogueWarps = 8;
#ifndef VQB_EARLY_RELEASE
#define VQB_EARLY_RELEASE 0
#endif
#ifndef VQB_ISSUERS       // whole-tile mode: MMA-issuing warps (tile t is issued by warp t % VQB_ISSUERS; developer A/B switch)
#define VQB_ISSUERS 2
#endif
constexpr int kIssuer2Warp = 2 + kEpilogueWarps;              // the second issuer sits after the epilogue warps
constexpr int kThreads = 64 + kEpilogueWarps * 32 + (VQB_ISSUERS > 1 ? 32 : 0);
constexpr uint32_t kStashBytes = kEpilogueWarps * 4096;  // winning-chunk stash: [warp][8 float4][32 lanes]
constexpr uint32_t kShareBytes = kEpilogueWarps * 32 * 16;  // per-row running best published to the other column half

struct TermTable {
  int n;
  int a[kMaxTerms];
  int b[kMaxTerms];
  int scale[kMaxTerms];  // 1: the accumulator is multiplied by 2^-kPairShift before this term is added
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
// D = D * 2^-11 + A.B  (scale-input-d): folds the 2^11 pre-scale of an fp16 pair's low plane back
__device__ __forceinline__ void umma_f16_scaled(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  static_assert(kPairShift == 11, "immediate below");
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 11;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc)
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// In-place bf16 -> fp16 conversion of a TMA-written shared-memory slot by one warp (any swizzle: element-wise),
// followed by the generic->async proxy fence that makes the result visible to tcgen05.mma.  Saturating: a bf16
// magnitude above 65504 becomes +-65504 (vqb_pack_rows' VQB_PLANES_F16 rescales such rows instead).
__device__ __forceinline__ void convert_slot_bf16_to_f16(uint32_t saddr, uint32_t bytes, int lane) {
  constexpr int U = 4;  // chunks in flight per lane: the slot is a multiple of 32 lanes x 16 B x 4 (128 rows x >= 16 columns)
  for (uint32_t off = (uint32_t)lane * 16; off < bytes; off += U * 32 * 16) {
    uint32_t w[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u)
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3])
                   : "r"(saddr + off + u * 512));
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float lo = __uint_as_float(w[u][i] << 16), hi = __uint_as_float(w[u][i] & 0xffff0000u);
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w[u][i]) : "f"(hi), "f"(lo));
      }
      asm volatile("st.shared.v4.b32 [%4], {%0, %1, %2, %3};" ::"r"(w[u][0]), "r"(w[u][1]), "r"(w[u][2]), "r"(w[u][3]),
                   "r"(saddr + off + u * 512)
                   : "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
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
#ifdef VQB_TIMELINE  // developer build: clock64 timeline of the first tiles of CTA 0 (tools/timeline.py)
__device__ long long g_ts[3][64][8];
#define TS(role, tile, ev) do { if (blockIdx.x == 0 && (tile) < 64) g_ts[role][tile][ev] = clock64(); } while (0)
#else
#define TS(role, tile, ev) do { } while (0)
#endif
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

// Predicated 128-byte stash of one thread's 32 chunk scores: shared layout [8 float4][32 lanes] per warp
// (lane-contiguous 16 B => conflict-free); only lanes whose running best just improved store.
__device__ __forceinline__ void stash_chunk(uint32_t pred, uint32_t saddr, const float (&s)[32]) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t"
      "@p st.shared.v4.f32 [%1], {%2, %3, %4, %5};\n\t"
      "@p st.shared.v4.f32 [%1+512], {%6, %7, %8, %9};\n\t"
      "@p st.shared.v4.f32 [%1+1024], {%10, %11, %12, %13};\n\t"
      "@p st.shared.v4.f32 [%1+1536], {%14, %15, %16, %17};\n\t"
      "@p st.shared.v4.f32 [%1+2048], {%18, %19, %20, %21};\n\t"
      "@p st.shared.v4.f32 [%1+2560], {%22, %23, %24, %25};\n\t"
      "@p st.shared.v4.f32 [%1+3072], {%26, %27, %28, %29};\n\t"
      "@p st.shared.v4.f32 [%1+3584], {%30, %31, %32, %33};\n\t}"
      ::"r"(pred), "r"(saddr), "f"(s[0]), "f"(s[1]), "f"(s[2]), "f"(s[3]), "f"(s[4]), "f"(s[5]), "f"(s[6]), "f"(s[7]),
      "f"(s[8]), "f"(s[9]), "f"(s[10]), "f"(s[11]), "f"(s[12]), "f"(s[13]), "f"(s[14]), "f"(s[15]), "f"(s[16]),
      "f"(s[17]), "f"(s[18]), "f"(s[19]), "f"(s[20]), "f"(s[21]), "f"(s[22]), "f"(s[23]), "f"(s[24]), "f"(s[25]),
      "f"(s[26]), "f"(s[27]), "f"(s[28]), "f"(s[29]), "f"(s[30]), "f"(s[31])
      : "memory");
}

// Branch-free running arg-max over a 32-column chunk held in registers: the thread keeps only
// (best score, first column of the chunk that produced it) and stashes that chunk's 32 scores; the
// position inside the chunk is resolved once per row tile at flush time (resolve_index).
// SIDE 1: score = acc - side[col] (L2), SIDE 2: score = acc * side[col];  MASK: columns >= n_valid (zero-padded rows of the last code
// tile) are excluded.
template <int SIDE, bool MASK>
__device__ __forceinline__ void chunk_scores(const uint32_t (&r)[32], uint32_t side_saddr, int n_valid, float (&s)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    if constexpr (SIDE == 1) {        // L2: score = <a,b> - 0.5|b|^2
      const float4 h = lds128(side_saddr + j * 4);  // warp-wide broadcast
      s[j] = __uint_as_float(r[j]) - h.x;
      s[j + 1] = __uint_as_float(r[j + 1]) - h.y;
      s[j + 2] = __uint_as_float(r[j + 2]) - h.z;
      s[j + 3] = __uint_as_float(r[j + 3]) - h.w;
    } else if constexpr (SIDE == 2) { // per-column scale: score = <a,b> * (1/|b|)  (un-normalised B rows)
      const float4 h = lds128(side_saddr + j * 4);
      s[j] = __uint_as_float(r[j]) * h.x;
      s[j + 1] = __uint_as_float(r[j + 1]) * h.y;
      s[j + 2] = __uint_as_float(r[j + 2]) * h.z;
      s[j + 3] = __uint_as_float(r[j + 3]) * h.w;
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
}
__device__ __forceinline__ float chunk_max(const float (&s)[32]) {
  float m[11];
#pragma unroll
  for (int j = 0; j < 10; ++j) m[j] = fmax3(s[3 * j], s[3 * j + 1], s[3 * j + 2]);
  m[10] = fmaxf(s[30], s[31]);
  const float m0 = fmax3(m[0], m[1], m[2]), m1 = fmax3(m[3], m[4], m[5]), m2 = fmax3(m[6], m[7], m[8]);
  return fmax3(fmax3(m0, m1, m2), m[9], m[10]);
}
template <bool CERT>
__device__ __forceinline__ void chunk_commit(float mx, const float (&s)[32], uint32_t col_base, uint32_t stash_saddr,
                                             float& best, uint32_t& best_col, float& second) {
  const bool better = mx > best;  // strict: an equal score later in the scan never displaces an earlier chunk
  // certified one-term mode: runner-up over the chunk maxima (the loser of every comparison); the runner-up INSIDE the
  // winning chunk is added at flush time from the stash
  if constexpr (CERT) second = fmaxf(second, fminf(best, mx));
  best = fmaxf(best, mx);
  best_col = better ? col_base : best_col;
  // warp-uniform skip: after the first few code tiles most chunks improve no row of the warp
  if (__any_sync(0xffffffffu, better)) stash_chunk((uint32_t)better, stash_saddr, s);
}
// Two adjacent chunks at once: their max trees are independent, so the scheduler can interleave the two
// dependent FMNMX3 chains (the reduction is latency-bound with 2 epilogue warps per SM sub-partition).
template <int SIDE, bool MASK, bool CERT>
__device__ __forceinline__ void pair_argmax(const uint32_t (&r0)[32], const uint32_t (&r1)[32], uint32_t side_saddr,
                                            uint32_t col_base, int nv0, int nv1, uint32_t stash_saddr, float& best,
                                            uint32_t& best_col, float& second) {
  float s0[32], s1[32];
  chunk_scores<SIDE, MASK>(r0, side_saddr, nv0, s0);
  chunk_scores<SIDE, MASK>(r1, side_saddr + 128, nv1, s1);
  const float mx0 = chunk_max(s0), mx1 = chunk_max(s1);
  chunk_commit<CERT>(mx0, s0, col_base, stash_saddr, best, best_col, second);
  chunk_commit<CERT>(mx1, s1, col_base + 32, stash_saddr, best, best_col, second);
}

// first position of `best` inside the stashed winning chunk (lowest index wins ties, like torch.argmin)
__device__ __forceinline__ uint32_t resolve_index(uint32_t stash_saddr, float best) {
  int j = 31;
#pragma unroll
  for (int q = 7; q >= 0; --q) {
    const float4 v = lds128(stash_saddr + q * 512);
    if (v.w == best) j = 4 * q + 3;
    if (v.z == best) j = 4 * q + 2;
    if (v.y == best) j = 4 * q + 1;
    if (v.x == best) j = 4 * q;
  }
  return (uint32_t)j;
}

// largest stashed score other than the one at position j (the runner-up inside the winning chunk)
__device__ __forceinline__ float stash_runner_up(uint32_t stash_saddr, int j) {
  float m = -INFINITY;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 v = lds128(stash_saddr + q * 512);
    m = fmaxf(m, 4 * q == j ? -INFINITY : v.x);
    m = fmaxf(m, 4 * q + 1 == j ? -INFINITY : v.y);
    m = fmaxf(m, 4 * q + 2 == j ? -INFINITY : v.z);
    m = fmaxf(m, 4 * q + 3 == j ? -INFINITY : v.w);
  }
  return m;
}

// One warp's 4 x 32 columns of a 128 x 256 accumulator; TMEM loads are double-buffered so that the load of
// chunk c+1 is in flight while chunk c is reduced.  The accumulator is released to the MMA warp as soon
// as the last load has landed in registers.
template <int SIDE, bool MASK, bool CERT>
__device__ __forceinline__ void tile_argmax(uint32_t taddr, uint32_t side_saddr, uint32_t gcol0, int b_rows,
                                            uint64_t* tmem_empty_bar, int lane, uint32_t stash_saddr, float& best,
                                            uint32_t& best_col, float& second) {
  uint32_t ra[32], rb[32];
  auto nv = [&](int c) -> int {
    if constexpr (!MASK) return 32;
    const int left = b_rows - (int)(gcol0 + 32 * c);
    return left >= 32 ? 32 : (left < 0 ? 0 : left);
  };
#if defined(VQB_ABLATE) && VQB_ABLATE == 1   // developer ablation: no TMEM reads, no scan (MMA / barrier chain alone)
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(tmem_empty_bar);
  return;
#endif
  uint32_t rc[32], rd[32];
#if VQB_EARLY_RELEASE   // all four loads up front, accumulator handed back before any reduction (developer A/B switch)
  tmem_ld32(taddr, ra);
  tmem_ld32(taddr + 32, rb);
  tmem_ld32(taddr + 64, rc);
  tmem_ld32(taddr + 96, rd);
  tmem_ld_wait(ra);
  tmem_ld_wait(rb);
  tmem_ld_wait(rc);
  tmem_ld_wait(rd);
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(tmem_empty_bar);
  pair_argmax<SIDE, MASK, CERT>(ra, rb, side_saddr, gcol0, nv(0), nv(1), stash_saddr, best, best_col, second);
  pair_argmax<SIDE, MASK, CERT>(rc, rd, side_saddr + 256, gcol0 + 64, nv(2), nv(3), stash_saddr, best, best_col, second);
  return;
#endif
  tmem_ld32(taddr, ra);
  tmem_ld32(taddr + 32, rb);
  tmem_ld_wait(ra);
  tmem_ld_wait(rb);
  tmem_ld32(taddr + 64, rc);
  tmem_ld32(taddr + 96, rd);
#if defined(VQB_ABLATE) && VQB_ABLATE == 2   // developer ablation: TMEM reads but no scan
  best = fmaxf(best, __uint_as_float(ra[0] ^ rb[0]));
#else
  pair_argmax<SIDE, MASK, CERT>(ra, rb, side_saddr, gcol0, nv(0), nv(1), stash_saddr, best, best_col, second);
#endif
  tmem_ld_wait(rc);
  tmem_ld_wait(rd);
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(tmem_empty_bar);  // every load of this warp is in registers: buffer is free
#if defined(VQB_ABLATE) && VQB_ABLATE == 2
  best = fmaxf(best, __uint_as_float(rc[0] ^ rd[0]));
#else
  pair_argmax<SIDE, MASK, CERT>(rc, rd, side_saddr + 256, gcol0 + 64, nv(2), nv(3), stash_saddr, best, best_col, second);
#endif
}

// Epilogue role: 8 warps drain every tile; warp%4 selects the TMEM lane quarter (hardware rule) and
// (warp-2)/4 the 128-column half of the accumulator.
template <int SIDE, bool CERT>
__device__ __forceinline__ void epilogue_loop(int warp, int lane, uint32_t tmem_base, uint8_t* stash_smem,
                                              uint8_t* share_smem, float* side_smem, uint64_t* acc_bar, int n_acc, uint64_t* tmem_empty, int t0,
                                              int t1, int b_tiles, int a_rows, int b_rows,
                                              const float* __restrict__ b_half_sqnorm, uint32_t b_index_offset,
                                              unsigned long long* __restrict__ keys,
                                              unsigned long long* __restrict__ second_keys) {
  const int ew = warp - 2;
  const int quarter = warp & 3;                // TMEM lanes [32*quarter, 32*quarter+32): the only ones this warp may read
  const uint32_t col0 = (uint32_t)(ew >> 2) * 128;
  const int gtid = (int)threadIdx.x - 64;       // 0..255 among the epilogue threads
  const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + col0;
  const uint32_t side_base = smem_u32(side_smem) + col0 * 4;
  const uint32_t stash_saddr = smem_u32(stash_smem) + (uint32_t)ew * 4096 + (uint32_t)lane * 16;
  const int row_in_tile = quarter * 32 + lane;
  const bool last_partial = b_tiles * BN > b_rows;
  float best = -INFINITY, second = -INFINITY;
  uint32_t best_col = 0xffffffffu;
  int at = t0 / b_tiles, bt = t0 - at * b_tiles;
  // "accumulator complete" barriers: tmem_full[2] in the k-blocked mode; in the whole-tile mode the pipeline stage's
  // own `empty` barrier (ONE tcgen05.commit per tile tells the producer and the epilogue): a cycle of n_acc barriers
  int acc_idx = 0;
  uint32_t acc_par = 0;
  // The two warps that own the two column halves of a row exchange their running best through shared memory
  // (no synchronisation: a stale entry is still a real score of the row).  A partner value from an EARLIER code
  // tile that beats mine makes my current candidate irrelevant and raises my threshold, which halves the number
  // of winning-chunk stashes (the stash is 16-22 % of the kernel, profiles/r1_assign_notes.md).  Entries are
  // tagged (row tile, code tile): only strictly earlier tiles of the same row tile are used, so on an exact
  // tie the lower code index still wins.
  const uint32_t my_share = smem_u32(share_smem) + ((uint32_t)ew * 32 + (uint32_t)lane) * 16;
  const uint32_t partner_share = smem_u32(share_smem) + ((uint32_t)(ew ^ 4) * 32 + (uint32_t)lane) * 16;

  // side term of this thread's column for the tile about to be staged, fetched one tile ahead so that its
  // L2 latency is not exposed in front of the barrier
  float side_val = 0.f;
  if constexpr (SIDE != 0) {
    if (t0 < t1) side_val = __ldg(b_half_sqnorm + bt * BN + gtid);
  }
  for (int t = t0; t < t1; ++t) {
    const uint32_t buf = (uint32_t)(t - t0) & 1u;
    {
      if constexpr (SIDE != 0) {
        // the 256 epilogue threads stage the 256 side terms of this code tile (vector padded with +inf); the
        // barrier also orders "everyone finished the tile that used this buffer before"
        side_smem[buf * BN + gtid] = side_val;
        if (t + 1 < t1) side_val = __ldg(b_half_sqnorm + (bt + 1 == b_tiles ? 0 : bt + 1) * BN + gtid);
        named_bar_sync(1, 256);
      }
      if constexpr (!CERT) {   // (the certified mode needs every half's own candidates: no exchange)
        uint32_t pb_, pat, pbt, pad;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(pb_), "=r"(pat), "=r"(pbt), "=r"(pad) : "r"(partner_share));
        const float pbest = __uint_as_float(pb_);
        if (pat == (uint32_t)at && pbt < (uint32_t)bt && pbest > best) {
          best = pbest;                        // the other half already holds a better candidate (lower index on ties):
          best_col = 0xffffffffu;              // mine is out of the race until a later chunk beats it
        }
      }
      mbar_wait(acc_bar + acc_idx, acc_par);
      if (warp == 2 && lane == 0) TS(2, t - t0, 0);
      tc_fence_after();
      const uint32_t taddr = taddr0 + buf * BN;
      const uint32_t side_saddr = side_base + buf * (BN * 4);
      const uint32_t gcol0 = (uint32_t)(bt * BN) + col0;
      if (last_partial && bt == b_tiles - 1)
        tile_argmax<SIDE, true, CERT>(taddr, side_saddr, gcol0, b_rows, tmem_empty + buf, lane, stash_saddr, best, best_col, second);
      else
        tile_argmax<SIDE, false, CERT>(taddr, side_saddr, gcol0, b_rows, tmem_empty + buf, lane, stash_saddr, best, best_col, second);
      if (warp == 2 && lane == 0) TS(2, t - t0, 1);
      if constexpr (!CERT)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_share), "r"(__float_as_uint(best)), "r"((uint32_t)at),
                     "r"((uint32_t)bt), "r"(0u)
                     : "memory");
      if (++acc_idx == n_acc) { acc_idx = 0; acc_par ^= 1; }
    }
    if (++bt == b_tiles || t + 1 == t1) {       // row tile finished (or this CTA's range ends): publish
      const int row = at * BM + row_in_tile;
      if (row < a_rows && best_col != 0xffffffffu) {
        const uint32_t j = resolve_index(stash_saddr, best);
        const unsigned long long kb = make_key(best, best_col + j + b_index_offset);
        if constexpr (CERT) {
          // runner-up = max(loser of every chunk comparison, second largest of the winning chunk); whoever loses the
          // race for keys[row] (this candidate or the previous holder) is a runner-up candidate as well
          second = fmaxf(second, stash_runner_up(stash_saddr, (int)j));
          const unsigned long long old = atomicMin(keys + row, kb);
          const unsigned long long loser = old > kb ? old : kb;
          const unsigned long long ks = make_key(second, 0xffffffffu);
          atomicMin(second_keys + row, ks < loser ? ks : loser);
        } else {
          atomicMin(keys + row, kb);
        }
      }
      best = -INFINITY;
      second = -INFINITY;
      best_col = 0xffffffffu;
      bt = 0;
      ++at;
      if (warp == 2 && lane == 0) TS(2, t - t0, 2);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// WHOLE = true : one pipeline stage holds every operand plane of one (row tile, code tile) work item
//                (Dp == BK <= 64): two barrier tests back to back, all term MMAs back to back, ONE commit per tile
//                (the stage's `empty` barrier also tells the epilogue that the accumulator is complete).
// WHOLE = false: classic k-blocked ring, one stage = one BK-wide slab of one plane pair (large D).
// MMA issue sequence of one whole-K tile for a term table known at compile time (NT terms, bit t of SCALE set when
// term t starts with the scale-input-d instruction).
template <int BK, int NT, uint32_t SCALE>
__device__ __forceinline__ void issue_terms(uint32_t d_tmem, uint64_t desc_hi, const uint32_t* a_lo, const uint32_t* b_lo,
                                            uint32_t a_add, uint32_t b_add, uint32_t idesc) {
#pragma unroll
  for (int term = 0; term < NT; ++term) {
    const uint64_t adesc = desc_hi | (uint64_t)(a_lo[term] + a_add);
    const uint64_t bdesc = desc_hi | (uint64_t)(b_lo[term] + b_add);
#pragma unroll
    for (int k = 0; k < BK / UMMA_K; ++k) {
      if (k == 0 && ((SCALE >> term) & 1u)) umma_f16_scaled(d_tmem, adesc, bdesc, idesc);
      else umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (term | k) != 0);
    }
  }
}

template <int BK, bool WHOLE>
__global__ void __launch_bounds__(kThreads, 1)
assign_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ TermTable terms, uint32_t idesc, int pa, int pb, int kblocks, int nstages,
                 int64_t a_rows,
                 int64_t a_rows_pad, int64_t b_rows, int64_t b_rows_pad, const float* __restrict__ b_half_sqnorm,
                 int side_mode, int64_t b_index_offset, unsigned long long* __restrict__ keys, int a_convert,
                 unsigned long long* __restrict__ second_keys, const int* __restrict__ a_rows_dev) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t kABytes = BM * BK * 2, kBBytes = BN * BK * 2;
  // WHOLE: a stage holds every B plane of one code tile; the row tile's A planes are RESIDENT in one of two
  // slots (loaded once per row tile, not once per work item).  k-blocked: a stage is one [A slab | B slab] pair.
  const uint32_t stage_bytes = WHOLE ? (uint32_t)pb * kBBytes : kABytes + kBBytes;
  const uint32_t a_slot_bytes = WHOLE ? (uint32_t)pa * kABytes : 0u;
  // carve: [A slots] | [stages] | stash[8 warps][4 KB] | share[8 warps][32 x 16 B] | side[2][256] | barriers | tmem ptr
  uint8_t* smem_a = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem = smem_a + 2 * (size_t)a_slot_bytes;
  uint8_t* stash_smem = smem + (size_t)nstages * stage_bytes;
  uint8_t* share_smem = stash_smem + kStashBytes;
  float* side_smem = reinterpret_cast<float*>(share_smem + kShareBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(side_smem + 2 * BN);
  uint64_t* empty_bar = full_bar + nstages;
  uint64_t* tmem_full = empty_bar + nstages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* a_full = tmem_empty + 2;
  uint64_t* a_empty = a_full + 2;
  uint64_t* a_ready = a_empty + 2;   // resident A tile converted to fp16 (a_convert mode)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_ready + 2);

  // PDL: barrier init / TMEM allocation below overlap the tail of the preceding launch (the operand packs);
  // global memory is first touched after pdl_wait().
  pdl_launch_dependents();
  if (threadIdx.x == 0) TS(0, 63, 3);   // kernel start (timeline builds only)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    for (int s = 0; s < nstages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + i, 1);
      mbar_init(tmem_empty + i, kEpilogueWarps);
      mbar_init(a_full + i, 1);
      mbar_init(a_empty + i, WHOLE ? VQB_ISSUERS : 1);   // every issuer releases every row tile
      mbar_init(a_ready + i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  if (warp >= 2 && warp < 2 + kEpilogueWarps) {  // epilogue: invalidate this thread's slot of the running-best exchange (tag = no row tile)
    const uint32_t slot = smem_u32(share_smem) + ((uint32_t)(warp - 2) * 32 + (uint32_t)lane) * 16;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "r"(0xff800000u), "r"(0xffffffffu), "r"(0u), "r"(0u) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  // The row count may live on the device (the exact re-run of the rows a one-term pass could not certify: their number
  // is only known to the preceding kernel); the TMA descriptor then covers the buffer's capacity.
  if (a_rows_dev != nullptr) {
    const int64_t n = *a_rows_dev;
    a_rows = n < a_rows ? n : a_rows;
  }
  const int64_t a_tiles = (a_rows + BM - 1) / BM, b_tiles = (b_rows + BN - 1) / BN;
  const int64_t total = a_tiles * b_tiles;
  const int64_t t0 = total * blockIdx.x / gridDim.x, t1 = total * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer (warp-converged, one elected lane issues) =====================
    if constexpr (WHOLE) {
      // The row tile's A planes are RESIDENT in one of two slots and loaded one row tile AHEAD (a few code tiles into
      // row tile r the slot of r-1 is free again), so neither the TMA latency nor the bf16 -> fp16 conversion of
      // zero-copy tokens (done here, by this warp, off the MMA thread's critical path) is exposed at a row-tile change.
      const int n_local = (int)(t1 - t0), b_tiles_i = (int)b_tiles;
      const int at0 = (int)(t0 / b_tiles);
      const int rt_count = n_local > 0 ? (int)((t1 - 1) / b_tiles) - at0 + 1 : 0;   // row tiles this CTA touches
      const int ahead = nstages + 1 < b_tiles_i - 1 ? nstages + 1 : b_tiles_i - 1;  // code tile at which A(r+1) is fetched
      const int ahead_cvt = ahead + 3 < b_tiles_i - 1 ? ahead + 3 : b_tiles_i - 1;  // ... and converted
      int bt = (int)(t0 - (int64_t)at0 * b_tiles), rt = 0, a_loaded = 0, a_converted = 0;
      int stage = 0;
      uint32_t phase = 0;
      auto load_a = [&]() {
        const int slot = a_loaded & 1;
        mbar_wait(a_empty + slot, (uint32_t)((a_loaded >> 1) & 1) ^ 1);   // MMAs of the slot's previous row tile retired
        if (elect_one()) {
          mbar_arrive_expect_tx(a_full + slot, a_slot_bytes);
          for (int p = 0; p < pa; ++p)
            tma_load_2d(smem_a + slot * a_slot_bytes + p * kABytes, &tmap_a, a_full + slot, 0,
                        (int)(p * a_rows_pad) + (at0 + a_loaded) * BM);
        }
        __syncwarp();
        ++a_loaded;
      };
      auto convert_a = [&]() {
        const int slot = a_converted & 1;
        mbar_wait(a_full + slot, (uint32_t)((a_converted >> 1) & 1));
        convert_slot_bf16_to_f16(smem_u32(smem_a) + (uint32_t)slot * a_slot_bytes, a_slot_bytes, lane);
        if (lane == 0) mbar_arrive(a_ready + slot);
        ++a_converted;
      };
      if (n_local > 0) load_a();
      for (int local = 0; local < n_local; ++local) {
        TS(0, local, 0);
        mbar_wait(empty_bar + stage, phase ^ 1);
        TS(0, local, 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar + stage, stage_bytes);
          uint8_t* sb = smem + (size_t)stage * stage_bytes;
          for (int p = 0; p < pb; ++p)
            tma_load_2d(sb + p * kBBytes, &tmap_b, full_bar + stage, 0, (int)(p * b_rows_pad) + bt * BN);
        }
        __syncwarp();
        TS(0, local, 2);
        if (a_loaded == rt + 1 && a_loaded < rt_count && bt >= ahead) load_a();
        if (a_convert && a_converted < a_loaded && (a_converted == rt || bt >= ahead_cvt)) convert_a();
        if (++stage == nstages) { stage = 0; phase ^= 1; }
        if (++bt == b_tiles_i) { bt = 0; ++rt; }
      }
    } else {
      int stage = 0;
      uint32_t phase = 0;
      int64_t at = t0 / b_tiles, bt = t0 - at * b_tiles;
      for (int64_t t = t0; t < t1; ++t) {
        const int a_row = (int)(at * BM), b_row = (int)(bt * BN);
#pragma unroll 1
        for (int term = 0; term < terms.n; ++term) {
          const int arow = (int)(terms.a[term] * a_rows_pad) + a_row, brow = (int)(terms.b[term] * b_rows_pad) + b_row;
#pragma unroll 1
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(empty_bar + stage, phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(full_bar + stage, stage_bytes);
              uint8_t* sa = smem + (size_t)stage * stage_bytes;
              tma_load_2d(sa, &tmap_a, full_bar + stage, kb * BK, arow);
              tma_load_2d(sa + kABytes, &tmap_b, full_bar + stage, kb * BK, brow);
            }
            __syncwarp();
            if (++stage == nstages) { stage = 0; phase ^= 1; }
          }
        }
        if (++bt == b_tiles) { bt = 0; ++at; }
      }
    }
  } else if (warp == 1 || warp == kIssuer2Warp) {
    // ===================== MMA issuer =====================
    if constexpr (WHOLE) {
      // This warp's serial instruction stream sets the tile period for D <= 64 (profiles/r2_notes.md section 10), so it
      // is kept to the two barrier tests (issued back to back), the MMAs with descriptors prepared outside the loop,
      // and ONE commit: the stage's `empty` barrier tells the producer "slot reusable" and the epilogue "accumulator
      // complete" at once.  (Running the loop on ONE thread instead of warp-converged + elected issue was 50 % slower:
      // divergent code cannot use the uniform datapath the tcgen05 operands live in.)
      {
        const uint32_t smem_base = smem_u32(smem);
        const uint64_t desc_hi = make_smem_desc<BK>(0);
        uint32_t a_lo[kMaxTerms], b_lo[kMaxTerms];
        uint32_t scale_mask = 0;
#pragma unroll
        for (int term = 0; term < kMaxTerms; ++term) {
          a_lo[term] = ((smem_u32(smem_a) + (uint32_t)terms.a[term] * kABytes) & 0x3ffff) >> 4;
          b_lo[term] = ((smem_base + (uint32_t)terms.b[term] * kBBytes) & 0x3ffff) >> 4;
          scale_mask |= (terms.scale[term] != 0 ? 1u : 0u) << term;
        }
        // opaque to the optimiser, or it re-reads the table from the constant bank inside the loop
        asm volatile("" : "+r"(scale_mask));
        const int n_terms = terms.n;
        const uint32_t a_slot_add = a_slot_bytes >> 4, stage_add = stage_bytes >> 4;
        const int n_local = (int)(t1 - t0), b_tiles_i = (int)b_tiles;
        uint64_t* const a_bar = a_convert ? a_ready : a_full;
        // Tile `local` belongs to issuer local % VQB_ISSUERS: with two issuers each owns one TMEM accumulator, and the
        // serial barrier-test / issue / commit stream of one tile overlaps the other issuer's.
        const int me = warp == 1 ? 0 : 1;
        const int at0 = (int)(t0 / b_tiles);
        const int rt_count = n_local > 0 ? (int)((t1 - 1) / b_tiles) - at0 + 1 : 0;   // row tiles this CTA touches
        int bt = (int)(t0 - (int64_t)at0 * b_tiles) + me, rt = 0, stage = me, released = 0, cur_rt = -1;
        uint32_t phase = 0, a_add = 0;
        while (bt >= b_tiles_i) { bt -= b_tiles_i; ++rt; }
        while (stage >= nstages) { stage -= nstages; phase ^= 1; }
        auto release_until = [&](int upto) {   // every issuer arrives once per row tile, in order (a_empty counts VQB_ISSUERS)
          while (released < upto) {
            if (elect_one()) umma_commit(a_empty + (released & 1));   // after this thread's MMAs on that slot (if any)
            __syncwarp();
            ++released;
          }
        };
        for (int local = me; local < n_local; local += VQB_ISSUERS) {
          const uint32_t buf = (uint32_t)local & 1u;
          const uint32_t e_par = (((uint32_t)local >> 1) & 1u) ^ 1u;
          TS(1, local, 0);
          const bool e_ok = mbar_try_wait(tmem_empty + buf, e_par);   // epilogue has drained this accumulator
          const bool f_ok = mbar_try_wait(full_bar + stage, phase);   // the code tile's planes have landed
          if (rt != cur_rt) {           // my first tile of a row tile: release the ones I am done with, wait for its A planes
            release_until(rt);
            mbar_wait(a_bar + (rt & 1), (uint32_t)((rt >> 1) & 1));
            a_add = (uint32_t)(rt & 1) * a_slot_add;
            cur_rt = rt;
          }
          if (!e_ok) mbar_wait(tmem_empty + buf, e_par);
          TS(1, local, 1);
          if (!f_ok) mbar_wait(full_bar + stage, phase);
          TS(1, local, 2);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BN;
          const uint32_t b_add = (uint32_t)stage * stage_add;
          if (elect_one()) {
            // straight-line issue for the two common tables (fp16 pair: lo' term, then the 2^-11-scaled hi term; one
            // plane): with the term count and the scale flags known at compile time the sequence is two adds per MMA
            if (n_terms == 2 && scale_mask == 2u) {
              issue_terms<BK, 2, 2u>(d_tmem, desc_hi, a_lo, b_lo, a_add, b_add, idesc);
            } else if (n_terms == 1) {
              issue_terms<BK, 1, 0u>(d_tmem, desc_hi, a_lo, b_lo, a_add, b_add, idesc);
            } else {
#pragma unroll
              for (int term = 0; term < kMaxTerms; ++term) {
                if (term < n_terms) {
                  const uint64_t adesc = desc_hi | (uint64_t)(a_lo[term] + a_add);
                  const uint64_t bdesc = desc_hi | (uint64_t)(b_lo[term] + b_add);
#pragma unroll
                  for (int k = 0; k < BK / UMMA_K; ++k) {
                    if (k == 0 && ((scale_mask >> term) & 1u)) umma_f16_scaled(d_tmem, adesc, bdesc, idesc);
                    else umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (term | k) != 0);
                  }
                }
              }
            }
            TS(1, local, 5);
            umma_commit(empty_bar + stage);  // these MMAs retired: stage reusable AND (one-commit mode) accumulator complete
            TS(1, local, 6);
#if !VQB_ONE_COMMIT
            umma_commit(tmem_full + buf);
#endif
            TS(1, local, 7);
          }
          __syncwarp();
          bt += VQB_ISSUERS;
          while (bt >= b_tiles_i) { bt -= b_tiles_i; ++rt; }
          stage += VQB_ISSUERS;
          while (stage >= nstages) { stage -= nstages; phase ^= 1; }
          TS(1, local, 3);
        }
        release_until(rt_count);
      }
    } else if (warp == 1) {
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
        const int nv = terms.n * kblocks;
        int term = 0, kb = 0;
#pragma unroll 1
        for (int v = 0; v < nv; ++v) {
          const bool rescale = kb == 0 && terms.scale[term] != 0;  // first MMA of a term that follows the 2^11-scaled ones
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
          if (elect_one()) {
            const uint64_t adesc = make_smem_desc<BK>(sa), bdesc = make_smem_desc<BK>(sa + kABytes);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {  // +32 bytes along K inside the swizzle atom: +2 in (addr >> 4)
              if (k == 0 && rescale) umma_f16_scaled(d_tmem, adesc, bdesc, idesc);
              else umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (v | k) != 0);
            }
            umma_commit(empty_bar + stage);
            if (v == nv - 1) umma_commit(tmem_full + buf);  // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++kb == kblocks) { kb = 0; ++term; }
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 2 + kEpilogueWarps) {
    // ===================== epilogue: fused arg-max =====================
#define VQB_EPI(SIDE_, CERT_)                                                                                       \
  epilogue_loop<SIDE_, CERT_>(warp, lane, tmem_base, stash_smem, share_smem, side_smem, (WHOLE && VQB_ONE_COMMIT) ? empty_bar : tmem_full,           \
                              (WHOLE && VQB_ONE_COMMIT) ? nstages : 2, tmem_empty, (int)t0,                                  \
                              (int)t1, (int)b_tiles, (int)a_rows, (int)b_rows, b_half_sqnorm, (uint32_t)b_index_offset, keys, \
                              second_keys)
    if (second_keys != nullptr) {   // certified one-term mode (cosine family: no side term, or the column scale)
      if (b_half_sqnorm == nullptr || side_mode == 0) VQB_EPI(0, true);
      else if (side_mode == 1) VQB_EPI(1, true);
      else VQB_EPI(2, true);
    } else {
      if (b_half_sqnorm == nullptr || side_mode == 0) VQB_EPI(0, false);
      else if (side_mode == 1) VQB_EPI(1, false);
      else VQB_EPI(2, false);
    }
#undef VQB_EPI
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

// plane-pair terms kept for bf16 planes: i + j <= max(pa, pb) - 1 (every dropped term is below 2^-24 relative
// for 3 planes); smallest terms first so that they are not absorbed by the large ones.
// fp16 pairs (hi, lo' = lo * 2^11): the lo' terms go first, then the accumulator is scaled by 2^-11 in the
// same instruction that adds the first hi term (lo'.lo' would be 2^-22 relative: dropped).
static TermTable make_terms(int pa, int pb) {
  TermTable t{};
  auto push = [&](int i, int j, int scale) {
    if (t.n < kMaxTerms) {
      t.a[t.n] = i;
      t.b[t.n] = j;
      t.scale[t.n] = scale;
      ++t.n;
    }
  };
  if (is_f16(pa)) {                    // fp16 family (both operands: checked by vqb_assign)
    if (is_f16x2(pa)) push(1, 0, 0);   // a_lo' . b_hi
    if (is_f16x2(pb)) push(0, 1, 0);   // a_hi  . b_lo'
    push(0, 0, t.n > 0 ? 1 : 0);       // D = D * 2^-11 + a_hi . b_hi
    return t;
  }
  const int order = (pa > pb ? pa : pb) - 1;
  for (int s = order; s >= 0; --s)
    for (int i = 0; i < pa; ++i) {
      const int j = s - i;
      if (j < 0 || j >= pb) continue;
      push(i, j, 0);
    }
  return t;
}

template <int BK, bool WHOLE>
static int launch(const void* a_planes, int pa, int64_t a_rows, int64_t a_pad, const void* b_planes, int pb,
                  int64_t b_rows, int64_t b_pad, int Dp, const float* h, int side_mode, int64_t off,
                  unsigned long long* keys, unsigned long long* second_keys, const int* a_rows_dev, cudaStream_t st) {
  // one bf16 plane (e.g. zero-copy tokens) against fp16 planes: the kernel converts the resident A tile in
  // shared memory (whole-tile mode only)
  const int a_convert = (pa == 1 && is_f16(pb)) ? 1 : 0;
  if (a_convert) {
    if (!WHOLE) {
      set_error("vqb_assign: bf16 rows against fp16 planes need D <= 64 (in-kernel conversion); pack them with VQB_PLANES_F16");
      return VQB_ERR_ARG;
    }
    pa = VQB_PLANES_F16;
  }
  const TermTable terms = make_terms(pa, pb);
  // kind::f16 instruction descriptor: D = f32, A / B = bf16 (1) or f16 (0), K-major both, N = 256, M = 128
  const uint32_t idesc = (1u << 4) | ((is_f16(pa) ? 0u : 1u) << 7) | ((is_f16(pb) ? 0u : 1u) << 10) |
                         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  pa = plane_count(pa);
  pb = plane_count(pb);
  CUtensorMap ma, mb;  // TMA moves 16-bit elements: the bf16 data type also carries the fp16 planes
  if (int e = make_operand_map(&ma, a_planes, pa * a_pad, Dp, BK, BM)) return e;
  if (int e = make_operand_map(&mb, b_planes, pb * b_pad, Dp, BK, BN)) return e;
  const uint32_t stage_bytes = WHOLE ? (uint32_t)(pb * BN) * BK * 2 : (uint32_t)(BM + BN) * BK * 2;
  const uint32_t a_resident = WHOLE ? 2u * (uint32_t)pa * BM * BK * 2 : 0u;
  int nstages = (int)((196608 - kStashBytes - a_resident) / stage_bytes);
  if (nstages > 8) nstages = 8;
  const size_t smem_bytes = 1024 + a_resident + (size_t)nstages * stage_bytes + kStashBytes + kShareBytes + 2 * BN * sizeof(float) +
                            (2 * nstages + 10) * 8 + 16;
  static bool attr_set[64] = {false};  // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    VQB_CUDA_OK(cudaFuncSetAttribute(assign_tc_kernel<BK, WHOLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int64_t total = ((a_rows + BM - 1) / BM) * ((b_rows + BN - 1) / BN);
  int grid = sm_count();
  if (total < grid) grid = (int)total;
  VQB_CUDA_OK(launch_pdl(assign_tc_kernel<BK, WHOLE>, grid, kThreads, smem_bytes, st, ma, mb, terms, idesc, pa, pb, Dp / BK,
                         nstages, a_rows, a_pad, b_rows, b_pad, h, side_mode, off, keys, a_convert, second_keys, a_rows_dev));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int assign_tc_launch(const void* a_planes, int pa, int64_t a_rows, int64_t a_pad, const void* b_planes, int pb,
                     int64_t b_rows, int64_t b_pad, int D, const float* h, int side_mode, int64_t off,
                     unsigned long long* keys, unsigned long long* second_keys, const int* a_rows_dev, cudaStream_t st) {
  const int Dp = (int)vqb_operand_dp(D);
  // resident-A mode: two A slots plus at least two whole-B-tile stages must fit in shared memory
  const bool whole =
      Dp <= 64 && (size_t)(2 * plane_count(pa) * BM + 2 * plane_count(pb) * BN) * Dp * 2 <= 196608 - kStashBytes;
#define VQB_LAUNCH(BK_, W_) \
  return launch<BK_, W_>(a_planes, pa, a_rows, a_pad, b_planes, pb, b_rows, b_pad, Dp, h, side_mode, off, keys, second_keys, \
                         a_rows_dev, st)
  if (Dp == 16) { if (whole) VQB_LAUNCH(16, true); VQB_LAUNCH(16, false); }
  if (Dp == 32) { if (whole) VQB_LAUNCH(32, true); VQB_LAUNCH(32, false); }
  if (whole) VQB_LAUNCH(64, true);
  VQB_LAUNCH(64, false);
#undef VQB_LAUNCH
}

}  // namespace vqb

#ifdef VQB_TIMELINE
extern "C" int vqb_debug_timeline(long long* out_host) {
  return (int)cudaMemcpyFromSymbol(out_host, vqb::g_ts, sizeof(vqb::g_ts));
}
#endif
