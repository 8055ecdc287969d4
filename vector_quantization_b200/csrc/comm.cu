// Multi-GPU exchange steps of the hot path over NVLink peer memory (one process per GPU), FUSED with the
// codebook update that consumes them: the kernel that reduces the per-rank statistics is the kernel that
// applies the update and publishes the new codebook rows — no separate collective launch, no NCCL on the
// data path.
//
// Every rank owns one "region" (cudaMalloc, exported with CUDA IPC) with the same layout; all ranks map all
// regions.  Two-shot scheme: rank r owns the code rows [r*K/w, (r+1)*K/w): it reads that slice of every peer's
// partial statistics over NVLink (fixed rank order => every replica gets bit-identical results), computes the
// updated rows, and stores them into EVERY peer's codebook (all-gather by remote stores).  Two flag barriers
// in peer memory bracket the exchange:
//   A  "my partial statistics are complete"            (before anyone reads them)
//   B  "I have read all I need and written all my rows" (before anyone consumes the codebook / reuses buffers)
// Flags carry a monotonically increasing epoch, so they never need resetting and a CUDA graph can replay the
// kernels.  All spins are bounded (trap after ~2 s): a missing peer is an error, never a hung GPU.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace vqb {

constexpr int kMaxWorld = VQB_COMM_MAX_WORLD;
constexpr int kBatch = 8;   // peers whose loads are issued back to back (one NVLink round trip per batch)

// Control block at offset 0 of every region; the peer table follows at byte 256.
struct CommCtl {
  uint32_t sig[2][kMaxWorld];  // sig[set][r] is written by rank r: the epoch it has reached on barrier `set`
  uint32_t epoch;              // collectives completed by THIS rank
  uint32_t ticket;             // blocks of the running collective that have finished (self-resetting)
  uint32_t ll_epoch;           // low-latency exchanges completed by THIS rank (its parity is the in-word flag)
};
static_assert(sizeof(CommCtl) <= 256, "control block");
constexpr size_t kPeerTableOff = 256;
static_assert(kPeerTableOff + kMaxWorld * sizeof(void*) <= VQB_COMM_HEADER_BYTES, "header");

// per-block copy of the peer table (filled by comm_begin): peer pointers come from shared memory, not from a
// dependent global load in front of every remote access
__device__ __forceinline__ char** peer_table() {
  __shared__ char* table[kMaxWorld];
  return table;
}

struct Comm {
  char* base;  // local region
  int rank, world;
  __device__ __forceinline__ char* peer(int r) const { return peer_table()[r]; }
  __device__ __forceinline__ CommCtl* ctl() const { return reinterpret_cast<CommCtl*>(base); }
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer-memory loads: system scope, so that no non-coherent cache level can serve a stale line
__device__ __forceinline__ float ld_peer(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_peer_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Barrier A: block 0 announces "everything this rank wrote BEFORE this kernel is complete" (the kernel boundary /
// griddepcontrol.wait made it visible); EVERY block then waits until all ranks have announced.
// (Measured on 2 x B200: polling with acquire loads + a fence per thread, as below, 21 us per exchange; relaxed
// polling + one fence per block, 35 us: the tight relaxed polls of ~600 blocks starve the flag line.)
__device__ __forceinline__ uint32_t comm_begin(const Comm& c) {
  __shared__ uint32_t s_epoch;
  CommCtl* ctl = c.ctl();
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(&ctl->epoch) + 1;
  if (threadIdx.x < c.world) peer_table()[threadIdx.x] = reinterpret_cast<char* const*>(c.base + kPeerTableOff)[threadIdx.x];
  __syncthreads();
  const uint32_t e = s_epoch;
  if (blockIdx.x == 0 && threadIdx.x < c.world) {
    __threadfence_system();
    st_release_sys(&reinterpret_cast<CommCtl*>(c.peer(threadIdx.x))->sig[0][c.rank], e);
  }
  if (threadIdx.x < c.world) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(&ctl->sig[0][threadIdx.x]) - e) < 0) {
      if (clock64() - t0 > 4000000000ll) {
        printf("vqb comm: rank %d timed out waiting for rank %d (barrier A, epoch %u)\n", c.rank, (int)threadIdx.x, e);
        __trap();
      }
    }
  }
  __syncthreads();
  return e;
}

// Barrier B: every block calls this after its last peer access.  The last block to arrive announces "this rank
// is done" to all peers, waits for all of them, and retires the epoch; the other blocks simply exit (the kernel
// boundary orders everything that follows behind the last block).
__device__ __forceinline__ void comm_end(const Comm& c, uint32_t e) {
  __shared__ bool s_last;
  CommCtl* ctl = c.ctl();
  __threadfence_system();   // my remote stores are performed before my arrival is counted
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t t = atomicAdd(&ctl->ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();   // acquire side of the ticket: the other blocks' stores precede the announcement
  if (threadIdx.x < c.world) {
    st_release_sys(&reinterpret_cast<CommCtl*>(c.peer(threadIdx.x))->sig[1][c.rank], e);
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(&ctl->sig[1][threadIdx.x]) - e) < 0) {
      if (clock64() - t0 > 4000000000ll) {
        printf("vqb comm: rank %d timed out waiting for rank %d (barrier B, epoch %u)\n", c.rank, (int)threadIdx.x, e);
        __trap();
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ctl->ticket = 0;
    *reinterpret_cast<volatile uint32_t*>(&ctl->epoch) = e;
    __threadfence();
  }
}

__device__ __forceinline__ void slice_of(int64_t total, const Comm& c, int64_t& lo, int64_t& hi) {
  const int64_t per = (total + c.world - 1) / c.world;
  lo = per * c.rank;
  if (lo > total) lo = total;
  hi = lo + per;
  if (hi > total) hi = total;
}

// ---- VQ-KD: all-reduce(SUM) of [K*D sums | K counts] fused with centroid / normalise / EMA / normalise --------
// vq/algorithms/vqkd/quantizers/callbacks.py:60-71,126-128,73-75 (the all_reduce at :63-64 and vq/utils.py:35)
// G lanes per code row, NPL elements per lane (d = lane + j*G): the reduced sums stay in registers, so every
// remote element crosses NVLink exactly once.
template <int G, int NPL>
__global__ void __launch_bounds__(256) comm_kmeans_ema_kernel(Comm c, size_t stats_off, size_t w_off, int64_t K, int D,
                                                              float decay, float omd) {
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t e = comm_begin(c);
  const int lane = threadIdx.x % G;
  const int64_t rows_per_block = blockDim.x / G;
  int64_t lo, hi;
  slice_of(K, c, lo, hi);
  const float* __restrict__ Wl = reinterpret_cast<const float*>(c.base + w_off);
  for (int64_t base = lo + blockIdx.x * rows_per_block; base < hi; base += (int64_t)gridDim.x * rows_per_block) {
    const int64_t k_raw = base + threadIdx.x / G;
    const bool valid = k_raw < hi;
    const int64_t k = valid ? k_raw : hi - 1;
    float cnt = 0.f, s[NPL], w[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) s[j] = 0.f;
    constexpr int B = NPL <= 4 ? kBatch : (NPL <= 8 ? 4 : (NPL <= 16 ? 2 : 1));   // B * NPL loads in flight per lane
    for (int r0 = 0; r0 < c.world; r0 += B) {        // all loads of a batch are issued before the first add
      float vc[B], v[B][NPL];
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const bool on = r0 + b < c.world;
        const float* __restrict__ sr = reinterpret_cast<const float*>(c.peer(on ? r0 + b : c.rank) + stats_off);
        vc[b] = on ? ld_peer(sr + K * (int64_t)D + k) : 0.f;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          const int d = lane + j * G;
          v[b][j] = (on && d < D) ? ld_peer(sr + k * D + d) : 0.f;
        }
      }
#pragma unroll
      for (int b = 0; b < B; ++b) {                  // fixed rank order: every replica computes bit-identical sums
        cnt += vc[b];
#pragma unroll
        for (int j = 0; j < NPL; ++j) s[j] += v[b][j];
      }
    }
    const bool occurred = cnt > 0.f;
    const float den = fmaxf(cnt, 1.f);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int d = lane + j * G;
      w[j] = d < D ? Wl[k * D + d] : 0.f;
      s[j] = d < D ? (occurred ? __fdiv_rn(s[j], den) : w[j]) : 0.f;   // centroid, or the old row of an unused code
      ss = fmaf(s[j], s[j], ss);
    }
    ss = group_sum<G>(ss);
    const float dn = fmaxf(sqrtf(ss), kNormEps);
    float ss2 = 0.f;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      s[j] = __fadd_rn(__fmul_rn(w[j], decay), __fmul_rn(__fdiv_rn(s[j], dn), omd));
      ss2 = fmaf(s[j], s[j], ss2);
    }
    ss2 = group_sum<G>(ss2);
    const float dn2 = fmaxf(sqrtf(ss2), kNormEps);
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int d = lane + j * G;
      if (valid && d < D) {
        const float out = __fdiv_rn(s[j], dn2);
        for (int r = 0; r < c.world; ++r) reinterpret_cast<float*>(c.peer(r) + w_off)[k * D + d] = out;
      }
    }
  }
  comm_end(c, e);
}

// ---- low-latency variant of the VQ-KD exchange (payloads up to a few MB) --------------------------------------
// The barrier protocol above costs six NVLink hops (flag, read round trip, write + ack, flag: ~20 us measured).
// Here every fp32 word is its OWN flag: its mantissa LSB carries the parity of the exchange epoch.  A staging slot is
// rewritten at every exchange, so it holds either the word of the previous exchange (other parity) or the current one:
// one bit distinguishes them, any 4-byte store is atomic, and the wire carries 1x the payload (an (value, epoch)
// 8-byte word, NCCL's "LL" format, was measured first: 2x the bytes, 20.7 us at 8 GPUs).  The price is the LSB of the
// transmitted values (truncation by at most one ulp, 6e-8 relative — the statistics are fp32 sums whose own rounding
// noise is larger; every replica stores the SAME truncated rows, so the codebooks stay bit-identical across ranks).
// No barriers, no fences, two one-way hops:
//   phase 1  every rank PUSHES the partial sums/counts of the rows it does not own into the owner's staging area
//   phase 2  the owner polls its staging area (local memory), reduces in fixed rank order, applies the k-means/EMA
//            update and pushes the new rows into every peer's row staging (its own codebook is written directly)
//   phase 3  every rank polls its row staging and writes the rows it does not own into its codebook
// Staging is single-buffered: a rank cannot start exchange e+1 before every owner has consumed its words of exchange e
// (phase 3 needs all owners' rows, which they send after consuming).  All blocks of the grid must be co-resident
// (a block spins on words that remote blocks produce); the host sizes the grid accordingly.
__device__ __forceinline__ uint32_t ll_word(float v, uint32_t e) { return (__float_as_uint(v) & ~1u) | (e & 1u); }
__device__ __forceinline__ void st_ll(uint32_t* p, float v, uint32_t e) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(ll_word(v, e)) : "memory");
}
__device__ __forceinline__ uint32_t ld_ll_raw(const uint32_t* p) {
  uint32_t w;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(w) : "l"(p) : "memory");
  return w;
}
// value of a low-latency word: `w` is what a first (batched) load returned; re-polls only if it had not arrived yet
__device__ __forceinline__ float ld_ll(const uint32_t* p, uint32_t w, uint32_t e, const Comm& c) {
  if (((w ^ e) & 1u) != 0u) {
    const long long t0 = clock64();
    do {
      w = ld_ll_raw(p);
      if (clock64() - t0 > 4000000000ll) {
        printf("vqb comm: rank %d timed out polling a low-latency word (exchange %u)\n", c.rank, e);
        __trap();
      }
    } while (((w ^ e) & 1u) != 0u);
  }
  return __uint_as_float(w & ~1u);
}

template <int G, int NPL>
__global__ void __launch_bounds__(256) comm_kmeans_ema_ll_kernel(Comm c, size_t stats_off, size_t w_off, size_t in_off,
                                                                 size_t out_off, int64_t K, int D, float decay, float omd) {
  __shared__ uint32_t s_epoch;
  pdl_wait();
  pdl_launch_dependents();
  CommCtl* ctl = c.ctl();
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(&ctl->ll_epoch) + 1;
  if (threadIdx.x < c.world) peer_table()[threadIdx.x] = reinterpret_cast<char* const*>(c.base + kPeerTableOff)[threadIdx.x];
  __syncthreads();
  const uint32_t e = s_epoch;
  const int per = (int)((K + c.world - 1) / c.world);   // rows per owner
  const int W1 = D + 1;                                  // staged words per row: D sums + the count
  const int lo = min((int)K, per * c.rank), hi = min((int)K, lo + per);
  const float* __restrict__ stats = reinterpret_cast<const float*>(c.base + stats_off);
  float* __restrict__ Wl = reinterpret_cast<float*>(c.base + w_off);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;

  // ---- phase 1: push my partials to the owners ----
  for (int idx0 = tid; idx0 < (int)K * W1; idx0 += 4 * nthreads) {   // 4 independent local loads in flight per thread
    float v[4];
    uint32_t* dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = idx0 + u * nthreads;
      dst[u] = nullptr;
      if (idx < (int)K * W1) {
        const int k = idx / W1, j = idx - k * W1;
        const int owner = k / per;
        if (owner != c.rank) {
          v[u] = j < D ? stats[(int64_t)k * D + j] : stats[K * (int64_t)D + k];
          dst[u] = reinterpret_cast<uint32_t*>(c.peer(owner) + in_off) + ((int64_t)c.rank * per + (k - owner * per)) * W1 + j;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (dst[u] != nullptr) st_ll(dst[u], v[u], e);
  }

  // ---- phase 2: reduce my rows in fixed rank order, update, publish ----
  const int lane = threadIdx.x % G;
  const int rows_per_block = blockDim.x / G;
  const uint32_t* __restrict__ stage_in = reinterpret_cast<const uint32_t*>(c.base + in_off);
  for (int base = lo + blockIdx.x * rows_per_block; base < hi; base += gridDim.x * rows_per_block) {
    const int k_raw = base + threadIdx.x / G;
    const bool valid = k_raw < hi;
    const int k = valid ? k_raw : hi - 1;
    float cnt = 0.f, s[NPL], w[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) s[j] = 0.f;
    constexpr int B = NPL <= 4 ? kBatch : (NPL <= 8 ? 4 : (NPL <= 16 ? 2 : 1));
    for (int r0 = 0; r0 < c.world; r0 += B) {          // every poll of a batch is in flight before the first check
      uint32_t wc[B], wv[B][NPL];
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const int r = r0 + b;
        if (r < c.world && r != c.rank) {
          const uint32_t* row = stage_in + ((int64_t)r * per + (k - lo)) * W1;
          wc[b] = ld_ll_raw(row + D);
#pragma unroll
          for (int j = 0; j < NPL; ++j) {
            const int d = lane + j * G;
            wv[b][j] = d < D ? ld_ll_raw(row + d) : 0u;
          }
        }
      }
#pragma unroll
      for (int b = 0; b < B; ++b) {                      // fixed rank order
        const int r = r0 + b;
        if (r >= c.world) continue;
        if (r == c.rank) {
          cnt += stats[K * (int64_t)D + k];
#pragma unroll
          for (int j = 0; j < NPL; ++j) {
            const int d = lane + j * G;
            if (d < D) s[j] += stats[(int64_t)k * D + d];
          }
        } else {
          const uint32_t* row = stage_in + ((int64_t)r * per + (k - lo)) * W1;
          cnt += ld_ll(row + D, wc[b], e, c);
#pragma unroll
          for (int j = 0; j < NPL; ++j) {
            const int d = lane + j * G;
            if (d < D) s[j] += ld_ll(row + d, wv[b][j], e, c);
          }
        }
      }
    }
    const bool occurred = cnt > 0.f;
    const float den = fmaxf(cnt, 1.f);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int d = lane + j * G;
      w[j] = d < D ? Wl[(int64_t)k * D + d] : 0.f;
      s[j] = d < D ? (occurred ? __fdiv_rn(s[j], den) : w[j]) : 0.f;
      ss = fmaf(s[j], s[j], ss);
    }
    ss = group_sum<G>(ss);
    const float dn = fmaxf(sqrtf(ss), kNormEps);
    float ss2 = 0.f;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      s[j] = __fadd_rn(__fmul_rn(w[j], decay), __fmul_rn(__fdiv_rn(s[j], dn), omd));
      ss2 = fmaf(s[j], s[j], ss2);
    }
    ss2 = group_sum<G>(ss2);
    const float dn2 = fmaxf(sqrtf(ss2), kNormEps);
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int d = lane + j * G;
      if (valid && d < D) {
        // the row every replica stores: the LSB is the wire flag, so it is cleared here as well (bit-identical replicas)
        const float out = __uint_as_float(__float_as_uint(__fdiv_rn(s[j], dn2)) & ~1u);
        Wl[(int64_t)k * D + d] = out;
        for (int r = 0; r < c.world; ++r)
          if (r != c.rank) st_ll(reinterpret_cast<uint32_t*>(c.peer(r) + out_off) + (int64_t)k * D + d, out, e);
      }
    }
  }

  // ---- phase 3: collect the rows of the other owners ----
  const uint32_t* __restrict__ stage_out = reinterpret_cast<const uint32_t*>(c.base + out_off);
  const int mine0 = lo * D, mine1 = hi * D;
  for (int idx0 = tid; idx0 < (int)K * D; idx0 += 4 * nthreads) {
    uint32_t w4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = idx0 + u * nthreads;
      const bool on = idx < (int)K * D && !(idx >= mine0 && idx < mine1);
      w4[u] = on ? ld_ll_raw(stage_out + idx) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = idx0 + u * nthreads;
      if (idx < (int)K * D && !(idx >= mine0 && idx < mine1)) Wl[idx] = ld_ll(stage_out + idx, w4[u], e, c);
    }
  }

  // ---- retire the epoch (last block) ----
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctl->ticket, 1u) == gridDim.x - 1) {
      ctl->ticket = 0;
      *reinterpret_cast<volatile uint32_t*>(&ctl->ll_epoch) = e;
      __threadfence();
    }
  }
}

// ---- CVQ-VAE: all-reduce of the usage counts + anchors fused with the probability EMA and the anchor blend ---
// vq/algorithms/cvqvae/quantizer_callback.py:88-103, anchors.py:50-67.
//   keys_off == SIZE_MAX (sync=False): anchors = mean over ranks of the per-rank nearest-token rows (anchors.py:64-67)
//   otherwise (sync=True): every rank holds its best (distance, global token index) key per code and that token's
//   row; the global nearest is the minimum key, its row is taken from the rank that owns it (anchors.py:50-57).
__global__ void __launch_bounds__(256) comm_cvq_update_kernel(Comm c, size_t counts_off, size_t anchors_off, size_t keys_off,
                                                              size_t w_off, size_t prob_off, int64_t K, int D, float decay,
                                                              float omd, float eps) {
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t e = comm_begin(c);
  int64_t lo, hi;
  slice_of(K, c, lo, hi);
  long long total_i = 0;
  for (int r0 = 0; r0 < c.world; r0 += kBatch) {
    unsigned long long v[kBatch];
#pragma unroll
    for (int b = 0; b < kBatch; ++b)
      v[b] = r0 + b < c.world ? ld_peer_u64(reinterpret_cast<const unsigned long long*>(c.peer(r0 + b) + counts_off) + K) : 0ull;
#pragma unroll
    for (int b = 0; b < kBatch; ++b) total_i += (long long)v[b];
  }
  const float total = (float)total_i;
  const bool minloc = keys_off != (size_t)-1;
  const float anchor_scale = minloc ? 1.f : 1.f / (float)c.world;
  const float* __restrict__ Wl = reinterpret_cast<const float*>(c.base + w_off);
  const float* __restrict__ Pl = reinterpret_cast<const float*>(c.base + prob_off);
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = lo + warp; k < hi; k += nwarps) {
    long long cnt = 0;
    unsigned long long best = ~0ull;
    int winner = -1;
    for (int r0 = 0; r0 < c.world; r0 += kBatch) {
      unsigned long long v[kBatch], kk[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        const bool on = r0 + b < c.world;
        v[b] = on ? ld_peer_u64(reinterpret_cast<const unsigned long long*>(c.peer(r0 + b) + counts_off) + k) : 0ull;
        kk[b] = (on && minloc) ? ld_peer_u64(reinterpret_cast<const unsigned long long*>(c.peer(r0 + b) + keys_off) + k) : ~0ull;
      }
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        cnt += (long long)v[b];
        if (kk[b] < best) { best = kk[b]; winner = r0 + b; }
      }
    }
    const float freq = __fdiv_rn((float)cnt, total);
    const float p = __fadd_rn(__fmul_rn(Pl[k], decay), __fmul_rn(freq, omd));
    float t = __fmul_rn(__fmul_rn(-p, (float)K), 10.f);
    t = __fsub_rn(__fdiv_rn(t, omd), eps);
    const float dec = __fsub_rn(1.f, expf(t));
    const float omdec = __fsub_rn(1.f, dec);
    __syncwarp();
    if (lane == 0)
      for (int r = 0; r < c.world; ++r) reinterpret_cast<float*>(c.peer(r) + prob_off)[k] = p;
    // In fp32 `dec` rounds to exactly 1 once p*K*10/(1-decay) exceeds ~17.3 (any code used at > 2 % of the uniform rate):
    // the blend is then W*1 + anchor*0 = W bit for bit, so neither the peers' anchor rows are read nor the unchanged row
    // is published — in a healthy codebook the 8 MB anchor exchange shrinks to the rows of the rarely used codes.
    if (omdec == 0.f) continue;
    for (int d = lane; d < D; d += 32) {
      float a = 0.f;
      if (minloc) {
        if (winner >= 0) a = ld_peer(reinterpret_cast<const float*>(c.peer(winner) + anchors_off) + k * D + d);
      } else {
        for (int r0 = 0; r0 < c.world; r0 += kBatch) {
          float v[kBatch];
#pragma unroll
          for (int b = 0; b < kBatch; ++b)
            v[b] = r0 + b < c.world ? ld_peer(reinterpret_cast<const float*>(c.peer(r0 + b) + anchors_off) + k * D + d) : 0.f;
#pragma unroll
          for (int b = 0; b < kBatch; ++b) a += v[b];
        }
      }
      a *= anchor_scale;
      const float out = __fadd_rn(__fmul_rn(Wl[k * D + d], dec), __fmul_rn(a, omdec));
      for (int r = 0; r < c.world; ++r) reinterpret_cast<float*>(c.peer(r) + w_off)[k * D + d] = out;
    }
  }
  comm_end(c, e);
}

// ---- packed (score, index) min-loc all-reduce over [n] keys (codebook shards) -------------------
__global__ void __launch_bounds__(256) comm_min_keys_kernel(Comm c, size_t keys_off, int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t e = comm_begin(c);
  int64_t lo, hi;
  slice_of(n, c, lo, hi);
  for (int64_t i = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long m = ~0ull;
    for (int r0 = 0; r0 < c.world; r0 += kBatch) {
      unsigned long long v[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; ++b)
        v[b] = r0 + b < c.world ? ld_peer_u64(reinterpret_cast<const unsigned long long*>(c.peer(r0 + b) + keys_off) + i) : ~0ull;
#pragma unroll
      for (int b = 0; b < kBatch; ++b) m = v[b] < m ? v[b] : m;
    }
    for (int r = 0; r < c.world; ++r) reinterpret_cast<unsigned long long*>(c.peer(r) + keys_off)[i] = m;
  }
  comm_end(c, e);
}

// ---- plain fp32 SUM all-reduce (k-means init rounds, CachedAnchor means) ---------------------------------------
__global__ void __launch_bounds__(256) comm_sum_f32_kernel(Comm c, size_t off, int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t e = comm_begin(c);
  int64_t lo, hi;
  slice_of((n + 3) / 4, c, lo, hi);   // in units of float4 (the buffer is padded to 16 B by the host layer)
  for (int64_t i = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = 0; r0 < c.world; r0 += kBatch) {
      float4 v[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; ++b)
        v[b] = r0 + b < c.world ? ld_peer4(reinterpret_cast<const float*>(c.peer(r0 + b) + off) + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int b = 0; b < kBatch; ++b) { s.x += v[b].x; s.y += v[b].y; s.z += v[b].z; s.w += v[b].w; }
    }
    for (int r = 0; r < c.world; ++r) reinterpret_cast<float4*>(c.peer(r) + off)[i] = s;
  }
  comm_end(c, e);
}

static inline int pow2_lanes_c(int n) {
  int g = 1;
  while (g < 32 && g < n) g <<= 1;
  return g;
}
static inline int blocks_for(int64_t items, int items_per_block) {
  int64_t b = (items + items_per_block - 1) / items_per_block;
  // The exchange is latency-bound (one NVLink round trip per loop iteration of a block): enough co-resident blocks
  // that every block makes ONE trip.  VQB_COMM_BLOCKS_PER_SM: developer knob.
  static int per_sm = 0;
  if (per_sm == 0) {
    const char* e = getenv("VQB_COMM_BLOCKS_PER_SM");
    per_sm = e ? atoi(e) : 4;
    if (per_sm < 1) per_sm = 1;
  }
  const int64_t cap = (int64_t)sm_count() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace vqb

using namespace vqb;

#define VQB_COMM_ARGS_OK(name)                                                                    \
  VQB_REQUIRE(region != nullptr, name ": null region");                                           \
  VQB_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, name ": bad rank/world")

extern "C" {

int vqb_comm_alloc(size_t bytes, void** region, unsigned char* handle_out) {
  VQB_REQUIRE(region && handle_out && bytes >= VQB_COMM_HEADER_BYTES, "vqb_comm_alloc: bad arguments");
  void* p = nullptr;
  VQB_CUDA_OK(cudaMalloc(&p, bytes));
  VQB_CUDA_OK(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == VQB_IPC_HANDLE_BYTES, "ipc handle size");
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return VQB_ERR_CUDA;
  }
  memcpy(handle_out, &h, sizeof(h));
  VQB_CUDA_OK(cudaDeviceSynchronize());
  *region = p;
  return VQB_OK;
}

int vqb_comm_open(const unsigned char* handle, void** peer_region) {
  VQB_REQUIRE(handle && peer_region, "vqb_comm_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  VQB_CUDA_OK(cudaIpcOpenMemHandle(peer_region, h, cudaIpcMemLazyEnablePeerAccess));
  return VQB_OK;
}

int vqb_comm_close(void* peer_region) {
  if (peer_region) VQB_CUDA_OK(cudaIpcCloseMemHandle(peer_region));
  return VQB_OK;
}

int vqb_comm_free(void* region) {
  if (region) VQB_CUDA_OK(cudaFree(region));
  return VQB_OK;
}

int vqb_comm_bind(void* region, const void* const* peer_regions_host, int rank, int world) {
  VQB_COMM_ARGS_OK("vqb_comm_bind");
  VQB_REQUIRE(peer_regions_host != nullptr && peer_regions_host[rank] == region, "vqb_comm_bind: peer table must hold the local region at [rank]");
  VQB_CUDA_OK(cudaMemcpy(static_cast<char*>(region) + kPeerTableOff, peer_regions_host, sizeof(void*) * world,
                         cudaMemcpyHostToDevice));
  VQB_CUDA_OK(cudaDeviceSynchronize());
  return VQB_OK;
}

int vqb_comm_kmeans_ema_update(void* region, int rank, int world, size_t stats_off, size_t w_off, size_t ll_in_off,
                               size_t ll_out_off, int64_t K, int D, float decay, float one_minus_decay, void* stream) {
  VQB_COMM_ARGS_OK("vqb_comm_kmeans_ema_update");
  VQB_REQUIRE(K >= 1 && D >= 1 && stats_off % 16 == 0 && w_off % 16 == 0, "vqb_comm_kmeans_ema_update: bad shape/offset");
  Comm c{static_cast<char*>(region), rank, world};
  VQB_REQUIRE(D <= 1024, "vqb_comm_kmeans_ema_update: D <= 1024");
  const bool ll = ll_in_off != (size_t)-1 && ll_out_off != (size_t)-1;
  VQB_REQUIRE(!ll || (ll_in_off % 4 == 0 && ll_out_off % 4 == 0 && K * (int64_t)(D + 1) < (1ll << 31)),
              "vqb_comm_kmeans_ema_update: bad low-latency staging");
  const int g = pow2_lanes_c(D);           // one lane per element up to 32: the exchange is latency-bound, go wide
  int npl = 1;
  while (g * npl < D) npl <<= 1;
  int blocks = blocks_for((K + world - 1) / world, 256 / g);
  if (ll) {   // every block must be resident (it spins on words produced by REMOTE blocks): 4 x 256 threads per SM
    blocks = sm_count() * 4;
  }
#define LAUNCH(G_, NPL_)                                                                                                   \
  if (ll)                                                                                                                  \
    VQB_CUDA_OK(launch_pdl(comm_kmeans_ema_ll_kernel<G_, NPL_>, blocks, 256, 0, (cudaStream_t)stream, c, stats_off, w_off, \
                           ll_in_off, ll_out_off, K, D, decay, one_minus_decay));                                         \
  else                                                                                                                     \
    VQB_CUDA_OK(launch_pdl(comm_kmeans_ema_kernel<G_, NPL_>, blocks, 256, 0, (cudaStream_t)stream, c, stats_off, w_off, K, D, \
                           decay, one_minus_decay))
  if (npl == 1) {
    switch (g) {
      case 1: LAUNCH(1, 1); break;
      case 2: LAUNCH(2, 1); break;
      case 4: LAUNCH(4, 1); break;
      case 8: LAUNCH(8, 1); break;
      case 16: LAUNCH(16, 1); break;
      default: LAUNCH(32, 1); break;
    }
  } else {
    switch (npl) {
      case 2: LAUNCH(32, 2); break;
      case 4: LAUNCH(32, 4); break;
      case 8: LAUNCH(32, 8); break;
      case 16: LAUNCH(32, 16); break;
      default: LAUNCH(32, 32); break;
    }
  }
#undef LAUNCH
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_comm_cvq_update(void* region, int rank, int world, size_t counts_off, size_t anchors_off, size_t keys_off,
                        size_t w_off, size_t prob_off, int64_t K, int D, float decay, float one_minus_decay, float eps,
                        void* stream) {
  VQB_COMM_ARGS_OK("vqb_comm_cvq_update");
  VQB_REQUIRE(K >= 1 && D >= 1, "vqb_comm_cvq_update: bad shape");
  Comm c{static_cast<char*>(region), rank, world};
  const int blocks = blocks_for((K + world - 1) / world, 8);   // one warp per code row
  VQB_CUDA_OK(launch_pdl(comm_cvq_update_kernel, blocks, 256, 0, (cudaStream_t)stream, c, counts_off, anchors_off, keys_off,
                         w_off, prob_off, K, D, decay, one_minus_decay, eps));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_comm_allreduce_min_keys(void* region, int rank, int world, size_t keys_off, int64_t n, void* stream) {
  VQB_COMM_ARGS_OK("vqb_comm_allreduce_min_keys");
  VQB_REQUIRE(n >= 1 && keys_off % 8 == 0, "vqb_comm_allreduce_min_keys: bad size/offset");
  Comm c{static_cast<char*>(region), rank, world};
  const int blocks = blocks_for((n + world - 1) / world, 256);
  VQB_CUDA_OK(launch_pdl(comm_min_keys_kernel, blocks, 256, 0, (cudaStream_t)stream, c, keys_off, n));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

int vqb_comm_allreduce_sum_f32(void* region, int rank, int world, size_t off, int64_t n, void* stream) {
  VQB_COMM_ARGS_OK("vqb_comm_allreduce_sum_f32");
  VQB_REQUIRE(n >= 1 && off % 16 == 0, "vqb_comm_allreduce_sum_f32: bad size/offset");
  Comm c{static_cast<char*>(region), rank, world};
  const int blocks = blocks_for(((n + 3) / 4 + world - 1) / world, 256);
  VQB_CUDA_OK(launch_pdl(comm_sum_f32_kernel, blocks, 256, 0, (cudaStream_t)stream, c, off, n));
  VQB_LAUNCH_OK();
  return VQB_OK;
}

}  // extern "C"
