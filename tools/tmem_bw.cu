#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int X, int INF>
__global__ void __launch_bounds__(512, 1) bw_kernel(int iters, uint32_t* out, long long* cycles) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 512 / X; c += INF) {
      uint32_t r[INF][X];
#pragma unroll
      for (int f = 0; f < INF; ++f) {
        const uint32_t a = base + (c + f) * X;
                if constexpr (X == 4) { asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[f][0]), "=r"(r[f][1]), "=r"(r[f][2]), "=r"(r[f][3]) : "r"(a) : "memory"); }
        if constexpr (X == 8) { asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(r[f][0]), "=r"(r[f][1]), "=r"(r[f][2]), "=r"(r[f][3]), "=r"(r[f][4]), "=r"(r[f][5]), "=r"(r[f][6]), "=r"(r[f][7]) : "r"(a) : "memory"); }
        if constexpr (X == 16) { asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" : "=r"(r[f][0]), "=r"(r[f][1]), "=r"(r[f][2]), "=r"(r[f][3]), "=r"(r[f][4]), "=r"(r[f][5]), "=r"(r[f][6]), "=r"(r[f][7]), "=r"(r[f][8]), "=r"(r[f][9]), "=r"(r[f][10]), "=r"(r[f][11]), "=r"(r[f][12]), "=r"(r[f][13]), "=r"(r[f][14]), "=r"(r[f][15]) : "r"(a) : "memory"); }
        if constexpr (X == 32) { asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(r[f][0]), "=r"(r[f][1]), "=r"(r[f][2]), "=r"(r[f][3]), "=r"(r[f][4]), "=r"(r[f][5]), "=r"(r[f][6]), "=r"(r[f][7]), "=r"(r[f][8]), "=r"(r[f][9]), "=r"(r[f][10]), "=r"(r[f][11]), "=r"(r[f][12]), "=r"(r[f][13]), "=r"(r[f][14]), "=r"(r[f][15]), "=r"(r[f][16]), "=r"(r[f][17]), "=r"(r[f][18]), "=r"(r[f][19]), "=r"(r[f][20]), "=r"(r[f][21]), "=r"(r[f][22]), "=r"(r[f][23]), "=r"(r[f][24]), "=r"(r[f][25]), "=r"(r[f][26]), "=r"(r[f][27]), "=r"(r[f][28]), "=r"(r[f][29]), "=r"(r[f][30]), "=r"(r[f][31]) : "r"(a) : "memory"); }
        if constexpr (X == 64) { asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];" : "=r"(r[f][0]), "=r"(r[f][1]), "=r"(r[f][2]), "=r"(r[f][3]), "=r"(r[f][4]), "=r"(r[f][5]), "=r"(r[f][6]), "=r"(r[f][7]), "=r"(r[f][8]), "=r"(r[f][9]), "=r"(r[f][10]), "=r"(r[f][11]), "=r"(r[f][12]), "=r"(r[f][13]), "=r"(r[f][14]), "=r"(r[f][15]), "=r"(r[f][16]), "=r"(r[f][17]), "=r"(r[f][18]), "=r"(r[f][19]), "=r"(r[f][20]), "=r"(r[f][21]), "=r"(r[f][22]), "=r"(r[f][23]), "=r"(r[f][24]), "=r"(r[f][25]), "=r"(r[f][26]), "=r"(r[f][27]), "=r"(r[f][28]), "=r"(r[f][29]), "=r"(r[f][30]), "=r"(r[f][31]), "=r"(r[f][32]), "=r"(r[f][33]), "=r"(r[f][34]), "=r"(r[f][35]), "=r"(r[f][36]), "=r"(r[f][37]), "=r"(r[f][38]), "=r"(r[f][39]), "=r"(r[f][40]), "=r"(r[f][41]), "=r"(r[f][42]), "=r"(r[f][43]), "=r"(r[f][44]), "=r"(r[f][45]), "=r"(r[f][46]), "=r"(r[f][47]), "=r"(r[f][48]), "=r"(r[f][49]), "=r"(r[f][50]), "=r"(r[f][51]), "=r"(r[f][52]), "=r"(r[f][53]), "=r"(r[f][54]), "=r"(r[f][55]), "=r"(r[f][56]), "=r"(r[f][57]), "=r"(r[f][58]), "=r"(r[f][59]), "=r"(r[f][60]), "=r"(r[f][61]), "=r"(r[f][62]), "=r"(r[f][63]) : "r"(a) : "memory"); }

      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int f = 0; f < INF; ++f)
#pragma unroll
        for (int j = 0; j < X; ++j) acc ^= r[f][j];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512u) : "memory");
}
template <int X, int INF>
void run(int nwarps) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 200;
  bw_kernel<X, INF><<<148, nwarps * 32>>>(iters, out, cyc);
  bw_kernel<X, INF><<<148, nwarps * 32>>>(iters, out, cyc);
  cudaDeviceSynchronize();
  long long c0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
  const double bytes_per_sm = (double)iters * nwarps * 65536.0;
  printf("x%-3d inflight=%d warps=%2d  %8.1f B/cycle/SM  (%s)\n", X, INF, nwarps, bytes_per_sm / c0, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<4,1>(4);
  run<4,1>(8);
  run<4,1>(16);
  run<4,4>(4);
  run<4,4>(8);
  run<4,4>(16);
  run<8,1>(4);
  run<8,1>(8);
  run<8,1>(16);
  run<8,2>(4);
  run<8,2>(8);
  run<8,2>(16);
  run<8,4>(4);
  run<8,4>(8);
  run<8,4>(16);
  run<16,1>(4);
  run<16,1>(8);
  run<16,1>(16);
  run<16,2>(4);
  run<16,2>(8);
  run<16,2>(16);
  run<16,4>(4);
  run<16,4>(8);
  run<16,4>(16);
  run<32,1>(4);
  run<32,1>(8);
  run<32,1>(16);
  run<32,2>(4);
  run<32,2>(8);
  run<32,2>(16);
  run<64,1>(4);
  run<64,1>(8);
  run<64,1>(16);

  return 0;
}
