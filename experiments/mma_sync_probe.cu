// Probe: legacy mma.sync throughput on sm_100a (tf32 m16n8k8, bf16 m16n8k16) vs FFMA, per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, int iters) {
  float acc[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x ^ 5u, 11u};
  float fa = threadIdx.x * 1e-3f, fb = 1.0001f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else if (MODE == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(acc[i][j], fb, fa);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  CK(cudaMalloc(&out, 148 * 1024 * 4 * 2)); CK(cudaMalloc(&cyc, 148 * 8 * 2));
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int threads : {128, 256, 512, 1024}) {
      if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
      if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
      if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
      CK(cudaDeviceSynchronize());
      long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
      const double warps = threads / 32.0;
      const double ops = warps * iters * 8;          // warp-level instructions (mode 2: x4 FFMA)
      if (mode < 2) {
        const double mac = (mode == 0 ? 16.0 * 8 * 8 : 16.0 * 8 * 16);
        printf("%s warps/SM=%2.0f: %.2f cyc per warp-MMA per SM, %.0f MAC/clk/SM\n", mode == 0 ? "mma.sync tf32 m16n8k8 " : "mma.sync bf16 m16n8k16", warps,
               c / ops, ops * mac / c);
      } else
        printf("FFMA                   warps/SM=%2.0f: %.0f FMA/clk/SM\n", warps, ops * 4 * 32 / c);
    }
  return 0;
}
