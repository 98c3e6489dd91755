// Throughput of legacy mma.sync (HMMA.16816.F32) on sm_100a, alone and next to an FFMA2 stream, as P-Net's conv2 (group B,
// 8 warps, 6 accumulator chains per warp) runs next to conv1 (group A, 8 warps of FFMA2).  One CTA on one SM.
// Prints cycles per HMMA at SM level (first start to last end over the HMMA warps).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 256

__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// warps [0, n_fma): FFMA2 stream (16 chains); warps [n_fma, n_fma + n_mma): HMMA with CHAINS independent accumulators
template <int CHAINS>
__global__ void __launch_bounds__(1024, 1) k(const float* in, float* out, unsigned long long* cyc, int n_fma, int n_mma, int with_lds) {
  __shared__ uint32_t sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * 2654435761u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float s = 0.f;
  __syncthreads();
  if (warp < n_fma) {
    unsigned long long d[16], xx, yy;
    const float x = in[lane], y = in[lane + 32];
    asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x), "f"(y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(yy) : "f"(y), "f"(x));
    for (int i = 0; i < 16; ++i) d[i] = xx + i;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER * 4; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d[i]) : "l"(xx), "l"(yy));
    }
    const long long t1 = clock64();
    if (lane == 0) { atomicMin(cyc + 3, (unsigned long long)t0); atomicMax(cyc + 4, (unsigned long long)t1); }
    for (int i = 0; i < 16; ++i) s += (float)(d[i] & 0xff);
  } else if (warp < n_fma + n_mma) {
    float acc[CHAINS][4];
    for (int c = 0; c < CHAINS; ++c) for (int e = 0; e < 4; ++e) acc[c][e] = 0.f;
    uint32_t a[4] = {sm[lane], sm[lane + 32], sm[lane + 64], sm[lane + 96]}, b0 = sm[lane + 128], b1 = sm[lane + 160];
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
      if (with_lds) {                      // 8 fragment loads per 6 HMMAs, like conv2
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = sm[(lane * 9 + it * 4 + q) & 2047];
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) mma(acc[c], a[0], a[1], a[2], a[3], b0, b1);
    }
    const long long t1 = clock64();
    if (lane == 0) { atomicMin(cyc + 1, (unsigned long long)t0); atomicMax(cyc + 2, (unsigned long long)t1); }
    for (int c = 0; c < CHAINS; ++c) s += acc[c][0] + acc[c][3];
  }
  out[threadIdx.x] = s;
}

int main() {
  float *in, *out; unsigned long long* cyc;
  cudaMalloc(&in, 1 << 16); cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 64);
  cudaMemset(in, 0, 1 << 16);
  struct Case { int n_fma, n_mma, lds, chains; };
  const Case cases[] = {{0, 4, 0, 6}, {0, 8, 0, 6}, {0, 16, 0, 6}, {0, 8, 0, 2}, {0, 8, 0, 12}, {0, 8, 1, 6}, {8, 8, 0, 6}, {8, 8, 1, 6}, {8, 8, 0, 12}, {8, 0, 0, 6}};
  for (const Case& c : cases) {
    unsigned long long h[5];
    for (int rep = 0; rep < 2; ++rep) {
      const unsigned long long init[5] = {0, ~0ull, 0, ~0ull, 0};
      cudaMemcpy(cyc, init, 40, cudaMemcpyHostToDevice);
      const int threads = 32 * (c.n_fma + c.n_mma);
      if (c.chains == 2) k<2><<<1, threads>>>(in, out, cyc, c.n_fma, c.n_mma, c.lds);
      else if (c.chains == 12) k<12><<<1, threads>>>(in, out, cyc, c.n_fma, c.n_mma, c.lds);
      else k<6><<<1, threads>>>(in, out, cyc, c.n_fma, c.n_mma, c.lds);
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, 40, cudaMemcpyDeviceToHost);
    }
    const double n_h = (double)ITER * 3 * c.chains * c.n_mma, n_f = (double)ITER * 4 * 16 * c.n_fma;
    printf("FFMA2 warps %d, HMMA warps %2d (chains %2d, lds %d): ", c.n_fma, c.n_mma, c.chains, c.lds);
    if (c.n_mma) printf("%.2f cycles per HMMA per SM (%.2f per warp)  ", (double)(h[2] - h[1]) / n_h, (double)(h[2] - h[1]) / (n_h / c.n_mma));
    if (c.n_fma) printf("| %.2f cycles per FFMA2 per SMSP", (double)(h[4] - h[3]) * 4 / n_f);
    printf("%s\n", cudaGetLastError() == cudaSuccess ? "" : " ERR");
  }
  return 0;
}
