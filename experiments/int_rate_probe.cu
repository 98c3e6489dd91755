// Reciprocal throughput of the integer / conversion instructions the pyramid kernel is built from (sm_100a), per SM
// sub-partition: which of them share the half-rate ALU pipe, and whether integer adds issued as IMAD (FMA pipe) run beside them.
// One CTA on one SM, W warps, every thread runs ITER x 16 independent chains.  Prints cycles per warp-instruction per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITER 512

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(const unsigned* in, unsigned* out, long long* cyc) {
  unsigned q[16], r[16];
  float f[16];
  for (int i = 0; i < 16; ++i) { q[i] = in[threadIdx.x + 32 * i] + i; r[i] = q[i] * 3 + 1; f[i] = (float)q[i]; }
  const unsigned x = in[threadIdx.x + 7] | 1, y = in[threadIdx.x + 9] + 2;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(q[i]) : "r"(x));                                  // IADD3
      if (MODE == 1) asm volatile("prmt.b32 %0, %0, %1, 0x4341;" : "+r"(q[i]) : "r"(x));                        // PRMT
      if (MODE == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(x), "r"(y));              // LOP3
      if (MODE == 3) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(q[i]) : "r"(x), "r"(y));                  // IMAD
      if (MODE == 4) {                                                                                           // PRMT + IADD3
        asm volatile("prmt.b32 %0, %0, %1, 0x4341;" : "+r"(q[i]) : "r"(x));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(y));
      }
      if (MODE == 5) {                                                                                           // PRMT + IMAD
        asm volatile("prmt.b32 %0, %0, %1, 0x4341;" : "+r"(q[i]) : "r"(x));
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(x), "r"(y));
      }
      if (MODE == 6) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[i]) : "r"(q[i] = q[i] + (unsigned)f[i]));    // I2F (+ F2I + IADD)
      if (MODE == 7) {                                                                                           // F2FP pack
        asm volatile("{ .reg .b32 t; cvt.rn.f16x2.f32 t, %0, %1; mov.b32 %0, t; }" : "+f"(f[i]) : "f"(f[(i + 1) & 15]));
      }
      if (MODE == 8) {                                                                                           // PRMT + 2 IMAD
        asm volatile("prmt.b32 %0, %0, %1, 0x4341;" : "+r"(q[i]) : "r"(x));
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(x), "r"(y));
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(q[(i + 8) & 15]) : "r"(y), "r"(x));
      }
      if (MODE == 9) {                                                                                           // PRMT + IADD3 + IADD3 (today's vertical pass mix)
        asm volatile("prmt.b32 %0, %0, %1, 0x4341;" : "+r"(q[i]) : "r"(x));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(y));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(q[(i + 8) & 15]) : "r"(x));
      }
      if (MODE == 10) asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(q[i]) : "r"(x));                           // what does ptxas make of x * 1 + acc ?
      if (MODE == 11) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 1) & 15]));                // FADD
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) { atomicMin((unsigned long long*)(cyc + 1), (unsigned long long)t0); atomicMax((unsigned long long*)(cyc + 2), (unsigned long long)t1); }
  unsigned s = 0;
  for (int i = 0; i < 16; ++i) s += q[i] + r[i] + (unsigned)f[i];
  out[threadIdx.x] = s;
}

int main() {
  unsigned *in, *out; long long* cyc;
  cudaMalloc(&in, 1 << 20); cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 24);
  cudaMemset(in, 1, 1 << 20);
  const char* names[12] = {"IADD", "PRMT", "LOP3", "IMAD r,r,acc", "PRMT + IADD (per pair)", "PRMT + IMAD (per pair)", "I2F + F2I + IADD (per triple)",
                          "F2FP.PACK", "PRMT + 2 IMAD (per triple)", "PRMT + 2 IADD (per triple)", "IMAD x*1+acc", "FADD"};
  const int ninstr[12] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
  for (int mode = 0; mode < 12; ++mode)
    for (int warps : {8, 16, 32}) {
      long long hh[3];
      for (int rep = 0; rep < 2; ++rep) {
        const long long init[3] = {0, 0x7fffffffffffffffLL, 0};
        cudaMemcpy(cyc, init, 24, cudaMemcpyHostToDevice);
        switch (mode) {
          case 0: k<0><<<1, 32 * warps>>>(in, out, cyc); break;
          case 1: k<1><<<1, 32 * warps>>>(in, out, cyc); break;
          case 2: k<2><<<1, 32 * warps>>>(in, out, cyc); break;
          case 3: k<3><<<1, 32 * warps>>>(in, out, cyc); break;
          case 4: k<4><<<1, 32 * warps>>>(in, out, cyc); break;
          case 5: k<5><<<1, 32 * warps>>>(in, out, cyc); break;
          case 6: k<6><<<1, 32 * warps>>>(in, out, cyc); break;
          case 7: k<7><<<1, 32 * warps>>>(in, out, cyc); break;
          case 8: k<8><<<1, 32 * warps>>>(in, out, cyc); break;
          case 9: k<9><<<1, 32 * warps>>>(in, out, cyc); break;
          case 10: k<10><<<1, 32 * warps>>>(in, out, cyc); break;
          default: k<11><<<1, 32 * warps>>>(in, out, cyc); break;
        }
        cudaDeviceSynchronize();
        cudaMemcpy(hh, cyc, 24, cudaMemcpyDeviceToHost);
      }
      const double n = (double)ITER * 16 * ninstr[mode];
      printf("%-32s warps/SMSP %d: %.2f cycles per SMSP per group%s\n", names[mode], warps / 4, (double)(hh[2] - hh[1]) * 4 / (warps * n),
             cudaGetLastError() == cudaSuccess ? "" : " ERR");
    }
  return 0;
}
