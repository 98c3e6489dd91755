// Reciprocal throughput of the fp32 FMA forms conv1 of P-Net can be built from, per SM sub-partition, on sm_100a.
// One CTA per launch on one SM, W warps (W / 4 per sub-partition), every thread runs ITER x 16 independent-chain FMAs.
// Prints cycles per warp-instruction per sub-partition:  cycles * 4 / (W * n_instr), cycles = first start to last end over all warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 512
struct KP { float w[64]; };

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(const float* in, float* out, long long* cyc, const __grid_constant__ KP kp) {
  float a[16];
  unsigned long long d[16];
  for (int i = 0; i < 16; ++i) { a[i] = in[threadIdx.x + 32 * i]; d[i] = pack2(a[i], a[i] + 1.f); }
  __shared__ unsigned sm[1024];
  unsigned q[16];
  for (int i = 0; i < 16; ++i) q[i] = threadIdx.x * 17 + i;
  sm[threadIdx.x] = threadIdx.x;
  float x = in[threadIdx.x + 7], y = in[threadIdx.x + 9];
  unsigned long long xx = pack2(x, y), yy = pack2(y, x);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if (MODE == 0) {              // FFMA, three register operands
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
    } else if (MODE == 1) {       // FFMA2, three 64-bit register operands
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(xx), "l"(yy));
    } else if (MODE == 2) {       // FFMA with a constant-bank multiplicand (kernel parameter, compile-time index)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], kp.w[i], y);
    } else if (MODE == 3) {       // FFMA accumulate form: d = x * c[] + d  (conv1's shape: acc += pixel * weight)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(x, kp.w[i], a[i]);
    } else if (MODE == 4) {       // FFMA2 accumulate form, all registers: d = xx * ww + d
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d[i]) : "l"(xx), "l"(yy));
    } else if (MODE == 5) {       // FFMA accumulate form, registers only
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(x, y, a[i]);
    } else if (MODE == 6) {       // FFMA with an immediate
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], 1.0009765625f, y);
    } else if (MODE == 8) {       // FFMA2 interleaved 1:1 with integer ALU work (does the packed FMA block the issue port?)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d[i]) : "l"(xx), "l"(yy));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(q[(i + 1) & 15]), "r"(it));
      }
    } else if (MODE == 9) {       // FFMA interleaved 1:1 with integer ALU work
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        a[i] = fmaf(x, y, a[i]);
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(q[(i + 1) & 15]), "r"(it));
      }
    } else if (MODE == 10) {      // FFMA2 interleaved 2:1 with shared-memory loads
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d[i]) : "l"(xx), "l"(yy));
        if (i & 1) q[i] += sm[(threadIdx.x + 32 * i + it) & 1023];
      }
    } else if (MODE == 7) {       // HFMA2 (two fp16 per lane)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        unsigned r = (unsigned)d[i], p = (unsigned)xx, q = (unsigned)yy;
        asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(r) : "r"(p), "r"(q));
        d[i] = r;
      }
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) { atomicMin((unsigned long long*)(cyc + 1), (unsigned long long)t0); atomicMax((unsigned long long*)(cyc + 2), (unsigned long long)t1); }
  float s = 0.f;
  for (int i = 0; i < 16; ++i) { s += a[i]; s += (float)(d[i] & 0xffff); s += (float)(q[i] & 3); }
  out[threadIdx.x] = s + x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 1 << 20); cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 24);
  cudaMemset(in, 0, 1 << 20);
  KP kp; for (int i = 0; i < 64; ++i) kp.w[i] = 1.f + i * 1e-3f;
  const char* names[11] = {"FFMA  r,r,r (d=a*x+y)", "FFMA2 r,r,r (d=d*xx+yy)", "FFMA  r,c[],r (d=a*c+y)", "FFMA  r,c[],acc (d=x*c+d)",
                          "FFMA2 acc (d=xx*yy+d)", "FFMA  acc (d=x*y+d)", "FFMA  r,imm,r", "HFMA2 acc", "FFMA2 + LOP3 1:1 (pairs)", "FFMA + LOP3 1:1 (pairs)", "FFMA2 + LDS 2:1 (per FFMA2)"};
  for (int mode = 0; mode < 11; ++mode)
    for (int warps : {4, 8, 16, 32}) {
      long long h = 0, hh[3];
      for (int rep = 0; rep < 2; ++rep) {
        const long long init[3] = {0, 0x7fffffffffffffffLL, 0};
        cudaMemcpy(cyc, init, 24, cudaMemcpyHostToDevice);
        switch (mode) {
          case 0: k<0><<<1, 32 * warps>>>(in, out, cyc, kp); break;
          case 1: k<1><<<1, 32 * warps>>>(in, out, cyc, kp); break;
          case 2: k<2><<<1, 32 * warps>>>(in, out, cyc, kp); break;
          case 3: k<3><<<1, 32 * warps>>>(in, out, cyc, kp); break;
          case 4: k<4><<<1, 32 * warps>>>(in, out, cyc, kp); break;
          case 5: k<5><<<1, 32 * warps>>>(in, out, cyc, kp); break;
          case 6: k<6><<<1, 32 * warps>>>(in, out, cyc, kp); break;
          case 7: k<7><<<1, 32 * warps>>>(in, out, cyc, kp); break;
          case 8: k<8><<<1, 32 * warps>>>(in, out, cyc, kp); break;
          case 9: k<9><<<1, 32 * warps>>>(in, out, cyc, kp); break;
          default: k<10><<<1, 32 * warps>>>(in, out, cyc, kp); break;
        }
        cudaDeviceSynchronize();
        cudaMemcpy(hh, cyc, 24, cudaMemcpyDeviceToHost);
        h = hh[2] - hh[1];                     // first warp's start to last warp's end
      }
      const double n = (double)ITER * 16;
      printf("%-28s warps/SMSP %d: %.2f cycles per warp-instruction per SMSP (%.0f cycles total)%s\n", names[mode], warps / 4,
             (double)h * 4 / (warps * n), (double)h, cudaGetLastError() == cudaSuccess ? "" : " ERR");
    }
  return 0;
}
