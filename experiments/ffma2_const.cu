#include <cstdint>
struct P { float w[27][12]; };
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__global__ void k(const float* in, float* out, const __grid_constant__ P p) {
  unsigned long long acc[4][5] = {};
  float v[30];
  for (int i = 0; i < 30; ++i) v[i] = in[threadIdx.x * 30 + i];
#pragma unroll
  for (int t = 0; t < 27; ++t) {
    unsigned long long w[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) w[c] = pack2(p.w[t][2 * c], p.w[t][2 * c + 1]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned long long vv = pack2(v[q + t], v[q + t]);
#pragma unroll
      for (int c = 0; c < 5; ++c) fma2(acc[q][c], vv, w[c]);
    }
  }
  for (int q = 0; q < 4; ++q) for (int j = 0; j < 5; ++j) {
    float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[q][j]));
    out[(threadIdx.x * 4 + q) * 10 + 2 * j] = a; out[(threadIdx.x * 4 + q) * 10 + 2 * j + 1] = b;
  }
}
