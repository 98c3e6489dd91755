#include <cstdint>
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__global__ void k(const float* in, const float4* w, float* out) {
  unsigned long long acc[4][2] = {};
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = in[threadIdx.x * 8 + i];
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    const float4 wv = w[t];
    const unsigned long long w0 = pack2(wv.x, wv.y), w1 = pack2(wv.z, wv.w);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned long long vv = pack2(v[q + t % 4], v[q + t % 4]);
      fma2(acc[q][0], vv, w0);
      fma2(acc[q][1], vv, w1);
    }
  }
  for (int q = 0; q < 4; ++q) for (int j = 0; j < 2; ++j) {
    float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[q][j]));
    out[(threadIdx.x * 4 + q) * 4 + 2 * j] = a; out[(threadIdx.x * 4 + q) * 4 + 2 * j + 1] = b;
  }
}
