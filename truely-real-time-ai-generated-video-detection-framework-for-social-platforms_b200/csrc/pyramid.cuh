// Pieces of the pyramid arithmetic shared by pyramid_sep_kernel (preproc.cu) and pyramid_stream_kernel (pyramid_stream.cu).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

constexpr int PYR_MAX_KH = 257;

__device__ __forceinline__ float div_small(float a, float b, float r) {
  const float q0 = __fmul_rn(a, r);
  const float e = __fmaf_rn(-q0, b, a);
  return __fmaf_rn(e, r, q0);
}

// Output formats of the pyramid kernels.  OUT = 0: planar fp32 [3][hs][pitch] (trl_pyramid, the 3-term P-Net).  OUT = 1: the
// all-tensor-pipe P-Net's input (pnet2.cu): per pixel 4 halves (B, G, R, 0) = 8 bytes in a "hi" image (fp16(v): the screen's
// operand) and a "lo" image (fp16(v - hi): with hi it restores v to 2^-23 relative for the exact re-evaluation,
// pnet_refine.cu); one 8-byte store each instead of three 4-byte planar stores.
template <int OUT> struct PyrOut;
template <> struct PyrOut<0> {
  typedef float T;
  static __device__ __forceinline__ void store(float* o, long long plane, float v0, float v1, float v2) {
    o[0] = v0; o[plane] = v1; o[2 * plane] = v2;
  }
};
template <> struct PyrOut<1> {
  typedef uint2 T;
  static __device__ __forceinline__ void store(uint2* o, long long lo_off, float v0, float v1, float v2) {
    const __half2 h01 = __floats2half2_rn(v0, v1);
    const __half2 h2 = __floats2half2_rn(v2, 0.f);
    const float2 f01 = __half22float2(h01);
    const float f2 = __low2float(h2);
    const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y);
    const __half2 l2 = __floats2half2_rn(v2 - f2, 0.f);
    o[0] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h2));
    o[lo_off] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l2));
  }
};

