// FaceNet kernels that do not run on the tensor pipe:
//   * stem conv2d_1a (3->32, 3x3 s2, K = 27) reading the uint8 crop directly (F.to_tensor's /255 is folded into
//     the weights on the host, so the input is exact and the crop never exists in float)
//   * MaxPool2d(3, 2) on NHWC bf16, writing into a channel slice (concat-free Mixed_6a / Mixed_7a)
//   * head: AdaptiveAvgPool2d(1) + last_linear + last_bn (folded) + F.normalize, fp32
//   * a direct-convolution implementation of ConvOp (fp32 accumulate, same epilogue as the tcgen05 kernel): the
//     on-device reference the tcgen05 path is validated against (cfg.facenet_impl = 1)
#include "facenet.cuh"

// ----------------------------------------------------------------------------- stem
// thread = one output pixel x 32 channels; weights [27][32] fp32 in shared memory
// `d_n` (optional, every kernel of the forward pass): the live batch size on the device -- the number of face-bearing
// crops after compaction (facenet_forward_valid), known only there; the launch is sized for the host-side bound `n`.
__global__ void __launch_bounds__(128) stem_conv_kernel(const uint8_t* __restrict__ crops, int n, int S, int Ho,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       bf16* __restrict__ out, const int* __restrict__ d_n) {
  if (d_n) n = min(n, *d_n);
  __shared__ __align__(16) float w_s[27 * 32];
  __shared__ float b_s[32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) w_s[i] = w[i];
  if (threadIdx.x < 32) b_s[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * Ho * Ho;
  if (pix >= total) return;
  const int img = (int)(pix / (Ho * Ho));
  const int rem = (int)(pix - (long long)img * Ho * Ho);
  const int oy = rem / Ho, ox = rem - oy * Ho;
  const uint8_t* src = crops + ((size_t)img * S * S + (size_t)(2 * oy) * S + 2 * ox) * 3;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = b_s[c];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float v = (float)src[(ky * S + kx) * 3 + ci];
        const float4* wr = reinterpret_cast<const float4*>(&w_s[((ky * 3 + kx) * 3 + ci) * 32]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 wv = wr[q];
          acc[q * 4 + 0] = fmaf(v, wv.x, acc[q * 4 + 0]);
          acc[q * 4 + 1] = fmaf(v, wv.y, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(v, wv.z, acc[q * 4 + 2]);
          acc[q * 4 + 3] = fmaf(v, wv.w, acc[q * 4 + 3]);
        }
      }
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)pix * 32);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaxf(acc[q * 8 + 0], 0.f), fmaxf(acc[q * 8 + 1], 0.f));
    __nv_bfloat162 p1 = __floats2bfloat162_rn(fmaxf(acc[q * 8 + 2], 0.f), fmaxf(acc[q * 8 + 3], 0.f));
    __nv_bfloat162 p2 = __floats2bfloat162_rn(fmaxf(acc[q * 8 + 4], 0.f), fmaxf(acc[q * 8 + 5], 0.f));
    __nv_bfloat162 p3 = __floats2bfloat162_rn(fmaxf(acc[q * 8 + 6], 0.f), fmaxf(acc[q * 8 + 7], 0.f));
    uint4 v;
    v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
    v.z = *reinterpret_cast<uint32_t*>(&p2); v.w = *reinterpret_cast<uint32_t*>(&p3);
    dst[q] = v;
  }
}

int launch_stem_conv(trl_ctx* c, const uint8_t* d_crops, int n, int S, const float* w, const float* bias, bf16* out,
                     int Ho, cudaStream_t s, const int* d_n) {
  const long long total = (long long)n * Ho * Ho;
  if (total == 0) return TRL_OK;
  stem_conv_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(d_crops, n, S, Ho, w, bias, out, d_n);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

// ----------------------------------------------------------------------------- maxpool 3x3 s2 (NHWC, 8 channels / thread)
__global__ void __launch_bounds__(256) maxpool_kernel(PoolOp op, int n, const int* __restrict__ d_n) {
  if (d_n) n = min(n, *d_n);
  const int cg = op.C / 8;
  const long long total = (long long)n * op.Hout * op.Wout * cg;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int g = (int)(i % cg);
  long long pix = i / cg;
  const int ox = (int)(pix % op.Wout); pix /= op.Wout;
  const int oy = (int)(pix % op.Hout);
  const int img = (int)(pix / op.Hout);
  float m[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) m[q] = -INFINITY;
  for (int dy = 0; dy < 3; ++dy)
    for (int dx = 0; dx < 3; ++dx) {
      const int y = 2 * oy + dy, x = 2 * ox + dx;   // no padding, floor mode: always in bounds
      const uint4 v = *reinterpret_cast<const uint4*>(op.in + (((size_t)img * op.Hin + y) * op.Win + x) * op.in_ctot +
                                                      op.in_coff + g * 8);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(h[q]);
        m[2 * q] = fmaxf(m[2 * q], f.x);
        m[2 * q + 1] = fmaxf(m[2 * q + 1], f.y);
      }
    }
  uint4 o;
  __nv_bfloat162 p0 = __floats2bfloat162_rn(m[0], m[1]), p1 = __floats2bfloat162_rn(m[2], m[3]);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(m[4], m[5]), p3 = __floats2bfloat162_rn(m[6], m[7]);
  o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
  o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
  *reinterpret_cast<uint4*>(op.out + (((size_t)img * op.Hout + oy) * op.Wout + ox) * op.out_ctot + op.out_coff + g * 8) = o;
}

int launch_maxpool(trl_ctx* c, const PoolOp& op, int n, cudaStream_t s, const int* d_n) {
  const long long total = (long long)n * op.Hout * op.Wout * (op.C / 8);
  if (total == 0) return TRL_OK;
  maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(op, n, d_n);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

// ----------------------------------------------------------------------------- head
// HEAD_G crops per CTA, 512 threads = 512 embedding dims.  w_t: [1792][512] fp32 (last_linear * last_bn scale, transposed).
constexpr int HEAD_G = 2;
__global__ void __launch_bounds__(512) head_kernel(const bf16* __restrict__ feat, int n, int hw, const float* __restrict__ w_t,
                                                  const float* __restrict__ bias, float* __restrict__ emb,
                                                  const int* __restrict__ d_n) {
  if (d_n) n = min(n, *d_n);
  if ((int)blockIdx.x * HEAD_G >= n) return;     // whole CTA beyond the live batch
  extern __shared__ __align__(16) float x_s[];   // [HEAD_G][1792]
  __shared__ float red[HEAD_G][16];
  const int n0 = blockIdx.x * HEAD_G;
  const float inv = 1.f / (float)hw;
  for (int i = threadIdx.x; i < HEAD_G * 1792; i += blockDim.x) {
    const int g = i / 1792, ch = i - g * 1792;
    float sacc = 0.f;
    if (n0 + g < n)
      for (int p = 0; p < hw; ++p) sacc += __bfloat162float(feat[((size_t)(n0 + g) * hw + p) * 1792 + ch]);
    x_s[i] = sacc * inv;
  }
  __syncthreads();
  const int o = threadIdx.x;
  float acc[HEAD_G];
#pragma unroll
  for (int g = 0; g < HEAD_G; ++g) acc[g] = 0.f;
  for (int k = 0; k < 1792; ++k) {
    const float wv = __ldg(w_t + (size_t)k * 512 + o);
#pragma unroll
    for (int g = 0; g < HEAD_G; ++g) acc[g] = fmaf(x_s[g * 1792 + k], wv, acc[g]);
  }
  const float bv = bias[o];
  float sq[HEAD_G];
#pragma unroll
  for (int g = 0; g < HEAD_G; ++g) { acc[g] += bv; sq[g] = acc[g] * acc[g]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int g = 0; g < HEAD_G; ++g) {
    float v = sq[g];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) red[g][warp] = v;
  }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < HEAD_G; ++g) {
    float tot = 0.f;
#pragma unroll
    for (int wq = 0; wq < 16; ++wq) tot += red[g][wq];
    const float nrm = fmaxf(sqrtf(tot), 1e-12f);     // F.normalize(p=2, eps=1e-12)
    if (n0 + g < n) emb[(size_t)(n0 + g) * 512 + o] = acc[g] / nrm;
  }
}

int launch_head(trl_ctx* c, const bf16* feat, int n, int hw, const float* w_t, const float* bias, float* emb, cudaStream_t s,
                const int* d_n) {
  if (n <= 0) return TRL_OK;
  static bool attr = false;
  if (!attr) {
    TRL_CUDA(c, cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HEAD_G * 1792 * 4));
    attr = true;
  }
  head_kernel<<<ceil_div(n, HEAD_G), 512, HEAD_G * 1792 * 4, s>>>(feat, n, hw, w_t, bias, emb, d_n);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

// ----------------------------------------------------------------------------- direct conv (validation path)
// thread = one output pixel x one output channel, 8-wide bf16 vector loads along Cin.
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvOp op, int n, const int* __restrict__ d_n) {
  if (d_n) n = min(n, *d_n);
  const long long total = (long long)n * op.Hout * op.Wout * op.Cout;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = (int)(i % op.Cout);
  long long pix = i / op.Cout;
  const long long m = pix;
  const int ox = (int)(pix % op.Wout); pix /= op.Wout;
  const int oy = (int)(pix % op.Hout);
  const int img = (int)(pix / op.Hout);
  float acc = 0.f;
  const bf16* wrow = op.w + (size_t)co * op.kh * op.kw * op.Cin;
  for (int ky = 0; ky < op.kh; ++ky) {
    const int iy = oy * op.stride + ky - op.pad_h;
    if (iy < 0 || iy >= op.Hin) continue;
    for (int kx = 0; kx < op.kw; ++kx) {
      const int ix = ox * op.stride + kx - op.pad_w;
      if (ix < 0 || ix >= op.Win) continue;
      const bf16* ip = op.in + (((size_t)img * op.Hin + iy) * op.Win + ix) * op.in_ctot + op.in_coff;
      const bf16* wp = wrow + (size_t)(ky * op.kw + kx) * op.Cin;
      for (int ci = 0; ci < op.Cin; ci += 8) {
        const uint4 a = *reinterpret_cast<const uint4*>(ip + ci);
        const uint4 b = *reinterpret_cast<const uint4*>(wp + ci);
        const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* bh = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 fa = __bfloat1622float2(ah[q]), fb = __bfloat1622float2(bh[q]);
          acc = fmaf(fa.x, fb.x, acc);
          acc = fmaf(fa.y, fb.y, acc);
        }
      }
    }
  }
  float v = acc + op.bias[co];
  if (op.epi != EPI_RELU) {
    const float x = __bfloat162float(op.resid[(size_t)m * op.Cout + co]);
    v = fmaf(op.scale, v, x);
  }
  if (op.epi != EPI_RESID) v = fmaxf(v, 0.f);
  for (int sidx = 0; sidx < op.nseg; ++sidx) {
    const OutSeg& sg = op.seg[sidx];
    if (co >= sg.n_begin && co < sg.n_end)
      sg.dst[(size_t)m * sg.dst_ctot + sg.dst_coff + (co - sg.n_begin)] = __float2bfloat16_rn(v);
  }
}

int launch_conv_simt(trl_ctx* c, const ConvOp& op, int n, cudaStream_t s, const int* d_n) {
  const long long total = (long long)n * op.Hout * op.Wout * op.Cout;
  if (total == 0) return TRL_OK;
  conv_simt_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(op, n, d_n);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}
