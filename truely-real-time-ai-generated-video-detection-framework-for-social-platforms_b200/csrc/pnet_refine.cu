// K3b: exact re-evaluation of the P-Net cells that the single-pass tensor-core screen (pnet2.cu, or pnet.cu with
// TERMS = 1) found at or near the detection threshold.
//
// upstream: models/mtcnn.py PNet.forward + models/utils/detect_face.py generateBoundingBox.
//
// Why: P-Net is 15 % of its FLOPs in fp32-faithful form on the tensor pipe only at three times the work (3-term fp16
// split).  A single fp16 pass is within ~1e-3 of the fp32 maps -- good enough to say "this cell is nowhere near 0.6",
// not good enough to decide the cells that are (the cascade's integer box truncations amplify last-bit differences:
// profiles/PROFILE_NOTES.md r02).  So the screen keeps every cell whose approximate probability is >= thr - margin, and this
// kernel recomputes exactly those cells -- a 12 x 12 x 3 receptive field each, 45 k MACs -- in plain fp32 FMA arithmetic,
// applies the real threshold and emits the candidate with the exact score and regression.  The candidate set and values
// are those of an fp32 P-Net (same bar as the 3-term kernel: maps within 2e-5 of the oracle), the work is proportional to
// the number of near-threshold cells (tens per frame on the bench clips, ~1e3 on textured video) instead of 167 k cells.
//
// One warp per screened cell, persistent grid-stride over the global screen list (count read on the device: no host
// synchronisation).  Weights live in shared memory (fp32, 26 KB per CTA), each warp has 3.4 KB of scratch:
//   input 12x12x3 -> conv1 (lane = pooled pixel, 4 conv positions x 10 channels) + PReLU + ceil-mode 2x2 max-pool
//   -> conv2 3x3x16 (lane = output, 4.5 per lane) + PReLU -> conv3 1x1x32 (lane = channel) + PReLU -> heads by shuffle
//   reduction -> softmax (ATen form), threshold, generateBoundingBox.
#include "common.cuh"
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

namespace pnet_refine {

constexpr int WARPS = 8;
constexpr int NTHREADS = 32 * WARPS;
// fp32 weight image (floats)
constexpr int RW1 = 0;                  // [27 = ci*9 + ky*3 + kx][12] (10 used)
constexpr int RB1 = RW1 + 27 * 12;      // [12] bias
constexpr int RA1 = RB1 + 12;           // [12] slope
constexpr int RW2 = RA1 + 12;           // [90 = (ky*3 + kx)*10 + ci][16]
constexpr int RB2 = RW2 + 90 * 16;
constexpr int RA2 = RB2 + 16;
constexpr int RW3 = RA2 + 16;           // [144 = (ky*3 + kx)*16 + ci][32]
constexpr int RB3 = RW3 + 144 * 32;
constexpr int RA3 = RB3 + 32;
constexpr int RW4 = RA3 + 32;           // [32][8]: conv4_1 (2), conv4_2 (4), 2 pads
constexpr int RB4 = RW4 + 32 * 8;       // [8]
constexpr int RTOTAL = RB4 + 8;
constexpr int SCR_IN = 0;               // per-warp scratch (floats): input [3][12][12]
constexpr int SCR_P1 = SCR_IN + 432;    // pooled conv1 [25][10]
constexpr int SCR_C2 = SCR_P1 + 250;    // conv2 [9][16]
constexpr int SCR = ((SCR_C2 + 144 + 3) / 4) * 4;
constexpr int SMEM_BYTES = (RTOTAL + WARPS * SCR) * 4;

struct Level {
  const void* img;       // FMT 0: fp32 planar [B][3][hs][pitch]; FMT 1: hi image, uint2 (4 halves BGR0) per pixel, rows of 2*pitch pixels
  const void* img_lo;    // FMT 1: lo image, same layout
  int hs, ws, pitch, oh, ow;   // pitch: floats per row (FMT 0) / pixel PAIRS per row (FMT 1)
  float scale;
};
struct Params {
  int n_levels;
  Level lv[TRL_MAX_SCALES];
  const ScreenEntry* screen;
  const int* screen_cnt;
  int screen_cap;
  float thr;
  int cap;               // candidates per (frame, level)
  Cand* cand;            // [B][n_levels][cap]
  int* cnt;              // [B][n_levels]
  CapFlag* capflag;
};

__device__ __forceinline__ float prelu(float v, float a) { return v > 0.f ? v : v * a; }

template <int FMT>
__global__ void __launch_bounds__(NTHREADS) refine_kernel(const float* __restrict__ wimg, const __grid_constant__ Params p) {
  extern __shared__ __align__(16) float sm[];
  float* w = sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* scr = sm + RTOTAL + warp * SCR;
  for (int i = tid; i < RTOTAL; i += NTHREADS) w[i] = __ldg(wimg + i);
  __syncthreads();
  const int total = min(__ldg(p.screen_cnt), p.screen_cap);
  const int nwarps = gridDim.x * WARPS;
  for (int e = blockIdx.x * WARPS + warp; e < total; e += nwarps) {
    const ScreenEntry se = p.screen[e];
    const int b = se.group / p.n_levels, lvl = se.group - b * p.n_levels;
    const Level& L = p.lv[lvl];
    const int oy = se.cell / L.ow, ox = se.cell - oy * L.ow;
    const int hs = L.hs, ws = L.ws;
    // ---- receptive field: input rows 2 oy .. 2 oy + 11, columns 2 ox .. 2 ox + 11 (zero outside the level; such pixels
    // only feed conv positions that ceil-mode pooling excludes)
    for (int i = lane; i < 144; i += 32) {
      const int r = i / 12, cx = i - r * 12;
      const int gy = 2 * oy + r, gx = 2 * ox + cx;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f;
      if (gy < hs && gx < ws) {
        if (FMT == 0) {
          const float* src = reinterpret_cast<const float*>(L.img) + ((size_t)b * 3 * hs + gy) * L.pitch + gx;
          const size_t plane = (size_t)hs * L.pitch;
          v0 = __ldg(src); v1 = __ldg(src + plane); v2 = __ldg(src + 2 * plane);
        } else {
          const size_t idx = ((size_t)b * hs + gy) * (2 * (size_t)L.pitch) + gx;
          const uint2 h = __ldg(reinterpret_cast<const uint2*>(L.img) + idx);
          const uint2 l = __ldg(reinterpret_cast<const uint2*>(L.img_lo) + idx);
          const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
          const float2 h2 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
          const float2 l01 = __half22float2(*reinterpret_cast<const __half2*>(&l.x));
          const float2 l2 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
          v0 = h01.x + l01.x; v1 = h01.y + l01.y; v2 = h2.x + l2.x;
        }
      }
      scr[SCR_IN + i] = v0; scr[SCR_IN + 144 + i] = v1; scr[SCR_IN + 288 + i] = v2;
    }
    __syncwarp();
    // ---- conv1 + PReLU + max-pool: lane < 25 owns pooled pixel (py, px) of the 5 x 5 pooled patch
    if (lane < 25) {
      const int py = lane / 5, px = lane - py * 5;
      float acc[4][10];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int co = 0; co < 10; ++co) acc[q][co] = 0.f;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        float patch[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int cx = 0; cx < 4; ++cx) patch[r][cx] = scr[SCR_IN + ci * 144 + (2 * py + r) * 12 + 2 * px + cx];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float* wr = w + RW1 + ((ci * 3 + ky) * 3 + kx) * 12;
#pragma unroll
            for (int co = 0; co < 10; ++co) {
              const float wv = wr[co];
#pragma unroll
              for (int q = 0; q < 4; ++q) acc[q][co] = fmaf(patch[(q >> 1) + ky][(q & 1) + kx], wv, acc[q][co]);
            }
          }
      }
      const int c1h = hs - 2, c1w = ws - 2;
      const int gy = 2 * (oy + py), gx = 2 * (ox + px);
#pragma unroll
      for (int co = 0; co < 10; ++co) {
        const float bias = w[RB1 + co], al = w[RA1 + co];
        float mm = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool ok = (gy + (q >> 1) < c1h) && (gx + (q & 1) < c1w);
          const float v = prelu(acc[q][co] + bias, al);
          mm = ok ? fmaxf(mm, v) : mm;
        }
        scr[SCR_P1 + lane * 10 + co] = (mm == -INFINITY) ? 0.f : mm;
      }
    }
    __syncwarp();
    // ---- conv2 (10 -> 16, 3x3 on the 5 x 5 pooled patch -> 3 x 3) + PReLU: output o = pos * 16 + co
#pragma unroll 1
    for (int o = lane; o < 144; o += 32) {
      const int pos = o >> 4, co = o & 15;
      const int y = pos / 3, x = pos - y * 3;
      float a = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float* in = scr + SCR_P1 + ((y + ky) * 5 + x + kx) * 10;
          const float* wr = w + RW2 + ((ky * 3 + kx) * 10) * 16 + co;
#pragma unroll
          for (int ci = 0; ci < 10; ++ci) a = fmaf(in[ci], wr[ci * 16], a);
        }
      scr[SCR_C2 + o] = prelu(a + w[RB2 + co], w[RA2 + co]);
    }
    __syncwarp();
    // ---- conv3 (16 -> 32, 3x3 on 3 x 3 -> 1 x 1) + PReLU: lane = output channel
    float a3 = 0.f;
#pragma unroll 4
    for (int k = 0; k < 144; ++k) a3 = fmaf(scr[SCR_C2 + k], w[RW3 + k * 32 + lane], a3);
    const float v3 = prelu(a3 + w[RB3 + lane], w[RA3 + lane]);
    // ---- heads: conv4_1 (2) and conv4_2 (4) over the 32 channels
    float h[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      float t = v3 * w[RW4 + lane * 8 + q];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) t += __shfl_xor_sync(0xffffffffu, t, s);
      h[q] = t + w[RB4 + q];
    }
    __syncwarp();
    if (lane == 0) {
      const float mx = fmaxf(h[0], h[1]);
      const float e0 = expf(h[0] - mx), e1 = expf(h[1] - mx);
      const float prob = __fdiv_rn(e1, e0 + e1);
      if (prob >= p.thr) {
        const int slot = atomicAdd(&p.cnt[se.group], 1);
        if (slot < p.cap) {
          Cand cd;
          cd.x1 = floorf(__fdiv_rn((float)(2 * ox + 1), L.scale));
          cd.y1 = floorf(__fdiv_rn((float)(2 * oy + 1), L.scale));
          cd.x2 = floorf(__fdiv_rn((float)(2 * ox + 12), L.scale));
          cd.y2 = floorf(__fdiv_rn((float)(2 * oy + 12), L.scale));
          cd.score = prob;
          cd.r0 = h[2]; cd.r1 = h[3]; cd.r2 = h[4]; cd.r3 = h[5];
          cd.key = (uint32_t)se.cell;
          p.cand[(size_t)se.group * p.cap + slot] = cd;
        } else if (p.capflag) {
          p.capflag->overflow = 1; p.capflag->stage = 1; p.capflag->frame = b;
          p.capflag->count = slot + 1; p.capflag->capacity = p.cap;
        }
      }
    }
  }
}

}  // namespace pnet_refine

// upstream layouts -> the refine kernel's fp32 weight image
int pnet_refine_pack_weights(trl_ctx* c, const float* h, size_t len) {
  using namespace pnet_refine;
  if (len != 6632) TRL_FAIL(c, TRL_E_INVALID, "pnet blob has %zu floats, expected 6632", len);
  std::vector<float> pk(RTOTAL, 0.f);
  const float* w1 = h;                   // [10][3][3][3]
  const float* b1 = w1 + 270;
  const float* a1 = b1 + 10;
  const float* w2 = a1 + 10;             // [16][10][3][3]
  const float* b2 = w2 + 1440;
  const float* a2 = b2 + 16;
  const float* w3 = a2 + 16;             // [32][16][3][3]
  const float* b3 = w3 + 4608;
  const float* a3 = b3 + 32;
  const float* w41 = a3 + 32;            // [2][32]
  const float* b41 = w41 + 64;
  const float* w42 = b41 + 2;            // [4][32]
  const float* b42 = w42 + 128;
  for (int co = 0; co < 10; ++co) {
    for (int k = 0; k < 27; ++k) pk[RW1 + k * 12 + co] = w1[co * 27 + k];
    pk[RB1 + co] = b1[co]; pk[RA1 + co] = a1[co];
  }
  for (int co = 0; co < 16; ++co) {
    for (int ci = 0; ci < 10; ++ci)
      for (int t = 0; t < 9; ++t) pk[RW2 + (t * 10 + ci) * 16 + co] = w2[(co * 10 + ci) * 9 + t];
    pk[RB2 + co] = b2[co]; pk[RA2 + co] = a2[co];
  }
  for (int co = 0; co < 32; ++co) {
    for (int ci = 0; ci < 16; ++ci)
      for (int t = 0; t < 9; ++t) pk[RW3 + (t * 16 + ci) * 32 + co] = w3[(co * 16 + ci) * 9 + t];
    pk[RB3 + co] = b3[co]; pk[RA3 + co] = a3[co];
    pk[RW4 + co * 8 + 0] = w41[co]; pk[RW4 + co * 8 + 1] = w41[32 + co];
    for (int j = 0; j < 4; ++j) pk[RW4 + co * 8 + 2 + j] = w42[j * 32 + co];
  }
  pk[RB4 + 0] = b41[0]; pk[RB4 + 1] = b41[1];
  for (int j = 0; j < 4; ++j) pk[RB4 + 2 + j] = b42[j];
  TRL_CUDA(c, cudaMalloc(&c->d_pnet_refine, RTOTAL * sizeof(float)));
  TRL_CUDA(c, cudaMemcpy(c->d_pnet_refine, pk.data(), RTOTAL * sizeof(float), cudaMemcpyHostToDevice));
  TRL_CUDA(c, cudaFuncSetAttribute(refine_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  TRL_CUDA(c, cudaFuncSetAttribute(refine_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  return TRL_OK;
}

// fmt 0: d_pyr = the padded fp32 planar pyramid (g.pitch, g.off); fmt 1: d_pyr / d_pyr_lo = the fp16 hi / lo pair images
// (g.pitch2, g.off2).  The screen list and its counter were filled by the screening P-Net launch on the same stream.
int launch_pnet_refine(trl_ctx* c, int fmt, const void* d_pyr, const void* d_pyr_lo, int B, const PyramidGeom& g, float thr,
                       const ScreenEntry* d_screen, const int* d_screen_cnt, int screen_cap, Cand* d_cand, int* d_cnt, int cap,
                       cudaStream_t s) {
  using namespace pnet_refine;
  if (B == 0 || g.n == 0) return TRL_OK;
  Params p{};
  p.n_levels = g.n;
  for (int k = 0; k < g.n; ++k) {
    Level& L = p.lv[k];
    L.hs = g.hs[k]; L.ws = g.ws[k]; L.oh = g.oh[k]; L.ow = g.ow[k]; L.scale = g.scale_f[k];
    if (fmt == 0) {
      L.img = reinterpret_cast<const float*>(d_pyr) + g.off[k] * B; L.img_lo = nullptr; L.pitch = g.pitch[k];
    } else {
      L.img = reinterpret_cast<const uint4*>(d_pyr) + g.off2[k] * B;
      L.img_lo = reinterpret_cast<const uint4*>(d_pyr_lo) + g.off2[k] * B;
      L.pitch = g.pitch2[k];
    }
  }
  p.screen = d_screen; p.screen_cnt = d_screen_cnt; p.screen_cap = screen_cap;
  p.thr = thr; p.cap = cap; p.cand = d_cand; p.cnt = d_cnt; p.capflag = c->d_cap;
  const int grid = 2 * c->num_sms;
  if (fmt == 0) refine_kernel<0><<<grid, NTHREADS, SMEM_BYTES, s>>>(c->d_pnet_refine, p);
  else refine_kernel<1><<<grid, NTHREADS, SMEM_BYTES, s>>>(c->d_pnet_refine, p);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}
