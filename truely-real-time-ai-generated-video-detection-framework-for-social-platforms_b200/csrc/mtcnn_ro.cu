// K8 / K9: R-Net (24x24) and O-Net (48x48) of the MTCNN cascade, fp32 on the FMA pipe.
// One CTA per candidate; all activations stay in shared memory, weights stream from L2 (they are shared by
// every CTA in flight), the first conv+pool layers are evaluated in row bands to bound shared memory.
//
// upstream: models/mtcnn.py RNet.forward / ONet.forward.  ceil_mode pooling clips the
// window at the border; the flatten before the first dense layer is (W, H, C) ordered upstream -- the host
// permutes the dense weight rows once so the kernel reads its [C][H][W] activations directly.
#include <algorithm>

#include "common.cuh"

// RO_FFMA2 = 1: packed fma.rn.f32x2 in the conv rows; 0: scalar FFMA.  Same IEEE results either way.  Measured on B200
// (experiments/fma_rate_probe.cu): FFMA issues one warp instruction per cycle and sub-partition, FFMA2 one per three
// cycles -- two FMAs in 3 cycles with two issue slots left for other pipes, against two FMAs in 2 cycles with none.
#ifndef RO_FFMA2
#define RO_FFMA2 1
#endif
// RO_CGP2: O-Net layers that compute two output-channel groups per pixel quad (conv_rows CGP = 2; same accumulation order,
// bit identical): 0 none, 1 conv1, 2 conv1 + conv2.  Measured on B200 (experiments/variants/ab_lib.sh), O-Net stage per step:
// 0.826 / 0.795 / 0.759 ms (720p bench clip), 18.13 / 17.43 / 16.63 ms (1080p clip batch).  Neutral for R-Net's conv2 (one item
// per thread leaves half of its 256 threads idle) and slower with four-row bands for O-Net's conv2 (0.781 ms): not used there.
#ifndef RO_CGP2
#define RO_CGP2 2
#endif

namespace ro {

__device__ __forceinline__ float prelu(float v, float a) { return v > 0.f ? v : v * a; }

// conv (valid, stride 1) + bias + PReLU for conv output rows [r0, r1): in [CIN][HIN][WIN] (smem) ->
// out [COUT][r1-r0][WOUT] (smem).  w: global [CIN*K*K][COUT].  item = 4 channels x 4 pixels of one row; a thread owns
// up to MAXI items.  The weights are staged through shared memory in slabs of CC input channels (cp.async, double
// buffered in wbuf): one coalesced L2 read per slab for the whole CTA instead of a dependent global load in front of
// every 16 FMAs of every thread (that chain, not the FMA pipe, set the single-candidate latency).
// The accumulation order over (ci, ky, kx) is unchanged.
// CGP: groups of 4 output channels per item (the same 4 pixels): the activation loads are shared by CGP x 4 channels, so the
// shared-memory loads per FMA drop by a third at CGP = 2 (the conv rows are bound by that traffic, profiles/r02d_rnet_full.md).
template <int CIN, int COUT, int K, int HIN, int WIN, int CC, int MAXI, int CGP = 1>
__device__ __forceinline__ void conv_rows(const float* __restrict__ in, float* __restrict__ out, int r0, int r1,
                                          const float* __restrict__ w, const float* __restrict__ bias,
                                          const float* __restrict__ alpha, float* __restrict__ wbuf) {
  constexpr int WOUT = WIN - K + 1;
  constexpr int PXG = (WOUT + 3) / 4;
  constexpr int CG = COUT / (4 * CGP);              // items along the channel axis
  static_assert(COUT % (4 * CGP) == 0, "conv_rows channel grouping");
  constexpr int SLAB = CC * K * K * COUT;          // floats per slab
  constexpr int NCH = CIN / CC;
  static_assert(COUT % 4 == 0 && CIN % CC == 0 && SLAB % 4 == 0, "conv_rows tiling");
  const int rows = r1 - r0;
  const int items = CG * rows * PXG;
  const uint32_t wb_s = (uint32_t)__cvta_generic_to_shared(wbuf);
  auto stage = [&](int ch) {
    const float* src = w + (size_t)ch * SLAB;
    const uint32_t dst = wb_s + (uint32_t)(ch & 1) * SLAB * 4u;
    for (int i = threadIdx.x; i < SLAB / 4; i += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + 4 * i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int base = 0; base < items; base += MAXI * (int)blockDim.x) {
    int cg[MAXI], row[MAXI], x0[MAXI];
    bool live[MAXI];
#if RO_FFMA2
    unsigned long long acc2[MAXI][CGP][4][2];
#else
    float accs[MAXI][CGP][4][4];
#endif
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int item = base + (int)threadIdx.x + i * (int)blockDim.x;
      live[i] = item < items;
      const int it = live[i] ? item : 0;
      cg[i] = it / (rows * PXG);
      const int rem = it - cg[i] * (rows * PXG);
      row[i] = rem / PXG;
      x0[i] = (rem - row[i] * PXG) * 4;
#pragma unroll
      for (int g = 0; g < CGP; ++g)
#pragma unroll
#if RO_FFMA2
        for (int px = 0; px < 4; ++px) acc2[i][g][px][0] = acc2[i][g][px][1] = 0ull;
#else
        for (int px = 0; px < 4; ++px)
#pragma unroll
          for (int j = 0; j < 4; ++j) accs[i][g][px][j] = 0.f;
#endif
    }
    stage(0);
#pragma unroll 1
    for (int ch = 0; ch < NCH; ++ch) {
      if (ch + 1 < NCH) {
        stage(ch + 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();                                   // slab ch is visible to every thread
      const float* ws = wbuf + (ch & 1) * SLAB;
#pragma unroll
      for (int i = 0; i < MAXI; ++i) {
        if (!live[i]) continue;
#pragma unroll 2
        for (int cl = 0; cl < CC; ++cl) {
          const int ci = ch * CC + cl;
#pragma unroll
          for (int ky = 0; ky < K; ++ky) {
            const float* ir = in + (ci * HIN + r0 + row[i] + ky) * WIN + x0[i];
            float v[4 + K - 1];
#pragma unroll
            for (int t = 0; t < 4 + K - 1; ++t) v[t] = (x0[i] + t < WIN) ? ir[t] : 0.f;
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
#pragma unroll
              for (int g = 0; g < CGP; ++g) {
                const float4 wv = *reinterpret_cast<const float4*>(ws + ((cl * K + ky) * K + kx) * COUT + (cg[i] * CGP + g) * 4);
#if RO_FFMA2
                const unsigned long long w01 = pack_f32x2(wv.x, wv.y), w23 = pack_f32x2(wv.z, wv.w);
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                  const unsigned long long vv = pack_f32x2(v[px + kx], v[px + kx]);
                  ffma2(acc2[i][g][px][0], vv, w01);
                  ffma2(acc2[i][g][px][1], vv, w23);
                }
#else
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                  const float vv = v[px + kx];
                  accs[i][g][px][0] = fmaf(vv, wv.x, accs[i][g][px][0]);
                  accs[i][g][px][1] = fmaf(vv, wv.y, accs[i][g][px][1]);
                  accs[i][g][px][2] = fmaf(vv, wv.z, accs[i][g][px][2]);
                  accs[i][g][px][3] = fmaf(vv, wv.w, accs[i][g][px][3]);
                }
#endif
              }
            }
          }
        }
      }
      __syncthreads();                                   // slab buffer ch & 1 may be overwritten by stage(ch + 2)
    }
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      if (!live[i]) continue;
#pragma unroll
      for (int g = 0; g < CGP; ++g) {
        float acc[4][4];
#pragma unroll
        for (int px = 0; px < 4; ++px) {
#if RO_FFMA2
          unpack_f32x2(acc2[i][g][px][0], acc[px][0], acc[px][1]);
          unpack_f32x2(acc2[i][g][px][1], acc[px][2], acc[px][3]);
#else
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[px][j] = accs[i][g][px][j];
#endif
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int co = (cg[i] * CGP + g) * 4 + j;
          const float b = __ldg(bias + co), a = __ldg(alpha + co);
#pragma unroll
          for (int px = 0; px < 4; ++px)
            if (x0[i] + px < WOUT) out[(co * rows + row[i]) * WOUT + x0[i] + px] = prelu(acc[px][j] + b, a);
        }
      }
    }
  }
}

// MaxPool2d(PK, PS, ceil_mode=True) for pooled rows [p0, p1) from a conv band that starts at conv row band_r0.
template <int C, int HC, int WC, int PK, int PS, int HP, int WP>
__device__ __forceinline__ void pool_rows(const float* __restrict__ band, int band_r0, int band_rows,
                                          float* __restrict__ out, int p0, int p1) {
  const int n = C * (p1 - p0) * WP;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i / ((p1 - p0) * WP);
    const int rem = i - c * ((p1 - p0) * WP);
    const int py = p0 + rem / WP, px = rem % WP;
    float m = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < PK; ++dy) {
      const int y = py * PS + dy;
      if (y >= HC) continue;
#pragma unroll
      for (int dx = 0; dx < PK; ++dx) {
        const int x = px * PS + dx;
        if (x >= WC) continue;
        m = fmaxf(m, band[(c * band_rows + (y - band_r0)) * WC + x]);
      }
    }
    out[(c * HP + py) * WP + px] = m;
  }
}

// conv + PReLU + ceil-mode maxpool, banded over PR pooled rows at a time
template <int CIN, int COUT, int K, int HIN, int WIN, int PK, int PS, int PR, int CC, int MAXI, int CGP = 1>
__device__ __forceinline__ void conv_pool(const float* in, float* band, float* out, const float* w, const float* b,
                                          const float* a, float* wbuf) {
  constexpr int HC = HIN - K + 1, WC = WIN - K + 1;
  constexpr int HP = (HC - PK + PS - 1) / PS + 1, WP = (WC - PK + PS - 1) / PS + 1;
  for (int p0 = 0; p0 < HP; p0 += PR) {
    const int p1 = min(p0 + PR, HP);
    const int r0 = p0 * PS, r1 = min((p1 - 1) * PS + PK, HC);
    conv_rows<CIN, COUT, K, HIN, WIN, CC, MAXI, CGP>(in, band, r0, r1, w, b, a, wbuf);
    __syncthreads();
    pool_rows<COUT, HC, WC, PK, PS, HP, WP>(band, r0, r1 - r0, out, p0, p1);
    __syncthreads();
  }
}

// dense + bias (+ PReLU): in_s[IN] -> out_s[OUT]; w global [IN][OUT]; blockDim.x must be a multiple of OUT
template <int IN, int OUT, bool PRELU>
__device__ __forceinline__ void dense(const float* __restrict__ in_s, float* __restrict__ out_s, float* __restrict__ part_s,
                                      const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ a) {
  const int ks_n = blockDim.x / OUT;
  const int o = threadIdx.x % OUT, ks = threadIdx.x / OUT;
  if (ks < ks_n) {
    const int k0 = (IN * ks) / ks_n, k1 = (IN * (ks + 1)) / ks_n;
    // 8 independent loads in flight per thread (the weight stream from L2 is the cost of this layer); the partial sums
    // are combined in a fixed order
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int k = k0;
    for (; k + 7 < k1; k += 8) {
      float wv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = __ldg(w + (size_t)(k + j) * OUT + o);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(in_s[k + j], wv[j], acc[j]);
    }
    for (; k < k1; ++k) acc[0] = fmaf(in_s[k], __ldg(w + (size_t)k * OUT + o), acc[0]);
    const float acc0 = (acc[0] + acc[1]) + (acc[2] + acc[3]), acc1 = (acc[4] + acc[5]) + (acc[6] + acc[7]);
    part_s[ks * OUT + o] = acc0 + acc1;
  }
  __syncthreads();
  if (threadIdx.x < OUT) {
    float v = __ldg(b + o);
    for (int q = 0; q < ks_n; ++q) v += part_s[q * OUT + o];
    out_s[o] = PRELU ? prelu(v, __ldg(a + o)) : v;
  }
  __syncthreads();
}

// heads: 6 outputs (2 class logits, 4 box offsets) = one warp each; w global [6][IN]
template <int IN>
__device__ __forceinline__ void heads(const float* __restrict__ in_s, const float* __restrict__ w, const float* __restrict__ b,
                                      float* __restrict__ h_s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 6) {
    float acc = 0.f;
    for (int k = lane; k < IN; k += 32) acc = fmaf(in_s[k], __ldg(w + warp * IN + k), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) h_s[warp] = acc + __ldg(b + warp);
  }
  __syncthreads();
}

__device__ __forceinline__ void write_outputs(const float* h_s, int slot, float* prob, float* reg) {
  if (threadIdx.x == 0) {
    const float mx = fmaxf(h_s[0], h_s[1]);
    const float e0 = expf(h_s[0] - mx), e1 = expf(h_s[1] - mx);
    prob[slot] = __fdiv_rn(e1, e0 + e1);
    reg[slot * 4 + 0] = h_s[2]; reg[slot * 4 + 1] = h_s[3]; reg[slot * 4 + 2] = h_s[4]; reg[slot * 4 + 3] = h_s[5];
  }
}

__device__ __forceinline__ bool slot_live(int slot, const int* d_count, int per_frame_cap) {
  if (per_frame_cap > 0) {
    const int b = slot / per_frame_cap;
    return slot - b * per_frame_cap < d_count[b];
  }
  return d_count == nullptr || slot < *d_count;
}

// ---------------------------------------------------------------- R-Net
// packed weights: w1[27][28] b1 a1 | w2[252][48] b2 a2 | w3[192][64] b3 a3 | w4[576][128] b4 a4 | wh[6][128] bh[6]
namespace r {
constexpr int W1 = 0, B1 = W1 + 27 * 28, A1 = B1 + 28;
constexpr int W2 = A1 + 28, B2 = W2 + 252 * 48, A2 = B2 + 48;
constexpr int W3 = A2 + 48, B3 = W3 + 192 * 64, A3 = B3 + 64;
constexpr int W4 = A3 + 64, B4 = W4 + 576 * 128, A4 = B4 + 128;
constexpr int WH = A4 + 128, BH = WH + 6 * 128;
constexpr int TOTAL = ((BH + 6 + 3) / 4) * 4;
// smem (floats)
constexpr int S_IN = 0;                    // 3*24*24 = 1728
constexpr int S_BAND = S_IN + 1728;        // max(28*9*22 = 5544, 48*9*9 = 3888)
constexpr int S_P1 = S_BAND + 5544;        // 28*11*11 = 3388
constexpr int S_P2 = S_P1 + 3388;          // 48*4*4 = 768
constexpr int S_C3 = S_P2 + 768;           // 64*3*3 = 576
constexpr int S_D4 = S_C3 + 576;           // 128
constexpr int S_PART = S_D4 + 128;         // 256
constexpr int S_H = S_PART + 256;          // 8
constexpr int S_WB = S_H + 8;              // weight slabs: 2 x max(27*28, 7*9*48, 8*4*64) = 2 x 3024
constexpr int S_TOTAL = S_WB + 2 * 3024;
}  // namespace r

// work distribution shared by both kernels: compact (grid-stride over live slots) when the per-frame counts fit the
// slot map, else one CTA per slot with an early exit
struct WorkList {
  bool compact;
  int total;
};
__device__ __forceinline__ WorkList worklist_init(const int* d_count, int per_frame_cap, int n_frames, int n_slots, int* s_pref) {
  WorkList wl;
  wl.compact = per_frame_cap > 0 && n_frames > 0 && n_frames <= SLOTMAP_MAX_FRAMES;
  wl.total = wl.compact ? slotmap_init(d_count, n_frames, per_frame_cap, s_pref) : n_slots;
  return wl;
}

__global__ void __launch_bounds__(256) rnet_kernel(const float* __restrict__ in, const float* __restrict__ wp,
                                                  const int* __restrict__ d_count, int per_frame_cap, int n_frames,
                                                  int n_slots, float* __restrict__ prob, float* __restrict__ reg) {
  using namespace r;
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_pref[SLOTMAP_MAX_FRAMES + 1];
  const WorkList wl = worklist_init(d_count, per_frame_cap, n_frames, n_slots, s_pref);
  for (int work = blockIdx.x; work < wl.total; work += gridDim.x) {
  const int slot = wl.compact ? slotmap_slot(s_pref, n_frames, per_frame_cap, work) : work;
  if (!wl.compact && !slot_live(slot, d_count, per_frame_cap)) continue;      // CTA uniform
  const float* src = in + (size_t)slot * 1728;
  for (int i = threadIdx.x; i < 1728 / 4; i += blockDim.x)
    reinterpret_cast<float4*>(sm + S_IN)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
  __syncthreads();
  conv_pool<3, 28, 3, 24, 24, 3, 2, 4, 3, 2>(sm + S_IN, sm + S_BAND, sm + S_P1, wp + W1, wp + B1, wp + A1, sm + S_WB);    // -> 28x11x11
  conv_pool<28, 48, 3, 11, 11, 3, 2, 4, 7, 2>(sm + S_P1, sm + S_BAND, sm + S_P2, wp + W2, wp + B2, wp + A2, sm + S_WB);   // -> 48x4x4
  conv_rows<48, 64, 2, 4, 4, 8, 1>(sm + S_P2, sm + S_C3, 0, 3, wp + W3, wp + B3, wp + A3, sm + S_WB);                    // -> 64x3x3
  __syncthreads();
  dense<576, 128, true>(sm + S_C3, sm + S_D4, sm + S_PART, wp + W4, wp + B4, wp + A4);
  heads<128>(sm + S_D4, wp + WH, wp + BH, sm + S_H);
  write_outputs(sm + S_H, slot, prob, reg);
  __syncthreads();
  }
}

#ifdef PNET_TIMING
__device__ unsigned long long g_onet_phase[8];
#define OT_MARK(i) do { __syncthreads(); if (threadIdx.x == 0) { const long long _t = clock64(); atomicAdd(&g_onet_phase[i], (unsigned long long)(_t - t_prev)); t_prev = _t; } } while (0)
#else
#define OT_MARK(i) do { } while (0)
#endif

// ---------------------------------------------------------------- O-Net
// packed: w1[27][32] | w2[288][64] | w3[576][64] | w4[256][128] | w5[1152][256] | wh[6][256] bh[6]  (+ b/a each)
namespace o {
constexpr int W1 = 0, B1 = W1 + 27 * 32, A1 = B1 + 32;
constexpr int W2 = A1 + 32, B2 = W2 + 288 * 64, A2 = B2 + 64;
constexpr int W3 = A2 + 64, B3 = W3 + 576 * 64, A3 = B3 + 64;
constexpr int W4 = A3 + 64, B4 = W4 + 256 * 128, A4 = B4 + 128;
constexpr int W5 = A4 + 128, B5 = W5 + 1152 * 256, A5 = B5 + 256;
constexpr int WH = A5 + 256, BH = WH + 6 * 256;
constexpr int TOTAL = ((BH + 6 + 3) / 4) * 4;
// smem (floats)
constexpr int S_IN = 0;                     // 3*48*48 = 6912; later p2 64*10*10 = 6400
constexpr int S_P1 = S_IN + 6912;           // 32*23*23 = 16928
constexpr int S_BAND = S_P1 + 16928;        // max(32*9*46 = 13248, 64*11*21 = 14784, 64*8*8 = 4096)
constexpr int S_P3 = S_BAND + 14784;        // 64*4*4 = 1024
constexpr int S_C4 = S_P3 + 1024;           // 128*3*3 = 1152
constexpr int S_D5 = S_C4 + 1152;           // 256
constexpr int S_PART = S_D5 + 256;          // 512
constexpr int S_H = S_PART + 512;
constexpr int S_WB = S_H + 8;               // weight slabs: 2 x max(27*32, 8*9*64, 8*4*128) = 2 x 4608
constexpr int S_TOTAL = S_WB + 2 * 4608;
}  // namespace o

__global__ void __launch_bounds__(512) onet_kernel(const float* __restrict__ in, const float* __restrict__ wp,
                                                  const int* __restrict__ d_count, int per_frame_cap, int n_frames,
                                                  int n_slots, float* __restrict__ prob, float* __restrict__ reg) {
  using namespace o;
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_pref[SLOTMAP_MAX_FRAMES + 1];
  const WorkList wl = worklist_init(d_count, per_frame_cap, n_frames, n_slots, s_pref);
  for (int work = blockIdx.x; work < wl.total; work += gridDim.x) {
  const int slot = wl.compact ? slotmap_slot(s_pref, n_frames, per_frame_cap, work) : work;
  if (!wl.compact && !slot_live(slot, d_count, per_frame_cap)) continue;      // CTA uniform
#ifdef PNET_TIMING
  long long t_prev = clock64();
#endif
  const float* src = in + (size_t)slot * 6912;
  for (int i = threadIdx.x; i < 6912 / 4; i += blockDim.x)
    reinterpret_cast<float4*>(sm + S_IN)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
  __syncthreads();
  OT_MARK(0);
#if RO_CGP2
  conv_pool<3, 32, 3, 48, 48, 3, 2, 4, 3, 1, 2>(sm + S_IN, sm + S_BAND, sm + S_P1, wp + W1, wp + B1, wp + A1, sm + S_WB);     // -> 32x23x23
#else
  conv_pool<3, 32, 3, 48, 48, 3, 2, 4, 3, 2>(sm + S_IN, sm + S_BAND, sm + S_P1, wp + W1, wp + B1, wp + A1, sm + S_WB);     // -> 32x23x23
#endif
  OT_MARK(1);
  float* p2 = sm + S_IN;                                                                                  // input is dead
#if RO_CGP2 >= 2
  conv_pool<32, 64, 3, 23, 23, 3, 2, 5, 8, 2, 2>(sm + S_P1, sm + S_BAND, p2, wp + W2, wp + B2, wp + A2, sm + S_WB);        // -> 64x10x10
#else
  conv_pool<32, 64, 3, 23, 23, 3, 2, 5, 8, 3>(sm + S_P1, sm + S_BAND, p2, wp + W2, wp + B2, wp + A2, sm + S_WB);           // -> 64x10x10
#endif
  OT_MARK(2);
  conv_pool<64, 64, 3, 10, 10, 2, 2, 4, 8, 1>(p2, sm + S_BAND, sm + S_P3, wp + W3, wp + B3, wp + A3, sm + S_WB);           // -> 64x4x4
  OT_MARK(3);
  conv_rows<64, 128, 2, 4, 4, 8, 1>(sm + S_P3, sm + S_C4, 0, 3, wp + W4, wp + B4, wp + A4, sm + S_WB);                     // -> 128x3x3
  __syncthreads();
  OT_MARK(4);
  dense<1152, 256, true>(sm + S_C4, sm + S_D5, sm + S_PART, wp + W5, wp + B5, wp + A5);
  OT_MARK(5);
  heads<256>(sm + S_D5, wp + WH, wp + BH, sm + S_H);
  write_outputs(sm + S_H, slot, prob, reg);
  OT_MARK(6);
#ifdef PNET_TIMING
  if (threadIdx.x == 0) atomicAdd(&g_onet_phase[7], 1ull);
#endif
  __syncthreads();
  }
}

}  // namespace ro
#ifdef PNET_TIMING
extern "C" void trl_debug_onet_timing(unsigned long long* out) {
  cudaMemcpyFromSymbol(out, ro::g_onet_phase, sizeof(unsigned long long) * 8);
  unsigned long long z[8] = {0};
  cudaMemcpyToSymbol(ro::g_onet_phase, z, sizeof(z));
}
#endif

// ---------------------------------------------------------------- host: weight packing + launch

static void pack_conv(std::vector<float>& pk, int off, const float* w, int cout, int cin, int k) {
  // upstream [cout][cin][k][k] -> [(ci,ky,kx)][cout]
  for (int co = 0; co < cout; ++co)
    for (int q = 0; q < cin * k * k; ++q) pk[off + q * cout + co] = w[co * cin * k * k + q];
}

// upstream dense weight [out][in] with in = (x*H + y)*C + c  ->  [ (c*H + y)*W + x ][out]
static void pack_dense_whc(std::vector<float>& pk, int off, const float* w, int out, int C, int H, int W) {
  const int in = C * H * W;
  for (int o = 0; o < out; ++o)
    for (int c = 0; c < C; ++c)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
          pk[off + ((c * H + y) * W + x) * out + o] = w[(size_t)o * in + (x * H + y) * C + c];
}

int ro_pack_weights(trl_ctx* c, const float* h_r, size_t rlen, const float* h_o, size_t olen) {
  using namespace ro;
  if (rlen != 100178) TRL_FAIL(c, TRL_E_INVALID, "rnet blob has %zu floats, expected 100178", rlen);
  if (olen != 389040) TRL_FAIL(c, TRL_E_INVALID, "onet blob has %zu floats, expected 389040", olen);
  {
    std::vector<float> pk(r::TOTAL, 0.f);
    const float* p = h_r;
    pack_conv(pk, r::W1, p, 28, 3, 3); p += 28 * 27;
    memcpy(&pk[r::B1], p, 28 * 4); p += 28; memcpy(&pk[r::A1], p, 28 * 4); p += 28;
    pack_conv(pk, r::W2, p, 48, 28, 3); p += 48 * 252;
    memcpy(&pk[r::B2], p, 48 * 4); p += 48; memcpy(&pk[r::A2], p, 48 * 4); p += 48;
    pack_conv(pk, r::W3, p, 64, 48, 2); p += 64 * 192;
    memcpy(&pk[r::B3], p, 64 * 4); p += 64; memcpy(&pk[r::A3], p, 64 * 4); p += 64;
    pack_dense_whc(pk, r::W4, p, 128, 64, 3, 3); p += 128 * 576;
    memcpy(&pk[r::B4], p, 128 * 4); p += 128; memcpy(&pk[r::A4], p, 128 * 4); p += 128;
    // dense5_1 [2][128], bias[2], dense5_2 [4][128], bias[4]
    memcpy(&pk[r::WH], p, 2 * 128 * 4); p += 256;
    memcpy(&pk[r::BH], p, 2 * 4); p += 2;
    memcpy(&pk[r::WH + 256], p, 4 * 128 * 4); p += 512;
    memcpy(&pk[r::BH + 2], p, 4 * 4); p += 4;
    TRL_CUDA(c, cudaMalloc(&c->d_rnet, pk.size() * 4));
    TRL_CUDA(c, cudaMemcpy(c->d_rnet, pk.data(), pk.size() * 4, cudaMemcpyHostToDevice));
  }
  {
    std::vector<float> pk(o::TOTAL, 0.f);
    const float* p = h_o;
    pack_conv(pk, o::W1, p, 32, 3, 3); p += 32 * 27;
    memcpy(&pk[o::B1], p, 32 * 4); p += 32; memcpy(&pk[o::A1], p, 32 * 4); p += 32;
    pack_conv(pk, o::W2, p, 64, 32, 3); p += 64 * 288;
    memcpy(&pk[o::B2], p, 64 * 4); p += 64; memcpy(&pk[o::A2], p, 64 * 4); p += 64;
    pack_conv(pk, o::W3, p, 64, 64, 3); p += 64 * 576;
    memcpy(&pk[o::B3], p, 64 * 4); p += 64; memcpy(&pk[o::A3], p, 64 * 4); p += 64;
    pack_conv(pk, o::W4, p, 128, 64, 2); p += 128 * 256;
    memcpy(&pk[o::B4], p, 128 * 4); p += 128; memcpy(&pk[o::A4], p, 128 * 4); p += 128;
    pack_dense_whc(pk, o::W5, p, 256, 128, 3, 3); p += 256 * 1152;
    memcpy(&pk[o::B5], p, 256 * 4); p += 256; memcpy(&pk[o::A5], p, 256 * 4); p += 256;
    // dense6_1 [2][256] + bias, dense6_2 [4][256] + bias, dense6_3 (landmarks) ignored
    memcpy(&pk[o::WH], p, 2 * 256 * 4); p += 512;
    memcpy(&pk[o::BH], p, 2 * 4); p += 2;
    memcpy(&pk[o::WH + 512], p, 4 * 256 * 4); p += 1024;
    memcpy(&pk[o::BH + 2], p, 4 * 4); p += 4;
    TRL_CUDA(c, cudaMalloc(&c->d_onet, pk.size() * 4));
    TRL_CUDA(c, cudaMemcpy(c->d_onet, pk.data(), pk.size() * 4, cudaMemcpyHostToDevice));
  }
  TRL_CUDA(c, cudaFuncSetAttribute(rnet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, r::S_TOTAL * 4));
  TRL_CUDA(c, cudaFuncSetAttribute(onet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, o::S_TOTAL * 4));
  return TRL_OK;
}

int launch_rnet_ex(trl_ctx* c, const float* d_in, int n_slots, const int* d_count, int per_frame_cap, float* d_prob,
                   float* d_reg, cudaStream_t s) {
  if (n_slots <= 0) return TRL_OK;
  const int n_frames = per_frame_cap > 0 ? n_slots / per_frame_cap : 0;
  const bool compact = per_frame_cap > 0 && n_frames <= SLOTMAP_MAX_FRAMES;
  const int grid = compact ? std::min(n_slots, c->num_sms * 6) : n_slots;       // ~37 KB of shared memory per CTA: 6 CTAs / SM
  ro::rnet_kernel<<<grid, 256, ro::r::S_TOTAL * 4, s>>>(d_in, c->d_rnet, d_count, per_frame_cap, n_frames, n_slots, d_prob, d_reg);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

int launch_onet_ex(trl_ctx* c, const float* d_in, int n_slots, const int* d_count, int per_frame_cap, float* d_prob,
                   float* d_reg, cudaStream_t s) {
  if (n_slots <= 0) return TRL_OK;
  const int n_frames = per_frame_cap > 0 ? n_slots / per_frame_cap : 0;
  const bool compact = per_frame_cap > 0 && n_frames <= SLOTMAP_MAX_FRAMES;
  const int grid = compact ? std::min(n_slots, c->num_sms) : n_slots;           // ~168 KB of shared memory per CTA: 1 CTA / SM
  ro::onet_kernel<<<grid, 512, ro::o::S_TOTAL * 4, s>>>(d_in, c->d_onet, d_count, per_frame_cap, n_frames, n_slots, d_prob, d_reg);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

int launch_rnet(trl_ctx* c, const float* d_in, int n_max, const int* d_count, float* d_prob, float* d_reg, cudaStream_t s) {
  return launch_rnet_ex(c, d_in, n_max, d_count, 0, d_prob, d_reg, s);
}
int launch_onet(trl_ctx* c, const float* d_in, int n_max, const int* d_count, float* d_prob, float* d_reg, cudaStream_t s) {
  return launch_onet_ex(c, d_in, n_max, d_count, 0, d_prob, d_reg, s);
}
