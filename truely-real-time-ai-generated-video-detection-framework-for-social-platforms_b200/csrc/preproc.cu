// K1 pyramid (area resample + normalise), K7 candidate crop-resample, K10 crop-align.
// All three are HBM/L2-bound byte kernels with exact integer window arithmetic.
//
// Parity notes (checked on CPU against torch 2.11 / OpenCV 4.13, see tests/test_host_numerics.py):
//  * F.interpolate(mode="area") == adaptive_avg_pool2d: window [floor(i*H/oh), ceil((i+1)*H/oh)),
//    value = (sum / kh) / kw in fp32 (two divisions, in that order) -- bit exact with ATen.
//  * (x - 127.5) * 0.0078125 : one rounding (the subtract), the multiply is exact.
//  * cv2.resize(INTER_LINEAR, uint8): 11-bit fixed-point coefficients, horizontal pass in int32,
//    vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2; x coefficients are clamped at the
//    borders, y rows are clipped instead (coefficients kept) -- bit exact with OpenCV.
#include <algorithm>

#include "common.cuh"

// ----------------------------------------------------------------------------- shared device helpers

// Sum of the u8 BGR window [y0,y1) x [x0,x1) of one frame, split over `grp` cooperating lanes
// (lane `sub` takes columns x0+sub, x0+sub+grp, ...).  Integer sums are exact, so the split is free.
__device__ __forceinline__ void window_sum(const uint8_t* __restrict__ frame, int W, int y0, int y1, int x0, int x1,
                                           int sub, int grp, int& s0, int& s1, int& s2) {
  s0 = s1 = s2 = 0;
  for (int y = y0; y < y1; ++y) {
    const uint8_t* row = frame + ((size_t)y * W) * 3;
    for (int x = x0 + sub; x < x1; x += grp) {
      const uint8_t* p = row + x * 3;
      s0 += p[0];
      s1 += p[1];
      s2 += p[2];
    }
  }
}

// Same sum over a BGRx (4 bytes / pixel) copy of the frame: one aligned 32-bit load per pixel instead of three byte
// loads (the LSU issue rate, not bandwidth, bounds these kernels), channels accumulated two at a time in 16-bit lanes.
// A lane adds at most ceil(kw/grp) <= 5 pixels per row, far below the 257 that would overflow a 16-bit lane.
__device__ __forceinline__ void window_sum_x(const uint32_t* __restrict__ frame, int W, int y0, int y1, int x0, int x1,
                                             int sub, int grp, int& s0, int& s1, int& s2) {
  s0 = s1 = s2 = 0;
  for (int y = y0; y < y1; ++y) {
    const uint32_t* row = frame + (size_t)y * W;
    uint32_t a = 0, b = 0;
    for (int x = x0 + sub; x < x1; x += grp) {
      const uint32_t w = __ldg(row + x);
      a += w & 0x00FF00FFu;          // B | R<<16
      b += (w >> 8) & 0x00FF00FFu;   // G | x<<16
    }
    s0 += (int)(a & 0xFFFFu);
    s2 += (int)(a >> 16);
    s1 += (int)(b & 0xFFFFu);
  }
}

// flat [n_px] BGR bytes -> BGRx words, 4 pixels (12 B in, 16 B out) per thread
__global__ void __launch_bounds__(256) bgr_to_bgrx_kernel(const uint8_t* __restrict__ in, uint32_t* __restrict__ out, size_t n_px) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t p = i * 4;
  if (p + 3 < n_px) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in + p * 3);
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    uint4 o;
    o.x = w0 & 0x00FFFFFFu;
    o.y = ((w0 >> 24) | (w1 << 8)) & 0x00FFFFFFu;
    o.z = ((w1 >> 16) | (w2 << 16)) & 0x00FFFFFFu;
    o.w = w2 >> 8;
    *reinterpret_cast<uint4*>(out + p) = o;
  } else {
    for (size_t q = p; q < n_px; ++q)
      out[q] = (uint32_t)in[q * 3] | ((uint32_t)in[q * 3 + 1] << 8) | ((uint32_t)in[q * 3 + 2] << 16);
  }
}

__device__ __forceinline__ float area_norm(int s, int kh, int kw) {
  float v = __fdiv_rn(__fdiv_rn((float)s, (float)kh), (float)kw);
  return __fmul_rn(__fsub_rn(v, 127.5f), 0.0078125f);
}

// ----------------------------------------------------------------------------- K1 pyramid

// Window tables (host computed once per frame shape): for level k, tab + tab_off[k] holds
// x0[ws], x1[ws], y0[hs], y1[hs]  -- no integer division on the device.
// One block = (level, tile of PYR_ROWS output rows, tile of 256/grp output columns); grid.y = frame.
// `grp` lanes (1..32, a power of two chosen from the window width) cooperate on one output pixel, so both the
// 2x2 windows of the finest level and the ~100x100 windows of the coarsest one stay coalesced.
constexpr int PYR_ROWS = 8;

__global__ void __launch_bounds__(256) pyramid_kernel(const uint32_t* __restrict__ frames, int H, int W, PyrParams p,
                                                     int lvl_first, const int* __restrict__ tab, float* __restrict__ out) {
  int lvl = lvl_first;
  while (lvl + 1 < p.n && (int)blockIdx.x >= p.blk_start[lvl + 1]) ++lvl;
  const int hs = p.hs[lvl], ws = p.ws[lvl];
  const int gsh = p.grp[lvl];                        // log2(lanes per pixel)
  const int grp = 1 << gsh;
  const int px_per_blk = 256 >> gsh;
  const int col_tiles = (ws + px_per_blk - 1) / px_per_blk;
  const int local = (int)blockIdx.x - p.blk_start[lvl];
  const int rt = local / col_tiles, ct = local - rt * col_tiles;
  const int b = blockIdx.y;
  const int ox = ct * px_per_blk + ((int)threadIdx.x >> gsh);
  const int sub = threadIdx.x & (grp - 1);
  const int* t = tab + p.tab_off[lvl];
  const bool vx = ox < ws;
  int x0 = 0, x1 = 0;
  if (vx) { x0 = __ldg(t + ox); x1 = __ldg(t + ws + ox); }
  const int kw = x1 - x0;
  const uint32_t* frame = frames + (size_t)b * H * W;
  const size_t plane = (size_t)hs * ws;
  float* obase = out + p.off[lvl] + (size_t)b * 3 * plane;
  const int oy_end = min(hs, (rt + 1) * PYR_ROWS);
  for (int oy = rt * PYR_ROWS; oy < oy_end; ++oy) {
    const int y0 = __ldg(t + 2 * ws + oy), y1 = __ldg(t + 2 * ws + hs + oy);
    int s0 = 0, s1 = 0, s2 = 0;
    if (vx) window_sum_x(frame, W, y0, y1, x0, x1, sub, grp, s0, s1, s2);
    for (int o = grp >> 1; o > 0; o >>= 1) {   // grp divides 32: groups never straddle a warp
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (vx && sub == 0) {
      float* o = obase + (size_t)oy * ws + ox;
      const int kh = y1 - y0;
      o[0] = area_norm(s0, kh, kw);
      o[plane] = area_norm(s1, kh, kw);
      o[2 * plane] = area_norm(s2, kh, kw);
    }
  }
}


// Fine levels (window at most WMAX x WMAX): one thread per output pixel, fully unrolled predicated window, 32-bit
// index arithmetic.  The generic lane-cooperative kernel above spends most of its instructions on loop control when
// the windows are 2x2..8x8, and these levels hold > 95 % of the pyramid's pixels.
template <int WMAX>
__global__ void __launch_bounds__(256) pyramid_fine_kernel(const uint32_t* __restrict__ frames, int H, int W, PyrParams p,
                                                          int lvl_first, int blk_first, const int* __restrict__ tab,
                                                          float* __restrict__ out) {
  int lvl = lvl_first;
  const int bx = (int)blockIdx.x + blk_first;
  while (lvl + 1 < p.n && bx >= p.blk_start[lvl + 1]) ++lvl;
  const int hs = p.hs[lvl], ws = p.ws[lvl];
  const int col_tiles = (ws + 255) >> 8;
  const int local = bx - p.blk_start[lvl];
  const int rt = local / col_tiles, ct = local - rt * col_tiles;
  const int ox = ct * 256 + (int)threadIdx.x;
  if (ox >= ws) return;
  const int* t = tab + p.tab_off[lvl];
  const int x0 = __ldg(t + ox), kw = __ldg(t + ws + ox) - x0;
  const uint32_t* frame = frames + (size_t)blockIdx.y * H * W;
  const int plane = hs * ws;
  float* obase = out + p.off[lvl] + (size_t)blockIdx.y * 3 * plane + ox;
  const float fkw = (float)kw;
  const int oy_end = min(hs, (rt + 1) * PYR_ROWS);
  for (int oy = rt * PYR_ROWS; oy < oy_end; ++oy) {
    const int y0 = __ldg(t + 2 * ws + oy), kh = __ldg(t + 2 * ws + hs + oy) - y0;
    const uint32_t* wp = frame + (y0 * W + x0);
    uint32_t a = 0, b = 0;
#pragma unroll
    for (int dy = 0; dy < WMAX; ++dy) {
      if (dy < kh) {
#pragma unroll
        for (int dx = 0; dx < WMAX; ++dx) {
          if (dx < kw) {
            const uint32_t w = __ldg(wp + dy * W + dx);
            a += w & 0x00FF00FFu;
            b += (w >> 8) & 0x00FF00FFu;
          }
        }
      }
    }
    // WMAX*WMAX <= 64 pixels: the 16-bit lanes cannot overflow
    const float fkh = (float)kh;
    float* o = obase + oy * ws;
    o[0] = __fmul_rn(__fsub_rn(__fdiv_rn(__fdiv_rn((float)(a & 0xFFFFu), fkh), fkw), 127.5f), 0.0078125f);
    o[plane] = __fmul_rn(__fsub_rn(__fdiv_rn(__fdiv_rn((float)(b & 0xFFFFu), fkh), fkw), 127.5f), 0.0078125f);
    o[2 * plane] = __fmul_rn(__fsub_rn(__fdiv_rn(__fdiv_rn((float)(a >> 16), fkh), fkw), 127.5f), 0.0078125f);
  }
}

static int pick_group_log2(int in_extent, int out_extent) {
  int kw = (in_extent + out_extent - 1) / out_extent + 1;   // upper bound of the window width
  int g = 0;
  while (g < 5 && (2 << g) <= kw) ++g;
  return g;
}

// adaptive_avg_pool2d windows: [floor(i*In/Out), ceil((i+1)*In/Out))
static void window_table(int In, int Out, std::vector<int>& v) {
  const size_t base = v.size();
  v.resize(base + 2 * (size_t)Out);
  for (int i = 0; i < Out; ++i) {
    v[base + i] = (int)(((long long)i * In) / Out);
    v[base + Out + i] = (int)(((long long)(i + 1) * In + Out - 1) / Out);
  }
}

int launch_pyramid(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const PyramidGeom& g, float* d_out,
                   cudaStream_t s) {
  if (g.n == 0 || B == 0) return TRL_OK;
  // tables are cached per frame shape
  if (c->pyr_tab_H != H || c->pyr_tab_W != W || c->d_pyr_tab == nullptr) {
    std::vector<int> tab;
    for (int k = 0; k < g.n; ++k) {
      c->pyr_tab_off[k] = (int)tab.size();
      window_table(W, g.ws[k], tab);
      window_table(H, g.hs[k], tab);
    }
    if (c->d_pyr_tab) { TRL_CUDA(c, cudaStreamSynchronize(s)); TRL_CUDA(c, cudaFree(c->d_pyr_tab)); c->d_pyr_tab = nullptr; }
    TRL_CUDA(c, cudaMalloc(&c->d_pyr_tab, tab.size() * sizeof(int)));
    TRL_CUDA(c, cudaMemcpy(c->d_pyr_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
    c->pyr_tab_H = H;
    c->pyr_tab_W = W;
  }
  // Levels are launched by class: fine levels (max window <= 3 / 4 / 6 / 8) with the unrolled one-thread-per-pixel
  // kernel, the coarse rest with the lane-cooperative kernel.  blk_start is a prefix over the levels of one class.
  PyrParams p{};
  p.n = g.n;
  int cls[TRL_MAX_SCALES];
  for (int k = 0; k < g.n; ++k) {
    p.hs[k] = g.hs[k];
    p.ws[k] = g.ws[k];
    p.off[k] = g.off[k] * B;
    p.grp[k] = pick_group_log2(W, g.ws[k]);
    p.tab_off[k] = c->pyr_tab_off[k];
    const int wmax = std::max((W + g.ws[k] - 1) / g.ws[k], (H + g.hs[k] - 1) / g.hs[k]) + 1;
    cls[k] = wmax <= 3 ? 3 : wmax <= 4 ? 4 : wmax <= 6 ? 6 : wmax <= 8 ? 8 : 0;
  }
  // BGR -> BGRx staging copy (frames are addressed flat, so any H*W*3 works as long as the batch base pointer is
  // 4-byte aligned; cudaMalloc / torch allocations are 256-byte aligned)
  const size_t n_px = (size_t)B * H * W;
  if ((reinterpret_cast<uintptr_t>(d_frames) & 3) != 0) TRL_FAIL(c, TRL_E_INVALID, "frames pointer must be 4-byte aligned");
  if (c->bgrx_cap < n_px) {
    if (c->d_bgrx) { TRL_CUDA(c, cudaFree(c->d_bgrx)); c->d_bgrx = nullptr; }
    TRL_CUDA(c, cudaMalloc(&c->d_bgrx, n_px * sizeof(uint32_t)));
    c->bgrx_cap = n_px;
  }
  bgr_to_bgrx_kernel<<<(unsigned)((n_px / 4 + 1 + 255) / 256), 256, 0, s>>>(d_frames, c->d_bgrx, n_px);
  TRL_LAUNCH_CHECK(c);
  // levels are ordered fine -> coarse, so each class is a contiguous run of levels
  int k = 0;
  while (k < g.n) {
    const int cl = cls[k];
    int k1 = k;
    int blocks = 0;
    while (k1 < g.n && cls[k1] == cl) {
      p.blk_start[k1] = blocks;
      blocks += cl ? ceil_div(g.hs[k1], PYR_ROWS) * ceil_div(g.ws[k1], 256)
                   : ceil_div(g.hs[k1], PYR_ROWS) * ceil_div(g.ws[k1], 256 >> p.grp[k1]);
      ++k1;
    }
    for (int q = k1; q <= g.n; ++q) p.blk_start[q] = blocks;      // sentinel for the level search
    PyrParams pc = p;
    // the level search in the kernels walks from the first level of the class: hide the finer ones
    for (int q = 0; q < k; ++q) pc.blk_start[q] = -1;
    dim3 grid(blocks, B);
    switch (cl) {
      case 3: pyramid_fine_kernel<3><<<grid, 256, 0, s>>>(c->d_bgrx, H, W, pc, k, 0, c->d_pyr_tab, d_out); break;
      case 4: pyramid_fine_kernel<4><<<grid, 256, 0, s>>>(c->d_bgrx, H, W, pc, k, 0, c->d_pyr_tab, d_out); break;
      case 6: pyramid_fine_kernel<6><<<grid, 256, 0, s>>>(c->d_bgrx, H, W, pc, k, 0, c->d_pyr_tab, d_out); break;
      case 8: pyramid_fine_kernel<8><<<grid, 256, 0, s>>>(c->d_bgrx, H, W, pc, k, 0, c->d_pyr_tab, d_out); break;
      default: pyramid_kernel<<<grid, 256, 0, s>>>(c->d_bgrx, H, W, pc, k, c->d_pyr_tab, d_out); break;
    }
    TRL_LAUNCH_CHECK(c);
    k = k1;
  }
  return TRL_OK;
}

// ----------------------------------------------------------------------------- K7 crop-resample

// One CTA per candidate slot.  grid = (n_max slots).  The crop is rows [y-1, ey) x cols [x-1, ex) (0-based),
// clipped by pad() upstream, never zero padded.  Output float32 [slot][3][S][S].
// d_count (optional) holds the live number of slots; slots >= *d_count exit.
__global__ void __launch_bounds__(256) crop_resample_kernel(const uint8_t* __restrict__ frames, int H, int W,
                                                           const int* __restrict__ pad4, const int* __restrict__ img,
                                                           const int* __restrict__ d_count, int per_frame_cap,
                                                           int S, float* __restrict__ out) {
  const int slot = blockIdx.x;
  int b, n_live;
  if (per_frame_cap > 0) {          // slot = frame * cap + i, counts per frame
    b = slot / per_frame_cap;
    n_live = d_count[b];
    if (slot - b * per_frame_cap >= n_live) return;
  } else {
    if (d_count && slot >= *d_count) return;
    b = img[slot];
  }
  const int y = pad4[slot * 4 + 0], ey = pad4[slot * 4 + 1], x = pad4[slot * 4 + 2], ex = pad4[slot * 4 + 3];
  float* o = out + (size_t)slot * 3 * S * S;
  const int ch = ey - (y - 1), cw = ex - (x - 1);
  if (ch <= 0 || cw <= 0) {   // degenerate: upstream skips such crops (detect_face.py stage 2/3 loop guard)
    for (int i = threadIdx.x; i < 3 * S * S; i += blockDim.x) o[i] = 0.f;
    return;
  }
  const uint8_t* frame = frames + (size_t)b * H * W * 3;
  int grp = 1;
  {
    int kw = (cw + S - 1) / S + 1;
    while (grp < 32 && grp * 2 <= kw) grp <<= 1;
  }
  const int sub = threadIdx.x % grp;
  const int per_iter = blockDim.x / grp;
  const int total = S * S;
  const int iters = (total + per_iter - 1) / per_iter;
  for (int it = 0; it < iters; ++it) {
    const int pix = it * per_iter + (int)threadIdx.x / grp;
    const bool valid = pix < total;
    int s0 = 0, s1 = 0, s2 = 0, kh = 1, kw = 1;
    if (valid) {
      const int oy = pix / S, ox = pix - oy * S;
      const int y0 = (oy * ch) / S, y1 = ((oy + 1) * ch + S - 1) / S;
      const int x0 = (ox * cw) / S, x1 = ((ox + 1) * cw + S - 1) / S;
      kh = y1 - y0;
      kw = x1 - x0;
      window_sum(frame, W, (y - 1) + y0, (y - 1) + y1, (x - 1) + x0, (x - 1) + x1, sub, grp, s0, s1, s2);
    }
    for (int d = grp >> 1; d > 0; d >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, d);
      s1 += __shfl_xor_sync(0xffffffffu, s1, d);
      s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    if (valid && sub == 0) {
      o[pix] = area_norm(s0, kh, kw);
      o[total + pix] = area_norm(s1, kh, kw);
      o[2 * total + pix] = area_norm(s2, kh, kw);
    }
  }
}

// d_count semantics: if per_frame_cap > 0, d_count is int[B] and slots are frame*cap+i; else a single int (or null).
int launch_crop_resample_ex(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const int* d_pad, const int* d_img,
                            const int* d_count, int per_frame_cap, int n_slots, int size, float* d_out, cudaStream_t s) {
  if (n_slots <= 0) return TRL_OK;
  crop_resample_kernel<<<n_slots, 256, 0, s>>>(d_frames, H, W, d_pad, d_img, d_count, per_frame_cap, size, d_out);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

int launch_crop_resample(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const int* d_pad, const int* d_img,
                         const int* d_count, int n_max, int size, float* d_out, cudaStream_t s) {
  return launch_crop_resample_ex(c, d_frames, B, H, W, d_pad, d_img, d_count, 0, n_max, size, d_out, s);
}

// ----------------------------------------------------------------------------- K10 crop-align

struct ResizeCoef {
  int idx;     // source index (already clamped for x; raw for y)
  int c0, c1;  // 11-bit fixed point coefficients
};

// OpenCV resize.cpp: fx = (float)((d + 0.5) * scale - 0.5); s = floor(fx); fx -= s; coefficients
// saturate_cast<short>(v * 2048) (round half to even).  `clamp_coef`: the x direction resets fx at the borders.
__device__ __forceinline__ ResizeCoef resize_coef(int d, int ssize, int dsize, bool clamp_coef) {
  const double scale = 1.0 / ((double)dsize / (double)ssize);
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int sidx = (int)floorf(f);
  f = __fsub_rn(f, (float)sidx);
  if (clamp_coef) {
    if (sidx < 0) { f = 0.f; sidx = 0; }
    if (sidx >= ssize - 1) { f = 0.f; sidx = ssize - 1; }
  }
  ResizeCoef r;
  r.idx = sidx;
  r.c1 = __float2int_rn(__fmul_rn(f, 2048.f));
  r.c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  return r;
}

// One CTA per frame.  Selects boxes[0] (already largest-first), truncates toward zero, clamps
// (server/model.py:49-53), and resamples the crop to SxS exactly like cv2.resize(INTER_LINEAR) on uint8.
__global__ void __launch_bounds__(256) crop_align_kernel(const uint8_t* __restrict__ frames, int H, int W,
                                                        const float* __restrict__ boxes, int box_stride,
                                                        const int* __restrict__ nfaces, int S,
                                                        int* __restrict__ box_int, uint8_t* __restrict__ valid,
                                                        uint8_t* __restrict__ crops) {
  const int b = blockIdx.x;
  __shared__ int sb[4];
  __shared__ int sok;
  if (threadIdx.x == 0) {
    int ok = 0;
    int x1 = 0, y1 = 0, x2 = 0, y2 = 0;
    if (nfaces[b] > 0) {
      const float* bx = boxes + (size_t)b * box_stride;
      x1 = (int)bx[0]; y1 = (int)bx[1]; x2 = (int)bx[2]; y2 = (int)bx[3];   // astype(int): toward zero
      x1 = max(0, x1); y1 = max(0, y1); x2 = min(W, x2); y2 = min(H, y2);
      ok = (x2 > x1 && y2 > y1) ? 1 : 0;
    }
    sb[0] = x1; sb[1] = y1; sb[2] = x2; sb[3] = y2;
    sok = ok;
    box_int[b * 4 + 0] = x1; box_int[b * 4 + 1] = y1; box_int[b * 4 + 2] = x2; box_int[b * 4 + 3] = y2;
    valid[b] = (uint8_t)ok;
  }
  __syncthreads();
  uint8_t* o = crops + (size_t)b * S * S * 3;
  if (!sok) {
    for (int i = threadIdx.x; i < S * S * 3; i += blockDim.x) o[i] = 0;
    return;
  }
  const int x1 = sb[0], y1 = sb[1], sw = sb[2] - sb[0], sh = sb[3] - sb[1];
  const uint8_t* src = frames + ((size_t)b * H * W + (size_t)y1 * W + x1) * 3;
  for (int pix = threadIdx.x; pix < S * S; pix += blockDim.x) {
    const int dy = pix / S, dx = pix - dy * S;
    const ResizeCoef cx = resize_coef(dx, sw, S, true);
    const ResizeCoef cy = resize_coef(dy, sh, S, false);
    const int xa = cx.idx, xb = min(cx.idx + 1, sw - 1);
    const int ya = min(max(cy.idx, 0), sh - 1), yb = min(max(cy.idx + 1, 0), sh - 1);
    const uint8_t* ra = src + (size_t)ya * W * 3;
    const uint8_t* rb = src + (size_t)yb * W * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int h0 = ra[xa * 3 + ch] * cx.c0 + ra[xb * 3 + ch] * cx.c1;
      const int h1 = rb[xa * 3 + ch] * cx.c0 + rb[xb * 3 + ch] * cx.c1;
      int v = (((cy.c0 * (h0 >> 4)) >> 16) + ((cy.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
      v = min(max(v, 0), 255);
      o[pix * 3 + ch] = (uint8_t)v;
    }
  }
}

int launch_crop_align(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const float* d_boxes, int box_stride,
                      const int* d_nfaces, int S, int* d_box_int, uint8_t* d_valid, uint8_t* d_crops, cudaStream_t s) {
  if (B <= 0) return TRL_OK;
  crop_align_kernel<<<B, 256, 0, s>>>(d_frames, H, W, d_boxes, box_stride, d_nfaces, S, d_box_int, d_valid, d_crops);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

// ----------------------------------------------------------------------------- pyramid geometry (host)

// upstream detect_face: m = 12/minsize; scales m*factor^k while min(h,w)*m*factor^k >= 12, all in doubles;
// level size (int(h*scale+1), int(w*scale+1)); P-Net map = ceil((hs-2)/2) - 4.
int compute_geometry(const trl_config_t& cfg, int H, int W, PyramidGeom* g) {
  g->n = 0;
  g->px_total = 0;
  if (H <= 0 || W <= 0 || cfg.min_face_size <= 0) return TRL_E_INVALID;
  const double m = 12.0 / (double)cfg.min_face_size;
  double minl = (double)(H < W ? H : W) * m;
  double scale_i = m;
  long long off = 0;
  while (minl >= 12.0) {
    if (g->n >= TRL_MAX_SCALES) return TRL_E_INVALID;
    const int k = g->n++;
    g->scale[k] = scale_i;
    g->scale_f[k] = (float)scale_i;
    g->hs[k] = (int)((double)H * scale_i + 1.0);
    g->ws[k] = (int)((double)W * scale_i + 1.0);
    g->oh[k] = (g->hs[k] - 2 + 1) / 2 - 4;
    g->ow[k] = (g->ws[k] - 2 + 1) / 2 - 4;
    g->off[k] = off;
    off += 3LL * g->hs[k] * g->ws[k];
    g->px_total += (long long)g->hs[k] * g->ws[k];
    scale_i = scale_i * cfg.factor;
    minl = minl * cfg.factor;
  }
  return TRL_OK;
}
