// K1 pyramid (area resample + normalise), K7 candidate crop-resample, K10 crop-align.
// All three are HBM/L2-bound byte kernels with exact integer window arithmetic.
//
// Parity notes (checked on CPU against torch 2.11 / OpenCV 4.13, see tests/test_oracle.py: the numpy restatements of
// interpolate(area) and cv2.resize are proven bit exact there, tests/test_gpu_stages.py then holds the kernels to them):
//  * F.interpolate(mode="area") == adaptive_avg_pool2d: window [floor(i*H/oh), ceil((i+1)*H/oh)),
//    value = (sum / kh) / kw in fp32 (two divisions, in that order) -- bit exact with ATen.
//  * (x - 127.5) * 0.0078125 : one rounding (the subtract), the multiply is exact.
//  * cv2.resize(INTER_LINEAR, uint8): 11-bit fixed-point coefficients, horizontal pass in int32,
//    vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2; x coefficients are clamped at the
//    borders, y rows are clipped instead (coefficients kept) -- bit exact with OpenCV.
#include <algorithm>
#include <string.h>

#include "common.cuh"
#include "pyramid.cuh"

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

__device__ __forceinline__ float area_norm(int s, int kh, int kw) {
  float v = __fdiv_rn(__fdiv_rn((float)s, (float)kh), (float)kw);
  return __fmul_rn(__fsub_rn(v, 127.5f), 0.0078125f);
}

// ----------------------------------------------------------------------------- K1 pyramid

// Separable exact area resample.  Window tables (host computed once per frame shape): for level k,
// tab + tab_off[k] holds ws pairs {x0 | kw << 16, bits of RN(1 / kw)}, then y0[hs], y1[hs] -- no integer
// division on the device.
//
// One CTA = (frame, level, R consecutive output rows); (level, first row) come from a per-CTA table.
//  pass 1 (vertical): the source rows are read as a flat array of 32-bit words straight from the packed BGR bytes
//    (no BGRx staging copy).  With 16-byte rows (every production shape) a thread owns four consecutive words: one
//    LDG.128 per source row; per word the raw word goes into one register and its odd bytes (PRMT) into a second one, the
//    even-byte lanes follow once per output row from  sum(w) - (odd << 8)  (mod 2^32: exact, both 16-bit lanes < 2^16).
//    Other shapes: word columns tid, tid+256, ... with even / odd byte lanes.  Integer sums are exact in any order.  The
//    column sums go to shared memory as a u16 array indexed by byte column.
//  pass 2 (horizontal): the CTA's rows are ONE flattened [rows][ws] index space, one thread per output pixel: kw column
//    sums per channel, (s / kh) / kw exactly like ATen's adaptive_avg_pool2d, normalise (one FMA), store planar fp32 or the
//    fp16 hi / lo pair pixel (coalesced along x).  Bound by the shared-memory wavefronts of its 2-byte loads and by integer
//    issue, not by HBM (profiles/r02d_pyr_full.md).
// A 16-bit lane holds at most 255 * kh, so kh <= 257 (frames up to ~4000 px on the short side with minsize 20).
//
// Division: a / b with b a small integer, r = RN(1/b): q0 = RN(a r), e = a - q0 b (exact, FMA), q = RN(q0 + e r)
// (Markstein's correction step, 3 instructions).  It equals the IEEE quotient for every operand this kernel can see:
// div_verify_kernel checks all s in [0, 255 kh kw] for every (kh, kw) pair of a level against __fdiv_rn when the
// tables are built, and a level that failed would use __fdiv_rn instead (p.fastdiv).
// pairs: int2 (kh, kw); bad[pair] set when the fast division differs from IEEE for some window sum
__global__ void __launch_bounds__(256) div_verify_kernel(const int2* __restrict__ pairs, int* __restrict__ bad) {
  const int2 k = pairs[blockIdx.y];
  const float fkh = (float)k.x, fkw = (float)k.y;
  const float rkh = __frcp_rn(fkh), rkw = __frcp_rn(fkw);
  const int n = 255 * k.x * k.y;
  bool ok = true;
  for (int s = blockIdx.x * 256 + threadIdx.x; s <= n; s += gridDim.x * 256) {
    const float ref = __fdiv_rn(__fdiv_rn((float)s, fkh), fkw);
    const float got = div_small(div_small((float)s, fkh, rkh), fkw, rkw);
    ok = ok && (__float_as_uint(ref) == __float_as_uint(got));
  }
  if (!ok) bad[blockIdx.y] = 1;
}

template <bool ALIGNED>
__device__ __forceinline__ uint32_t load_word(const uint8_t* __restrict__ frames, size_t byte_off, size_t total_bytes) {
  if (ALIGNED) return __ldg(reinterpret_cast<const uint32_t*>(frames + byte_off));
  // frames is 4-byte aligned; the word may straddle two aligned words and the batch may end inside it
  const size_t a = byte_off & ~(size_t)3;
  const uint32_t sh = (uint32_t)(byte_off & 3) * 8u;
  if (a + 8 <= total_bytes) {
    const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t*>(frames + a));
    const uint32_t w1 = __ldg(reinterpret_cast<const uint32_t*>(frames + a + 4));
    return __funnelshift_r(w0, w1, sh);
  }
  uint32_t v = 0;
  for (int i = 0; i < 4; ++i)
    if (byte_off + i < total_bytes) v |= (uint32_t)__ldg(frames + byte_off + i) << (8 * i);
  return v;
}

// Pass 2 of one output row when every window of the level is KMIN or KMIN + 1 columns wide: the KMIN certain columns are
// summed unconditionally, the possible extra one under a predicate -- no per-pixel loop, no divergence between neighbouring
// pixels whose windows differ by one column (the general loop spent ~60 of its ~140 warp instructions per 32 pixels on the
// branch ladder of its unrolled variable-trip loop).  Same integer sums, same two divisions: bit identical.
// Output formats of the pyramid kernel.  OUT = 0: planar fp32 [3][hs][pitch] (trl_pyramid, the 3-term P-Net).  OUT = 1: the
// all-tensor-pipe P-Net's input (pnet2.cu): per pixel 4 halves (B, G, R, 0) = 8 bytes in a "hi" image (fp16(v): the screen's
// operand) and a "lo" image (fp16(v - hi): with hi it restores v to 2^-23 relative for the exact re-evaluation,
// pnet_refine.cu); one 8-byte store each instead of three 4-byte planar stores.
// The CTA's rows are one flattened [rows][ws] index space (idx -> row by a multiply-high with floor(2^32 / ws) + 1): on a
// 769-pixel row a quarter of the lanes of a row-by-row loop idle (769 = 3 x 256 + 1), and the per-row constants {kh, RN(1 / kh)}
// come from shared memory instead of being rebuilt per thread and row.  (a - 127.5) * 2^-7 is one FMA: scaling by a power of
// two commutes with the rounding of the subtraction.
template <int KMIN, int OUT>
__device__ __forceinline__ void hpass_fixed(const uint16_t* __restrict__ v16, int vpitch, const float2* __restrict__ rowc,
                                            const int2* __restrict__ tw, int ws, int nrows, uint32_t magic, int tid, bool fast,
                                            typename PyrOut<OUT>::T* __restrict__ obase, int pitch, long long plane) {
  const uint32_t total = (uint32_t)(nrows * ws);
#pragma unroll 2
  for (uint32_t idx = tid; idx < total; idx += 256) {
    const uint32_t row = __umulhi(idx, magic);
    const uint32_t col = idx - row * (uint32_t)ws;
    const float2 rc = rowc[row];
    const float fkh = rc.x, rkh = rc.y;
    const int2 e = __ldg(tw + col);
    const int xw = e.x;
    const float rkw = __int_as_float(e.y);
    const int kw = xw >> 16;
    const uint16_t* vp = v16 + row * vpitch + 3 * (xw & 0xFFFF);
    int s0 = vp[0], s1 = vp[1], s2 = vp[2];
#pragma unroll
    for (int x = 1; x < KMIN; ++x) { s0 += vp[3 * x]; s1 += vp[3 * x + 1]; s2 += vp[3 * x + 2]; }
    const bool wide = kw > KMIN;
    if (wide) { s0 += vp[3 * KMIN]; s1 += vp[3 * KMIN + 1]; s2 += vp[3 * KMIN + 2]; }
    const float fkw = wide ? (float)(KMIN + 1) : (float)KMIN;
    float a0, a1, a2;
    if (fast) {
      a0 = div_small(div_small((float)s0, fkh, rkh), fkw, rkw);
      a1 = div_small(div_small((float)s1, fkh, rkh), fkw, rkw);
      a2 = div_small(div_small((float)s2, fkh, rkh), fkw, rkw);
    } else {
      a0 = __fdiv_rn(__fdiv_rn((float)s0, fkh), fkw);
      a1 = __fdiv_rn(__fdiv_rn((float)s1, fkh), fkw);
      a2 = __fdiv_rn(__fdiv_rn((float)s2, fkh), fkw);
    }
    PyrOut<OUT>::store(obase + row * pitch + col, plane, __fmaf_rn(a0, 0.0078125f, -0.99609375f),
                       __fmaf_rn(a1, 0.0078125f, -0.99609375f), __fmaf_rn(a2, 0.0078125f, -0.99609375f));
  }
}

template <int MODE, int OUT>     // MODE 0: byte-aligned rows, 1: 4-byte aligned rows, 2: 16-byte aligned rows (one 128-bit load per thread and row)
__global__ void __launch_bounds__(256) pyramid_sep_kernel(const uint8_t* __restrict__ frames, int H, int W, size_t total_bytes,
                                                         const __grid_constant__ PyrParams p, const int* __restrict__ tab,
                                                         const int2* __restrict__ blk_tab,
                                                         typename PyrOut<OUT>::T* __restrict__ out, long long lo_off) {
  extern __shared__ __align__(16) uint32_t vs[];      // [R][nw] x 2 words = u16 column sums [R][4*nw]
  const int2 bi = __ldg(blk_tab + blockIdx.x);        // (level, first output row) of this CTA: one load instead of a search
  const int lvl = bi.x, j0 = bi.y;
  const int hs = p.hs[lvl], ws = p.ws[lvl], R = p.rows[lvl];
  const int nrows = min(R, hs - j0);
  const int b = blockIdx.y;
  const int rowbytes = 3 * W;
  const int nw = (rowbytes + 3) >> 2;
  const int* t = tab + p.tab_off[lvl];
  const int* ty0 = t + 2 * ws + j0;
  const int* ty1 = ty0 + hs;
  const size_t fbase = (size_t)b * H * rowbytes;
  const int tid = threadIdx.x;
  __shared__ float2 rowc[8];                          // {kh, RN(1 / kh)} of the CTA's rows (R <= 4), read after the barrier below
  if (tid < nrows) {
    const float fkh = (float)(__ldg(ty1 + tid) - __ldg(ty0 + tid));
    rowc[tid] = make_float2(fkh, __frcp_rn(fkh));
  }

  constexpr bool ALIGNED = MODE >= 1;
  if (MODE == 2) {
    // 16-byte rows: thread owns 4 consecutive words of a row -> one LDG.128 per source row, two STS.128 per output row
    const int nq = nw >> 2;                                  // uint4 per row
    for (int c0 = 0; c0 < nq; c0 += 256) {
      const int q = c0 + tid;
      if (q >= nq) continue;
      for (int jj = 0; jj < nrows; ++jj) {
        const int y0 = __ldg(ty0 + jj), y1 = __ldg(ty1 + jj);
        // per source row and word: the raw word into one register, its odd bytes (PRMT) into a second one; the even-byte
        // lanes follow once per output row from  sum(w) - (odd << 8)  (mod 2^32; exact, both 16-bit lanes of the result are
        // < 2^16).  Three instructions per word and row instead of four, and two rows fold into one three-input add.
        uint32_t sw0 = 0, ao0 = 0, sw1 = 0, ao1 = 0, sw2 = 0, ao2 = 0, sw3 = 0, ao3 = 0;
        const uint4* row = reinterpret_cast<const uint4*>(frames + fbase) + (size_t)y0 * nq + q;
        for (int y = y0; y < y1; ++y, row += nq) {
          const uint4 w = __ldg(row);
          sw0 += w.x; ao0 += __byte_perm(w.x, 0, 0x4341);
          sw1 += w.y; ao1 += __byte_perm(w.y, 0, 0x4341);
          sw2 += w.z; ao2 += __byte_perm(w.z, 0, 0x4341);
          sw3 += w.w; ao3 += __byte_perm(w.w, 0, 0x4341);
        }
        const uint32_t ae0 = sw0 - (ao0 << 8), ae1 = sw1 - (ao1 << 8), ae2 = sw2 - (ao2 << 8), ae3 = sw3 - (ao3 << 8);
        uint4* vrow = reinterpret_cast<uint4*>(vs) + jj * (nw >> 1) + 2 * q;      // 2 words of sums per source word
        vrow[0] = make_uint4(__byte_perm(ae0, ao0, 0x5410), __byte_perm(ae0, ao0, 0x7632),
                             __byte_perm(ae1, ao1, 0x5410), __byte_perm(ae1, ao1, 0x7632));
        vrow[1] = make_uint4(__byte_perm(ae2, ao2, 0x5410), __byte_perm(ae2, ao2, 0x7632),
                             __byte_perm(ae3, ao3, 0x5410), __byte_perm(ae3, ao3, 0x7632));
      }
    }
  } else
  for (int c0 = 0; c0 < nw; c0 += 1024) {
    const int q = c0 + tid;
    const bool v0 = q < nw, v1 = q + 256 < nw, v2 = q + 512 < nw, v3 = q + 768 < nw;
    for (int jj = 0; jj < nrows; ++jj) {
      const int y0 = __ldg(ty0 + jj), y1 = __ldg(ty1 + jj);
      uint32_t ae0 = 0, ao0 = 0, ae1 = 0, ao1 = 0, ae2 = 0, ao2 = 0, ae3 = 0, ao3 = 0;
      if (ALIGNED) {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(frames + fbase) + (size_t)y0 * (rowbytes >> 2) + q;
        for (int y = y0; y < y1; ++y, row += rowbytes >> 2) {
          const uint32_t w0 = v0 ? __ldg(row) : 0u;
          const uint32_t w1 = v1 ? __ldg(row + 256) : 0u;
          const uint32_t w2 = v2 ? __ldg(row + 512) : 0u;
          const uint32_t w3 = v3 ? __ldg(row + 768) : 0u;
          ae0 += w0 & 0x00FF00FFu; ao0 += __byte_perm(w0, 0, 0x4341);
          ae1 += w1 & 0x00FF00FFu; ao1 += __byte_perm(w1, 0, 0x4341);
          ae2 += w2 & 0x00FF00FFu; ao2 += __byte_perm(w2, 0, 0x4341);
          ae3 += w3 & 0x00FF00FFu; ao3 += __byte_perm(w3, 0, 0x4341);
        }
      } else {
        size_t off = fbase + (size_t)y0 * rowbytes + 4 * (size_t)q;
        for (int y = y0; y < y1; ++y, off += rowbytes) {
          const uint32_t w0 = v0 ? load_word<false>(frames, off, total_bytes) : 0u;
          const uint32_t w1 = v1 ? load_word<false>(frames, off + 1024, total_bytes) : 0u;
          const uint32_t w2 = v2 ? load_word<false>(frames, off + 2048, total_bytes) : 0u;
          const uint32_t w3 = v3 ? load_word<false>(frames, off + 3072, total_bytes) : 0u;
          ae0 += w0 & 0x00FF00FFu; ao0 += __byte_perm(w0, 0, 0x4341);
          ae1 += w1 & 0x00FF00FFu; ao1 += __byte_perm(w1, 0, 0x4341);
          ae2 += w2 & 0x00FF00FFu; ao2 += __byte_perm(w2, 0, 0x4341);
          ae3 += w3 & 0x00FF00FFu; ao3 += __byte_perm(w3, 0, 0x4341);
        }
      }
      // bytes (0,2) live in ae, (1,3) in ao -> u16 sums in byte order
      uint2* vrow = reinterpret_cast<uint2*>(vs) + jj * nw + q;
      if (v0) vrow[0] = make_uint2(__byte_perm(ae0, ao0, 0x5410), __byte_perm(ae0, ao0, 0x7632));
      if (v1) vrow[256] = make_uint2(__byte_perm(ae1, ao1, 0x5410), __byte_perm(ae1, ao1, 0x7632));
      if (v2) vrow[512] = make_uint2(__byte_perm(ae2, ao2, 0x5410), __byte_perm(ae2, ao2, 0x7632));
      if (v3) vrow[768] = make_uint2(__byte_perm(ae3, ao3, 0x5410), __byte_perm(ae3, ao3, 0x7632));
    }
  }
  __syncthreads();

  const uint16_t* v16 = reinterpret_cast<const uint16_t*>(vs);
  typedef typename PyrOut<OUT>::T OT;
  const int pitch = p.pitch[lvl];                     // OUT 0: floats per row; OUT 1: pixels per row (2 x pair pitch)
  const long long plane = OUT == 0 ? (long long)hs * pitch : lo_off;
  OT* obase = out + p.off[lvl] + (size_t)b * (OUT == 0 ? 3 : 1) * hs * pitch + (size_t)j0 * pitch;
  const bool fast = p.fastdiv[lvl] != 0;
  if (p.kwmin[lvl] > 0) {
    const int2* tw = reinterpret_cast<const int2*>(t);
    const uint32_t magic = p.magic[lvl];
    switch (p.kwmin[lvl]) {
      case 1: hpass_fixed<1, OUT>(v16, 4 * nw, rowc, tw, ws, nrows, magic, tid, fast, obase, pitch, plane); break;
      case 2: hpass_fixed<2, OUT>(v16, 4 * nw, rowc, tw, ws, nrows, magic, tid, fast, obase, pitch, plane); break;
      case 3: hpass_fixed<3, OUT>(v16, 4 * nw, rowc, tw, ws, nrows, magic, tid, fast, obase, pitch, plane); break;
      case 4: hpass_fixed<4, OUT>(v16, 4 * nw, rowc, tw, ws, nrows, magic, tid, fast, obase, pitch, plane); break;
      default: hpass_fixed<5, OUT>(v16, 4 * nw, rowc, tw, ws, nrows, magic, tid, fast, obase, pitch, plane); break;
    }
    return;
  }
  for (int jj = 0; jj < nrows; ++jj) {
    const float fkh = rowc[jj].x, rkh = rowc[jj].y;
    const uint16_t* vr = v16 + jj * (4 * nw);
    // per-thread running pointers: one 64-bit add per plane and step instead of rebuilding three addresses per pixel
    OT* o0 = obase + jj * pitch + tid;
    const int2* tw = reinterpret_cast<const int2*>(t) + tid;      // {x0 | kw << 16, bits of RN(1 / kw)}
    for (int i = tid; i < ws; i += 256, o0 += 256, tw += 256) {
      const int2 e = __ldg(tw);
      const int xw = e.x;
      const float rkw = __int_as_float(e.y);
      const int kw = xw >> 16;
      const uint16_t* vp = vr + 3 * (xw & 0xFFFF);
      int s0 = 0, s1 = 0, s2 = 0;
      for (int x = 0; x < kw; ++x, vp += 3) { s0 += vp[0]; s1 += vp[1]; s2 += vp[2]; }
      const float fkw = (float)kw;
      float a0, a1, a2;
      if (fast) {
        a0 = div_small(div_small((float)s0, fkh, rkh), fkw, rkw);
        a1 = div_small(div_small((float)s1, fkh, rkh), fkw, rkw);
        a2 = div_small(div_small((float)s2, fkh, rkh), fkw, rkw);
      } else {
        a0 = __fdiv_rn(__fdiv_rn((float)s0, fkh), fkw);
        a1 = __fdiv_rn(__fdiv_rn((float)s1, fkh), fkw);
        a2 = __fdiv_rn(__fdiv_rn((float)s2, fkh), fkw);
      }
      PyrOut<OUT>::store(o0, plane, __fmul_rn(__fsub_rn(a0, 127.5f), 0.0078125f), __fmul_rn(__fsub_rn(a1, 127.5f), 0.0078125f),
                         __fmul_rn(__fsub_rn(a2, 127.5f), 0.0078125f));
    }
  }
  // pad columns [ws, pitch) stay untouched: readers clip at ws
}

// adaptive_avg_pool2d windows: [floor(i*In/Out), ceil((i+1)*In/Out))
static void window_table(int In, int Out, std::vector<int>& lo, std::vector<int>& hi) {
  lo.resize(Out);
  hi.resize(Out);
  for (int i = 0; i < Out; ++i) {
    lo[i] = (int)(((long long)i * In) / Out);
    hi[i] = (int)(((long long)(i + 1) * In + Out - 1) / Out);
  }
}

// Builds (and caches per frame shape) the window tables and verifies the fast division for every (kh, kw) pair.
static int build_pyramid_tables(trl_ctx* c, int H, int W, const PyramidGeom& g, cudaStream_t s) {
  if (c->pyr_tab_H == H && c->pyr_tab_W == W && c->d_pyr_tab != nullptr) return TRL_OK;
  if (W > 65535) TRL_FAIL(c, TRL_E_INVALID, "frame width %d too large", W);
  std::vector<int> tab;
  std::vector<int2> pairs;
  std::vector<int> pair_level;
  for (int k = 0; k < g.n; ++k) {
    c->pyr_tab_off[k] = (int)tab.size();
    std::vector<int> x0, x1, y0, y1;
    window_table(W, g.ws[k], x0, x1);
    window_table(H, g.hs[k], y0, y1);
    for (int i = 0; i < g.ws[k]; ++i) {                  // interleaved pairs: one 64-bit load per output pixel
      tab.push_back(x0[i] | ((x1[i] - x0[i]) << 16));
      const float r = 1.0f / (float)(x1[i] - x0[i]);     // IEEE: correctly rounded reciprocal
      int bits;
      memcpy(&bits, &r, 4);
      tab.push_back(bits);
    }
    tab.insert(tab.end(), y0.begin(), y0.end());
    tab.insert(tab.end(), y1.begin(), y1.end());
    std::vector<int> khs, kws;
    for (int i = 0; i < g.hs[k]; ++i) khs.push_back(y1[i] - y0[i]);
    for (int i = 0; i < g.ws[k]; ++i) kws.push_back(x1[i] - x0[i]);
    std::sort(khs.begin(), khs.end()); khs.erase(std::unique(khs.begin(), khs.end()), khs.end());
    std::sort(kws.begin(), kws.end()); kws.erase(std::unique(kws.begin(), kws.end()), kws.end());
    // adaptive-pool windows of one level are kwmin or kwmin + 1 wide; the fine levels (87 % of the pixels) have kwmin <= 5
    c->pyr_kwmin[k] = (kws.back() <= kws.front() + 1 && kws.front() >= 1 && kws.front() <= 5) ? kws.front() : 0;
    for (int a : khs)
      for (int b : kws) { pairs.push_back(make_int2(a, b)); pair_level.push_back(k); }
  }
  if (c->d_pyr_tab) { TRL_CUDA(c, cudaStreamSynchronize(s)); TRL_CUDA(c, cudaFree(c->d_pyr_tab)); c->d_pyr_tab = nullptr; }
  TRL_CUDA(c, cudaMalloc(&c->d_pyr_tab, tab.size() * sizeof(int)));
  TRL_CUDA(c, cudaMemcpy(c->d_pyr_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
  // exhaustive check of the 3-instruction division on this shape's windows (a few million quotients, once per shape)
  int2* d_pairs = nullptr;
  int* d_bad = nullptr;
  TRL_CUDA(c, cudaMalloc(&d_pairs, pairs.size() * sizeof(int2)));
  TRL_CUDA(c, cudaMalloc(&d_bad, pairs.size() * sizeof(int)));
  TRL_CUDA(c, cudaMemcpy(d_pairs, pairs.data(), pairs.size() * sizeof(int2), cudaMemcpyHostToDevice));
  TRL_CUDA(c, cudaMemset(d_bad, 0, pairs.size() * sizeof(int)));
  div_verify_kernel<<<dim3(64, (unsigned)pairs.size()), 256, 0, s>>>(d_pairs, d_bad);
  TRL_LAUNCH_CHECK(c);
  std::vector<int> bad(pairs.size());
  TRL_CUDA(c, cudaMemcpyAsync(bad.data(), d_bad, pairs.size() * sizeof(int), cudaMemcpyDeviceToHost, s));
  TRL_CUDA(c, cudaStreamSynchronize(s));
  cudaFree(d_pairs);
  cudaFree(d_bad);
  for (int k = 0; k < g.n; ++k) c->pyr_fastdiv[k] = 1;
  for (size_t i = 0; i < pairs.size(); ++i)
    if (bad[i]) c->pyr_fastdiv[pair_level[i]] = 0;
  c->pyr_tab_H = H;
  c->pyr_tab_W = W;
  return TRL_OK;
}

// `padded`: rows at g.pitch[k] floats and levels at g.off[k]*B (the cascade's internal layout, 16-byte aligned rows
// for P-Net's cp.async staging); otherwise the compact layout of the trl_pyramid stage entry point.
static int launch_pyramid_any(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const PyramidGeom& g, float* d_out,
                              bool padded, uint4* d_hi, uint4* d_lo, cudaStream_t s);

int launch_pyramid(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const PyramidGeom& g, float* d_out,
                   bool padded, cudaStream_t s) {
  return launch_pyramid_any(c, d_frames, B, H, W, g, d_out, padded, nullptr, nullptr, s);
}

int launch_pyramid_pairs(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const PyramidGeom& g, uint4* d_hi, uint4* d_lo,
                         cudaStream_t s) {
  return launch_pyramid_any(c, d_frames, B, H, W, g, nullptr, true, d_hi, d_lo, s);
}

static int launch_pyramid_any(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const PyramidGeom& g, float* d_out,
                              bool padded, uint4* d_hi, uint4* d_lo, cudaStream_t s) {
  if (g.n == 0 || B == 0) return TRL_OK;
  const bool pairs = d_hi != nullptr;
  int rc = build_pyramid_tables(c, H, W, g, s);
  if (rc != TRL_OK) return rc;
  if ((reinterpret_cast<uintptr_t>(d_frames) & 3) != 0) TRL_FAIL(c, TRL_E_INVALID, "frames pointer must be 4-byte aligned");
  const int nw = (3 * W + 3) / 4;
  PyrParams p{};
  p.n = g.n;
  int blocks = 0, rmax = 1;
  long long off = 0;
  for (int k = 0; k < g.n; ++k) {
    const int kh = (H + g.hs[k] - 1) / g.hs[k] + 1;
    if (kh > PYR_MAX_KH) TRL_FAIL(c, TRL_E_INVALID, "pyramid window of %d rows exceeds %d (frame %dx%d)", kh, PYR_MAX_KH, H, W);
    p.hs[k] = g.hs[k];
    p.ws[k] = g.ws[k];
    p.pitch[k] = pairs ? 2 * g.pitch2[k] : padded ? g.pitch[k] : g.ws[k];
    p.off[k] = pairs ? 2 * g.off2[k] * B : padded ? g.off[k] * B : off * B;
    off += 3LL * g.hs[k] * g.ws[k];
    p.tab_off[k] = c->pyr_tab_off[k];
    p.fastdiv[k] = c->pyr_fastdiv[k];
    p.kwmin[k] = c->pyr_kwmin[k];
    p.magic[k] = (unsigned)((1ull << 32) / (unsigned)g.ws[k]) + 1u;
    // rows per CTA: about 8-12 source rows of work, bounded by 64 KB of column sums
    int R = (int)(8.0 * g.hs[k] / H);
    R = std::max(1, std::min(R, 4));
    while (R > 1 && (size_t)R * nw * 8 > 64 * 1024) --R;
    p.rows[k] = R;
    rmax = std::max(rmax, R);
    p.blk_start[k] = blocks;
    blocks += ceil_div(g.hs[k], R);
  }
  for (int q = g.n; q <= TRL_MAX_SCALES; ++q) p.blk_start[q] = blocks;
  const size_t smem = (size_t)rmax * nw * 8;
  if (smem > 200 * 1024) TRL_FAIL(c, TRL_E_INVALID, "frame width %d too large for the pyramid kernel", W);
  const size_t total = (size_t)B * H * W * 3;
  const bool aligned = (3 * W) % 4 == 0;
  if (!c->pyr_smem_set) {
    // The opt-in is a property of the kernel function, shared by every context of the process: always raise it to the
    // kernel's ceiling (a per-context "largest size so far" let a second context with smaller frames lower it under the
    // first one's feet -> invalid-argument launches).  The launch itself still asks only for what this shape needs.
    const int cap = 200 * 1024;
    TRL_CUDA(c, cudaFuncSetAttribute(pyramid_sep_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    TRL_CUDA(c, cudaFuncSetAttribute(pyramid_sep_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    TRL_CUDA(c, cudaFuncSetAttribute(pyramid_sep_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    TRL_CUDA(c, cudaFuncSetAttribute(pyramid_sep_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    TRL_CUDA(c, cudaFuncSetAttribute(pyramid_sep_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    TRL_CUDA(c, cudaFuncSetAttribute(pyramid_sep_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    c->pyr_smem_set = cap;
  }
  // (level, first row) of every CTA of one frame, rebuilt when the frame shape (hence the row grouping) changes
  const long long blk_key = ((long long)H << 40) ^ ((long long)W << 16) ^ (long long)blocks;
  if (c->d_pyr_blk == nullptr || c->pyr_blk_key != blk_key) {
    std::vector<int2> blks;
    for (int k = 0; k < g.n; ++k)
      for (int j0 = 0; j0 < g.hs[k]; j0 += p.rows[k]) blks.push_back(make_int2(k, j0));
    if (c->d_pyr_blk) { TRL_CUDA(c, cudaStreamSynchronize(s)); TRL_CUDA(c, cudaFree(c->d_pyr_blk)); c->d_pyr_blk = nullptr; }
    TRL_CUDA(c, cudaMalloc(&c->d_pyr_blk, blks.size() * sizeof(int2)));
    TRL_CUDA(c, cudaMemcpy(c->d_pyr_blk, blks.data(), blks.size() * sizeof(int2), cudaMemcpyHostToDevice));
    c->pyr_blk_key = blk_key;
  }
  const int2* blk_tab = reinterpret_cast<const int2*>(c->d_pyr_blk);
  dim3 grid(blocks, B);
  const bool rows16 = (3 * W) % 16 == 0 && (reinterpret_cast<uintptr_t>(d_frames) & 15) == 0;
  if (pairs) {
    uint2* hi = reinterpret_cast<uint2*>(d_hi);
    const long long lo_off = reinterpret_cast<uint2*>(d_lo) - hi;
    if (rows16) pyramid_sep_kernel<2, 1><<<grid, 256, smem, s>>>(d_frames, H, W, total, p, c->d_pyr_tab, blk_tab, hi, lo_off);
    else if (aligned) pyramid_sep_kernel<1, 1><<<grid, 256, smem, s>>>(d_frames, H, W, total, p, c->d_pyr_tab, blk_tab, hi, lo_off);
    else pyramid_sep_kernel<0, 1><<<grid, 256, smem, s>>>(d_frames, H, W, total, p, c->d_pyr_tab, blk_tab, hi, lo_off);
  } else if (rows16) pyramid_sep_kernel<2, 0><<<grid, 256, smem, s>>>(d_frames, H, W, total, p, c->d_pyr_tab, blk_tab, d_out, 0);
  else if (aligned) pyramid_sep_kernel<1, 0><<<grid, 256, smem, s>>>(d_frames, H, W, total, p, c->d_pyr_tab, blk_tab, d_out, 0);
  else pyramid_sep_kernel<0, 0><<<grid, 256, smem, s>>>(d_frames, H, W, total, p, c->d_pyr_tab, blk_tab, d_out, 0);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

// ----------------------------------------------------------------------------- K7 crop-resample

// One CTA per candidate slot.  grid = (n_max slots).  The crop is rows [y-1, ey) x cols [x-1, ex) (0-based),
// clipped by pad() upstream, never zero padded.  Output float32 [slot][3][S][S].
// d_count (optional) holds the live number of slots; slots >= *d_count exit.
__global__ void __launch_bounds__(256) crop_resample_kernel(const uint8_t* __restrict__ frames, int H, int W,
                                                           const int* __restrict__ pad4, const int* __restrict__ img,
                                                           const int* __restrict__ d_count, int per_frame_cap,
                                                           int n_frames, int n_slots, int split, int S,
                                                           float* __restrict__ out) {
  // per-frame lists: grid-stride over the compacted live slots (common.cuh slot map); flat lists: one CTA per slot
  __shared__ int s_pref[SLOTMAP_MAX_FRAMES + 1];
  const bool compact = per_frame_cap > 0 && n_frames > 0 && n_frames <= SLOTMAP_MAX_FRAMES;
  // a candidate is split over `split` CTAs (interleaved pixel iterations): with a few hundred candidates per launch the
  // stage time is the latency of one candidate, not throughput
  const int total_work = (compact ? slotmap_init(d_count, n_frames, per_frame_cap, s_pref) : n_slots) * split;
  for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
  const int part = work % split;
  const int slot = compact ? slotmap_slot(s_pref, n_frames, per_frame_cap, work / split) : work / split;
  int b, n_live;
  if (per_frame_cap > 0) {          // slot = frame * cap + i, counts per frame
    b = slot / per_frame_cap;
    n_live = d_count[b];
    if (slot - b * per_frame_cap >= n_live) continue;
  } else {
    if (d_count && slot >= *d_count) continue;
    b = img[slot];
  }
  const int y = pad4[slot * 4 + 0], ey = pad4[slot * 4 + 1], x = pad4[slot * 4 + 2], ex = pad4[slot * 4 + 3];
  float* o = out + (size_t)slot * 3 * S * S;
  const int ch = ey - (y - 1), cw = ex - (x - 1);
  if (ch <= 0 || cw <= 0) {   // degenerate: upstream skips such crops (detect_face.py stage 2/3 loop guard)
    if (part == 0)
      for (int i = threadIdx.x; i < 3 * S * S; i += blockDim.x) o[i] = 0.f;
    continue;
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
  for (int it = part; it < iters; it += split) {
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
}

// d_count semantics: if per_frame_cap > 0, d_count is int[B] and slots are frame*cap+i; else a single int (or null).
int launch_crop_resample_ex(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const int* d_pad, const int* d_img,
                            const int* d_count, int per_frame_cap, int n_slots, int size, float* d_out, cudaStream_t s) {
  if (n_slots <= 0) return TRL_OK;
  const int n_frames = per_frame_cap > 0 ? n_slots / per_frame_cap : 0;
  const bool compact = per_frame_cap > 0 && n_frames <= SLOTMAP_MAX_FRAMES;
  const int split = compact ? (size == 24 ? 4 : 8) : 1;
  const long long slots = (long long)n_slots * split;
  const int grid = compact ? (int)(slots < c->num_sms * 8 ? slots : c->num_sms * 8) : n_slots;
  crop_resample_kernel<<<grid, 256, 0, s>>>(d_frames, H, W, d_pad, d_img, d_count, per_frame_cap, n_frames, n_slots, split, size, d_out);
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

// ----------------------------------------------------------------------------- K10, mode B: extract_face (INTER_AREA)
// Upstream facenet_pytorch extract_face (SURVEY.md Appendix A "Mode-B extras"; oracle/mode_b.py): margin-adjusted box in
// fp32, truncated and clamped, crop, cv2.resize(..., (S, S), interpolation=cv2.INTER_AREA) on uint8 -- bit exact with
// OpenCV's three code paths (oracle.mode_b.resize_area_u8 is the numpy restatement this mirrors, pinned against cv2):
//   both axes shrink by integers : block sums, (s + 2) >> 2 for 2x2, else saturate(rint(s * (1.f / area)))
//   both axes shrink             : separable fp32 weights from computeResizeAreaTab (double -> float), accumulated in table
//                                  order with separate multiplies and adds, saturate(rint(sum))
//   an axis is enlarged          : the 11-bit fixed-point bilinear resize with area-mode coefficients
// One CTA per face.  keep_all: face i of the batch = box j of frame b through the prefix `face_off` (faces_prefix_kernel).
constexpr int AREA_MAX_S = 256;

struct AreaTab {      // one destination index: up to three weights over source indices [s_first .. s_last]
  int s_mid0, s_mid1; // full cells [s_mid0, s_mid1)
  float a_first, a_mid, a_last;   // a_first applies to s_mid0 - 1 (if has_first), a_last to s_mid1 (if has_last)
  int has_first, has_last;
};

__device__ __forceinline__ AreaTab area_tab(int d, int ssize, double scale) {
  AreaTab t;
  const double fsx1 = (double)d * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, (double)ssize - fsx1);
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  t.s_mid0 = sx1; t.s_mid1 = sx2;
  t.has_first = ((double)sx1 - fsx1 > 1e-3) ? 1 : 0;
  t.has_last = (fsx2 - (double)sx2 > 1e-3) ? 1 : 0;
  t.a_first = (float)(((double)sx1 - fsx1) / cell);
  t.a_mid = (float)(1.0 / cell);
  t.a_last = (float)(fmin(fmin(fsx2 - (double)sx2, 1.0), cell) / cell);
  return t;
}

// bilinear coefficients in area mode (cv::resize with INTER_AREA when an axis is enlarged)
__device__ __forceinline__ ResizeCoef resize_coef_area(int d, int ssize, double scale, double inv_scale) {
  int sidx = (int)floor((double)d * scale);
  float f = (float)((double)(d + 1) - (double)(sidx + 1) * inv_scale);
  f = f <= 0.f ? 0.f : __fsub_rn(f, floorf(f));
  if (sidx < 0) { f = 0.f; sidx = 0; }
  if (sidx >= ssize - 1) { f = 0.f; sidx = ssize - 1; }
  ResizeCoef r;
  r.idx = sidx;
  r.c1 = __float2int_rn(__fmul_rn(f, 2048.f));
  r.c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  return r;
}

// face_off[b] = number of faces of frames < b (each clipped to cap); face_off[B] = total.  One CTA.
__global__ void __launch_bounds__(256) faces_prefix_kernel(const int* __restrict__ nfaces, int B, int cap, int max_faces,
                                                          int* __restrict__ face_off, CapFlag* capflag) {
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 256) {
    const int i = base + threadIdx.x;
    int v = i < B ? min(max(nfaces[i], 0), cap) : 0;
    // inclusive scan over the 256 threads (warp scans + warp totals)
    __shared__ int wsum[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    int add = carry;
    for (int q = 0; q < w; ++q) add += wsum[q];
    if (i < B) face_off[i] = add + x - v;
    __syncthreads();
    if (threadIdx.x == 255) carry = add + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    face_off[B] = carry;
    if (carry > max_faces && capflag) { capflag->overflow = 1; capflag->stage = 6; capflag->frame = 0; capflag->count = carry; capflag->capacity = max_faces; }
  }
}

__global__ void __launch_bounds__(256) crop_area_kernel(const uint8_t* __restrict__ frames, int B, int H, int W,
                                                       const float* __restrict__ boxes, int box_stride,
                                                       const int* __restrict__ nfaces, const int* __restrict__ face_off,
                                                       int max_faces, int S, int margin, int* __restrict__ box_int,
                                                       uint8_t* __restrict__ valid, int* __restrict__ face_frame,
                                                       uint8_t* __restrict__ crops) {
  __shared__ int sb[6];          // x1, y1, x2, y2, ok, frame
  __shared__ AreaTab xt[AREA_MAX_S], yt[AREA_MAX_S];
  __shared__ ResizeCoef xc[AREA_MAX_S], yc[AREA_MAX_S];
  const int i = blockIdx.x;
  if (threadIdx.x == 0) {
    int b = i, j = 0, present = 0;
    if (face_off) {                                  // keep_all: face i -> (frame, box) by binary search in the prefix
      const int total = min(face_off[B], max_faces);
      if (i < total) {
        int lo = 0, hi = B;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (face_off[mid] <= i) lo = mid; else hi = mid; }
        b = lo; j = i - face_off[lo]; present = 1;
      } else {
        b = -1;
      }
    } else {
      present = nfaces[b] > 0 ? 1 : 0;
    }
    int x1 = 0, y1 = 0, x2 = 0, y2 = 0, ok = 0;
    if (present) {
      const float* bx = boxes + (size_t)b * box_stride + (size_t)j * 5;
      // margin = [margin * (x2 - x1) / (S - margin), margin * (y2 - y1) / (S - margin)], float32 scalars upstream
      const float mx = __fdiv_rn(__fmul_rn((float)margin, __fsub_rn(bx[2], bx[0])), (float)(S - margin));
      const float my = __fdiv_rn(__fmul_rn((float)margin, __fsub_rn(bx[3], bx[1])), (float)(S - margin));
      x1 = (int)fmaxf(__fsub_rn(bx[0], __fmul_rn(mx, 0.5f)), 0.f);
      y1 = (int)fmaxf(__fsub_rn(bx[1], __fmul_rn(my, 0.5f)), 0.f);
      x2 = (int)fminf(__fadd_rn(bx[2], __fmul_rn(mx, 0.5f)), (float)W);
      y2 = (int)fminf(__fadd_rn(bx[3], __fmul_rn(my, 0.5f)), (float)H);
      ok = (x2 > x1 && y2 > y1) ? 1 : 0;
    }
    sb[0] = x1; sb[1] = y1; sb[2] = x2; sb[3] = y2; sb[4] = ok; sb[5] = b;
    if (b >= 0) {
      box_int[i * 4 + 0] = x1; box_int[i * 4 + 1] = y1; box_int[i * 4 + 2] = x2; box_int[i * 4 + 3] = y2;
      valid[i] = (uint8_t)ok;
      if (face_frame) face_frame[i] = b;
    }
  }
  __syncthreads();
  if (sb[5] < 0) return;                              // beyond the number of faces of this batch
  uint8_t* o = crops + (size_t)i * S * S * 3;
  if (!sb[4]) {
    for (int k = threadIdx.x; k < S * S * 3; k += blockDim.x) o[k] = 0;
    return;
  }
  const int b = sb[5], x1 = sb[0], y1 = sb[1], sw = sb[2] - sb[0], sh = sb[3] - sb[1];
  const uint8_t* src = frames + ((size_t)b * H * W + (size_t)y1 * W + x1) * 3;
  const size_t rowb = (size_t)W * 3;
  const double inv_x = (double)S / (double)sw, inv_y = (double)S / (double)sh;
  const double scale_x = 1.0 / inv_x, scale_y = 1.0 / inv_y;
  if (scale_x >= 1.0 && scale_y >= 1.0) {
    const int isx = (int)rint(scale_x), isy = (int)rint(scale_y);
    const double eps = 2.220446049250313e-16;
    if (fabs(scale_x - (double)isx) < eps && fabs(scale_y - (double)isy) < eps) {
      // integer ratios (resizeAreaFast_)
      const float fscale = __fdiv_rn(1.f, (float)(isx * isy));
      const bool two = isx == 2 && isy == 2;
      for (int pix = threadIdx.x; pix < S * S; pix += blockDim.x) {
        const int dy = pix / S, dx = pix - dy * S;
        int s0 = 0, s1 = 0, s2 = 0;
        for (int ky = 0; ky < isy; ++ky) {
          const uint8_t* r = src + (size_t)(dy * isy + ky) * rowb + (size_t)dx * isx * 3;
          for (int kx = 0; kx < isx; ++kx, r += 3) { s0 += r[0]; s1 += r[1]; s2 += r[2]; }
        }
        int v0, v1, v2;
        if (two) { v0 = (s0 + 2) >> 2; v1 = (s1 + 2) >> 2; v2 = (s2 + 2) >> 2; }
        else {
          v0 = __float2int_rn(__fmul_rn((float)s0, fscale));
          v1 = __float2int_rn(__fmul_rn((float)s1, fscale));
          v2 = __float2int_rn(__fmul_rn((float)s2, fscale));
        }
        o[pix * 3 + 0] = (uint8_t)min(max(v0, 0), 255);
        o[pix * 3 + 1] = (uint8_t)min(max(v1, 0), 255);
        o[pix * 3 + 2] = (uint8_t)min(max(v2, 0), 255);
      }
      return;
    }
    // general area path (resizeArea_<uchar, float>)
    for (int d = threadIdx.x; d < S; d += blockDim.x) { xt[d] = area_tab(d, sw, scale_x); yt[d] = area_tab(d, sh, scale_y); }
    __syncthreads();
    for (int pix = threadIdx.x; pix < S * S; pix += blockDim.x) {
      const int dy = pix / S, dx = pix - dy * S;
      const AreaTab tx = xt[dx], ty = yt[dy];
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f;
      const int ya = ty.s_mid0 - ty.has_first, yb = ty.s_mid1 + ty.has_last;      // source rows [ya, yb)
      for (int sy = ya; sy < yb; ++sy) {
        const float beta = (sy < ty.s_mid0) ? ty.a_first : (sy < ty.s_mid1 ? ty.a_mid : ty.a_last);
        const uint8_t* r = src + (size_t)sy * rowb;
        float b0 = 0.f, b1 = 0.f, b2 = 0.f;
        const int xa = tx.s_mid0 - tx.has_first, xb = tx.s_mid1 + tx.has_last;
        for (int sx = xa; sx < xb; ++sx) {
          const float alpha = (sx < tx.s_mid0) ? tx.a_first : (sx < tx.s_mid1 ? tx.a_mid : tx.a_last);
          const uint8_t* q = r + (size_t)sx * 3;
          b0 = __fadd_rn(b0, __fmul_rn((float)q[0], alpha));
          b1 = __fadd_rn(b1, __fmul_rn((float)q[1], alpha));
          b2 = __fadd_rn(b2, __fmul_rn((float)q[2], alpha));
        }
        sum0 = __fadd_rn(sum0, __fmul_rn(beta, b0));
        sum1 = __fadd_rn(sum1, __fmul_rn(beta, b1));
        sum2 = __fadd_rn(sum2, __fmul_rn(beta, b2));
      }
      o[pix * 3 + 0] = (uint8_t)min(max(__float2int_rn(sum0), 0), 255);
      o[pix * 3 + 1] = (uint8_t)min(max(__float2int_rn(sum1), 0), 255);
      o[pix * 3 + 2] = (uint8_t)min(max(__float2int_rn(sum2), 0), 255);
    }
    return;
  }
  // an axis is enlarged: bilinear fixed point with area-mode coefficients
  for (int d = threadIdx.x; d < S; d += blockDim.x) {
    xc[d] = resize_coef_area(d, sw, scale_x, inv_x);
    yc[d] = resize_coef_area(d, sh, scale_y, inv_y);
  }
  __syncthreads();
  for (int pix = threadIdx.x; pix < S * S; pix += blockDim.x) {
    const int dy = pix / S, dx = pix - dy * S;
    const ResizeCoef cx = xc[dx], cy = yc[dy];
    const int xa = cx.idx, xb = min(cx.idx + 1, sw - 1);
    const int ya = cy.idx, yb = min(cy.idx + 1, sh - 1);
    const uint8_t* ra = src + (size_t)ya * rowb;
    const uint8_t* rb = src + (size_t)yb * rowb;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int h0 = ra[xa * 3 + ch] * cx.c0 + ra[xb * 3 + ch] * cx.c1;
      const int h1 = rb[xa * 3 + ch] * cx.c0 + rb[xb * 3 + ch] * cx.c1;
      int v = (((cy.c0 * (h0 >> 4)) >> 16) + ((cy.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
      o[pix * 3 + ch] = (uint8_t)min(max(v, 0), 255);
    }
  }
}

int launch_extract_face(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const float* d_boxes, int box_stride,
                        const int* d_nfaces, int S, int margin, int* d_box_int, uint8_t* d_valid, uint8_t* d_crops,
                        cudaStream_t s) {
  if (B <= 0) return TRL_OK;
  if (S < 1 || S > AREA_MAX_S || margin < 0 || margin >= S) TRL_FAIL(c, TRL_E_INVALID, "extract_face: image size %d / margin %d unsupported", S, margin);
  crop_area_kernel<<<B, 256, 0, s>>>(d_frames, B, H, W, d_boxes, box_stride, d_nfaces, nullptr, B, S, margin, d_box_int, d_valid,
                                     nullptr, d_crops);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

int launch_extract_faces_all(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const float* d_boxes, int box_stride,
                             int box_cap, const int* d_nfaces, int S, int margin, int max_faces, int* d_face_off,
                             int* d_face_frame, int* d_box_int, uint8_t* d_valid, uint8_t* d_crops, cudaStream_t s) {
  if (B <= 0 || max_faces <= 0) return TRL_OK;
  if (S < 1 || S > AREA_MAX_S || margin < 0 || margin >= S) TRL_FAIL(c, TRL_E_INVALID, "extract_face: image size %d / margin %d unsupported", S, margin);
  faces_prefix_kernel<<<1, 256, 0, s>>>(d_nfaces, B, box_cap, max_faces, d_face_off, c->d_cap);
  TRL_LAUNCH_CHECK(c);
  crop_area_kernel<<<max_faces, 256, 0, s>>>(d_frames, B, H, W, d_boxes, box_stride, d_nfaces, d_face_off, max_faces, S, margin,
                                             d_box_int, d_valid, d_face_frame, d_crops);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

// ----------------------------------------------------------------------------- pyramid geometry (host)

// upstream detect_face: m = 12/minsize; scales m*factor^k while min(h,w)*m*factor^k >= 12, all in doubles;
// level size (int(h*scale+1), int(w*scale+1)); P-Net map = ceil((hs-2)/2) - 4.
int compute_geometry(const trl_config_t& cfg, int H, int W, PyramidGeom* g) {
  g->n = 0;
  g->px_total = 0;
  g->floats_total = 0;
  g->pairs_total = 0;
  if (H <= 0 || W <= 0 || cfg.min_face_size <= 0) return TRL_E_INVALID;
  const double m = 12.0 / (double)cfg.min_face_size;
  double minl = (double)(H < W ? H : W) * m;
  double scale_i = m;
  long long off = 0, off2 = 0;
  while (minl >= 12.0) {
    if (g->n >= TRL_MAX_SCALES) return TRL_E_INVALID;
    const int k = g->n++;
    g->scale[k] = scale_i;
    g->scale_f[k] = (float)scale_i;
    g->hs[k] = (int)((double)H * scale_i + 1.0);
    g->ws[k] = (int)((double)W * scale_i + 1.0);
    g->oh[k] = (g->hs[k] - 2 + 1) / 2 - 4;
    g->ow[k] = (g->ws[k] - 2 + 1) / 2 - 4;
    g->pitch[k] = (g->ws[k] + 3) & ~3;               // rows padded to 16 bytes (internal cascade layout)
    g->off[k] = off;
    off += 3LL * g->hs[k] * g->pitch[k];
    g->pitch2[k] = (g->ws[k] + 1) / 2;
    g->off2[k] = off2;
    off2 += (long long)g->hs[k] * g->pitch2[k];
    g->pairs_total = off2;
    g->px_total += (long long)g->hs[k] * g->ws[k];
    g->floats_total = off;
    scale_i = scale_i * cfg.factor;
    minl = minl * cfg.factor;
  }
  return TRL_OK;
}
