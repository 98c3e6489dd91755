// K1 for the cascade (pnet_precision 3): the fp16 hi / lo pair images of every pyramid level, for frames whose width is a
// multiple of 16 pixels (every production shape: 360p, 720p, 1080p, ...).  Same window tables, the same exact integer window
// sums and the same two IEEE divisions as pyramid_sep_kernel (preproc.cu) -- bit identical, tests/test_gpu_pnet_hybrid.py --
// organised around what bounded that kernel on B200 (ncu, profiles/r02c_pyr_full.md: ALU pipe 64 % busy at its 2-cycle
// issue rate, L1TEX wavefronts 82 %, 56 % of the shared-memory wavefronts bank conflicts of the 2-byte loads of pass 2,
// 10 % of all instructions in the per-CTA level search, a quarter of pass 2's lanes idle because 769 = 3 x 256 + 1):
//
//  * one CTA = (frame, level, R output rows); the (level, first row) of a CTA comes from a table, not from a search.
//  * pass 1 (vertical), one warp per (output row, 160-pixel segment): a lane owns 16 bytes of the source row (one coalesced
//    LDG.128 per row) and adds, per source row and word, the raw word into one register and its odd bytes (PRMT) into a second
//    one -- the even-byte lanes follow once per output row from  sum(w) - (odd << 8)  (mod 2^32; exact, both 16-bit lanes of
//    the result are < 2^16): three instructions per word and row instead of four, and the compiler folds two rows into one
//    three-input add.
//  * transposition, same warp: the interleaved u16 sums go through a warp-private scratch line and come back as one
//    {B, G, R, 0} u16x4 per source column (three conflict-free LDS.32 and one STS.128 per pixel pair).
//  * pass 2 (horizontal): an output pixel reads one LDS.64 per window column (all three channels at once; consecutive
//    threads = consecutive pixels of the flattened [rows][ws] index space of the CTA, so no lanes idle on a 769-pixel row and
//    a warp's loads span a few hundred bytes), adds packed 16-bit lanes, converts with the 2^23 trick (no I2F on the
//    quarter-rate pipe), divides, normalises with one FMA and splits into fp16 hi + lo exactly like PyrOut<1>.
//
// A persistent variant that streams the source rows through a shared-memory ring with bulk copies (one producer thread,
// mbarrier pairs, 4-pixel ownership, no transposition) is kept as experiments/variants/pyramid_stream.cu: bit identical,
// fewer L1 wavefronts, but with two 9-warp CTAs per SM it is latency bound (issue slots 49 % busy) and slower (2.0-2.7 ms
// against 1.46 ms per 225 720p frames).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "pyramid.cuh"

namespace pyrp {

constexpr int SEG_PX = 160;                // pixels per warp task: 30 lanes x 16 bytes
constexpr int SCR_BYTES = 30 * 32;         // scratch line per warp: the 480 interleaved u16 sums of a segment

struct Params {
  int hs[TRL_MAX_SCALES], ws[TRL_MAX_SCALES], pitch[TRL_MAX_SCALES];
  long long off[TRL_MAX_SCALES];           // pixel offset of level k in the hi image (already multiplied by B)
  int tab_off[TRL_MAX_SCALES];
  int fastdiv[TRL_MAX_SCALES];
  int kwmin[TRL_MAX_SCALES];               // > 0: windows of kwmin or kwmin + 1 columns and packed 16-bit lanes cannot overflow
  unsigned magic_ws[TRL_MAX_SCALES];       // floor(2^32 / ws) + 1
  int rows[TRL_MAX_SCALES];                // output rows per CTA of level k
  int H, W;
  int RMAX;                                // rows of the column-sum buffer
  int nseg;                                // 160-pixel segments per row
  unsigned magic_nseg;                     // floor(2^32 / nseg) + 1
};

__device__ __forceinline__ float u16_to_float(uint32_t packed, uint32_t sel) {
  return __fsub_rn(__uint_as_float(__byte_perm(packed, 0x4B000000u, sel)), 8388608.f);
}

template <int KMIN>      // > 0: windows of KMIN or KMIN + 1 columns in packed 16-bit lanes; 0: general
__device__ __forceinline__ void hslot(const uint2* __restrict__ brow, const int2 e, float fkh, float rkh, bool fast,
                                      uint2* __restrict__ o, long long lo_off) {
  const int kw = e.x >> 16;
  const float rkw = __int_as_float(e.y);
  const uint2* vp = brow + (e.x & 0xFFFF);
  float f0, f1, f2, fkw;
  if (KMIN > 0) {
    uint2 a = vp[0];
#pragma unroll
    for (int x = 1; x < KMIN; ++x) { const uint2 w = vp[x]; a.x += w.x; a.y += w.y; }
    const bool wide = kw > KMIN;
    if (wide) { const uint2 w = vp[KMIN]; a.x += w.x; a.y += w.y; }
    fkw = wide ? (float)(KMIN + 1) : (float)KMIN;
    f0 = u16_to_float(a.x, 0x7610);
    f1 = u16_to_float(a.x, 0x7632);
    f2 = u16_to_float(a.y, 0x7610);
  } else {
    uint32_t s0 = 0, s1 = 0, s2 = 0;
    for (int x = 0; x < kw; ++x) { const uint2 w = vp[x]; s0 += w.x & 0xFFFFu; s1 += w.x >> 16; s2 += w.y; }
    f0 = (float)s0; f1 = (float)s1; f2 = (float)s2;
    fkw = (float)kw;
  }
  float a0, a1, a2;
  if (fast) {
    a0 = div_small(div_small(f0, fkh, rkh), fkw, rkw);
    a1 = div_small(div_small(f1, fkh, rkh), fkw, rkw);
    a2 = div_small(div_small(f2, fkh, rkh), fkw, rkw);
  } else {
    a0 = __fdiv_rn(__fdiv_rn(f0, fkh), fkw);
    a1 = __fdiv_rn(__fdiv_rn(f1, fkh), fkw);
    a2 = __fdiv_rn(__fdiv_rn(f2, fkh), fkw);
  }
  // (a - 127.5) * 2^-7 == fma(a, 2^-7, -127.5 * 2^-7): scaling by a power of two commutes with the rounding of the subtraction
  PyrOut<1>::store(o, lo_off, __fmaf_rn(a0, 0.0078125f, -0.99609375f), __fmaf_rn(a1, 0.0078125f, -0.99609375f),
                   __fmaf_rn(a2, 0.0078125f, -0.99609375f));
}

template <int KMIN>
__device__ __forceinline__ void hpass(const uint2* __restrict__ bsum, int cw, const float2* __restrict__ rowc,
                                      const int2* __restrict__ tw, int ws, int nrows, uint32_t magic, bool fast,
                                      uint2* __restrict__ obase, int pitch, long long lo_off) {
  const uint32_t total = (uint32_t)(nrows * ws);
#pragma unroll 2
  for (uint32_t idx = threadIdx.x; idx < total; idx += 256) {
    const uint32_t row = __umulhi(idx, magic);
    const uint32_t col = idx - row * (uint32_t)ws;
    const float2 rc = rowc[row];
    hslot<KMIN>(bsum + row * cw, __ldg(tw + col), rc.x, rc.y, fast, obase + row * pitch + col, lo_off);
  }
}

__global__ void __launch_bounds__(256) pyramid_pairs_kernel(const uint4* __restrict__ frames, const __grid_constant__ Params p,
                                                           const int* __restrict__ tab, const int2* __restrict__ blk_tab,
                                                           uint2* __restrict__ out, long long lo_off) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int2 bi = __ldg(blk_tab + blockIdx.x);
  const int lvl = bi.x, j0 = bi.y;
  const int W = p.W;
  const int hs = p.hs[lvl], ws = p.ws[lvl];
  const int nrows = min(p.rows[lvl], hs - j0);
  const int nq = (3 * W) >> 4;                                  // uint4 per source row
  const int* t = tab + p.tab_off[lvl];
  const int* ty0 = t + 2 * ws + j0;
  const int* ty1 = ty0 + hs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  uint2* bsum = reinterpret_cast<uint2*>(smem);                 // [RMAX][W] {B, G, R, 0} u16 column sums
  unsigned char* scr = smem + (size_t)p.RMAX * W * 8 + warp * SCR_BYTES;
  float2* rowc = reinterpret_cast<float2*>(smem + (size_t)p.RMAX * W * 8 + 8 * SCR_BYTES);     // {kh, RN(1 / kh)} per row
  if (tid < nrows) {
    const float fkh = (float)(__ldg(ty1 + tid) - __ldg(ty0 + tid));
    rowc[tid] = make_float2(fkh, __frcp_rn(fkh));
  }

  const int nseg = p.nseg;
  const uint4* fbase = frames + (size_t)blockIdx.y * p.H * nq;
  const int ntask = nrows * nseg;
  for (int task = warp; task < ntask; task += 8) {
    const int jj = (int)__umulhi((uint32_t)task, p.magic_nseg);
    const int seg = task - jj * nseg;
    const int q = seg * 30 + lane;
    const int y0 = __ldg(ty0 + jj), y1 = __ldg(ty1 + jj);
    if (lane < 30 && q < nq) {
      uint32_t sw0 = 0, sw1 = 0, sw2 = 0, sw3 = 0, ao0 = 0, ao1 = 0, ao2 = 0, ao3 = 0;
      const uint4* row = fbase + (size_t)y0 * nq + q;
      for (int y = y0; y < y1; ++y, row += nq) {
        const uint4 w = __ldg(row);
        sw0 += w.x; ao0 += __byte_perm(w.x, 0, 0x4341);
        sw1 += w.y; ao1 += __byte_perm(w.y, 0, 0x4341);
        sw2 += w.z; ao2 += __byte_perm(w.z, 0, 0x4341);
        sw3 += w.w; ao3 += __byte_perm(w.w, 0, 0x4341);
      }
      const uint32_t ae0 = sw0 - (ao0 << 8), ae1 = sw1 - (ao1 << 8), ae2 = sw2 - (ao2 << 8), ae3 = sw3 - (ao3 << 8);
      uint4* sl = reinterpret_cast<uint4*>(scr + lane * 32);
      sl[0] = make_uint4(__byte_perm(ae0, ao0, 0x5410), __byte_perm(ae0, ao0, 0x7632),
                         __byte_perm(ae1, ao1, 0x5410), __byte_perm(ae1, ao1, 0x7632));
      sl[1] = make_uint4(__byte_perm(ae2, ao2, 0x5410), __byte_perm(ae2, ao2, 0x7632),
                         __byte_perm(ae3, ao3, 0x5410), __byte_perm(ae3, ao3, 0x7632));
    }
    __syncwarp();
    // pixel pair pr of the segment = u16 sums [6 pr, 6 pr + 6) = three words at byte 12 pr: (B0 G0) (R0 B1) (G1 R1)
    const int npairs = min(SEG_PX, W - seg * SEG_PX) >> 1;
    uint4* brow = reinterpret_cast<uint4*>(bsum + (size_t)jj * W + seg * SEG_PX);
    const uint32_t* sp = reinterpret_cast<const uint32_t*>(scr) + 3 * lane;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int pr = lane + 32 * r;
      if (pr < npairs) {
        const uint32_t w0 = sp[96 * r], w1 = sp[96 * r + 1], w2 = sp[96 * r + 2];
        brow[pr] = make_uint4(w0, w1 & 0xFFFFu, __byte_perm(w1, w2, 0x5432), w2 >> 16);
      }
    }
    __syncwarp();
  }
  __syncthreads();

  const int pitch = p.pitch[lvl];                               // pixels per output row
  uint2* obase = out + p.off[lvl] + ((size_t)blockIdx.y * hs + j0) * pitch;
  const int2* tw = reinterpret_cast<const int2*>(t);
  const bool fast = p.fastdiv[lvl] != 0;
  const uint32_t magic = p.magic_ws[lvl];
  switch (p.kwmin[lvl]) {
    case 1: hpass<1>(bsum, W, rowc, tw, ws, nrows, magic, fast, obase, pitch, lo_off); break;
    case 2: hpass<2>(bsum, W, rowc, tw, ws, nrows, magic, fast, obase, pitch, lo_off); break;
    case 3: hpass<3>(bsum, W, rowc, tw, ws, nrows, magic, fast, obase, pitch, lo_off); break;
    case 4: hpass<4>(bsum, W, rowc, tw, ws, nrows, magic, fast, obase, pitch, lo_off); break;
    case 5: hpass<5>(bsum, W, rowc, tw, ws, nrows, magic, fast, obase, pitch, lo_off); break;
    default: hpass<0>(bsum, W, rowc, tw, ws, nrows, magic, fast, obase, pitch, lo_off); break;
  }
}

}  // namespace pyrp

bool pyramid_pairs_fast_eligible(int W, const void* d_frames) {
  return W >= 16 && W % 16 == 0 && W <= 16384 && (reinterpret_cast<uintptr_t>(d_frames) & 15) == 0;
}

// window tables (c->d_pyr_tab, c->pyr_*) are built by the caller (preproc.cu::build_pyramid_tables)
int launch_pyramid_pairs_fast(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const PyramidGeom& g, uint4* d_hi, uint4* d_lo,
                              cudaStream_t s) {
  using namespace pyrp;
  Params p{};
  int RMAX = c->pyr_rmax > 0 ? c->pyr_rmax : 4;
  while (RMAX > 1 && (size_t)RMAX * W * 8 > 40 * 1024) --RMAX;
  const size_t smem = (size_t)RMAX * W * 8 + 8 * SCR_BYTES + 8 * sizeof(float2);
  if (smem > 200 * 1024) TRL_FAIL(c, TRL_E_INVALID, "frame width %d too large for the pyramid kernel", W);
  p.H = H; p.W = W; p.RMAX = RMAX;
  p.nseg = (W + SEG_PX - 1) / SEG_PX;
  p.magic_nseg = (unsigned)((1ull << 32) / (unsigned)p.nseg) + 1u;
  std::vector<int2> blks;
  for (int k = 0; k < g.n; ++k) {
    p.hs[k] = g.hs[k]; p.ws[k] = g.ws[k];
    p.pitch[k] = 2 * g.pitch2[k];
    p.off[k] = 2 * g.off2[k] * B;
    p.tab_off[k] = c->pyr_tab_off[k];
    p.fastdiv[k] = c->pyr_fastdiv[k];
    const int khmax = (H + g.hs[k] - 1) / g.hs[k] + 1;
    if (khmax > PYR_MAX_KH) TRL_FAIL(c, TRL_E_INVALID, "pyramid window of %d rows exceeds %d (frame %dx%d)", khmax, PYR_MAX_KH, H, W);
    p.kwmin[k] = (c->pyr_kwmin[k] > 0 && 255 * khmax * (c->pyr_kwmin[k] + 1) <= 65535) ? c->pyr_kwmin[k] : 0;
    p.magic_ws[k] = (unsigned)((1ull << 32) / (unsigned)g.ws[k]) + 1u;
    // rows per CTA: about 8-12 source rows of work
    int R = (int)(8.0 * g.hs[k] / H);
    R = std::max(1, std::min(R, RMAX));
    p.rows[k] = R;
    for (int j0 = 0; j0 < g.hs[k]; j0 += R) blks.push_back(make_int2(k, j0));
  }
  const long long key = ((long long)H << 40) ^ ((long long)W << 20) ^ ((long long)blks.size() << 4) ^ (long long)g.n;
  if (c->d_pyrp_blk == nullptr || c->pyrp_blk_key != key) {
    if (c->d_pyrp_blk) { TRL_CUDA(c, cudaStreamSynchronize(s)); TRL_CUDA(c, cudaFree(c->d_pyrp_blk)); c->d_pyrp_blk = nullptr; }
    TRL_CUDA(c, cudaMalloc(&c->d_pyrp_blk, blks.size() * sizeof(int2)));
    TRL_CUDA(c, cudaMemcpy(c->d_pyrp_blk, blks.data(), blks.size() * sizeof(int2), cudaMemcpyHostToDevice));
    c->pyrp_blk_key = key;
  }
  if (!c->pyrp_smem_set) {
    TRL_CUDA(c, cudaFuncSetAttribute(pyramid_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    c->pyrp_smem_set = 1;
  }
  uint2* hi = reinterpret_cast<uint2*>(d_hi);
  const long long lo_off = reinterpret_cast<uint2*>(d_lo) - hi;
  pyramid_pairs_kernel<<<dim3((unsigned)blks.size(), B), 256, smem, s>>>(reinterpret_cast<const uint4*>(d_frames), p, c->d_pyr_tab,
                                                                         reinterpret_cast<const int2*>(c->d_pyrp_blk), hi, lo_off);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}
