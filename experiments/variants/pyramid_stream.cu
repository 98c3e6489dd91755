// K1 for the cascade (pnet_precision 3): the fp16 hi / lo pair images of every pyramid level, for frames whose width is a
// multiple of 16 pixels (every production shape: 360p, 720p, 1080p, ...).  Same window tables, the same exact integer window
// sums and the same two IEEE divisions as pyramid_sep_kernel (preproc.cu) -- bit identical, tests/test_gpu_pnet_hybrid.py --
// organised around what bounded that kernel on B200 (ncu, profiles/r02c_pyr_full.md: ALU pipe 64 % busy at its 2-cycle
// issue rate, L1TEX wavefronts 82 %, 56 % of the shared-memory wavefronts bank conflicts of 2-byte loads, 184 k small CTAs
// per launch whose prologue was 10 % of all instructions):
//
//  * persistent: 2 CTAs per SM walk the (frame, level, R output rows) units of the whole batch, unit u -> CTA u mod grid.
//  * the source rows of a unit stream through a ring of shared-memory stages filled by 1-D bulk copies (cp.async.bulk, one
//    elected producer thread, mbarrier full / empty pairs): no global-load instructions, no L1 tag traffic, and the producer
//    runs ahead over unit boundaries, so HBM / L2 latency never shows.
//  * pass 1 (vertical): a consumer thread owns 4 source pixels (3 words) of one output row and adds, per source row and
//    word, the raw word into one register and its odd bytes (PRMT) into a second one; the even-byte lanes follow once per
//    output row from  sum(w) - (odd << 8)  (mod 2^32; exact, both 16-bit lanes of the result are < 2^16).  Three
//    instructions per word and row, one of them bound to the ALU pipe (was four of four).  The three LDS.32 of a warp are
//    conflict free (stride of 3 banks).  The sums leave the registers as {B, G, R, 0} u16x4 per source column.
//    Fine levels ("resident" units: all source rows of the unit fit the ring) run item by item over the ring; coarse levels
//    (one output row per unit, windows of up to 257 rows) accumulate stage by stage.
//  * pass 2 (horizontal): an output pixel reads one LDS.64 per window column (all three channels at once), adds packed
//    16-bit lanes, converts with the 2^23 trick (no I2F on the quarter-rate pipe), divides / normalises / splits into
//    fp16 hi + lo exactly like PyrOut<1>.  Consecutive threads = consecutive pixels of the flattened [rows][ws] index
//    space of the unit, so a warp's loads span a few hundred bytes.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "pyramid.cuh"

namespace pyrs {

constexpr int NCONS = 256;            // consumer threads (8 warps); warp 8 is the producer
constexpr int MAXNI = 5;              // (output row, 4-pixel group) items per consumer thread and unit
constexpr int NSTAGE_MAX = 16;

struct Params {
  int n_levels;
  int hs[TRL_MAX_SCALES], ws[TRL_MAX_SCALES], pitch[TRL_MAX_SCALES];
  long long off[TRL_MAX_SCALES];           // pixel offset of level k in the hi image (already multiplied by B)
  int tab_off[TRL_MAX_SCALES];
  int fastdiv[TRL_MAX_SCALES];
  int kwmin[TRL_MAX_SCALES];               // > 0: windows of kwmin or kwmin + 1 columns and packed 16-bit lanes cannot overflow
  unsigned magic_ws[TRL_MAX_SCALES];       // floor(2^32 / ws) + 1
  int rows[TRL_MAX_SCALES];                // output rows per unit of level k
  int resident[TRL_MAX_SCALES];            // 1: every unit of level k fits the ring (item-by-item pass 1); 0: one row per unit, stage by stage
  int H, W, B;
  int RMAX;                                // largest rows[k]: rows of the column-sum buffer
  int SR, NS;                              // source rows per stage, stages
  int nblk;                                // units per frame
  unsigned magic_g;                        // floor(2^32 / (W / 4)) + 1
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) break;
    if (clock64() - t0 > 4000000000LL) __trap();        // ~2 s: a lost arrival must not hang the device
  }
}
// the producer thread waits for a free stage most of the time: back off so that its polling does not take issue slots
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) break;
    __nanosleep(200);
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory"); }

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
                                      uint2* __restrict__ obase, int pitch, long long lo_off, int tid) {
  const uint32_t total = (uint32_t)(nrows * ws);
#pragma unroll 2
  for (uint32_t idx = tid; idx < total; idx += NCONS) {
    const uint32_t row = __umulhi(idx, magic);
    const uint32_t col = idx - row * (uint32_t)ws;
    const float2 rc = rowc[row];
    hslot<KMIN>(bsum + row * cw, __ldg(tw + col), rc.x, rc.y, fast, obase + row * pitch + col, lo_off);
  }
}

// byte sums of the three words [B0 G0 R0 B1] [G1 R1 B2 G2] [R2 B3 G3 R3] of four pixels -> four {B, G, R, 0} u16x4 columns.
// sw = sum of the raw words, ao = sums of the odd bytes (2 x u16); the even bytes follow from sw - (ao << 8).
__device__ __forceinline__ void store_sums(uint2* dst2, uint32_t sw0, uint32_t sw1, uint32_t sw2, uint32_t o0, uint32_t o1, uint32_t o2) {
  const uint32_t ae0 = sw0 - (o0 << 8), ae1 = sw1 - (o1 << 8), ae2 = sw2 - (o2 << 8);
  uint4* dst = reinterpret_cast<uint4*>(dst2);
  dst[0] = make_uint4(__byte_perm(ae0, o0, 0x5410), ae0 >> 16,                       // B0 G0 | R0
                      __byte_perm(o0, ae1, 0x5432), o1 & 0xFFFFu);                   // B1 G1 | R1
  dst[1] = make_uint4(__byte_perm(ae1, o1, 0x7632), ae2 & 0xFFFFu,                   // B2 G2 | R2
                      __byte_perm(o2, ae2, 0x7610), o2 >> 16);                       // B3 G3 | R3
}

__global__ void __launch_bounds__(NCONS + 32, 2) pyramid_stream_kernel(const uint8_t* __restrict__ frames,
                                                                     const __grid_constant__ Params p,
                                                                     const int* __restrict__ tab, const int2* __restrict__ blk_tab,
                                                                     uint2* __restrict__ out, long long lo_off) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int W = p.W, SR = p.SR, NS = p.NS;
  const int rowbytes = 3 * W;
  const int stage_bytes = SR * rowbytes;
  const int ring_bytes = NS * stage_bytes;
  unsigned char* ring = smem;                                                       // [NS][SR][3 W] raw source rows
  uint2* bsum = reinterpret_cast<uint2*>(smem + ring_bytes);                         // [RMAX][W] {B, G, R, 0} u16 column sums
  float2* rowc = reinterpret_cast<float2*>(smem + ring_bytes + (size_t)p.RMAX * W * 8);   // {kh, RN(1 / kh)} per row
  uint64_t* full = reinterpret_cast<uint64_t*>(rowc + 8);
  uint64_t* empty = full + NSTAGE_MAX;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCONS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int n_units = p.B * p.nblk;
  int b = 0, blk = blockIdx.x;
  while (blk >= p.nblk) { blk -= p.nblk; ++b; }
  int st = 0;
  uint32_t ph = 0;                       // ring position and phase, advanced identically by the producer and the consumers

  if (warp == NCONS / 32) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int2 bi = __ldg(blk_tab + blk);
        const int lvl = bi.x, j0 = bi.y;
        const int hs = p.hs[lvl];
        const int nrows = min(p.rows[lvl], hs - j0);
        const int* ty0 = tab + p.tab_off[lvl] + 2 * p.ws[lvl] + j0;
        const int ya = __ldg(ty0), yb = __ldg(ty0 + hs + nrows - 1);
        const uint8_t* src = frames + ((size_t)b * p.H + ya) * rowbytes;
        for (int sy = ya; sy < yb; sy += SR, src += stage_bytes) {
          const uint32_t bytes = (uint32_t)(min(SR, yb - sy) * rowbytes);
          mbar_wait_sleep(empty + st, ph ^ 1);
          const uint32_t bar = smem_u32(full + st);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(ring + (size_t)st * stage_bytes)), "l"(src), "r"(bytes), "r"(bar) : "memory");
          if (++st == NS) { st = 0; ph ^= 1; }
        }
        blk += gridDim.x;
        while (blk >= p.nblk) { blk -= p.nblk; ++b; }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int G = W >> 2;
  for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
    const int2 bi = __ldg(blk_tab + blk);
    const int lvl = bi.x, j0 = bi.y;
    const int hs = p.hs[lvl], ws = p.ws[lvl];
    const int nrows = min(p.rows[lvl], hs - j0);
    const int* t = tab + p.tab_off[lvl];
    const int* ty0 = t + 2 * ws + j0;
    const int* ty1 = ty0 + hs;
    const int ya = __ldg(ty0), yb = __ldg(ty1 + nrows - 1);
    if (tid < nrows) {
      const float fkh = (float)(__ldg(ty1 + tid) - __ldg(ty0 + tid));
      rowc[tid] = make_float2(fkh, __frcp_rn(fkh));
    }

    if (p.resident[lvl]) {
      // every source row of the unit is (about to be) in the ring: wait for its stages, then one item after the other
      const int nst = (yb - ya + SR - 1) / SR;
      {
        int s2 = st; uint32_t ph2 = ph;
        for (int k = 0; k < nst; ++k) { mbar_wait(full + s2, ph2); if (++s2 == NS) { s2 = 0; ph2 ^= 1; } }
      }
      const unsigned char* row0 = ring + (size_t)st * stage_bytes;
      const unsigned char* ring_end = ring + ring_bytes;
      const int items = nrows * G;
#pragma unroll 1
      for (int item = tid; item < items; item += NCONS) {
        const int jj = (int)__umulhi((uint32_t)item, p.magic_g);
        const int g = item - jj * G;
        const int y0 = __ldg(ty0 + jj), y1 = __ldg(ty1 + jj);
        const unsigned char* q = row0 + (size_t)(y0 - ya) * rowbytes + 12 * g;
        if (q >= ring_end) q -= ring_bytes;
        uint32_t sw0 = 0, sw1 = 0, sw2 = 0, ao0 = 0, ao1 = 0, ao2 = 0;
        for (int y = y0; y < y1; ++y) {
          const uint32_t* qw = reinterpret_cast<const uint32_t*>(q);
          const uint32_t w0 = qw[0], w1 = qw[1], w2 = qw[2];
          sw0 += w0; ao0 += __byte_perm(w0, 0, 0x4341);
          sw1 += w1; ao1 += __byte_perm(w1, 0, 0x4341);
          sw2 += w2; ao2 += __byte_perm(w2, 0, 0x4341);
          q += rowbytes;
          if (q >= ring_end) q -= ring_bytes;
        }
        store_sums(bsum + (size_t)jj * W + 4 * g, sw0, sw1, sw2, ao0, ao1, ao2);
      }
      __syncwarp();
      for (int k = 0; k < nst; ++k) {
        if (lane == 0) mbar_arrive(empty + st);
        if (++st == NS) { st = 0; ph ^= 1; }
      }
    } else {
      // one output row per unit, its window streams through the ring: thread owns the 4-pixel groups tid, tid + 256, ...
      uint32_t sw[MAXNI][3], ao[MAXNI][3];
#pragma unroll
      for (int i = 0; i < MAXNI; ++i) sw[i][0] = sw[i][1] = sw[i][2] = ao[i][0] = ao[i][1] = ao[i][2] = 0;
      for (int sy = ya; sy < yb; sy += SR) {
        mbar_wait(full + st, ph);
        const int nr = min(SR, yb - sy);
        const unsigned char* sbase = ring + (size_t)st * stage_bytes + 12 * tid;
#pragma unroll
        for (int i = 0; i < MAXNI; ++i) {
          if (tid + NCONS * i < G) {
            const unsigned char* q = sbase + 12 * NCONS * i;
            for (int r = 0; r < nr; ++r, q += rowbytes) {
              const uint32_t* qw = reinterpret_cast<const uint32_t*>(q);
              const uint32_t w0 = qw[0], w1 = qw[1], w2 = qw[2];
              sw[i][0] += w0; ao[i][0] += __byte_perm(w0, 0, 0x4341);
              sw[i][1] += w1; ao[i][1] += __byte_perm(w1, 0, 0x4341);
              sw[i][2] += w2; ao[i][2] += __byte_perm(w2, 0, 0x4341);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + st);
        if (++st == NS) { st = 0; ph ^= 1; }
      }
#pragma unroll
      for (int i = 0; i < MAXNI; ++i)
        if (tid + NCONS * i < G) store_sums(bsum + 4 * (tid + NCONS * i), sw[i][0], sw[i][1], sw[i][2], ao[i][0], ao[i][1], ao[i][2]);
    }
    cons_sync();

    const int pitch = p.pitch[lvl];
    uint2* obase = out + p.off[lvl] + ((size_t)b * hs + j0) * pitch;
    const int2* tw = reinterpret_cast<const int2*>(t);
    const bool fast = p.fastdiv[lvl] != 0;
    const uint32_t magic = p.magic_ws[lvl];
    switch (p.kwmin[lvl]) {
      case 1: hpass<1>(bsum, W, rowc, tw, ws, nrows, magic, fast, obase, pitch, lo_off, tid); break;
      case 2: hpass<2>(bsum, W, rowc, tw, ws, nrows, magic, fast, obase, pitch, lo_off, tid); break;
      case 3: hpass<3>(bsum, W, rowc, tw, ws, nrows, magic, fast, obase, pitch, lo_off, tid); break;
      case 4: hpass<4>(bsum, W, rowc, tw, ws, nrows, magic, fast, obase, pitch, lo_off, tid); break;
      case 5: hpass<5>(bsum, W, rowc, tw, ws, nrows, magic, fast, obase, pitch, lo_off, tid); break;
      default: hpass<0>(bsum, W, rowc, tw, ws, nrows, magic, fast, obase, pitch, lo_off, tid); break;
    }
    cons_sync();                          // bsum / rowc are rewritten by the next unit

    blk += gridDim.x;
    while (blk >= p.nblk) { blk -= p.nblk; ++b; }
  }
}

}  // namespace pyrs

bool pyramid_stream_eligible(int W, const void* d_frames) {
  return W >= 16 && W % 16 == 0 && W <= 8192 && (reinterpret_cast<uintptr_t>(d_frames) & 15) == 0;
}

// source rows a unit of R output rows starting at row j0 of a level with hs rows reads (adaptive-pool windows, preproc.cu::window_table)
static int unit_src_rows(int H, int hs, int j0, int R) {
  const int j1 = std::min(j0 + R, hs) - 1;
  const int ya = (int)(((long long)j0 * H) / hs);
  const int yb = (int)(((long long)(j1 + 1) * H + hs - 1) / hs);
  return yb - ya;
}

// window tables (c->d_pyr_tab, c->pyr_*) are built by the caller (preproc.cu::build_pyramid_tables)
int launch_pyramid_stream(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const PyramidGeom& g, uint4* d_hi, uint4* d_lo,
                          cudaStream_t s) {
  using namespace pyrs;
  Params p{};
  const int G = W / 4;
  if ((G + NCONS - 1) / NCONS > MAXNI) TRL_FAIL(c, TRL_E_INVALID, "frame width %d too large for the streaming pyramid kernel", W);
  // shared memory of one CTA (two per SM): column sums of up to RMAX output rows (<= 40 KB) + the ring of source rows
  int RMAX = 4;
  while (RMAX > 1 && (size_t)RMAX * W * 8 > 40 * 1024) RMAX >>= 1;
  const size_t fixed = (size_t)RMAX * W * 8 + 8 * sizeof(float2) + 2 * NSTAGE_MAX * sizeof(uint64_t);
  const size_t budget = 111 * 1024;
  if (fixed + 4 * (size_t)3 * W > budget) TRL_FAIL(c, TRL_E_INVALID, "frame width %d too large for the streaming pyramid kernel", W);
  int ring_rows = (int)((budget - fixed) / ((size_t)3 * W));
  int SR = ring_rows >= 16 ? 2 : 1;
  int NS = std::min(NSTAGE_MAX, ring_rows / SR);
  if (ring_rows / SR > NSTAGE_MAX) { SR = (ring_rows + NSTAGE_MAX - 1) / NSTAGE_MAX; NS = std::min(NSTAGE_MAX, ring_rows / SR); }
  const size_t smem = (size_t)NS * SR * 3 * W + fixed;

  p.n_levels = g.n; p.H = H; p.W = W; p.B = B; p.RMAX = RMAX; p.SR = SR; p.NS = NS;
  p.magic_g = (unsigned)((1ull << 32) / (unsigned)G) + 1u;
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
    // the largest R whose units leave at least two stages of the ring to the producer's prefetch; none: stream one row at a time
    int R = RMAX, resident = 0;
    for (; R >= 1; R >>= 1) {
      int worst = 0;
      for (int j0 = 0; j0 < g.hs[k]; j0 += R) worst = std::max(worst, unit_src_rows(H, g.hs[k], j0, R));
      if ((worst + SR - 1) / SR <= NS - 2) { resident = 1; break; }
    }
    if (!resident) R = 1;
    p.rows[k] = R;
    p.resident[k] = resident;
    for (int j0 = 0; j0 < g.hs[k]; j0 += R) blks.push_back(make_int2(k, j0));
  }
  p.nblk = (int)blks.size();
  const long long key = ((long long)H << 40) ^ ((long long)W << 20) ^ ((long long)p.nblk << 4) ^ (long long)g.n;
  if (c->d_pyrs_blk == nullptr || c->pyrs_blk_key != key) {
    if (c->d_pyrs_blk) { TRL_CUDA(c, cudaStreamSynchronize(s)); TRL_CUDA(c, cudaFree(c->d_pyrs_blk)); c->d_pyrs_blk = nullptr; }
    TRL_CUDA(c, cudaMalloc(&c->d_pyrs_blk, blks.size() * sizeof(int2)));
    TRL_CUDA(c, cudaMemcpy(c->d_pyrs_blk, blks.data(), blks.size() * sizeof(int2), cudaMemcpyHostToDevice));
    c->pyrs_blk_key = key;
  }
  if (!c->pyrs_smem_set) {
    TRL_CUDA(c, cudaFuncSetAttribute(pyramid_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    c->pyrs_smem_set = 1;
  }
  const long long units = (long long)B * p.nblk;
  const int grid = (int)std::min<long long>(units, 2LL * c->num_sms);
  uint2* hi = reinterpret_cast<uint2*>(d_hi);
  const long long lo_off = reinterpret_cast<uint2*>(d_lo) - hi;
  pyramid_stream_kernel<<<grid, NCONS + 32, smem, s>>>(d_frames, p, c->d_pyr_tab, reinterpret_cast<const int2*>(c->d_pyrs_blk), hi, lo_off);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}
