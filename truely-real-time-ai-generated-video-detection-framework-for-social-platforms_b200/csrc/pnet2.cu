// K2 (screen): P-Net entirely on the 5th-generation tensor cores -- conv1, conv2 and conv3 as tcgen05.mma implicit GEMMs in one
// persistent kernel, single fp16 pass, fp32 accumulation in TMEM.  It does not emit candidates: it marks the cells whose
// approximate face probability is >= thr - TRL_SCREEN_MARGIN; pnet_refine.cu re-evaluates exactly those cells in fp32 and
// applies the real threshold (the "hybrid" P-Net, trl_config_t.pnet_precision = 3).
//
// upstream: models/mtcnn.py PNet.forward.
//
// Why it is shaped like this (experiments/umma_desc_probe.cu, profiles/PROFILE_NOTES.md r02h): with 10 / 16 / 32 output
// channels every MMA is tiny (N <= 64) and costs ~40-48 cycles whatever N is, because the tensor pipe has to fetch the
// 128 x 16 A tile from shared memory (4 KB at ~100 B/clk).  So the cost of a layer is its number of MMA instructions, and
// the layout below exists to minimise that number:
//   * an M tile is a 2-D patch, not a run of flat pixels: 8 columns x 16 row PAIRS.  The no-swizzle K-major descriptor
//     addresses row m at start + (m / 8) SBO + (m % 8) 16 B for any SBO, so the sixteen 8-row groups of a tile are 16
//     image rows two apart (SBO = 2 row pitches).
//   * two output rows per MMA: output rows 2j and 2j + 1 share the input rows 2j + 1, 2j + 2, so the B operand of input-row
//     tap t = 0..3 is [weights of dy = t for the even row | weights of dy = t - 1 for the odd row] (N doubles, 4 row taps
//     instead of 2 x 3; the half-empty taps 0 and 3 run at half N into one half of the accumulator).
//   * conv1 (3 channels): the pyramid kernel writes the level as pixel PAIRS of (B, G, R, 0) halves (16 bytes).  With
//     LBO = 16 B the second K chunk of operand row c is the next pair, so one K = 16 operand row is pixels 2c .. 2c + 3 with
//     no im2col copy, and the B operand holds the filter twice -- shifted by 0 and by 1 pixel -- so that one MMA yields
//     the conv outputs x = 2c and x = 2c + 1: with the row trick one N = 64 accumulator row is exactly the 2 x 2 pooling
//     window of pooled pixel (R, c) x 16 channels, and max-pooling is thread local.
//   * conv2 (10 channels): K = 3 kx x 10 ch = 30 per row tap fits two MMAs: [kx0 ch0-7 | kx1 ch0-7] (LBO = 16 B again)
//     and [kx2 ch0-7 | ch8-9 of kx0, kx1, kx2] -- the conv1 epilogue writes channels 8, 9 of every pixel into the three
//     pixel slots that see it (a 16-byte side plane).
//   MMAs per 28 x 28-cell tile: conv1 8 tiles x 4, conv2 4 x 8, conv3 4 x 12 = 112 (~4.9 k tensor-pipe cycles) against
//   ~1500 HMMA + 90 UTCHMMA (~9.4 k cycles) per 20 x 28 cells in the 3-term kernel (pnet.cu).
//
// Roles (640 threads, one CTA per SM): warp 0 TMA producer (one 3-D box per tile, zero fill outside the level, double
// buffered), warp 1 MMA issue (one elected thread, uniform registers), warps 4-7 conv1 epilogue (TMEM -> pool, bias, PReLU
// -> fp16 operand planes), warps 8-11 conv2 epilogue, warps 12-19 conv3 epilogue + conv4_1 logit difference -> screen list.
// Issue order per iteration: conv2(k-1), conv1(k), conv3(k-1): each epilogue runs under the next layer's MMAs, and because the
// tensor pipe executes in order the conv1 / conv2 operand planes need no double buffering (an epilogue can only start on
// a commit that was issued after the last reader of the plane it overwrites).
// TMEM (512 columns): conv1 ring 2 x 64, conv2 4 x 32, conv3 4 x 64.
#include <cuda.h>
#include <algorithm>

#include "common.cuh"
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

namespace pnet2 {

constexpr int T = 28;                       // output cells per tile side
constexpr int IN_ROWS = 2 * T + 10;         // 66 input rows
constexpr int IN_PAIRS = T + 5;             // 33 pixel pairs per row
constexpr int IN_ROWB = IN_PAIRS * 16;      // 528 bytes
constexpr int IN_BYTES = IN_ROWS * IN_ROWB; // bytes one TMA box delivers
constexpr int IN_STRIDE = ((IN_BYTES + 127) / 128) * 128;
constexpr int PL_ROWS = 34;                 // rows of the p1 / c2 operand planes (2 * 15 + 3 is the last row a row tap reads)
constexpr int PL_PITCH = 32;                // pixels per row (16 bytes each)
constexpr int PL_BYTES = PL_ROWS * PL_PITCH * 16 + 64;   // + the two pixels the kx taps of the last pixel read

// weights: fp16 K-major core matrices [k-chunk][row][8 halves]; row blocks in the order dy = 2, 1, 0 (see row taps below)
constexpr int W1_ROWS = 96;                 // 3 dy x [even-x 16 | odd-x 16]
constexpr int W2_ROWS = 48;                 // 3 dy x 16, two matrices (a, b)
constexpr int W3_ROWS = 96;                 // 3 dy x 32, three matrices (kx)
constexpr int W1_OFF = 0;
constexpr int W2_OFF = W1_OFF + 2 * W1_ROWS * 16;
constexpr int W3_OFF = W2_OFF + 2 * 2 * W2_ROWS * 16;
constexpr int W_BYTES = W3_OFF + 3 * 2 * W3_ROWS * 16;   // 15360
constexpr int IN_OFF = ((W_BYTES + 127) / 128) * 128;
constexpr int P1A_OFF = IN_OFF + 2 * IN_STRIDE;
constexpr int P1B_OFF = P1A_OFF + PL_BYTES;
constexpr int P1_STRIDE = 2 * PL_BYTES;              // the p1 planes are double buffered (tile k uses buffer k & 1)
constexpr int C0_OFF = P1A_OFF + 2 * P1_STRIDE;
constexpr int C1_OFF = C0_OFF + PL_BYTES;
constexpr int TILES_OFF = C1_OFF + PL_BYTES;         // tile table: 16 bytes per tile of a frame
constexpr int SMEM_BYTES = TILES_OFF;                // + 16 * tiles per frame at launch
constexpr int SMEM_CAP = 227 * 1024 - 1024;          // dynamic shared memory ceiling (the barriers are static shared memory)
constexpr int MAX_TILES_PER_FRAME = (SMEM_CAP - SMEM_BYTES) / 16;
static_assert(SMEM_BYTES <= 200 * 1024, "shared memory budget");

constexpr int MAX_TILES_PARAM = 4096;   // tiles per frame the group table in the kernel parameters holds (8 KB)
constexpr int TMEM_COLS = 512;
constexpr int ACC1_COL = 0, ACC2_COL = 128, ACC3_COL = 256;
#ifndef PNET2_E1_SETS
#define PNET2_E1_SETS 2
#endif
constexpr int E1_SETS = PNET2_E1_SETS;       // conv1 epilogue sets of four warps (M tiles are dealt round robin)
#ifndef PNET2_E3_SETS
#define PNET2_E3_SETS 1
#endif
constexpr int E3_SETS = PNET2_E3_SETS;       // conv3 epilogue sets of four warps (set s takes the column groups q = s, s + E3_SETS, ..)
#ifndef PNET2_BACKOFF
#define PNET2_BACKOFF 0                     // ns the epilogue warps sleep between barrier polls (they share issue slots with the MMA threads)
#endif
constexpr int W_TMA = 0, W_MMA = 1, W_MMA3 = 2, W_E1 = 4, W_E2 = W_E1 + 4 * E1_SETS, W_E3 = W_E2 + 4, N_WARPS = W_E3 + 4 * E3_SETS;
constexpr int NTHREADS = 32 * N_WARPS;
constexpr float ACT_MAX = 60000.f;          // fp16 operand range guard

struct Level {
  int hs, ws, oh, ow;
  int tiles_x, tiles;
  float* logit;        // optional (stage entry point): conv4_1 logit difference per cell, [B][oh][ow]
};

struct Params {
  CUtensorMap tmap[TRL_MAX_SCALES];   // per level: u32 view {2 ws, hs, B} of the hi pair image, box {132, 66, 1}
  int n_levels;
  int blocks;          // tiles per frame (all levels)
  int n_frames;
  const int4* tiles;   // [blocks] tile table of one frame
  uint16_t groups[MAX_TILES_PARAM];   // per tile of a frame: live groups nq1 | nrb << 4 | nq2 << 8 | nq3 << 12 -- in the parameter
                       // (constant) bank so that the MMA-issue thread's control flow and operands stay in uniform registers
  Level lv[TRL_MAX_SCALES];
  float logit_lo;      // screen: logit(thr - margin)
  ScreenEntry* screen;
  int* screen_cnt;
  int screen_cap;
  CapFlag* capflag;
  // Epilogue constants.  PReLU is evaluated as c1 v + c2 |v| with c1 = (1 + a) / 2, c2 = (1 - a) / 2: two packed instructions
  // per channel pair instead of compare / multiply / select per channel (this kernel only screens: the extra rounding is
  // far inside the margin).  conv1 / conv2: packed fp16 pairs (the results are stored as fp16 anyway); conv3: fp32 pairs,
  // already multiplied by the conv4_1 logit-difference weight dw, so the logit is sum k1 u + k2 |u| with u = acc + b3.
  float b1[10], a1[10];
  float b2[16], a2[16];
  float b3[32], k1[32], k2[32];
  float db;
  int conv1_monotone;
};

#ifdef PNET_TIMING
// cycles summed over CTAs: [0] MMA thread total, [1..6] its waits (in_full, p1_ready, acc2_empty, acc1_empty, c2_ready, acc3_empty),
// [7] tiles, [8] E1 warp 0 wait, [9] E1 busy, [10] E2 wait, [11] E2 busy, [12] E3 wait, [13] E3 busy, [14] TMA wait
__device__ unsigned long long g_pnet2_phase[24];   // [16] conv2 issue, [17] conv1 issue, [18] conv3 issue (MMA thread, waits excluded)
#define P2T(...) __VA_ARGS__
#else
#define P2T(...)
#endif
__device__ __forceinline__ float prelu(float v, float a) { return v > 0.f ? v : v * a; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred px;\nelect.sync _|px, %1;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
#ifdef PNET_TIMING
__device__ long long g_dummy_sink;
#define umma_commit_t(bar, acc) do { const long long _t = clock64(); umma_commit(bar); acc += clock64() - _t; } while (0)
#else
#define umma_commit_t(bar, acc) umma_commit(bar)
#endif
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  if (ok) return;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) break;
    if (PNET2_BACKOFF > 0 && (threadIdx.x >> 5) >= W_E1) __nanosleep(PNET2_BACKOFF);
    if (clock64() - t0 > 4000000000LL) __trap();        // ~2 s: a lost arrival must not hang the device
  }
}
#ifdef PNET_TIMING
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, long long& acc) {
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}
#else
#define mbar_wait_t(bar, parity, acc) mbar_wait(bar, parity)
#endif
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// ten consecutive columns (the live conv1 channels of one 16-column block)
__device__ __forceinline__ void tmem_ld10(uint32_t taddr, float (&v)[10]) {
  uint32_t r[10];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[8]), "=r"(r[9]) : "r"(taddr + 8u) : "memory");
#pragma unroll
  for (int i = 0; i < 10; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ __half2 as_h2(uint32_t u) { return *reinterpret_cast<const __half2*>(&u); }
__device__ __forceinline__ uint32_t as_u32(__half2 h) { return *reinterpret_cast<const uint32_t*>(&h); }
// PReLU of a packed fp16 pair: c1 v + c2 |v|
__device__ __forceinline__ __half2 prelu_h2(__half2 v, uint32_t c1, uint32_t c2) {
  return __hfma2(v, as_h2(c1), __hmul2(__habs2(v), as_h2(c2)));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Tiles: the per-frame tile table (level, origin, live extent of every tile of every level; built on the host, copied to shared
// memory once per CTA) + a cursor (frame, tile-in-frame) that advances by the grid size -- no division or level search per tile
// (the MMA-issue thread spent 600 cycles per decode on them).
struct TileRef { int b, lvl, oy0, ox0, rows, cols; };
struct TileCursor { int b, blk; };
__device__ __forceinline__ TileCursor cursor_init(int blocks) {
  TileCursor c;
  c.b = (int)blockIdx.x / blocks;
  c.blk = (int)blockIdx.x - c.b * blocks;
  return c;
}
__device__ __forceinline__ void cursor_next(TileCursor& c, int blocks) {
  c.blk += (int)gridDim.x;
  while (c.blk >= blocks) { c.blk -= blocks; ++c.b; }
}
__device__ __forceinline__ TileRef tile_at(const int4* tile_s, const TileCursor& c) {
  const int4 rec = tile_s[c.blk];            // x = level | rows << 8 | cols << 16, y = oy0, z = ox0
  TileRef r;
  r.b = c.b; r.lvl = rec.x & 0xFF; r.rows = (rec.x >> 8) & 0xFF; r.cols = (rec.x >> 16) & 0xFF;
  r.oy0 = rec.y; r.ox0 = rec.z;
  return r;
}
// live 8-column groups of the three layers and conv1 row blocks (pooled rows in blocks of 16) of a tile
__device__ __forceinline__ int ncg3(const TileRef& t) { return (t.cols + 7) >> 3; }
__device__ __forceinline__ int ncg2(const TileRef& t) { return (t.cols + 2 + 7) >> 3; }
__device__ __forceinline__ int ncg1(const TileRef& t) { return (t.cols + 4 + 7) >> 3; }
__device__ __forceinline__ int nrb1(const TileRef& t) { return t.rows + 4 > 16 ? 2 : 1; }

__global__ void __launch_bounds__(NTHREADS, 1) pnet2_kernel(const uint32_t* __restrict__ wpacked, const __grid_constant__ Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t in_full[2], in_empty[2];
  // conv1 accumulators: two TMEM slots, but EIGHT barrier pairs used round robin -- with more epilogue sets than slots a set's
  // consecutive uses of one barrier would be two phases apart, which a parity wait cannot tell from zero phases
  __shared__ __align__(8) uint64_t acc1_full[8], acc1_empty[8];
  __shared__ __align__(8) uint64_t p1_ready[2], acc2_full[4], acc2_empty, c2_ready, acc3_full[4], acc3_empty;   // p1_ready: one per p1 buffer
  // (conv1 and its epilogue may finish tile k + 1 before the conv2 issue waits for tile k: one barrier would be two phases ahead)
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  P2T(long long tcm = 0; long long tdec = 0; long long ti2 = 0; long long ti1 = 0; long long ti3 = 0; long long tq = 0; long long tw0 = 0; long long tw1 = 0; long long tw2 = 0; long long tw3 = 0; long long tw4 = 0; long long tw5 = 0; long long tstart = clock64();)

  // zero the operand planes once (slack rows / pad halves are read by dead accumulator rows and must stay finite), load weights
  for (int i = tid; i < (TILES_OFF - IN_OFF) / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem + IN_OFF)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < W_BYTES / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(wpacked) + i);
  const int4* tile_s = reinterpret_cast<const int4*>(smem + TILES_OFF);
  for (int i = tid; i < p.blocks; i += NTHREADS) reinterpret_cast<int4*>(smem + TILES_OFF)[i] = __ldg(p.tiles + i);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); }
      for (int i = 0; i < 8; ++i) { mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], 128); }
      for (int i = 0; i < 4; ++i) { mbar_init(&acc2_full[i], 1); mbar_init(&acc3_full[i], 1); }
      mbar_init(&p1_ready[0], 128 * E1_SETS); mbar_init(&p1_ready[1], 128 * E1_SETS); mbar_init(&acc2_empty, 128); mbar_init(&c2_ready, 128); mbar_init(&acc3_empty, 128 * E3_SETS);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // weights / zeroed planes -> visible to the UMMA and TMA proxies
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int total = p.blocks * p.n_frames;
  const int n_my = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // tiles of this CTA (>= 1: grid <= total)

  if (warp == W_TMA) {
    // =================================================================== TMA producer
    if (elect_one()) {
      TileCursor cur = cursor_init(p.blocks);
      for (int i = 0; i < n_my; ++i, cursor_next(cur, p.blocks)) {
        const TileRef tr = tile_at(tile_s, cur);
        const int buf = i & 1;
        if (i >= 2) mbar_wait_t(&in_empty[buf], ((i >> 1) - 1) & 1, tw0);
        const uint32_t bar = smem_u32(&in_full[buf]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)IN_BYTES) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(smem + IN_OFF + buf * IN_STRIDE)), "l"(reinterpret_cast<uint64_t>(&p.tmap[tr.lvl])), "r"(bar),
                       "r"(4 * tr.ox0), "r"(2 * tr.oy0), "r"(tr.b) : "memory");
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    // =================================================================== MMA issue (one elected thread)
    // This thread is the kernel's critical path: ~105 MMAs, 17 commits and a dozen barrier waits per tile from ONE thread.  The
    // first version spent 1300 instructions per tile here (descriptors rebuilt in vector registers and moved to uniform ones
    // with R2UR for every group; 600-cycle tile decodes) and took 9.5 k cycles per tile for 4.5 k cycles of tensor work.  Now
    // everything an MMA needs is a compile-time offset from a descriptor template that the compiler can keep in the uniform
    // datapath: the shared-memory window base is uniform by construction, the TMEM base is 0 because the CTA owns all 512
    // columns (checked once), and the group loops are fully unrolled.
    const uint32_t sbase = smem_u32(smem);
    constexpr uint32_t tm = 0u;
    if (tmem != 0u) __trap();
    constexpr uint32_t ID16 = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);      // f32 accumulate, f16 x f16, K-major, M = 128
    constexpr uint32_t ID32 = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t ID64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    if (elect_one()) {
      // Descriptor templates; a descriptor's low 14 bits are (address >> 4), so moving an operand by `bytes` is a 64-bit add of
      // bytes / 16 (every address here is < 256 KB: no carry into the LBO field).
      const uint64_t a1_t = umma_desc(sbase + IN_OFF, 16u, 2u * IN_ROWB);                       // conv1 A: [pair c | pair c + 1], rows two apart
      const uint64_t a2a_t = umma_desc(sbase + P1A_OFF, 16u, 2u * PL_PITCH * 16u);              // conv2 A (a): [px n | px n + 1], channels 0-7
      const uint64_t a2b_t = umma_desc(sbase + P1A_OFF + 32u, (uint32_t)(P1B_OFF - P1A_OFF) - 32u, 2u * PL_PITCH * 16u);   // (b): [px n + 2 | side plane of px n]
      const uint64_t a3_t = umma_desc(sbase + C0_OFF, (uint32_t)(C1_OFF - C0_OFF), 2u * PL_PITCH * 16u);   // conv3 A: [ch 0-7 | ch 8-15]
      const uint64_t b1_t = umma_desc(sbase + W1_OFF, W1_ROWS * 16u, 128u);
      const uint64_t b2a_t = umma_desc(sbase + W2_OFF, W2_ROWS * 16u, 128u);
      const uint64_t b2b_t = umma_desc(sbase + W2_OFF + 2u * W2_ROWS * 16u, W2_ROWS * 16u, 128u);
      const uint64_t b3_t = umma_desc(sbase + W3_OFF, W3_ROWS * 16u, 128u);
      uint32_t u1 = 0;                                   // conv1 accumulator ring use count
      TileCursor cur = cursor_init(p.blocks);
      int g_cur = 0, g_prev = 0;                         // live groups of tile i / i - 1: nq1 | nrb << 4 | nq2 << 8 | nq3 << 12
#pragma unroll 1
      for (int i = 0; i <= n_my; ++i) {
        g_prev = g_cur;
        if (i < n_my) {
          g_cur = p.groups[cur.blk];
          cursor_next(cur, p.blocks);
        }
        const bool do1 = i < n_my;
        // ---- conv1 of tile i: M tile = 8 pooled columns x 16 pooled rows; input-row taps t = 1, 0, 2, 3 (tap 1 first: it writes
        // all 64 columns).  It advances only as fast as its epilogue sets free the two accumulator slots.  Its epilogue writes
        // the p1 planes of buffer i & 1; the last MMAs that read that buffer (conv2 of tile i - 2) were issued by this thread
        // before any conv1 MMA of tile i, so the epilogue cannot start before they have completed.
        const int buf = i & 1;
        const int nq1 = do1 ? (g_cur & 15) : 0, nrb = (g_cur >> 4) & 15;
        const uint64_t a1_b = a1_t + (uint64_t)(buf * (IN_STRIDE / 16));
        if (do1) mbar_wait_t(&in_full[buf], (i >> 1) & 1, tw0);
        auto conv1_tile = [&](int rb, int q) {             // called with literals: the operand offsets fold to constants
          if (rb < nrb && q < nq1) {
            const uint32_t slot = u1 & 1u;
            if (u1 >= 2) mbar_wait_t(&acc1_empty[(u1 - 2) & 7u], ((u1 - 2) >> 3) & 1, tw3);   // the slot's previous user has been read out
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            P2T(tq = clock64();)
            const uint32_t d = tm + ACC1_COL + 64u * slot;
            const uint64_t a = a1_b + (uint64_t)(32 * rb * (IN_ROWB / 16) + 8 * q);
            umma_f16(d, a + 1 * (IN_ROWB / 16), b1_t + 32, ID64, 0u);
            umma_f16(d, a, b1_t + 64, ID32, 1u);
            umma_f16(d, a + 2 * (IN_ROWB / 16), b1_t, ID64, 1u);
            umma_f16(d + 32u, a + 3 * (IN_ROWB / 16), b1_t, ID32, 1u);
            umma_commit(smem_u32(&acc1_full[u1 & 7u]));
            ++u1;
            P2T(ti1 += clock64() - tq;)
          }
        };
        conv1_tile(0, 0); conv1_tile(0, 1); conv1_tile(0, 2); conv1_tile(0, 3);
        conv1_tile(1, 0); conv1_tile(1, 1); conv1_tile(1, 2); conv1_tile(1, 3);
        if (do1) umma_commit(smem_u32(&in_empty[buf]));     // every MMA that reads this input buffer has completed
        // ---- conv2 of tile i - 1 (its conv1 epilogue had the whole conv1 phase of tile i to finish): per 8-column group one M
        // tile of 16 row pairs; row taps t = 1, 0, 2, 3, two MMAs each.  B rows are blocks dy2 | dy1 | dy0; tap t multiplies
        // [dy = t | dy = t - 1] (taps 0 and 3: one block, half N, one half of D)
        if (i >= 1) {
          const int k = i - 1;
          mbar_wait_t(&p1_ready[k & 1], (k >> 1) & 1, tw1);
          if (k >= 1) mbar_wait_t(&acc2_empty, (k - 1) & 1, tw2);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int nq = (g_prev >> 8) & 15;
          const uint64_t pb = (uint64_t)((k & 1) * (P1_STRIDE / 16));
          P2T(tq = clock64();)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q < nq) {
              const uint32_t d = tm + ACC2_COL + 32u * q;
              const uint64_t aq = pb + (uint64_t)(8 * q);                             // 8 pixels = 128 bytes
              // tap 1 first: it writes all 32 columns (accumulate = 0)
              umma_f16(d, a2a_t + aq + 1 * PL_PITCH, b2a_t + 16, ID32, 0u);
              umma_f16(d, a2b_t + aq + 1 * PL_PITCH, b2b_t + 16, ID32, 1u);
              umma_f16(d, a2a_t + aq, b2a_t + 32, ID16, 1u);
              umma_f16(d, a2b_t + aq, b2b_t + 32, ID16, 1u);
              umma_f16(d, a2a_t + aq + 2 * PL_PITCH, b2a_t, ID32, 1u);
              umma_f16(d, a2b_t + aq + 2 * PL_PITCH, b2b_t, ID32, 1u);
              umma_f16(d + 16u, a2a_t + aq + 3 * PL_PITCH, b2a_t, ID16, 1u);
              umma_f16(d + 16u, a2b_t + aq + 3 * PL_PITCH, b2b_t, ID16, 1u);
            }
            umma_commit(smem_u32(&acc2_full[q]));          // dead groups too: keeps the barrier phases aligned with k
          }
          P2T(ti2 += clock64() - tq;)
        }
      }
      P2T(atomicAdd(&g_pnet2_phase[0], (unsigned long long)(clock64() - tstart)); atomicAdd(&g_pnet2_phase[1], (unsigned long long)tw0);
          atomicAdd(&g_pnet2_phase[2], (unsigned long long)tw1); atomicAdd(&g_pnet2_phase[3], (unsigned long long)tw2);
          atomicAdd(&g_pnet2_phase[4], (unsigned long long)tw3); atomicAdd(&g_pnet2_phase[5], (unsigned long long)tw4);
          atomicAdd(&g_pnet2_phase[6], (unsigned long long)tw5); atomicAdd(&g_pnet2_phase[7], (unsigned long long)n_my);
          atomicAdd(&g_pnet2_phase[16], (unsigned long long)ti2); atomicAdd(&g_pnet2_phase[17], (unsigned long long)ti1); atomicAdd(&g_pnet2_phase[18], (unsigned long long)ti3);
          atomicAdd(&g_pnet2_phase[19], (unsigned long long)tcm); atomicAdd(&g_pnet2_phase[20], (unsigned long long)tdec);)
    }
    __syncwarp();
  } else if (warp == W_MMA3) {
    // =================================================================== MMA issue, conv3 (a second elected thread)
    // conv3 has its own issue thread: one thread cannot issue the ~105 MMAs of a tile as fast as the tensor pipe executes them.
    // Ordering against the other issuer is by barriers only: c2_ready (conv2 epilogue done) before, acc3_full after -- the
    // conv2 epilogue of the NEXT tile waits for the last acc3_full of this one before it overwrites the c2 planes.
    const uint32_t sbase = smem_u32(smem);
    constexpr uint32_t tm = 0u;
    constexpr uint32_t ID32 = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t ID64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    if (elect_one()) {
      const uint64_t a3_t = umma_desc(sbase + C0_OFF, (uint32_t)(C1_OFF - C0_OFF), 2u * PL_PITCH * 16u);   // conv3 A: [ch 0-7 | ch 8-15]
      const uint64_t b3_t = umma_desc(sbase + W3_OFF, W3_ROWS * 16u, 128u);
      TileCursor cur = cursor_init(p.blocks);
#pragma unroll 1
      for (int k = 0; k < n_my; ++k) {
        const int n3 = (p.groups[cur.blk] >> 12) & 15;
        cursor_next(cur, p.blocks);
        mbar_wait_t(&c2_ready, k & 1, tw4);
        if (k >= 1) mbar_wait_t(&acc3_empty, (k - 1) & 1, tw5);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        P2T(tq = clock64();)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < n3) {
            const uint32_t d = tm + ACC3_COL + 64u * q;
            const uint64_t aq = a3_t + (uint64_t)(8 * q);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint64_t bk = b3_t + (uint64_t)(kx * 2 * W3_ROWS);
              umma_f16(d, aq + (1 * PL_PITCH + kx), bk + 32, ID64, kx == 0 ? 0u : 1u);
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint64_t bk = b3_t + (uint64_t)(kx * 2 * W3_ROWS);
              umma_f16(d, aq + kx, bk + 64, ID32, 1u);
              umma_f16(d, aq + (2 * PL_PITCH + kx), bk, ID64, 1u);
              umma_f16(d + 32u, aq + (3 * PL_PITCH + kx), bk, ID32, 1u);
            }
          }
          umma_commit(smem_u32(&acc3_full[q]));          // dead groups too: keeps the barrier phases aligned with the tile count
        }
        P2T(ti3 += clock64() - tq;)
      }
      P2T(atomicAdd(&g_pnet2_phase[5], (unsigned long long)tw4); atomicAdd(&g_pnet2_phase[6], (unsigned long long)tw5);
          atomicAdd(&g_pnet2_phase[18], (unsigned long long)ti3); atomicAdd(&g_pnet2_phase[21], (unsigned long long)(clock64() - tstart));)
    }
    __syncwarp();
  } else if (warp >= W_E1 && warp < W_E2) {
    // =================================================================== conv1 epilogue: pool, bias, PReLU -> p1 planes
    // E1_SETS sets of four warps take the M tiles round robin, so several are in flight (one set needs ~900 cycles per M
    // tile, the MMAs of one take 176)
    const int lg = warp & 3, set = (warp - W_E1) >> 2;
    int turn = 0;                                            // u1 % E1_SETS
    const int g = 4 * lg + (lane >> 3), ci = lane & 7;       // accumulator row = 8 g + ci: pooled row g of the block, column ci of the group
    uint32_t u1 = 0;
    TileCursor cur = cursor_init(p.blocks);
    for (int k = 0; k < n_my; ++k, cursor_next(cur, p.blocks)) {
      const TileRef tr = tile_at(tile_s, cur);
      const Level& Lv = p.lv[tr.lvl];
      const int c1h = Lv.hs - 2, c1w = Lv.ws - 2;
      const int nq = ncg1(tr), nrb = nrb1(tr);
      for (int rb = 0; rb < nrb; ++rb)
        for (int q = 0; q < nq; ++q, ++u1) {
          const uint32_t slot = u1 & 1u;
          const bool mine = turn == set;
          turn = turn + 1 == E1_SETS ? 0 : turn + 1;
          if (!mine) continue;
          mbar_wait_t(&acc1_full[u1 & 7u], (u1 >> 3) & 1, tw0);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + ACC1_COL + 64u * slot;
          float v[4][10];                     // [2 * (conv row parity) + (conv x parity)][channel]
          tmem_ld10(taddr, v[0]); tmem_ld10(taddr + 16, v[1]); tmem_ld10(taddr + 32, v[2]); tmem_ld10(taddr + 48, v[3]);
          tmem_ld_wait();
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(&acc1_empty[u1 & 7u]);
          const int R = 16 * rb + g, c = 8 * q + ci;
          const int gy = 2 * (tr.oy0 + R), gx = 2 * (tr.ox0 + c);
          float m[10];
          if (p.conv1_monotone && gy + 1 < c1h && gx + 1 < c1w) {
            // interior pixel, every slope >= 0: bias + PReLU are non-decreasing and commute with the max
#pragma unroll
            for (int co = 0; co < 10; ++co)
              m[co] = prelu(fmaxf(fmaxf(v[0][co], v[1][co]), fmaxf(v[2][co], v[3][co])) + p.b1[co], p.a1[co]);
          } else {
#pragma unroll
            for (int co = 0; co < 10; ++co) {
              float mm = -INFINITY;
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                const bool ok = (gy + (qq >> 1) < c1h) && (gx + (qq & 1) < c1w);
                const float x = prelu(v[qq][co] + p.b1[co], p.a1[co]);
                mm = ok ? fmaxf(mm, x) : mm;
              }
              m[co] = (mm == -INFINITY) ? 0.f : mm;
            }
          }
          const int n = R * PL_PITCH + c;
          uint8_t* p1 = smem + (k & 1) * P1_STRIDE;
          *reinterpret_cast<uint4*>(p1 + P1A_OFF + n * 16) =
              make_uint4(pack_h2(m[0], m[1]), pack_h2(m[2], m[3]), pack_h2(m[4], m[5]), pack_h2(m[6], m[7]));
          // channels 8, 9 go to the side plane of the three pixels whose kx taps see this one
          const uint32_t w89 = pack_h2(m[8], m[9]);
          uint32_t* side = reinterpret_cast<uint32_t*>(p1 + P1B_OFF) + n * 4;
          side[0] = w89;
          if (c >= 1) side[-4 + 1] = w89;
          if (c >= 2) side[-8 + 2] = w89;
        }
      // The other set may own the tile's last M tile: wait for that one too before arriving, so that no warp can arrive for
      // the next tile before every warp has arrived for this one (conv1 of the next tile is issued behind conv2 of this one,
      // which waits for p1_ready).
      mbar_wait_t(&acc1_full[(u1 - 1) & 7u], ((u1 - 1) >> 3) & 1, tw0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // p1 planes -> visible to the UMMA proxy
      mbar_arrive(&p1_ready[k & 1]);
    }
  } else if (warp >= W_E2 && warp < W_E3) {
    // =================================================================== conv2 epilogue: bias, PReLU -> c2 planes
    const int lg = warp & 3;
    const int g = 4 * lg + (lane >> 3), ci = lane & 7;       // accumulator row: row pair g (rows 2g, 2g + 1), column ci of the group
    TileCursor cur = cursor_init(p.blocks);
    for (int k = 0; k < n_my; ++k, cursor_next(cur, p.blocks)) {
      const TileRef tr = tile_at(tile_s, cur);
      const int nq = ncg2(tr);
      // conv3 of the previous tile (issued by the other MMA thread) must have finished reading the c2 planes
      if (k >= 1) mbar_wait_t(&acc3_full[3], (k - 1) & 1, tw0);
      for (int q = 0; q < nq; ++q) {
        mbar_wait_t(&acc2_full[q], k & 1, tw0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + ACC2_COL + 32u * q;
        float v[2][16];
        tmem_ld16(taddr, v[0]); tmem_ld16(taddr + 16, v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float x[16];
#pragma unroll
          for (int co = 0; co < 16; ++co) {
            x[co] = prelu(v[r][co] + p.b2[co], p.a2[co]);
          }
          const int n = (2 * g + r) * PL_PITCH + 8 * q + ci;
          *reinterpret_cast<uint4*>(smem + C0_OFF + n * 16) =
              make_uint4(pack_h2(x[0], x[1]), pack_h2(x[2], x[3]), pack_h2(x[4], x[5]), pack_h2(x[6], x[7]));
          *reinterpret_cast<uint4*>(smem + C1_OFF + n * 16) =
              make_uint4(pack_h2(x[8], x[9]), pack_h2(x[10], x[11]), pack_h2(x[12], x[13]), pack_h2(x[14], x[15]));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&acc2_empty);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&c2_ready);
    }
  } else if (warp >= W_E3) {
    // =================================================================== conv3 epilogue: bias, PReLU, conv4_1 logit difference -> screen
    const int lg = warp & 3, set3 = (warp - W_E3) >> 2;
    const int g = 4 * lg + (lane >> 3), ci = lane & 7;
    TileCursor cur = cursor_init(p.blocks);
    for (int k = 0; k < n_my; ++k, cursor_next(cur, p.blocks)) {
      const TileRef tr = tile_at(tile_s, cur);
      const Level& Lv = p.lv[tr.lvl];
      const int nq = ncg3(tr);
      for (int q = 0; q < 4; ++q) {
        // dead groups are waited for as well (their barrier is committed empty): a warp without work must not run ahead and
        // arrive on acc3_empty for the next tile before the others have arrived for this one
        mbar_wait_t(&acc3_full[q], k & 1, tw0);
        if (q >= nq || (q % E3_SETS) != set3) continue;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + ACC3_COL + 64u * q;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float dsum = p.db;
          float v[2][16];
          tmem_ld16(taddr + 32 * r, v[0]); tmem_ld16(taddr + 32 * r + 16, v[1]);
          tmem_ld_wait();
          // logit = db + sum_c k1_c u_c + k2_c |u_c| with u = acc + b3 (PReLU(u) dw = k1 u + k2 |u|): four instructions per channel
          // with constant-bank operands, two independent accumulation chains
          float d2 = 0.f;
#pragma unroll
          for (int c0 = 0; c0 < 32; ++c0) {
            const float u = v[c0 >> 4][c0 & 15] + p.b3[c0];
            dsum = fmaf(u, p.k1[c0], dsum);
            d2 = fmaf(fabsf(u), p.k2[c0], d2);
          }
          dsum += d2;
          const int row = 2 * g + r, col = 8 * q + ci;
          const int oy = tr.oy0 + row, ox = tr.ox0 + col;
          const bool live = row < tr.rows && col < tr.cols;
          if (live && Lv.logit) Lv.logit[((size_t)tr.b * Lv.oh + oy) * Lv.ow + ox] = dsum;
          const bool hit = live && p.screen && dsum >= p.logit_lo;
          const uint32_t mask = __ballot_sync(0xffffffffu, hit);
          if (mask) {
            int base = 0;
            if (lane == (__ffs(mask) - 1)) base = atomicAdd(p.screen_cnt, __popc(mask));
            base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
            if (hit) {
              const int slot = base + __popc(mask & ((1u << lane) - 1u));
              if (slot < p.screen_cap) {
                p.screen[slot] = ScreenEntry{tr.b * p.n_levels + tr.lvl, oy * Lv.ow + ox};
              } else if (p.capflag) {
                p.capflag->overflow = 1; p.capflag->stage = 1; p.capflag->frame = tr.b;
                p.capflag->count = slot + 1; p.capflag->capacity = p.screen_cap;
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&acc3_empty);
    }
  }
  P2T(if (lane == 0 && warp != W_MMA) {
        const unsigned long long tot = (unsigned long long)(clock64() - tstart), w = (unsigned long long)tw0;
        const int slot = warp == W_E1 ? 8 : warp == W_E2 ? 10 : warp == W_E3 ? 12 : warp == W_TMA ? 14 : -1;
        if (slot >= 0) { atomicAdd(&g_pnet2_phase[slot], w); if (slot != 14) atomicAdd(&g_pnet2_phase[slot + 1], tot - w); }
      })
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

}  // namespace pnet2
#ifdef PNET_TIMING
extern "C" void trl_debug_pnet2_timing(unsigned long long* out) {
  cudaMemcpyFromSymbol(out, pnet2::g_pnet2_phase, sizeof(unsigned long long) * 24);
  unsigned long long z[24] = {0};
  cudaMemcpyToSymbol(pnet2::g_pnet2_phase, z, sizeof(z));
}
#endif

// ---- host side

static uint16_t h16(float x) {
  const __half h = __float2half_rn(x);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

// upstream layouts -> the fp16 core-matrix images of the three B operands + the epilogue constants (kept in the context, copied
// into the kernel parameters at launch)
int pnet2_pack_weights(trl_ctx* c, const float* h, size_t len) {
  using namespace pnet2;
  if (len != 6632) TRL_FAIL(c, TRL_E_INVALID, "pnet blob has %zu floats, expected 6632", len);
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
  std::vector<uint16_t> img(W_BYTES / 2, 0);
  // element (row, k) of a matrix with `rows` rows at byte offset `off`: [k / 8][row][k % 8]
  auto at = [&](int off, int rows, int row, int k) -> uint16_t& { return img[off / 2 + ((k >> 3) * rows + row) * 8 + (k & 7)]; };
  for (int dy = 0; dy < 3; ++dy) {
    const int blk = 2 - dy;              // row blocks in the order dy = 2, 1, 0
    // conv1: rows [even-x 16 | odd-x 16]; K = 4 pixels (2c .. 2c + 3) x (B, G, R, 0)
    for (int co = 0; co < 10; ++co)
      for (int ci = 0; ci < 3; ++ci)
        for (int kx = 0; kx < 3; ++kx) {
          const uint16_t w = h16(w1[co * 27 + ci * 9 + dy * 3 + kx]);
          at(W1_OFF, W1_ROWS, blk * 32 + co, 4 * kx + ci) = w;                 // output x = 2c
          at(W1_OFF, W1_ROWS, blk * 32 + 16 + co, 4 * (kx + 1) + ci) = w;      // output x = 2c + 1
        }
    // conv2: matrix a = [kx0 ch0-7 | kx1 ch0-7], matrix b = [kx2 ch0-7 | ch8, ch9 of kx0, kx1, kx2, 0, 0]
    for (int co = 0; co < 16; ++co)
      for (int ci = 0; ci < 10; ++ci)
        for (int kx = 0; kx < 3; ++kx) {
          const uint16_t w = h16(w2[(co * 10 + ci) * 9 + dy * 3 + kx]);
          const int row = blk * 16 + co;
          if (ci < 8) {
            if (kx < 2) at(W2_OFF, W2_ROWS, row, 8 * kx + ci) = w;
            else at(W2_OFF + 2 * W2_ROWS * 16, W2_ROWS, row, ci) = w;
          } else {
            at(W2_OFF + 2 * W2_ROWS * 16, W2_ROWS, row, 8 + 2 * kx + (ci - 8)) = w;
          }
        }
    // conv3: one matrix per kx, K = 16 input channels
    for (int co = 0; co < 32; ++co)
      for (int ci = 0; ci < 16; ++ci)
        for (int kx = 0; kx < 3; ++kx)
          at(W3_OFF + kx * 2 * W3_ROWS * 16, W3_ROWS, blk * 32 + co, ci) = h16(w3[(co * 16 + ci) * 9 + dy * 3 + kx]);
  }
  TRL_CUDA(c, cudaMalloc(&c->d_pnet2_packed, W_BYTES));
  TRL_CUDA(c, cudaMemcpy(c->d_pnet2_packed, img.data(), W_BYTES, cudaMemcpyHostToDevice));
  float* e = c->h_pnet2_epi;
  memcpy(e + 0, b1, 40); memcpy(e + 10, a1, 40);
  memcpy(e + 20, b2, 64); memcpy(e + 36, a2, 64);
  memcpy(e + 52, b3, 128); memcpy(e + 84, a3, 128);
  for (int ci = 0; ci < 32; ++ci) e[116 + ci] = w41[32 + ci] - w41[ci];
  e[148] = b41[1] - b41[0];
  // Range of the fp16 operands, bounded on the host instead of checked per value on the device: inputs are in [-1, 1], so
  // |conv1| <= sum |w1| + |b1|, PReLU scales by at most max(1, |a|), and so on for conv2.  The hybrid P-Net is only offered
  // when both bounds are far inside fp16's range (trl_create falls back to the 3-term kernel otherwise).
  float bound1 = 0.f, bound2 = 0.f;
  for (int co = 0; co < 10; ++co) {
    float sa = fabsf(b1[co]);
    for (int k = 0; k < 27; ++k) sa += fabsf(w1[co * 27 + k]);
    bound1 = fmaxf(bound1, sa * fmaxf(1.f, fabsf(a1[co])));
  }
  for (int co = 0; co < 16; ++co) {
    float sa = fabsf(b2[co]);
    for (int k = 0; k < 90; ++k) sa += fabsf(w2[co * 90 + k]) * bound1;
    bound2 = fmaxf(bound2, sa * fmaxf(1.f, fabsf(a2[co])));
  }
  c->pnet2_range_ok = (bound1 < ACT_MAX && bound2 < ACT_MAX && isfinite(bound1) && isfinite(bound2)) ? 1 : 0;
  TRL_CUDA(c, cudaFuncSetAttribute(pnet2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CAP));
  return TRL_OK;
}

// All levels of B frames.  d_logit (optional): the conv4_1 logit difference of every cell, levels concatenated, [B][oh][ow]
// each (stage entry point trl_pnet_screen_maps); d_screen (optional): the screen list.
int launch_pnet2(trl_ctx* c, const uint4* d_pyr_hi, int B, const PyramidGeom& g, float thr_lo, ScreenEntry* d_screen,
                 int* d_screen_cnt, int screen_cap, float* d_logit, cudaStream_t s) {
  using namespace pnet2;
  if (B == 0 || g.n == 0) return TRL_OK;
  if (!c->d_pnet2_packed) TRL_FAIL(c, TRL_E_STATE, "P-Net weights not loaded");
  Params p{};
  p.n_levels = g.n;
  int blocks = 0;
  long long logit_off = 0;
  std::vector<int4> table;
  for (int k = 0; k < g.n; ++k) {
    Level& L = p.lv[k];
    L.hs = g.hs[k]; L.ws = g.ws[k]; L.oh = g.oh[k]; L.ow = g.ow[k];
    L.tiles_x = L.ow > 0 ? ceil_div(L.ow, T) : 0;
    L.tiles = (L.oh > 0 && L.ow > 0) ? L.tiles_x * ceil_div(L.oh, T) : 0;
    L.logit = d_logit ? d_logit + logit_off : nullptr;
    if (L.oh > 0 && L.ow > 0) logit_off += (long long)B * L.oh * L.ow;
    for (int t = 0; t < L.tiles; ++t) {
      const int ty = t / L.tiles_x, tx = t - ty * L.tiles_x;
      const int rows = std::min(T, L.oh - ty * T), cols = std::min(T, L.ow - tx * T);
      if ((int)table.size() < MAX_TILES_PARAM)
        p.groups[table.size()] = (uint16_t)(((cols + 4 + 7) >> 3) | ((rows + 4 > 16 ? 2 : 1) << 4) | (((cols + 2 + 7) >> 3) << 8) | (((cols + 7) >> 3) << 12));
      table.push_back(make_int4(k | (rows << 8) | (cols << 16), ty * T, tx * T, 0));
    }
    blocks += L.tiles;
    const unsigned long long dims[3] = {2ull * L.ws, (unsigned long long)L.hs, (unsigned long long)B};
    const unsigned long long strides[2] = {(unsigned long long)g.pitch2[k] * 16, (unsigned long long)L.hs * g.pitch2[k] * 16};
    const unsigned box[3] = {(unsigned)(IN_PAIRS * 4), (unsigned)IN_ROWS, 1u};
    int rc = tma_encode_tiled_f32(c, &p.tmap[k], d_pyr_hi + g.off2[k] * B, 3, dims, strides, box);
    if (rc != TRL_OK) return rc;
  }
  p.blocks = blocks; p.n_frames = B;
  if (blocks == 0) return TRL_OK;
  if (blocks > MAX_TILES_PARAM) TRL_FAIL(c, TRL_E_INVALID, "frame has %d P-Net tiles, the group table holds %d", blocks, MAX_TILES_PARAM);
  if (blocks > MAX_TILES_PER_FRAME) TRL_FAIL(c, TRL_E_INVALID, "frame has %d P-Net tiles, the tile table holds %d", blocks, MAX_TILES_PER_FRAME);
  // the tile table of this frame geometry lives in the context; it is rebuilt (stream ordered) only when the geometry changes
  if (c->pnet2_tiles_n != blocks || c->pnet2_tiles_key != ((long long)g.hs[0] << 32 | (unsigned)g.ws[0]) || !c->d_pnet2_tiles) {
    if (c->d_pnet2_tiles) { TRL_CUDA(c, cudaStreamSynchronize(s)); TRL_CUDA(c, cudaFree(c->d_pnet2_tiles)); c->d_pnet2_tiles = nullptr; }
    TRL_CUDA(c, cudaMalloc(&c->d_pnet2_tiles, table.size() * sizeof(int4)));
    TRL_CUDA(c, cudaMemcpy(c->d_pnet2_tiles, table.data(), table.size() * sizeof(int4), cudaMemcpyHostToDevice));
    c->pnet2_tiles_n = blocks;
    c->pnet2_tiles_key = ((long long)g.hs[0] << 32 | (unsigned)g.ws[0]);
  }
  p.tiles = reinterpret_cast<const int4*>(c->d_pnet2_tiles);
  // prob >= thr_lo  <=>  logit difference >= log(thr_lo / (1 - thr_lo)); thresholds at or below the margin screen every cell
  p.logit_lo = thr_lo <= 0.f ? -INFINITY : thr_lo >= 1.f ? INFINITY : logf(thr_lo / (1.f - thr_lo));
  p.screen = d_screen; p.screen_cnt = d_screen_cnt; p.screen_cap = screen_cap; p.capflag = c->d_cap;
  const float* e = c->h_pnet2_epi;
  const float *b1 = e + 0, *a1 = e + 10, *b2 = e + 20, *a2 = e + 36, *b3 = e + 52, *a3 = e + 84, *dw = e + 116;
  memcpy(p.b1, b1, 40); memcpy(p.a1, a1, 40);
  memcpy(p.b2, b2, 64); memcpy(p.a2, a2, 64);
  for (int ci = 0; ci < 32; ++ci) {
    p.b3[ci] = b3[ci];
    p.k1[ci] = dw[ci] * 0.5f * (1.f + a3[ci]);
    p.k2[ci] = dw[ci] * 0.5f * (1.f - a3[ci]);
  }
  p.db = e[148];
  p.conv1_monotone = 1;
  for (int co = 0; co < 10; ++co) if (!(a1[co] >= 0.f)) p.conv1_monotone = 0;
  const long long total = (long long)blocks * B;
  const int grid = (int)(total < c->num_sms ? total : c->num_sms);
  pnet2_kernel<<<grid, NTHREADS, SMEM_BYTES + blocks * 16, s>>>(c->d_pnet2_packed, p);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

int launch_pnet2_screen(trl_ctx* c, const uint4* d_pyr_hi, int B, const PyramidGeom& g, float thr_lo, ScreenEntry* d_screen,
                        int* d_screen_cnt, int screen_cap, cudaStream_t s) {
  return launch_pnet2(c, d_pyr_hi, B, g, thr_lo, d_screen, d_screen_cnt, screen_cap, nullptr, s);
}
