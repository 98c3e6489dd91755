// K2 + K3: fused P-Net (conv1+PReLU+maxpool, conv2+PReLU, conv3+PReLU, conv4_1 -> softmax, conv4_2) and
// generateBoundingBox.  One CTA computes a 20x28 tile of output cells at a time; every intermediate lives in shared memory, the
// only HBM traffic is the input tile and the (rare) candidates.
//
// upstream: models/mtcnn.py PNet.forward, models/utils/detect_face.py generateBoundingBox.
//
// Persistent and warp specialised: ONE 544-thread CTA per SM walks the list of tiles of every level and frame.
//   group A (warps 0-7)  : stages the fp32 input tile (one 3-D TMA box, zero filled out of bounds, double buffered; cp.async when
//                          the rows are not 16-byte aligned) and runs conv1 + PReLU + 2x2 max-pool on the FMA pipe (packed FFMA2),
//                          writing the pooled tile as scaled fp16 hi / lo planes (double buffered towards group B);
//   group B (warps 8-15) : conv2 as an implicit GEMM on the legacy tensor path (mma.sync.m16n8k16 -> HMMA.16816.F32), its
//                          bias / PReLU / split epilogue into the UMMA operand planes, and the conv3 + heads epilogue
//                          (tcgen05.ld -> bias, PReLU, conv4_1 / conv4_2, softmax, generateBoundingBox, candidate append);
//   warp 16              : issues conv3 as tcgen05.mma implicit GEMMs (A = the conv2 tile in shared memory read through
//                          shifted no-swizzle K-major descriptors, one per filter tap; fp32 accumulators in TMEM).
// conv1 is 15 % of the MACs, conv2 + conv3 85 %.  The tensor-pipe layers keep fp32-level accuracy with a 3-term fp16 split:
// every operand x is scaled by a power of two S and stored as hi = fp16(S x), lo = fp16(S x - hi);
//     S_a S_w sum(a w) ~= sum(a_hi w_hi) + sum(a_lo w_hi) + sum(a_hi w_lo)          (fp32 accumulate)
// drops only a_lo w_lo (2^-22 relative) and lo's own rounding (2^-22 relative; the scale keeps lo out of fp16's
// subnormal range for |x| >= 2^-9, below that the absolute error is < 1e-9).  The result is unscaled exactly in the
// epilogue.  Against the fp32 oracle the maps agree to ~1e-6 (tests/test_gpu_stages.py bar: 2e-5), so the cascade
// sees the same candidates.  The operands are split once where they are produced (weights on the host, activations in the
// producing layer's epilogue), not in the MMA loop.  Activations beyond +-1000 would overflow the scaled fp16 hi part:
// the kernel raises the capacity flag (stage 5) instead of continuing silently.
// A single-pass variant (TERMS = 1: hi parts only, trl_config_t.pnet_precision = 1) exists as a measured experiment: 15 %
// faster, maps within ~1e-4, but the cascade is chaotic in the last bits (integer truncation of boxes) -- face counts, boxes
// and embeddings left the north-star tolerances on the bundled and the 1080p clips, so it is not the default
// (profiles/PROFILE_NOTES.md r02).
// Roofs (profiles/r02_pnet_full.md, experiments/fma_rate_probe.cu, hmma_rate_probe.cu): per tile the FMA pipe is busy
// ~9 k of 18.5 k cycles (conv1's FFMA2 stream + the two epilogues), the tensor pipe ~9 k (1476 HMMA at one per 2.7 cycles
// next to an FFMA2 stream + 90 UTCHMMA at ~60 cycles), the issue slots 65 %: no single pipe is saturated, the two groups are
// latency bound at 4.25 warps per scheduler (216 KB of shared memory pin the kernel to one CTA per SM).  HBM is not the
// roof: the kernel reads the pyramid exactly once.
#include <cuda.h>

#include "common.cuh"
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

namespace pnet {

// Tile shape.  Group A's conv1 sets the pace of the kernel and runs one pooled pixel per thread in rounds of 256:
// 20 x 28 cells need (20 + 4) x (28 + 4) = 768 pooled pixels = exactly three full rounds for 560 cells, and conv3's
// 20 x 32 flat pixels are exactly five M tiles.  16 x 32 (720 pooled pixels = three rounds, the last one 81 % full, for
// 512 cells; 4.5 M tiles) measured 10.35-10.6 ms per 450 720p frames against 9.74 ms for 20 x 28.
#ifndef PNET_TOY
#define PNET_TOY 20
#define PNET_TOX 28
#endif
#ifndef PNET_C1_FFMA2
#define PNET_C1_FFMA2 1
#endif
#ifndef PNET_C2_PIPE
#define PNET_C2_PIPE 0
#endif
constexpr int TOY = PNET_TOY, TOX = PNET_TOX;   // output cells per CTA
constexpr int P1H = TOY + 4, P1W = TOX + 4;  // pooled conv1 tile (24 x 32 for the 20 x 28 tile), flat pixel index n = row * P1W + col
constexpr int C2TILES_ = ((TOY + 2) * P1W + 15) / 16;
// pixels per p1 plane: the tile + slack read by conv2's last (partial) M tile, rounded up to 8 mod 32 (776 for 20 x 28)
constexpr int P1PL = ((C2TILES_ * 16 + 2 * P1W + 2 - 8 + 31) / 32) * 32 + 8;
static_assert(P1PL >= P1H * P1W && P1PL >= C2TILES_ * 16 + 2 * P1W + 2, "p1 plane covers the tile and conv2's slack reads");
                                             // spreads the four channel-pair planes a warp reads at once over the banks
constexpr int P1WORDS = 5;                   // 10 channels = 5 half2 planes [pair][pixel] (hi set, lo set)
constexpr int C2H = TOY + 2, C2P = P1W;      // conv2 tile: 22 rows at the same pitch as its input (flat indexing)
constexpr int C2PX = C2H * C2P;              // 704 pixels (columns 30, 31 of a row are computed but never read)
constexpr int C2TILES = (C2PX + 15) / 16;    // 44 M tiles of two 8-pixel segments
// conv2 output = conv3's UMMA A operand: four planes [hi k0, hi k1, lo k0, lo k1] of [pixel][8 channels = 16 B]
// (no-swizzle K-major core matrices: 8 consecutive pixels x 16 B; k-half stride = plane, 8-pixel stride = 128 B)
constexpr int C3TILES = (TOY * C2P + 127) / 128;   // conv3 M tiles of 128 flat pixels (20 rows x pitch 32 = 640 = exactly 5 tiles)
constexpr int C2NP = ((C3TILES * 128 + 2 * C2P + 2 + 15) / 16) * 16;   // pixels per plane: computed + slack read by the last M tile (720 for 20 x 28)
static_assert(C2NP >= C2PX, "c2 plane covers the conv2 tile");
constexpr int C2PLANE = C2NP * 4;            // words per plane
constexpr int TMEM_COLS = 512;
// DEFER_EPILOGUE = true: two accumulator sets of 256 columns (tiles 0-2 x 64, tiles 3-4 x 32 with a third MMA per tap) and
// the head epilogue of tile k-1 runs under the MMAs of tile k.  Measured slower on B200 (15.1 vs 12.0 ms per 450 frames):
// the MMAs are limited by shared-memory operand bandwidth, which the epilogue does not relieve, and the third MMA adds
// 12 % tensor work.  false: one set of five 64-column accumulators, epilogue per tile as its commit arrives.
constexpr bool DEFER_EPILOGUE = false;
__host__ __device__ constexpr bool tile_n64(int tile) { return DEFER_EPILOGUE ? tile < 3 : true; }
__host__ __device__ constexpr uint32_t tile_col(int set, int tile) {
  return DEFER_EPILOGUE ? 256u * (uint32_t)(set & 1) + (tile < 3 ? 64u * tile : 192u + 32u * (tile - 3)) : 64u * tile;
}
constexpr int NB_WARPS = 8;                  // group B (conv2 + heads) warps; 12 was measured no faster (11.46 vs 11.35 ms per 450 frames)
constexpr int NA_WARPS = 8;                  // group A (staging + conv1) warps; 12 was measured slower (10.84 vs 10.58 ms)
constexpr int NA_THREADS = 32 * NA_WARPS;
constexpr int NB_THREADS = 32 * NB_WARPS;
constexpr int ISSUE_WARP = NA_WARPS + NB_WARPS;
constexpr int NTHREADS = 32 * (ISSUE_WARP + 1);   // group A warps: staging + conv1, then group B: conv2 + heads, last warp: conv3 MMA issue
static_assert(NB_WARPS % 4 == 0 && NA_WARPS % 4 == 0, "group B covers the four TMEM lane groups evenly and starts at a multiple of 4");
constexpr int INH = 2 * TOY + 10, INW = 2 * TOX + 10;   // 50 x 66 input tile
constexpr int INP = (INW + 3) / 4 * 4;       // row pitch: 16-byte multiple (68 for 20 x 28)
static_assert(C3TILES * 64 <= 512, "conv3 accumulators fit TMEM");

constexpr float SA = 64.f;                   // activation scale (p1 and c2 tiles)
constexpr float ACT_MAX = 1000.f;            // |activation| bound for the fp16 hi part (64 * 1000 < 65504)

// packed weights (32-bit words)
constexpr int W1 = 0;                 // fp32 [27][12]
constexpr int B1 = W1 + 27 * 12;      // [12]
constexpr int A1 = B1 + 12;           // [12]
constexpr int W2 = A1 + 12;           // half2 B fragments: [6 k-steps][hi, lo][32 lanes][n0 b0, n0 b1, n1 b0, n1 b1]
constexpr int T2 = W2 + 6 * 2 * 32 * 4;   // int[48]: p1 word offset (plane + tap shift) of channel pair P (see conv2_pair())
constexpr int B2 = T2 + 48;           // [16]
constexpr int A2 = B2 + 16;           // [16]
constexpr int W3 = A2 + 16;           // UMMA B operand: [9 taps][k-half][64 rows: w_hi[32], w_lo[32]][8 channels as fp16]
constexpr int SC = W3 + 9 * 2 * 64 * 4;   // [4]: 1 / (SA * S_w2)
constexpr int WTOTAL = SC + 4;
static_assert(WTOTAL % 4 == 0 && W2 % 4 == 0 && W3 % 4 == 0, "16-byte alignment of the operand arrays");

constexpr int SM_IN = ((3 * INH * INP + 31) / 32) * 32;      // 10208 words, padded to 128 bytes (TMA destination alignment)
constexpr int IN_BYTES = 3 * INH * INP * 4;                  // bytes one TMA box delivers
constexpr int IN_OFF = ((WTOTAL + 31) / 32) * 32;            // input tiles start 128-byte aligned behind the weights
constexpr int SM_C2 = 4 * C2PLANE;                   // 11520 words
constexpr int SM_P1 = 2 * P1WORDS * P1PL;            // 7760 words
constexpr int SMEM_WORDS = IN_OFF + 2 * SM_IN + 2 * SM_P1 + SM_C2;    // weights, 2 input tiles, 2 pooled conv1 tiles, conv2 planes
constexpr int SMEM_BYTES = SMEM_WORDS * 4;           // 216 KB -> one persistent CTA per SM
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct Level {
  const float* in;     // [B][3][hs][pitch]
  int hs, ws, pitch, oh, ow;
  int tiles_x, tiles;  // tiles per frame
  float scale;
  float* prob;         // optional maps
  float* reg;
  Cand* cand;          // optional: candidate list of (frame, level): cand + (b*n_levels + lvl)*cap
  int* cnt;
};

struct Params {
  CUtensorMap tmap[TRL_MAX_SCALES];   // per level: fp32 {ws, hs, 3 B} view of the pyramid, box {INP, INH, 3} = {68, 50, 3}, zero fill out of bounds
  int use_tma;         // 1: every level has a tensor map (16-byte aligned rows); 0: cp.async staging
  int n_levels;
  int blocks;          // tiles per frame (all levels)
  int n_frames;
  int blk_start[TRL_MAX_SCALES + 1];
  Level lv[TRL_MAX_SCALES];
  float thr;
  int cap;
  CapFlag* capflag;
  float head[32][8];   // per conv3 channel: conv4_1 (2), conv4_2 (4), conv3 bias, conv3 PReLU slope -- read as constant-bank FFMA operands
  float head_bias[8];
  float inv3;          // 1 / (SA * S_w3)
  float conv1_bias[10], conv1_slope[10];
  int conv1_monotone;  // 1: every conv1 PReLU slope is >= 0 (max-pool may run before bias + PReLU, exactly)
  ScreenEntry* screen; // non-null: screening launch (hybrid P-Net) -- cells with prob >= thr go to this list for pnet_refine.cu
  int* screen_cnt;
  int screen_cap;
};

#ifdef PNET_TIMING
__device__ unsigned long long g_pnet_phase[8];
#endif
__device__ __forceinline__ float prelu(float v, float a) { return v > 0.f ? v : v * a; }
// (x0, x1), already scaled -> packed fp16 hi and lo = fp16(x - hi)
__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// ---- tcgen05 (conv3): D[128 px x N] (TMEM, fp32) += A[128 px x 16 ch] (smem) * B[N x 16 ch] (smem), fp16 operands
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred px;\nelect.sync _|px, %1;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}
// no-swizzle K-major descriptor: start >> 4 | (k-half stride >> 4) << 16 | (8-row stride >> 4) << 32 | version 1 << 46
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  if (ok) return;                                      // fast path: no clock read
  const long long t0 = clock64();
  while (true) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) break;
    if (clock64() - t0 > 4000000000LL) __trap();        // ~2 s: a lost commit must not hang the device
  }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// tile id -> (frame, level, tile origin); uniform per call
struct TileRef { int b, lvl, oy0, ox0; };
__device__ __forceinline__ TileRef decode_tile(const Params& p, int id) {
  TileRef r;
  r.b = id / p.blocks;
  const int blk = id - r.b * p.blocks;
  int lvl = 0;
  while (lvl + 1 < p.n_levels && blk >= p.blk_start[lvl + 1]) ++lvl;
  const int tile = blk - p.blk_start[lvl];
  const int ty = tile / p.lv[lvl].tiles_x;
  r.lvl = lvl; r.oy0 = ty * TOY; r.ox0 = (tile - ty * p.lv[lvl].tiles_x) * TOX;
  return r;
}
// conv2 M tiles per group-B warp (see the conv2 loop), one nibble per warp, warp 0 lowest: 5 5 5 5 6 6 6 6 = 44 for the
// 20 x 28 tile (warps 0-3 carry three conv3 head epilogues per tile, warps 4-7 two; the reverse split measured 9.91
// against 9.74 ms).  On the 16 x 32 tile (41 M tiles) 4 4 5 6 5 6 5 6 gave 10.35 ms, an even split 10.79 ms.
#ifndef PNET_C2_COUNTS
#define PNET_C2_COUNTS (PNET_TOY == 20 && PNET_TOX == 28 ? 0x66665555u : 0x65656544u)
#endif
static_assert(C2TILES_ == (C2PX + 15) / 16, "");
__host__ __device__ constexpr int c2_count(int warp) { return (int)((PNET_C2_COUNTS >> (4 * warp)) & 0xFu); }
__host__ __device__ constexpr int c2_first(int warp) {
  int f = 0;
  for (int w = 0; w < warp; ++w) f += c2_count(w);
  return f;
}
static_assert(NB_WARPS == 8 && c2_first(NB_WARPS) == C2TILES, "the conv2 tile split covers every M tile once");
__host__ __device__ constexpr int c2_max_count() {
  int m = 0;
  for (int w = 0; w < 8; ++w) m = c2_count(w) > m ? c2_count(w) : m;
  return m;
}
constexpr int C2PASSES = (c2_max_count() + 2) / 3;     // conv2 passes of up to three M tiles per warp
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent, warp-specialised: one CTA per SM walks the tile list.  Warps 0-7 (group A) stage the input tile
// (cp.async, double buffered) and run conv1 on the FMA pipe; warps 8-15 (group B) run conv2 (mma.sync), issue conv3
// (tcgen05) and do the head epilogue.  The pooled conv1 tile is double buffered between the groups (full / empty
// mbarriers), so conv1 of tile k+1 overlaps conv2 / conv3 of tile k and the FMA pipe, the tensor pipe and the
// issue slots are busy at the same time.  Weights are loaded once per CTA.
// TERMS = 3: the 3-term fp16 operand split (fp32-equivalent maps).  TERMS = 1: single-pass fp16 operands (hi parts only,
// fp32 accumulate) for conv2 / conv3 -- a third of the tensor work, maps within ~1e-3 of the fp32 oracle instead of 2e-6;
// selected by trl_config_t.pnet_precision and judged at cascade level (same face count, IoU >= 0.95).
template <int TERMS>
__global__ void __launch_bounds__(NTHREADS, 1) pnet_kernel(const float* __restrict__ wpacked, const __grid_constant__ Params p) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t mma_bar[2 * C3TILES];   // [accumulator set][conv3 M tile]: tcgen05.commit arrives
  __shared__ __align__(8) uint64_t p1_full[2], p1_empty[2];
  __shared__ __align__(8) uint64_t in_full[2];         // input tile landed (TMA complete_tx)
  __shared__ __align__(8) uint64_t c2_full;            // conv2 output of the current tile is complete (NB_THREADS arrivals)
  __shared__ uint32_t tmem_slot;
  float* w_s = smem;
  float* in_base = smem + IN_OFF;                                             // [2][SM_IN] fp32 input tiles
  uint32_t* p1_base = reinterpret_cast<uint32_t*>(in_base + 2 * SM_IN);       // [2][hi, lo][5 pairs][P1PL]
  uint32_t* c2_s = p1_base + 2 * SM_P1;                                       // [hi k0, hi k1, lo k0, lo k1][C2NP][4]
  const uint32_t* wu = reinterpret_cast<const uint32_t*>(w_s);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);    // warp-uniform copy (keeps the MMA issue path in uniform registers)
  bool range_bad = false;
#ifdef PNET_TIMING
  long long tw = 0, tc1 = 0, tc2 = 0, tc3 = 0, tmma = 0, tmark;
#endif

  if (warp_u == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 2 * C3TILES; ++i)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mma_bar[i])), "r"(1u) : "memory");
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&p1_full[i])), "r"((uint32_t)NA_THREADS) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&p1_empty[i])), "r"((uint32_t)NB_THREADS) : "memory");
      }
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&c2_full)), "r"((uint32_t)NB_THREADS) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&in_full[0])), "r"(1u) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&in_full[1])), "r"(1u) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  // weights (once per CTA) + the slack pixels of both p1 buffers (read by conv2's last M tile, never written by conv1)
  {
    const uint32_t w_dst = smem_u32(w_s);
    for (int i = tid; i < WTOTAL / 4; i += NTHREADS)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(w_dst + 16u * i), "l"(wpacked + 4 * i) : "memory");
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    for (int i = tid; i < 2 * 2 * P1WORDS * (P1PL - P1H * P1W); i += NTHREADS) {
      const int per = P1PL - P1H * P1W;
      const int plane = i / per, off = i - plane * per;                       // plane = (buf * 2 + (hi | lo)) * 5 + pair
      p1_base[plane * P1PL + P1H * P1W + off] = 0u;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int total = p.blocks * p.n_frames;

  if (warp_u < NA_WARPS) {
    // =================================================================== group A: input staging + conv1
    // the tile is staged with cp.async (LDGSTS): every copy of a thread is in flight at once; out-of-image elements
    // are zero filled (src-size 0)
    auto issue_load = [&](int id, int buf) {
      const TileRef tr = decode_tile(p, id);
      if (p.use_tma) {
        // one elected thread: a single 3-D TMA box {76 columns, 42 rows, 3 channel planes}; rows / columns beyond the
        // level are zero filled by the TMA unit
        if (tid == 0) {
          const uint32_t bar = smem_u32(&in_full[buf]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)IN_BYTES) : "memory");
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                       ::"r"(smem_u32(in_base + buf * SM_IN)), "l"(reinterpret_cast<uint64_t>(&p.tmap[tr.lvl])), "r"(bar),
                         "r"(2 * tr.ox0), "r"(2 * tr.oy0), "r"(3 * tr.b) : "memory");
        }
        return;
      }
      const Level& Lv = p.lv[tr.lvl];
      const int hs = Lv.hs, ws = Lv.ws, pitch = Lv.pitch;
      const float* src = Lv.in + (size_t)tr.b * 3 * hs * pitch;
      const int iy0 = 2 * tr.oy0, ix0 = 2 * tr.ox0;
      const uint32_t a_dst = smem_u32(in_base + buf * SM_IN);
      if ((pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(Lv.in) & 15) == 0) {
        // 16-byte rows (the cascade's padded pyramid): 19 chunks per tile row, partial chunks zero filled by src-size
        constexpr int CH = INP / 4;
        for (int i = tid; i < 3 * INH * CH; i += NA_THREADS) {
          const int rr = i / CH, j = i - rr * CH;          // rr = ci * INH + r
          const int ci = rr / INH, r = rr - ci * INH;
          const int gy = iy0 + r, gx = ix0 + 4 * j;
          const int nb = gy < hs ? min(max(ws - gx, 0), 4) * 4 : 0;
          const float* gp = nb ? src + ((size_t)ci * hs + gy) * pitch + gx : src;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                       ::"r"(a_dst + 16u * (uint32_t)i), "l"(gp), "r"(nb) : "memory");
        }
      } else {
        for (int i = tid; i < 3 * INH * INW; i += NA_THREADS) {
          const int ci = i / (INH * INW);
          const int r = (i - ci * INH * INW) / INW;
          const int cx = i - ci * INH * INW - r * INW;
          const int gy = iy0 + r, gx = ix0 + cx;
          const bool ok = gy < hs && gx < ws;
          const float* gp = ok ? src + ((size_t)ci * hs + gy) * pitch + gx : src;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;"
                       ::"r"(a_dst + 4u * ((ci * INH + r) * INP + cx)), "l"(gp), "r"(ok ? 4 : 0) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if ((int)blockIdx.x < total) issue_load(blockIdx.x, 0);
    int k = 0;
    for (int id = blockIdx.x; id < total; id += gridDim.x, ++k) {
#ifdef PNET_TIMING
      tmark = clock64();
#endif
      if (p.use_tma) mbar_wait(&in_full[k & 1], (k >> 1) & 1);
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(NA_THREADS) : "memory");      // tile k landed for all of A; all of A is done reading tile k-1
      if (id + (int)gridDim.x < total) issue_load(id + gridDim.x, (k + 1) & 1);
      if (k >= 2) mbar_wait(&p1_empty[k & 1], ((k >> 1) - 1) & 1);      // group B is done with tile k-2's p1
#ifdef PNET_TIMING
      tw += clock64() - tmark; tmark = clock64();
#endif
      const TileRef tr = decode_tile(p, id);
      const Level& Lv = p.lv[tr.lvl];
      const int oy0 = tr.oy0, ox0 = tr.ox0;
      const float* in_s = in_base + (k & 1) * SM_IN;
      uint32_t* p1_s = p1_base + (k & 1) * SM_P1;
      // ---- conv1 (3->10, 3x3) + PReLU + maxpool(2,2,ceil): one pooled pixel x 10 channels per item, written as
      // scaled fp16 hi / lo channel pairs
      const int c1h = Lv.hs - 2, c1w = Lv.ws - 2;     // valid conv1 extent (ceil-mode pooling clips to it)
#pragma unroll 1
      for (int item = tid; item < P1H * P1W; item += NA_THREADS) {
      asm volatile("" ::: "memory");          // keep the weight loads inside the loop (hoisted, they spill)
      const int py = item / P1W, px = item - py * P1W;
      float patch[3][4][4];
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float2 v0 = *reinterpret_cast<const float2*>(&in_s[(ci * INH + 2 * py + r) * INP + 2 * px]);
          const float2 v1 = *reinterpret_cast<const float2*>(&in_s[(ci * INH + 2 * py + r) * INP + 2 * px + 2]);
          patch[ci][r][0] = v0.x; patch[ci][r][1] = v0.y; patch[ci][r][2] = v1.x; patch[ci][r][3] = v1.y;
        }
      float accf[4][10];
#if PNET_C1_FFMA2
      // packed fp32 pairs: FFMA2 (fma.rn.f32x2) does two channels per issued instruction, same rounding as FFMA.  On B200
      // an FFMA2 holds the FMA pipe for 3 cycles (experiments/fma_rate_probe.cu: two FMAs per 3 cycles, against one FFMA
      // per cycle) but leaves two issue slots to the other pipes and to group B, which is what this kernel is short of.
      unsigned long long acc[4][5];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int cp = 0; cp < 5; ++cp) acc[q][cp] = 0ull;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float* wr = &w_s[W1 + ((ci * 3 + ky) * 3 + kx) * 12];
            const float4 wa = *reinterpret_cast<const float4*>(wr);
            const float4 wb = *reinterpret_cast<const float4*>(wr + 4);
            const float2 wc = *reinterpret_cast<const float2*>(wr + 8);
            const unsigned long long w[5] = {pack_f32x2(wa.x, wa.y), pack_f32x2(wa.z, wa.w), pack_f32x2(wb.x, wb.y),
                                             pack_f32x2(wb.z, wb.w), pack_f32x2(wc.x, wc.y)};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float v = patch[ci][(q >> 1) + ky][(q & 1) + kx];
              const unsigned long long vv = pack_f32x2(v, v);
#pragma unroll
              for (int cp = 0; cp < 5; ++cp) ffma2(acc[q][cp], vv, w[cp]);
            }
          }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int cp = 0; cp < 5; ++cp) unpack_f32x2(acc[q][cp], accf[q][2 * cp], accf[q][2 * cp + 1]);
#else
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int co = 0; co < 10; ++co) accf[q][co] = 0.f;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float* wr = &w_s[W1 + ((ci * 3 + ky) * 3 + kx) * 12];
            const float4 wa = *reinterpret_cast<const float4*>(wr);
            const float4 wb = *reinterpret_cast<const float4*>(wr + 4);
            const float2 wc = *reinterpret_cast<const float2*>(wr + 8);
            const float w[10] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float v = patch[ci][(q >> 1) + ky][(q & 1) + kx];
#pragma unroll
              for (int co = 0; co < 10; ++co) accf[q][co] = fmaf(v, w[co], accf[q][co]);
            }
          }
#endif
      const int gy = 2 * (oy0 + py), gx = 2 * (ox0 + px);
      float m[10];
      if (p.conv1_monotone && gy + 1 < c1h && gx + 1 < c1w) {
        // interior pixel and every PReLU slope >= 0: bias add and PReLU are non-decreasing (rounding included), so
        // they commute exactly with the 2x2 max -- one bias / PReLU per channel instead of four
#pragma unroll
        for (int co = 0; co < 10; ++co) {
          const float mx = fmaxf(fmaxf(accf[0][co], accf[1][co]), fmaxf(accf[2][co], accf[3][co]));
          m[co] = prelu(mx + p.conv1_bias[co], p.conv1_slope[co]);
          range_bad |= fabsf(m[co]) > ACT_MAX;
        }
      } else {
#pragma unroll
        for (int co = 0; co < 10; ++co) {
          const float bias = p.conv1_bias[co], al = p.conv1_slope[co];
          float mm = -INFINITY;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool ok = (gy + (q >> 1) < c1h) && (gx + (q & 1) < c1w);
            const float v = prelu(accf[q][co] + bias, al);
            mm = ok ? fmaxf(mm, v) : mm;
          }
          m[co] = (mm == -INFINITY) ? 0.f : mm;
          range_bad |= fabsf(m[co]) > ACT_MAX;
        }
      }
#pragma unroll
      for (int cp = 0; cp < 5; ++cp) {
        if (TERMS == 3) {
          uint32_t hi, lo;
          split_h2(m[2 * cp] * SA, m[2 * cp + 1] * SA, hi, lo);
          p1_s[cp * P1PL + item] = hi;
          p1_s[(P1WORDS + cp) * P1PL + item] = lo;
        } else {
          const __half2 h = __floats2half2_rn(m[2 * cp] * SA, m[2 * cp + 1] * SA);
          p1_s[cp * P1PL + item] = *reinterpret_cast<const uint32_t*>(&h);
        }
      }
    }
      mbar_arrive(&p1_full[k & 1]);
#ifdef PNET_TIMING
      tc1 += clock64() - tmark;
#endif
    }
  } else if (warp_u == ISSUE_WARP) {
    // =================================================================== conv3 MMA issue (one elected thread)
    // kept off the compute warps: the issuing thread is held back by the MMA queue for as long as the MMAs run
    int k = 0;
    for (int id = blockIdx.x; id < total; id += gridDim.x, ++k) {
      mbar_wait(&c2_full, k & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      {
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
        const uint32_t a_base = __shfl_sync(0xffffffffu, smem_u32(c2_s), 0);
        const uint32_t b_base = __shfl_sync(0xffffffffu, smem_u32(w_s + W3), 0);
        const uint32_t bar_base = __shfl_sync(0xffffffffu, smem_u32(&mma_bar[(k & 1) * C3TILES]), 0);
        constexpr uint32_t IDESC32 = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);   // f32 accumulate, f16 x f16, K-major
        constexpr uint32_t IDESC64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint32_t APL = C2PLANE * 4;                                                // plane stride in bytes
        if (elect_one()) {
#pragma unroll
          for (int tile = 0; tile < C3TILES; ++tile) {
            const uint32_t d = tm + tile_col(k, tile);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t a_off = (uint32_t)(tile * 128 + (tap / 3) * C2P + (tap % 3)) * 16u;
              const uint64_t a_hi = umma_desc(a_base + a_off, APL, 128u);
              const uint64_t a_lo = umma_desc(a_base + 2u * APL + a_off, APL, 128u);
              const uint64_t b_hl = umma_desc(b_base + (uint32_t)tap * 2048u, 1024u, 128u);        // rows 0-31 w_hi, 32-63 w_lo
              if (TERMS == 1) {
                umma_f16(d, a_hi, b_hl, IDESC32, tap > 0 ? 1u : 0u);        // rows 0-31 of the B tile: w_hi
              } else if (tile_n64(tile)) {
                umma_f16(d, a_hi, b_hl, IDESC64, tap > 0 ? 1u : 0u);
                umma_f16(d, a_lo, b_hl, IDESC32, 1u);
              } else {
                const uint64_t b_lo = umma_desc(b_base + (uint32_t)tap * 2048u + 512u, 1024u, 128u);
                umma_f16(d, a_hi, b_hl, IDESC32, tap > 0 ? 1u : 0u);
                umma_f16(d, a_lo, b_hl, IDESC32, 1u);
                umma_f16(d, a_hi, b_lo, IDESC32, 1u);
              }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_base + 8u * tile) : "memory");
          }
        }
        __syncwarp();
      }
    }
  } else {
    // =================================================================== group B: conv2, heads
    const int warp = warp_u - NA_WARPS;
    // head epilogue of the tile whose conv3 went into accumulator set kk & 1 (thread = pixel: 32 channels -> bias,
    // PReLU, heads, softmax, candidate append; no cross-lane traffic)
    auto conv3_epilogue = [&](const TileRef& tr, int kk) {
      const Level& Lv = p.lv[tr.lvl];
      const int b = tr.b, lvl = tr.lvl, oy0 = tr.oy0, ox0 = tr.ox0;
      const int lg = warp & 3;                        // TMEM lane group this warp may read
      for (int tile = warp >> 2; tile < C3TILES; tile += NB_WARPS / 4) {
#ifdef PNET_TIMING
        const long long tq = clock64();
#endif
        mbar_wait(&mma_bar[(kk & 1) * C3TILES + tile], (kk >> 1) & 1);
#ifdef PNET_TIMING
        tmma += clock64() - tq;
#endif
        if (tile * 128 + lg * 32 >= TOY * C2P) break;   // flat pixels >= TOY * C2P do not exist (the half tile of 16 x 32)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + tile_col(kk, tile);
      float h[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 16) {
        float dh[16];
        tmem_ld16(taddr + c0, dh);
        if (TERMS == 3 && tile_n64(tile)) {             // N = 64 tiles: add the a_hi x w_lo column block
          float dl[16];
          tmem_ld16(taddr + 32 + c0, dl);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) dh[j] += dl[j];
        } else {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float* hw = p.head[c0 + j];                                  // compile-time index: c[0x0][..] operands, no LDS
          const float v = prelu(fmaf(dh[j], p.inv3, hw[6]), hw[7]);
#pragma unroll
          for (int q = 0; q < 6; ++q) h[q] = fmaf(v, hw[q], h[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 6; ++q) h[q] += p.head_bias[q];
      const int n = tile * 128 + lg * 32 + lane;
      const int row = n / C2P, col = n - row * C2P;
      const int oy = oy0 + row, ox = ox0 + col;
      if (row < TOY && col < TOX && oy < Lv.oh && ox < Lv.ow) {
        // softmax over (h0, h1), class 1 -- same form as ATen's softmax (subtract max, exp, normalise)
        const float mx = fmaxf(h[0], h[1]);
        const float e0 = expf(h[0] - mx), e1 = expf(h[1] - mx);
        const float prob = __fdiv_rn(e1, e0 + e1);
        const size_t cell = (size_t)oy * Lv.ow + ox;
        if (Lv.prob) {
          const size_t plane = (size_t)Lv.oh * Lv.ow;
          Lv.prob[(size_t)b * plane + cell] = prob;
          float* rg = Lv.reg + (size_t)b * 4 * plane + cell;
          rg[0] = h[2]; rg[plane] = h[3]; rg[2 * plane] = h[4]; rg[3 * plane] = h[5];
        }
        if (p.screen) {
          if (prob >= p.thr) {
            const int slot = atomicAdd(p.screen_cnt, 1);
            if (slot < p.screen_cap) {
              p.screen[slot] = ScreenEntry{b * p.n_levels + lvl, (int)cell};
            } else if (p.capflag) {
              p.capflag->overflow = 1; p.capflag->stage = 1; p.capflag->frame = b;
              p.capflag->count = slot + 1; p.capflag->capacity = p.screen_cap;
            }
          }
        } else if (Lv.cand && prob >= p.thr) {
          // generateBoundingBox: q1 = floor((2*c + 1)/scale), q2 = floor((2*c + 12)/scale)  (fp32, true division)
          const int slotbase = b * p.n_levels + lvl;
          const int slot = atomicAdd(&Lv.cnt[slotbase], 1);
          if (slot < p.cap) {
            Cand cd;
            cd.x1 = floorf(__fdiv_rn((float)(2 * ox + 1), Lv.scale));
            cd.y1 = floorf(__fdiv_rn((float)(2 * oy + 1), Lv.scale));
            cd.x2 = floorf(__fdiv_rn((float)(2 * ox + 12), Lv.scale));
            cd.y2 = floorf(__fdiv_rn((float)(2 * oy + 12), Lv.scale));
            cd.score = prob;
            cd.r0 = h[2]; cd.r1 = h[3]; cd.r2 = h[4]; cd.r3 = h[5];
            cd.key = (uint32_t)cell;
            Lv.cand[(size_t)slotbase * p.cap + slot] = cd;
          } else if (p.capflag) {
            p.capflag->overflow = 1; p.capflag->stage = 1; p.capflag->frame = b;
            p.capflag->count = slot + 1; p.capflag->capacity = p.cap;
          }
        }
      }
      }
    };
    TileRef prev{};
    int k = 0;
    for (int id = blockIdx.x; id < total; id += gridDim.x, ++k) {
#ifdef PNET_TIMING
      tmark = clock64();
#endif
      mbar_wait(&p1_full[k & 1], (k >> 1) & 1);
      if (k >= 1) mbar_wait(&mma_bar[((k - 1) & 1) * C3TILES + C3TILES - 1], ((k - 1) >> 1) & 1);      // conv3 of tile k-1 has consumed c2
#ifdef PNET_TIMING
      tw += clock64() - tmark; tmark = clock64();
#endif
      const TileRef tr = decode_tile(p, id);
      const uint32_t* p1_s = p1_base + (k & 1) * SM_P1;
      // ---- conv2 (10->16, 3x3) + PReLU on the tensor pipe (mma.sync).
      // Flat implicit GEMM: output pixel n = row * 36 + col reads input pixels n + ky * 36 + kx, so an M tile is any
      // two 8-pixel runs of the flat index (41 tiles cover the 18 x 36 tile, columns 34/35 are computed but unused).
      // K = 90 as 45 channel pairs P = ky*15 + kx*5 + cp (+3 zero-weight pads) = 6 k16 steps; the word offset of pair
      // P relative to the pixel comes from T2.  Fragment coordinates (PTX m16n8k16): g = lane/4, t = lane%4
      //   A: a0 (px g, pair t)  a1 (px g+8, pair t)  a2 (px g, pair t+4)  a3 (px g+8, pair t+4)
      //   B: b0 (pair t, n g)   b1 (pair t+4, n g)        C: c0,c1 (px g, n 2t, 2t+1)  c2,c3 (px g+8, ..)
      {
        const int* tab = reinterpret_cast<const int*>(w_s + T2);
    const uint32_t* p1h = p1_s;
    const uint32_t* p1l = p1_s + P1WORDS * P1PL;
    const float inv = w_s[SC + 0];
#pragma unroll 1
    for (int pass = 0; pass < C2PASSES; ++pass) {
      // M tiles c2_first(warp) + 3 pass + q, q < 3.  The 41 tiles are not dealt evenly: warps 0-1 also carry three conv3
      // head epilogues per tile (M tiles 0, 2 and the half tile 4) where the other warps carry two, and the slowest warp
      // of group B sets the pace of the whole CTA, so they get 4 conv2 tiles and the others 6 or 5.
      const int mt0 = c2_first(warp) + 3 * pass;
      const int nq = max(0, min(3, c2_count(warp) - 3 * pass));      // live M tiles of this warp in this pass (warp uniform)
      if (nq == 0) break;
      int base[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) base[q] = min(mt0 + q, C2TILES - 1) * 16 + g;
      float acc[3][2][4];
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[q][j][e] = 0.f;
      // Software pipelined over the six k-steps: the A fragments (and weights) of step s + 1 are requested before the MMAs
      // of step s are issued, so the shared-memory latency of a step hides under the previous step's tensor work
      // (PNET_C2_PIPE = 0 keeps the straight load-then-multiply order).
      uint32_t ah[2][3][4], al[2][3][4];
      uint4 bhv[2], blv[2];
      auto load_step = [&](int s, int buf) {
        const int o0 = tab[8 * s + t], o1 = tab[8 * s + t + 4];
        bhv[buf] = *reinterpret_cast<const uint4*>(&wu[W2 + ((2 * s + 0) * 32 + lane) * 4]);
        blv[buf] = make_uint4(0u, 0u, 0u, 0u);
        if (TERMS == 3) blv[buf] = *reinterpret_cast<const uint4*>(&wu[W2 + ((2 * s + 1) * 32 + lane) * 4]);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          if (q >= nq) continue;               // warp uniform
          ah[buf][q][0] = p1h[base[q] + o0];
          ah[buf][q][1] = p1h[base[q] + 8 + o0];
          ah[buf][q][2] = p1h[base[q] + o1];
          ah[buf][q][3] = p1h[base[q] + 8 + o1];
          if (TERMS == 3) {
            al[buf][q][0] = p1l[base[q] + o0];
            al[buf][q][1] = p1l[base[q] + 8 + o0];
            al[buf][q][2] = p1l[base[q] + o1];
            al[buf][q][3] = p1l[base[q] + 8 + o1];
          }
        }
      };
      if (PNET_C2_PIPE) load_step(0, 0);
#pragma unroll
      for (int s = 0; s < 6; ++s) {
        const int cb = PNET_C2_PIPE ? (s & 1) : 0;
        if (PNET_C2_PIPE) { if (s + 1 < 6) load_step(s + 1, cb ^ 1); }
        else load_step(s, 0);
        const uint4 bh = bhv[cb], bl = blv[cb];
        // the three terms of one accumulator are dependent: issue the 6 independent accumulators between them
        if (TERMS == 3) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (q < nq) {
              mma_f16(acc[q][0], al[cb][q][0], al[cb][q][1], al[cb][q][2], al[cb][q][3], bh.x, bh.y);
              mma_f16(acc[q][1], al[cb][q][0], al[cb][q][1], al[cb][q][2], al[cb][q][3], bh.z, bh.w);
            }
          }
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (q < nq) {
              mma_f16(acc[q][0], ah[cb][q][0], ah[cb][q][1], ah[cb][q][2], ah[cb][q][3], bl.x, bl.y);
              mma_f16(acc[q][1], ah[cb][q][0], ah[cb][q][1], ah[cb][q][2], ah[cb][q][3], bl.z, bl.w);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          if (q < nq) {
            mma_f16(acc[q][0], ah[cb][q][0], ah[cb][q][1], ah[cb][q][2], ah[cb][q][3], bh.x, bh.y);
            mma_f16(acc[q][1], ah[cb][q][0], ah[cb][q][1], ah[cb][q][2], ah[cb][q][3], bh.z, bh.w);
          }
        }
      }
      // epilogue: unscale, bias, PReLU, split; lane holds channels (2t, 2t+1) + 8 j of pixels g and g + 8.
      const float bias0 = w_s[B2 + 2 * t], bias1 = w_s[B2 + 2 * t + 1], bias2 = w_s[B2 + 8 + 2 * t], bias3 = w_s[B2 + 9 + 2 * t];
      const float al0 = w_s[A2 + 2 * t], al1 = w_s[A2 + 2 * t + 1], al2 = w_s[A2 + 8 + 2 * t], al3 = w_s[A2 + 9 + 2 * t];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int mt = mt0 + q;
        if (q >= nq) break;                       // warp uniform
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int n = mt * 16 + 8 * h + g;
          if (n < C2PX) {
            const float v0 = prelu(fmaf(acc[q][0][2 * h], inv, bias0), al0);
            const float v1 = prelu(fmaf(acc[q][0][2 * h + 1], inv, bias1), al1);
            const float v2 = prelu(fmaf(acc[q][1][2 * h], inv, bias2), al2);
            const float v3 = prelu(fmaf(acc[q][1][2 * h + 1], inv, bias3), al3);
            range_bad |= fmaxf(fmaxf(fabsf(v0), fabsf(v1)), fmaxf(fabsf(v2), fabsf(v3))) > ACT_MAX;
            if (TERMS == 3) {
              uint32_t h0, l0, h1, l1;
              split_h2(v0 * SA, v1 * SA, h0, l0);
              split_h2(v2 * SA, v3 * SA, h1, l1);
              c2_s[0 * C2PLANE + n * 4 + t] = h0;      // hi, channels 0-7   (a warp writes 32 consecutive words)
              c2_s[1 * C2PLANE + n * 4 + t] = h1;      // hi, channels 8-15
              c2_s[2 * C2PLANE + n * 4 + t] = l0;
              c2_s[3 * C2PLANE + n * 4 + t] = l1;
            } else {
              const __half2 h0 = __floats2half2_rn(v0 * SA, v1 * SA), h1 = __floats2half2_rn(v2 * SA, v3 * SA);
              c2_s[0 * C2PLANE + n * 4 + t] = *reinterpret_cast<const uint32_t*>(&h0);
              c2_s[1 * C2PLANE + n * 4 + t] = *reinterpret_cast<const uint32_t*>(&h1);
            }
          }
        }
      }
    }
      }
      mbar_arrive(&p1_empty[k & 1]);
#ifdef PNET_TIMING
      tc2 += clock64() - tmark; tmark = clock64();
#endif

      // ---- conv3 (16->32, 3x3) as tcgen05 implicit GEMMs + PReLU + heads.
      // Flat indexing again: output pixel n = row * 36 + col reads conv2 pixels n + ky * 36 + kx, so the A operand
      // of filter tap (ky, kx) is the same shared-memory image with its start address moved by (ky * 36 + kx) * 16
      // bytes -- no im2col copy.  One M tile = 128 flat pixels (5 tiles cover 16 rows x pitch 36), one MMA = one tap
      // x 16 channels (K = 16).  The 3-term split a_hi w_hi + a_hi w_lo + a_lo w_hi costs two MMAs per tap on tiles
      // 0-2 (B = [w_hi | w_lo] as N = 64, then a_lo x w_hi as N = 32 into the first 32 columns; the epilogue adds the
      // two column blocks) and three N = 32 MMAs per tap on tiles 3-4, so one tile set needs 256 TMEM columns and the
      // accumulators are double buffered: the MMAs of tile k run while the epilogue of tile k-1 is computed.
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // c2 planes -> visible to the UMMA proxy
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // this thread's earlier epilogue reads of D are done
      mbar_arrive(&c2_full);                                            // -> warp 16 issues the conv3 MMAs of this tile
      if (DEFER_EPILOGUE) {
        if (k >= 1) conv3_epilogue(prev, k - 1);
        prev = tr;
      } else {
        conv3_epilogue(tr, k);
      }
#ifdef PNET_TIMING
      tc3 += clock64() - tmark;
#endif
    }
    if (DEFER_EPILOGUE && k >= 1) conv3_epilogue(prev, k - 1);
  }
  if (range_bad && p.capflag) {
    p.capflag->overflow = 1; p.capflag->stage = 5; p.capflag->frame = 0;
    p.capflag->count = 0; p.capflag->capacity = (int)ACT_MAX;
  }
#ifdef PNET_TIMING
  if (tid == 0) {
    atomicAdd(&g_pnet_phase[0], (unsigned long long)tw); atomicAdd(&g_pnet_phase[1], (unsigned long long)tc1);
  }
  if (tid == NA_THREADS) {
    atomicAdd(&g_pnet_phase[2], (unsigned long long)tw); atomicAdd(&g_pnet_phase[3], (unsigned long long)tc2);
    atomicAdd(&g_pnet_phase[4], (unsigned long long)tc3);
    atomicAdd(&g_pnet_phase[6], (unsigned long long)tmma);
    atomicAdd(&g_pnet_phase[5], (unsigned long long)((total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x));
  }
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp_u == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

}  // namespace pnet
#ifdef PNET_TIMING
extern "C" void trl_debug_pnet_timing(unsigned long long* out) {
  cudaMemcpyFromSymbol(out, pnet::g_pnet_phase, sizeof(unsigned long long) * 8);
  unsigned long long z[8] = {0};
  cudaMemcpyToSymbol(pnet::g_pnet_phase, z, sizeof(z));
}
#endif

// ---- host-side packing

// conv2's K ordering: channel pair P = ky*15 + kx*5 + cp holds channels (2cp, 2cp+1) of tap (ky, kx); P >= 45 are
// zero-weight pads that read pair P - 45 (any finite word).
static void conv2_pair(int P, int* ky, int* kx, int* cp, bool* used) {
  *used = P < 45;
  const int q = *used ? P : P - 45;
  *ky = q / 15; *kx = (q % 15) / 5; *cp = q % 5;
}

// power-of-two scale that brings max|w| just below 2^14 (fp16 max is 65504): keeps w's lo part out of the subnormals
static float weight_scale(const float* w, int n) {
  float m = 0.f;
  for (int i = 0; i < n; ++i) m = fmaxf(m, fabsf(w[i]));
  if (!(m > 0.f) || !isfinite(m)) return 1.f;
  int e;
  frexpf(m, &e);                        // m = f * 2^e, f in [0.5, 1)
  return ldexpf(1.f, 14 - e);
}

static uint32_t pack_h2(float x0, float x1) {
  const __half a = __float2half_rn(x0), b = __float2half_rn(x1);
  uint16_t ua, ub;
  memcpy(&ua, &a, 2);
  memcpy(&ub, &b, 2);
  return (uint32_t)ua | ((uint32_t)ub << 16);
}
// scaled pair -> hi word and lo word
static void split_pair(float x0, float x1, float scale, uint32_t* hi, uint32_t* lo) {
  const float s0 = x0 * scale, s1 = x1 * scale;         // power-of-two scale: exact
  const float h0 = __half2float(__float2half_rn(s0)), h1 = __half2float(__float2half_rn(s1));
  *hi = pack_h2(s0, s1);
  *lo = pack_h2(s0 - h0, s1 - h1);
}

// upstream layouts -> packed shared-memory image
int pnet_pack_weights(trl_ctx* c, const float* h, size_t len) {
  using namespace pnet;
  if (len != 6632) TRL_FAIL(c, TRL_E_INVALID, "pnet blob has %zu floats, expected 6632", len);
  std::vector<float> pk(WTOTAL, 0.f);
  uint32_t* pu = reinterpret_cast<uint32_t*>(pk.data());
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
  for (int co = 0; co < 10; ++co)
    for (int k = 0; k < 27; ++k) pk[W1 + k * 12 + co] = w1[co * 27 + k];
  for (int co = 0; co < 10; ++co) { pk[B1 + co] = b1[co]; pk[A1 + co] = a1[co]; }
  const float s2 = weight_scale(w2, 1440), s3 = weight_scale(w3, 4608);
  // conv2 B fragments (m16n8k16: b0 = pair t, b1 = pair t + 4 of the k-step; n = g + 8 j) + the pair offset table
  for (int s = 0; s < 6; ++s)
    for (int lane = 0; lane < 32; ++lane) {
      const int g = lane >> 2, t = lane & 3;
      for (int i = 0; i < 2; ++i) {        // b0 / b1
        int ky, kx, cp;
        bool used;
        conv2_pair(8 * s + 4 * i + t, &ky, &kx, &cp, &used);
        for (int j = 0; j < 2; ++j) {
          const int co = 8 * j + g;
          const float x0 = used ? w2[co * 90 + (2 * cp) * 9 + ky * 3 + kx] : 0.f;
          const float x1 = used ? w2[co * 90 + (2 * cp + 1) * 9 + ky * 3 + kx] : 0.f;
          uint32_t hi, lo;
          split_pair(x0, x1, s2, &hi, &lo);
          pu[W2 + ((2 * s + 0) * 32 + lane) * 4 + 2 * j + i] = hi;
          pu[W2 + ((2 * s + 1) * 32 + lane) * 4 + 2 * j + i] = lo;
        }
        if (g == 0) {
          const int off = cp * P1PL + ky * P1W + kx;
          memcpy(&pk[T2 + 8 * s + 4 * i + t], &off, sizeof(int));
        }
      }
    }
  for (int co = 0; co < 16; ++co) { pk[B2 + co] = b2[co]; pk[A2 + co] = a2[co]; }
  // conv3 UMMA B operand: [tap][k-half][row: w_hi of channel co = row (0-31), w_lo of co = row - 32][8 input channels]
  for (int tap = 0; tap < 9; ++tap)
    for (int kh = 0; kh < 2; ++kh)
      for (int co = 0; co < 32; ++co)
        for (int cp = 0; cp < 4; ++cp) {
          const int ci = 8 * kh + 2 * cp;
          uint32_t hi, lo;
          split_pair(w3[co * 144 + ci * 9 + tap], w3[co * 144 + (ci + 1) * 9 + tap], s3, &hi, &lo);
          pu[W3 + ((tap * 2 + kh) * 64 + co) * 4 + cp] = hi;
          pu[W3 + ((tap * 2 + kh) * 64 + 32 + co) * 4 + cp] = lo;
        }
  // conv3 epilogue + head constants travel in the kernel parameters (constant bank operands)
  for (int ci = 0; ci < 32; ++ci) {
    float* hw = c->h_pnet_head + ci * 8;
    hw[0] = w41[0 * 32 + ci];
    hw[1] = w41[1 * 32 + ci];
    for (int j = 0; j < 4; ++j) hw[2 + j] = w42[j * 32 + ci];
    hw[6] = b3[ci];
    hw[7] = a3[ci];
  }
  c->h_pnet_head[256] = b41[0]; c->h_pnet_head[257] = b41[1];
  for (int j = 0; j < 4; ++j) c->h_pnet_head[258 + j] = b42[j];
  c->h_pnet_head[264] = 1.f / (SA * s3);
  for (int co = 0; co < 10; ++co) { c->h_pnet_head[265 + co] = b1[co]; c->h_pnet_head[275 + co] = a1[co]; }
  pk[SC + 0] = 1.f / (SA * s2);
  TRL_CUDA(c, cudaMalloc(&c->d_pnet_packed, WTOTAL * sizeof(float)));
  TRL_CUDA(c, cudaMemcpy(c->d_pnet_packed, pk.data(), WTOTAL * sizeof(float), cudaMemcpyHostToDevice));
  TRL_CUDA(c, cudaFuncSetAttribute(pnet::pnet_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  TRL_CUDA(c, cudaFuncSetAttribute(pnet::pnet_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  return TRL_OK;
}

static void fill_head(const trl_ctx* c, pnet::Params* p) {
  memcpy(p->head, c->h_pnet_head, sizeof(p->head));
  memset(p->head_bias, 0, sizeof(p->head_bias));
  memcpy(p->head_bias, c->h_pnet_head + 256, 6 * sizeof(float));
  p->inv3 = c->h_pnet_head[264];
  memcpy(p->conv1_bias, c->h_pnet_head + 265, 10 * sizeof(float));
  memcpy(p->conv1_slope, c->h_pnet_head + 275, 10 * sizeof(float));
  p->conv1_monotone = 1;
  for (int co = 0; co < 10; ++co) if (!(p->conv1_slope[co] >= 0.f)) p->conv1_monotone = 0;
}

// persistent grid: one CTA per SM (or per tile when there are fewer tiles)
static int grid_for(const trl_ctx* c, int blocks, int B) {
  const long long total = (long long)blocks * B;
  return (int)(total < c->num_sms ? total : c->num_sms);
}

int launch_pnet_maps(trl_ctx* c, const float* d_in, int B, int hs, int ws, float* d_prob, float* d_reg, cudaStream_t s) {
  using namespace pnet;
  Params p{};
  p.n_levels = 1;
  Level& L = p.lv[0];
  L.in = d_in; L.hs = hs; L.ws = ws; L.pitch = ws;
  L.oh = (hs - 2 + 1) / 2 - 4; L.ow = (ws - 2 + 1) / 2 - 4;
  if (L.oh <= 0 || L.ow <= 0) TRL_FAIL(c, TRL_E_INVALID, "pnet input %dx%d too small", hs, ws);
  L.tiles_x = ceil_div(L.ow, TOX);
  L.tiles = L.tiles_x * ceil_div(L.oh, TOY);
  L.scale = 1.f; L.prob = d_prob; L.reg = d_reg; L.cand = nullptr; L.cnt = nullptr;
  p.blk_start[0] = 0; p.blk_start[1] = L.tiles;
  p.blocks = L.tiles; p.n_frames = B;
  fill_head(c, &p);
  // stage entry point: arbitrary row pitch -- TMA staging when the rows are 16-byte aligned, else cp.async
  p.use_tma = ((ws & 3) == 0 && (reinterpret_cast<uintptr_t>(d_in) & 15) == 0) ? 1 : 0;
  if (p.use_tma) {
    const unsigned long long dims[3] = {(unsigned long long)ws, (unsigned long long)hs, (unsigned long long)3 * B};
    const unsigned long long strides[2] = {(unsigned long long)ws * 4, (unsigned long long)hs * ws * 4};
    const unsigned box[3] = {(unsigned)INP, (unsigned)INH, 3u};
    int rc = tma_encode_tiled_f32(c, &p.tmap[0], d_in, 3, dims, strides, box);
    if (rc != TRL_OK) return rc;
  }
  p.thr = 2.f; p.cap = 0; p.capflag = c->d_cap;
  if (B == 0) return TRL_OK;
  if (c->cfg.pnet_precision == 1) pnet_kernel<1><<<grid_for(c, L.tiles, B), NTHREADS, SMEM_BYTES, s>>>(c->d_pnet_packed, p);
  else pnet_kernel<3><<<grid_for(c, L.tiles, B), NTHREADS, SMEM_BYTES, s>>>(c->d_pnet_packed, p);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

static int launch_pnet_levels(trl_ctx* c, const float* d_pyr, int B, const PyramidGeom& g, float thr, Cand* d_cand,
                              int* d_cnt, int cap, ScreenEntry* d_screen, int* d_screen_cnt, int screen_cap, cudaStream_t s);

int launch_pnet_candidates(trl_ctx* c, const float* d_pyr, int B, const PyramidGeom& g, float thr, Cand* d_cand,
                           int* d_cnt, int cap, cudaStream_t s) {
  return launch_pnet_levels(c, d_pyr, B, g, thr, d_cand, d_cnt, cap, nullptr, nullptr, 0, s);
}

// screening launch of the hybrid P-Net on the fp32 pyramid: the single-pass variant of this kernel appends every cell with
// prob >= thr_lo to the screen list (pnet_precision = 2; pnet2.cu is the faster screen, pnet_precision = 3)
int launch_pnet_screen_v1(trl_ctx* c, const float* d_pyr, int B, const PyramidGeom& g, float thr_lo, ScreenEntry* d_screen,
                          int* d_screen_cnt, int screen_cap, cudaStream_t s) {
  return launch_pnet_levels(c, d_pyr, B, g, thr_lo, nullptr, nullptr, 0, d_screen, d_screen_cnt, screen_cap, s);
}

static int launch_pnet_levels(trl_ctx* c, const float* d_pyr, int B, const PyramidGeom& g, float thr, Cand* d_cand,
                              int* d_cnt, int cap, ScreenEntry* d_screen, int* d_screen_cnt, int screen_cap, cudaStream_t s) {
  using namespace pnet;
  Params p{};
  p.screen = d_screen; p.screen_cnt = d_screen_cnt; p.screen_cap = screen_cap;
  p.n_levels = g.n;
  int blocks = 0;
  for (int k = 0; k < g.n; ++k) {
    Level& L = p.lv[k];
    L.in = d_pyr + g.off[k] * B;
    L.hs = g.hs[k]; L.ws = g.ws[k]; L.pitch = g.pitch[k]; L.oh = g.oh[k]; L.ow = g.ow[k];
    L.tiles_x = ceil_div(L.ow, TOX);
    L.tiles = L.tiles_x * ceil_div(L.oh, TOY);
    L.scale = g.scale_f[k];
    L.prob = nullptr; L.reg = nullptr; L.cand = d_cand; L.cnt = d_cnt;
    p.blk_start[k] = blocks;
    blocks += L.tiles;
  }
  p.blk_start[g.n] = blocks;
  p.thr = thr; p.cap = cap; p.capflag = c->d_cap;
  p.blocks = blocks; p.n_frames = B;
  fill_head(c, &p);
  p.use_tma = 1;
  for (int k = 0; k < g.n && p.use_tma; ++k) {
    const Level& L = p.lv[k];
    if ((L.pitch & 3) != 0 || (reinterpret_cast<uintptr_t>(L.in) & 15) != 0) { p.use_tma = 0; break; }
    const unsigned long long dims[3] = {(unsigned long long)L.ws, (unsigned long long)L.hs, (unsigned long long)3 * B};
    const unsigned long long strides[2] = {(unsigned long long)L.pitch * 4, (unsigned long long)L.hs * L.pitch * 4};
    const unsigned box[3] = {(unsigned)INP, (unsigned)INH, 3u};
    int rc = tma_encode_tiled_f32(c, &p.tmap[k], L.in, 3, dims, strides, box);
    if (rc != TRL_OK) return rc;
  }
  if (blocks == 0 || B == 0) return TRL_OK;
  if (c->cfg.pnet_precision == 1 || d_screen) pnet_kernel<1><<<grid_for(c, blocks, B), NTHREADS, SMEM_BYTES, s>>>(c->d_pnet_packed, p);
  else pnet_kernel<3><<<grid_for(c, blocks, B), NTHREADS, SMEM_BYTES, s>>>(c->d_pnet_packed, p);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}
