// K2 + K3: fused P-Net (conv1+PReLU+maxpool, conv2+PReLU, conv3+PReLU, conv4_1 -> softmax, conv4_2) and
// generateBoundingBox, fp32 on the FMA pipe.  One CTA computes a 16x32 tile of output cells; every intermediate
// lives in shared memory, the only HBM traffic is the input tile and the (rare) candidates.
//
// upstream: models/mtcnn.py PNet.forward, models/utils/detect_face.py generateBoundingBox.
//
// conv1/conv2 run on the FP32 FMA pipe: thread tiles are register blocked (4 px x 4 channels) with weights broadcast
// from shared memory by LDS.128.  conv3 (63 % of the FLOPs) runs on the tensor pipe as an implicit GEMM
// (M = 16 pixels of one output row, N = 8 channels, K = 8 = one filter tap x 8 input channels) with
// mma.sync.m16n8k8 TF32 and the 3xTF32 split (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, fp32 accumulate), which keeps
// fp32-level accuracy (measured max |err| 1e-5 on |sum| ~ 4.5 over K = 144, experiments/umma_probe.cu) so the
// cascade sees the same candidates as the fp32 reference.  tcgen05 was measured and rejected for this layer: with
// N = 32 output channels an SS-mode UMMA is bound by re-reading the A tile from shared memory (~73 cycles per
// M128 x N32 x K8 step, experiments/umma_probe.cu), no faster than mma.sync once the 3x split is paid.
// The two CTAs of an SM overlap one CTA's FMA-pipe stages with the other's tensor-pipe stage.
// FLOP roof, not HBM, binds this kernel.
#include "common.cuh"
#include <string.h>

namespace pnet {

constexpr int TOY = 16, TOX = 32;            // output cells per CTA
constexpr int P1H = TOY + 4, P1W = TOX + 4;  // pooled conv1 tile (20 x 36)
constexpr int P1P = 40;                      // pitch (conv2's last 8-pixel segment over-reads to col 41 = next row)
constexpr int P1PLANE = P1H * P1P + 8;       // 808 = 8 mod 32: four channels of one tap land in four bank octets
constexpr int C2H = TOY + 2, C2W = 36;       // conv2 tile rows x stored cols (34 valid + 2 slack)
constexpr int C2SEG = 5;                     // conv2 row = five 8-pixel segments (40 computed columns)
constexpr int C2P = 36;
constexpr int INH = 2 * TOY + 10, INW = 2 * TOX + 10;   // 42 x 74 input tile
constexpr int INP = 76;

// packed weights (floats), k = (ci*3+ky)*3+kx
constexpr int W1 = 0;                 // [27][12]
constexpr int B1 = W1 + 27 * 12;      // [12]
constexpr int A1 = B1 + 12;           // [12]
constexpr int W2 = A1 + 12;           // mma B fragments: [12 k-steps][32 lanes][b0 n0, b0 n1, b1 n0, b1 n1]
constexpr int T2 = W2 + 12 * 32 * 4;  // int[96]: pooled-tile offset of k index 8s + t (+4), see conv2_k()
constexpr int B2 = T2 + 96;
constexpr int A2 = B2 + 16;
constexpr int W3 = A2 + 16;           // mma B fragments: [18 k-steps][2][32 lanes][4 n-tiles], k = tap*16 + ci
constexpr int B3 = W3 + 144 * 32;
constexpr int A3 = B3 + 32;
constexpr int WH = A3 + 32;           // [32][8]: cols 0,1 conv4_1; 2..5 conv4_2
constexpr int BH = WH + 32 * 8;       // [8]
constexpr int WTOTAL = BH + 8;        // 6948 floats
static_assert(WTOTAL % 4 == 0, "float4 copy");

constexpr int SM_IN = 3 * INH * INP;          // 9576
constexpr int SM_C2 = 16 * C2H * C2P;         // 10368   (aliases the input tile)
constexpr int SM_A = SM_C2 > SM_IN ? SM_C2 : SM_IN;
constexpr int SM_P1 = 10 * P1PLANE;           // 8080
constexpr int SMEM_FLOATS = WTOTAL + SM_A + SM_P1;
constexpr int SMEM_BYTES = SMEM_FLOATS * 4;   // ~100 KB -> 2 CTAs / SM

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
  int n_levels;
  int blk_start[TRL_MAX_SCALES + 1];
  Level lv[TRL_MAX_SCALES];
  float thr;
  int cap;
  CapFlag* capflag;
};

#ifdef PNET_TIMING
__device__ unsigned long long g_pnet_phase[8];
#endif
__device__ __forceinline__ float prelu(float v, float a) { return v > 0.f ? v : v * a; }

// x = hi + lo with hi exactly representable in TF32 (round to nearest, ties away; two integer ops instead of the
// five-instruction NaN-safe expansion of cvt.rna.tf32.f32 -- activations and weights are finite).  The tensor core
// drops lo's low 13 bits, an error of 2^-21 |x|.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 2) pnet_kernel(const float* __restrict__ wpacked, const __grid_constant__ Params p) {
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;
  float* a_s = smem + WTOTAL;          // input tile, later conv2 output
  float* p1_s = a_s + SM_A;            // pooled conv1, later head partials
  const int tid = threadIdx.x;
#ifdef PNET_TIMING
  long long tph[5]; tph[0] = clock64();
#endif

  int lvl = 0;
  while (lvl + 1 < p.n_levels && (int)blockIdx.x >= p.blk_start[lvl + 1]) ++lvl;
  const Level& L = p.lv[lvl];
  const int b = blockIdx.y;
  const int tile = (int)blockIdx.x - p.blk_start[lvl];
  const int ty = tile / L.tiles_x, tx = tile - ty * L.tiles_x;
  const int oy0 = ty * TOY, ox0 = tx * TOX;
  const int hs = L.hs, ws = L.ws;

  // ---- stage weights and the input tile with cp.async (LDGSTS): every copy of a thread is in flight at once, so
  // the tile costs one L2 round trip instead of one per unrolled load group; out-of-image elements are zero filled
  // (src-size 0).
  {
    const uint32_t w_dst = (uint32_t)__cvta_generic_to_shared(w_s);
    for (int i = tid; i < WTOTAL / 4; i += 256)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(w_dst + 16u * i), "l"(wpacked + 4 * i) : "memory");
    const int pitch = L.pitch;
    const float* src = L.in + (size_t)b * 3 * hs * pitch;
    const int iy0 = 2 * oy0, ix0 = 2 * ox0;
    const uint32_t a_dst = (uint32_t)__cvta_generic_to_shared(a_s);
    if ((pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(L.in) & 15) == 0) {
      // 16-byte rows (the cascade's padded pyramid): 19 chunks per tile row, partial chunks zero filled by src-size
      constexpr int CH = INP / 4;
      for (int i = tid; i < 3 * INH * CH; i += 256) {
        const int rr = i / CH, j = i - rr * CH;          // rr = ci * INH + r
        const int ci = rr / INH, r = rr - ci * INH;
        const int gy = iy0 + r, gx = ix0 + 4 * j;
        const int nb = gy < hs ? min(max(ws - gx, 0), 4) * 4 : 0;
        const float* gp = nb ? src + ((size_t)ci * hs + gy) * pitch + gx : src;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                     ::"r"(a_dst + 16u * (uint32_t)i), "l"(gp), "r"(nb) : "memory");
      }
    } else {
      for (int i = tid; i < 3 * INH * INW; i += 256) {
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
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
#ifdef PNET_TIMING
  tph[1] = clock64();
#endif

  // ---- conv1 (3->10, 3x3) + PReLU + maxpool(2,2,ceil): one pooled pixel x 10 channels per item
  {
    const int c1h = hs - 2, c1w = ws - 2;     // valid conv1 extent (ceil-mode pooling clips to it)
    for (int item = tid; item < P1H * P1W; item += 256) {
      const int py = item / P1W, px = item - py * P1W;
      float patch[3][4][4];
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float2 v0 = *reinterpret_cast<const float2*>(&a_s[(ci * INH + 2 * py + r) * INP + 2 * px]);
          const float2 v1 = *reinterpret_cast<const float2*>(&a_s[(ci * INH + 2 * py + r) * INP + 2 * px + 2]);
          patch[ci][r][0] = v0.x; patch[ci][r][1] = v0.y; patch[ci][r][2] = v1.x; patch[ci][r][3] = v1.y;
        }
      float acc[4][12];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int co = 0; co < 12; ++co) acc[q][co] = 0.f;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float* wr = &w_s[W1 + ((ci * 3 + ky) * 3 + kx) * 12];
            const float4 wa = *reinterpret_cast<const float4*>(wr);
            const float4 wb = *reinterpret_cast<const float4*>(wr + 4);
            const float4 wc = *reinterpret_cast<const float4*>(wr + 8);
            const float w[12] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y, wc.z, wc.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float v = patch[ci][(q >> 1) + ky][(q & 1) + kx];
#pragma unroll
              for (int co = 0; co < 10; ++co) acc[q][co] = fmaf(v, w[co], acc[q][co]);
            }
          }
      const int gy = 2 * (oy0 + py), gx = 2 * (ox0 + px);
#pragma unroll
      for (int co = 0; co < 10; ++co) {
        const float bias = w_s[B1 + co], al = w_s[A1 + co];
        float m = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool ok = (gy + (q >> 1) < c1h) && (gx + (q & 1) < c1w);
          const float v = prelu(acc[q][co] + bias, al);
          m = ok ? fmaxf(m, v) : m;
        }
        p1_s[co * P1PLANE + py * P1P + px] = (m == -INFINITY) ? 0.f : m;
      }
    }
    // slack columns read by conv2's last pixel group
    for (int i = tid; i < 10 * P1H * (P1P - P1W); i += 256) {
      const int co = i / (P1H * (P1P - P1W));
      const int r = (i / (P1P - P1W)) % P1H;
      const int cx = P1W + i % (P1P - P1W);
      p1_s[co * P1PLANE + r * P1P + cx] = 0.f;
    }
  }
  __syncthreads();
#ifdef PNET_TIMING
  tph[2] = clock64();
#endif

  // ---- conv2 (10->16, 3x3) + PReLU on the tensor pipe (3xTF32), output over the dead input tile.
  // M tile = two 8-pixel row segments (18 rows x 5 segments = 45 tiles), N = 2 x 8 channels, K = 96 (90 used):
  // k index 8s + t (+4) -> (ci, ky, kx) by conv2_k(): the four k of one A load share kx and differ in (ci + ky) mod 4,
  // so with planes 808 floats apart they read four different bank octets (conflict free); offsets come from T2.
  {
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int* tab = reinterpret_cast<const int*>(w_s + T2);
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int mt0 = (pass * 8 + warp) * 3;
      if (mt0 >= 45) break;                       // warp uniform
      int pa[3][2], row[3][2], col[3][2];
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int seg = 2 * (mt0 + q) + h;
          row[q][h] = seg / C2SEG;
          col[q][h] = 8 * (seg - row[q][h] * C2SEG) + g;
          pa[q][h] = row[q][h] * P1P + col[q][h];
        }
      float acc[3][2][4];
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[q][j][e] = 0.f;
#pragma unroll 2
      for (int s = 0; s < 12; ++s) {
        const int o0 = tab[8 * s + t], o1 = tab[8 * s + t + 4];
        const float4 w = *reinterpret_cast<const float4*>(&w_s[W2 + (s * 32 + lane) * 4]);
        uint32_t bh[4], bl[4];
        split_tf32(w.x, bh[0], bl[0]);
        split_tf32(w.y, bh[1], bl[1]);
        split_tf32(w.z, bh[2], bl[2]);
        split_tf32(w.w, bh[3], bl[3]);
        uint32_t ah[3][4], al[3][4];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          split_tf32(p1_s[pa[q][0] + o0], ah[q][0], al[q][0]);
          split_tf32(p1_s[pa[q][1] + o0], ah[q][1], al[q][1]);
          split_tf32(p1_s[pa[q][0] + o1], ah[q][2], al[q][2]);
          split_tf32(p1_s[pa[q][1] + o1], ah[q][3], al[q][3]);
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int j = 0; j < 2; ++j) mma_tf32(acc[q][j], al[q], bh[j], bh[2 + j]);
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int j = 0; j < 2; ++j) mma_tf32(acc[q][j], ah[q], bl[j], bl[2 + j]);
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int j = 0; j < 2; ++j) mma_tf32(acc[q][j], ah[q], bh[j], bh[2 + j]);
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (mt0 + q >= 45) break;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int co = 8 * j + 2 * t + e;
            const float bias = w_s[B2 + co], al2 = w_s[A2 + co];
#pragma unroll
            for (int h = 0; h < 2; ++h)
              if (col[q][h] < C2W)
                a_s[(co * C2H + row[q][h]) * C2P + col[q][h]] = prelu(acc[q][j][2 * h + e] + bias, al2);
          }
      }
    }
  }
  __syncthreads();
#ifdef PNET_TIMING
  tph[3] = clock64();
#endif

  // ---- conv3 (16->32, 3x3) on the tensor pipe + PReLU + heads.
  // warp w owns output rows 2w, 2w+1; one pass = one row = two 16-pixel M tiles x four 8-channel N tiles.
  // fragment coordinates (PTX m16n8k8): g = lane/4, t = lane%4
  //   A: a0 (px g, k t)  a1 (px g+8, k t)  a2 (px g, k t+4)  a3 (px g+8, k t+4)
  //   B: b0 (k t, n g)   b1 (k t+4, n g)          C: c0 (px g, n 2t) c1 (px g, n 2t+1) c2/c3 (px g+8, ..)
  // conv2 planes are 18*36 = 648 floats apart (= 8 mod 32 banks), so the 32 lanes of an A load hit 32 banks.
  {
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int row = 2 * warp + pass;
      const float* arow = a_s + (t * C2H + row) * C2P + g;
      const float4* wf = reinterpret_cast<const float4*>(w_s + W3) + lane;
      float acc[2][4][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.f;
#pragma unroll 1
      for (int ky = 0; ky < 3; ++ky) {
        // 6 k-steps per filter row: (kx, channel half); unrolled so the offsets are immediates, the outer loop
        // stays rolled to bound the number of live shared-memory loads (no spills at 128 registers)
#pragma unroll
        for (int s6 = 0; s6 < 6; ++s6) {
          const int kx = s6 >> 1;
          const int koff = ((s6 & 1) * 8 * C2H) * C2P + kx;
          const float4 w0 = wf[(2 * s6) * 32], w1 = wf[(2 * s6 + 1) * 32];
          const float bw[2][4] = {{w0.x, w0.y, w0.z, w0.w}, {w1.x, w1.y, w1.z, w1.w}};
          uint32_t bh[2][4], bl[2][4];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) split_tf32(bw[i][j], bh[i][j], bl[i][j]);
          uint32_t ah[2][4], al[2][4];
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            const float* ap = arow + koff + 16 * m;
            split_tf32(ap[0], ah[m][0], al[m][0]);
            split_tf32(ap[8], ah[m][1], al[m][1]);
            split_tf32(ap[4 * C2H * C2P], ah[m][2], al[m][2]);
            split_tf32(ap[4 * C2H * C2P + 8], ah[m][3], al[m][3]);
          }
          // the three terms of one accumulator are dependent: issue the 8 independent accumulators between them
#pragma unroll
          for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_tf32(acc[m][j], al[m], bh[0][j], bh[1][j]);
#pragma unroll
          for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_tf32(acc[m][j], ah[m], bl[0][j], bl[1][j]);
#pragma unroll
          for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_tf32(acc[m][j], ah[m], bh[0][j], bh[1][j]);
        }
        arow += C2P;
        wf += 12 * 32;
      }
      // epilogue: bias + PReLU, heads (6 outputs over 32 channels; this thread holds 8 channels of 2 pixels per tile)
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        float hp[2][6];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int q = 0; q < 6; ++q) hp[i][q] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int co = 8 * j + 2 * t + e;
            const float bias = w_s[B3 + co], al = w_s[A3 + co];
            const float4 ha = *reinterpret_cast<const float4*>(&w_s[WH + co * 8]);
            const float2 hb = *reinterpret_cast<const float2*>(&w_s[WH + co * 8 + 4]);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float v = prelu(acc[m][j][2 * i + e] + bias, al);
              hp[i][0] = fmaf(v, ha.x, hp[i][0]);
              hp[i][1] = fmaf(v, ha.y, hp[i][1]);
              hp[i][2] = fmaf(v, ha.z, hp[i][2]);
              hp[i][3] = fmaf(v, ha.w, hp[i][3]);
              hp[i][4] = fmaf(v, hb.x, hp[i][4]);
              hp[i][5] = fmaf(v, hb.y, hp[i][5]);
            }
          }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            hp[i][q] += __shfl_xor_sync(0xffffffffu, hp[i][q], 1);
            hp[i][q] += __shfl_xor_sync(0xffffffffu, hp[i][q], 2);
          }
        // lanes t = 0 / 1 finish pixels g / g+8 of this tile
        if (t < 2) {
          float h[6];
#pragma unroll
          for (int q = 0; q < 6; ++q) h[q] = (t == 0 ? hp[0][q] : hp[1][q]) + w_s[BH + q];
          const int oy = oy0 + row, ox = ox0 + 16 * m + g + 8 * t;
          if (oy < L.oh && ox < L.ow) {
            // softmax over (h0, h1), class 1 -- same form as ATen's softmax (subtract max, exp, normalise)
            const float mx = fmaxf(h[0], h[1]);
            const float e0 = expf(h[0] - mx), e1 = expf(h[1] - mx);
            const float prob = __fdiv_rn(e1, e0 + e1);
            const size_t cell = (size_t)oy * L.ow + ox;
            if (L.prob) {
              const size_t plane = (size_t)L.oh * L.ow;
              L.prob[(size_t)b * plane + cell] = prob;
              float* rg = L.reg + (size_t)b * 4 * plane + cell;
              rg[0] = h[2]; rg[plane] = h[3]; rg[2 * plane] = h[4]; rg[3 * plane] = h[5];
            }
            if (L.cand && prob >= p.thr) {
              // generateBoundingBox: q1 = floor((2*c + 1)/scale), q2 = floor((2*c + 12)/scale)  (fp32, true division)
              const int slotbase = b * p.n_levels + lvl;
              const int slot = atomicAdd(&L.cnt[slotbase], 1);
              if (slot < p.cap) {
                Cand cd;
                cd.x1 = floorf(__fdiv_rn((float)(2 * ox + 1), L.scale));
                cd.y1 = floorf(__fdiv_rn((float)(2 * oy + 1), L.scale));
                cd.x2 = floorf(__fdiv_rn((float)(2 * ox + 12), L.scale));
                cd.y2 = floorf(__fdiv_rn((float)(2 * oy + 12), L.scale));
                cd.score = prob;
                cd.r0 = h[2]; cd.r1 = h[3]; cd.r2 = h[4]; cd.r3 = h[5];
                cd.key = (uint32_t)cell;
                L.cand[(size_t)slotbase * p.cap + slot] = cd;
              } else if (p.capflag) {
                p.capflag->overflow = 1; p.capflag->stage = 1; p.capflag->frame = b;
                p.capflag->count = slot + 1; p.capflag->capacity = p.cap;
              }
            }
          }
        }
      }
    }
  }
#ifdef PNET_TIMING
  tph[4] = clock64();
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) atomicAdd(&g_pnet_phase[i], (unsigned long long)(tph[i + 1] - tph[i]));
    atomicAdd(&g_pnet_phase[4], 1ull);
  }
#endif
}

}  // namespace pnet
#ifdef PNET_TIMING
extern "C" void trl_debug_pnet_timing(unsigned long long* out) {
  cudaMemcpyFromSymbol(out, pnet::g_pnet_phase, sizeof(unsigned long long) * 8);
  unsigned long long z[8] = {0};
  cudaMemcpyToSymbol(pnet::g_pnet_phase, z, sizeof(z));
}
#endif

// conv2's K ordering: group G = 2*kstep + (0: a0/a1/b0, 1: a2/a3/b1), slot t = lane % 4.
//   G < 18 : tap G/2, channels 4*(G%2) + t               (channels 0..7)
//   G >= 18: kx = (G-18)/2; channels 8,9 x ky 0..2, two zero-weight pads
// Every group has one kx and four distinct (ci + ky) mod 4 -> four bank octets.  Returns false for the pads.
static bool conv2_k(int G, int t, int* ci, int* ky, int* kx) {
  if (G < 18) {
    const int tap = G / 2;
    *ci = 4 * (G % 2) + t; *ky = tap / 3; *kx = tap % 3;
    return true;
  }
  *kx = (G - 18) / 2;
  if ((G - 18) % 2 == 0) {
    *ci = 8 + (t & 1); *ky = (t < 2) ? 0 : 2;            // (8,0) (9,0) (8,2) (9,2): octets 0,1,2,3
    return true;
  }
  if (t < 2) { *ci = 8 + t; *ky = 1; return true; }       // (8,1) (9,1): octets 1,2
  *ci = (t == 2) ? 0 : 3; *ky = 0;                        // pads (zero weights): octets 0,3
  return false;
}

// upstream layouts -> packed shared-memory image
int pnet_pack_weights(trl_ctx* c, const float* h, size_t len) {
  using namespace pnet;
  if (len != 6632) TRL_FAIL(c, TRL_E_INVALID, "pnet blob has %zu floats, expected 6632", len);
  std::vector<float> pk(WTOTAL, 0.f);
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
  // conv2 as mma.m16n8k8 B fragments + the k -> pooled-tile offset table
  for (int s = 0; s < 12; ++s)
    for (int lane = 0; lane < 32; ++lane) {
      const int g = lane >> 2, t = lane & 3;
      for (int i = 0; i < 2; ++i) {
        int ci, ky, kx;
        const bool used = conv2_k(2 * s + i, t, &ci, &ky, &kx);
        for (int j = 0; j < 2; ++j)
          pk[W2 + (s * 32 + lane) * 4 + 2 * i + j] = used ? w2[(8 * j + g) * 90 + ci * 9 + ky * 3 + kx] : 0.f;
        if (g == 0) {
          const int off = ci * P1PLANE + ky * P1P + kx;
          memcpy(&pk[T2 + 8 * s + 4 * i + t], &off, sizeof(int));
        }
      }
    }
  for (int co = 0; co < 16; ++co) { pk[B2 + co] = b2[co]; pk[A2 + co] = a2[co]; }
  // conv3 as mma.m16n8k8 B fragments: k-step s = (tap, channel half), k = 8s + t (+4) -> ci = (s&1)*8 + t (+4)
  for (int s = 0; s < 18; ++s)
    for (int i = 0; i < 2; ++i)
      for (int lane = 0; lane < 32; ++lane)
        for (int j = 0; j < 4; ++j) {
          const int g = lane >> 2, t = lane & 3;
          const int tap = s >> 1, ci = (s & 1) * 8 + t + 4 * i, co = 8 * j + g;
          pk[W3 + ((2 * s + i) * 32 + lane) * 4 + j] = w3[co * 144 + ci * 9 + tap];
        }
  for (int co = 0; co < 32; ++co) { pk[B3 + co] = b3[co]; pk[A3 + co] = a3[co]; }
  for (int ci = 0; ci < 32; ++ci) {
    pk[WH + ci * 8 + 0] = w41[0 * 32 + ci];
    pk[WH + ci * 8 + 1] = w41[1 * 32 + ci];
    for (int j = 0; j < 4; ++j) pk[WH + ci * 8 + 2 + j] = w42[j * 32 + ci];
  }
  pk[BH + 0] = b41[0]; pk[BH + 1] = b41[1];
  for (int j = 0; j < 4; ++j) pk[BH + 2 + j] = b42[j];
  TRL_CUDA(c, cudaMalloc(&c->d_pnet_packed, WTOTAL * sizeof(float)));
  TRL_CUDA(c, cudaMemcpy(c->d_pnet_packed, pk.data(), WTOTAL * sizeof(float), cudaMemcpyHostToDevice));
  TRL_CUDA(c, cudaFuncSetAttribute(pnet::pnet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  return TRL_OK;
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
  p.thr = 2.f; p.cap = 0; p.capflag = nullptr;
  pnet_kernel<<<dim3(L.tiles, B), 256, SMEM_BYTES, s>>>(c->d_pnet_packed, p);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

int launch_pnet_candidates(trl_ctx* c, const float* d_pyr, int B, const PyramidGeom& g, float thr, Cand* d_cand,
                           int* d_cnt, int cap, cudaStream_t s) {
  using namespace pnet;
  Params p{};
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
  if (blocks == 0 || B == 0) return TRL_OK;
  pnet_kernel<<<dim3(blocks, B), 256, SMEM_BYTES, s>>>(c->d_pnet_packed, p);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}
