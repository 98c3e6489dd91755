// K11: InceptionResnetV1 convolutions as bf16 implicit GEMM on the 5th-gen tensor cores (sm_100a).
//
//   D[128 x BN] (fp32, TMEM) += A[128 x BK] (smem, K-major, TMA) * W[BN x BK]^T (smem, K-major, TMA)
//
// * A is never materialised: for filter tap (ky,kx) and channel block c0 the 128 rows of the tile are one 4-D TMA box
//   {BK channels, box_w, box_h, box_n} of the NHWC input at (c0, ox0*s+kx-pw, oy0*s+ky-ph, n0); out-of-bounds
//   elements (the zero padding) are filled by the TMA unit, strided convs use the map's element strides.
//   1x1 stride-1 convs use a flat (channels x pixels) view, so tiles carry no spatial padding at all.
// * tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) is issued by one thread; accumulators live in TMEM.
// * warp 0 = TMA producer of the A operand, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue (two warps per
//   TMEM lane group, one per column half: tcgen05.ld -> bias / residual / ReLU -> bf16 -> channel slice of the
//   destination: concat-free Inception blocks), warp 10 = TMA producer of the weight and residual tiles.
// * smem ring of `stages` {A,B} tiles guarded by full/empty mbarriers; tcgen05.commit releases a stage.
// Every mbarrier wait is bounded (trap after ~2 s) so a protocol bug surfaces as a CUDA error, never as a hang.
#include "facenet.cuh"

namespace umma {

constexpr int BLOCK_M = 128;
constexpr int NUM_THREADS = 352;     // warp 0 TMA (A operand), warp 1 MMA, warps 2-9 epilogue (two warps per TMEM lane group:
                                     // column halves), warp 10 TMA (weights + residual): two issue threads halve the per-k-iteration issue time

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;       // fast path: no clock read (the producer / MMA issue loops are latency critical)
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred px;\nelect.sync _|px, %1;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1 for swizzled K-major) | [32,46) SBO>>4 (8 rows * row bytes) | [46,48) version=1
//   | [61,64) layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// tile id -> output tile: N tiles fastest, so CTAs running side by side share the same A rows through L2
struct TileCoord { int ox0, oy0, n0, n_off; };
__device__ __forceinline__ TileCoord tile_coord(const ConvOp& op, int t, int n_ntiles, int tiles_w, int tiles_h) {
  const int nt = t % n_ntiles, mt = t / n_ntiles;
  const int tw = mt % tiles_w, th = (mt / tiles_w) % tiles_h, tn = mt / (tiles_w * tiles_h);
  TileCoord c;
  c.ox0 = tw * op.box_w; c.oy0 = th * op.box_h; c.n0 = tn * op.box_n; c.n_off = nt * op.block_n;
  return c;
}

// Persistent: one CTA per SM walks the tile list.  The TMA producer runs ahead across tile boundaries (the smem ring
// never drains), the MMA issuer alternates between two TMEM accumulators, and the epilogue warps drain accumulator
// i while the MMAs of the next tile fill accumulator i^1 -- load, MMA and epilogue of consecutive tiles overlap.
#ifdef PNET_TIMING
__device__ unsigned long long g_umma_phase[8];
#endif

__global__ void __launch_bounds__(NUM_THREADS, 1) conv_umma_kernel(const __grid_constant__ ConvOp op, int n_img, int tiles_w,
                                                                   int tiles_h, int n_ntiles, int total_tiles, int stages,
                                                                   uint32_t acc_cols, const int* __restrict__ d_n) {
  // shared memory: [stages x A][stages x B][barriers, tmem slot, row table][output tile][residual tile]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int BK = op.block_k, BN = op.block_n;
  const uint32_t a_bytes = BLOCK_M * BK * 2, b_bytes = BN * BK * 2;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + stages * a_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + stages * b_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;      // [2] accumulator complete (tcgen05.commit)
  uint64_t* tempty_bar = tfull_bar + 2;          // [2] accumulator drained (128 epilogue threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* rfull_bar = reinterpret_cast<uint64_t*>(tmem_slot + 2);          // [2] residual tile landed (TMA)
  uint64_t* rempty_bar = rfull_bar + 2;                                       // [2] residual tile consumed (256 epilogue threads)
  float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(rempty_bar + 2) + 15) & ~(uintptr_t)15);   // [Cout] bias, once per CTA
  // residual tiles: [2][BN / 64] boxes of 128 rows x 64 channels (128 B rows, SWIZZLE_128B), 1024-byte aligned
  uint8_t* rt = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sbias + op.Cout) + 1023) & ~(uintptr_t)1023);
  const bool has_resid = op.epi != EPI_RELU;
  const uint32_t rt_bytes = (uint32_t)(BN / 64) * 16384u;                     // one residual tile

  // warp index as a warp-uniform value: the TMA / MMA issue loops then live in uniform registers (no R2UR per instruction)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&op.tmap_a);
    prefetch_tmap(&op.tmap_w);
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }      // full: A producer + B producer
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 256); mbar_init(&rfull_bar[i], 1); mbar_init(&rempty_bar[i], 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2u * acc_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < op.Cout; i += NUM_THREADS) sbias[i] = __ldg(op.bias + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  // Programmatic dependent launch: everything above (barriers, TMEM, bias -- none of it produced by the previous
  // layer) may run while the previous kernel of the stream is still finishing; the activations may not.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (d_n) {
    // live batch size known only on the device (face-bearing crops after compaction): shrink the tile list to it.
    // The launch (grid, tensor maps) is sized for the host-side bound; rows beyond the live batch are never computed.
    const int ne = min(n_img, *reinterpret_cast<const volatile int*>(d_n));
    n_img = ne;
    int tiles_n = 1;
    if (op.flat) tiles_w = (int)(((long long)ne * op.Hout * op.Wout + 127) / 128);
    else tiles_n = (ne + op.box_n - 1) / op.box_n;
    total_tiles = tiles_w * tiles_h * tiles_n * n_ntiles;
  }

  const int cin_blocks = op.Cin / BK;
  const int k_iters = op.kh * op.kw * cin_blocks;

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer, A operand
      const uint32_t box_rows = (uint32_t)(op.box_w * op.box_h * op.box_n);
      const uint32_t tx_bytes = box_rows * BK * 2;
      int stage = 0;
      uint32_t phase = 0;
      int tl = 0;
#ifdef PNET_TIMING
      long long p_wait = 0, p_total = clock64();
#endif
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
        const TileCoord tc = tile_coord(op, t, n_ntiles, tiles_w, tiles_h);
        const int cx = tc.ox0 * op.stride - op.pad_w, cy = tc.oy0 * op.stride - op.pad_h;
        int cb = 0, ky = 0, kx = 0;                     // incremental (tap, channel block) counters: no divisions in the issue loop
        for (int it = 0; it < k_iters; ++it) {
#ifdef PNET_TIMING
          const long long tq = clock64();
#endif
          mbar_wait(&empty_bar[stage], phase ^ 1);
#ifdef PNET_TIMING
          p_wait += clock64() - tq;
#endif
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          tma_load_4d(smem_u32(smem_a + stage * a_bytes), &op.tmap_a, &full_bar[stage], cb * BK, cx + kx, cy + ky, tc.n0);
          if (++stage == stages) { stage = 0; phase ^= 1; }
          if (++cb == cin_blocks) { cb = 0; if (++kx == op.kw) { kx = 0; ++ky; } }
        }
      }
#ifdef PNET_TIMING
      atomicAdd(&g_umma_phase[0], (unsigned long long)(clock64() - p_total));
      atomicAdd(&g_umma_phase[1], (unsigned long long)p_wait);
      atomicAdd(&g_umma_phase[6], (unsigned long long)tl * k_iters);
#endif
    }
  } else if (warp == 10) {
    if (elect_one()) {
      // ===== TMA producer, weights (and the residual tile of the up-projections)
      int stage = 0;
      uint32_t phase = 0;
      int tl = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
        const TileCoord tc = tile_coord(op, t, n_ntiles, tiles_w, tiles_h);
        if (has_resid) {
          // the tile's residual rows (flat 1x1 layers only): fetched by TMA a tile ahead of its epilogue
          const int rb = tl & 1;
          mbar_wait(&rempty_bar[rb], ((uint32_t)(tl >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&rfull_bar[rb], rt_bytes);
          for (int j = 0; j < BN / 64; ++j)
            tma_load_2d(smem_u32(rt + rb * rt_bytes + j * 16384), &op.tmap_r, &rfull_bar[rb], tc.n_off + 64 * j, tc.ox0);
        }
        int kcol = 0;                                   // K coordinate of the weight tile: (tap * Cin + cb * BK) = it * BK
        for (int it = 0; it < k_iters; ++it, kcol += BK) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], b_bytes);
          tma_load_2d(smem_u32(smem_b + stage * b_bytes), &op.tmap_w, &full_bar[stage], kcol, tc.n_off);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
      const uint32_t layout = (BK == 64) ? 2u : 4u;
      const uint32_t sbo = 8u * (uint32_t)BK * 2u;
      int stage = 0;
      uint32_t phase = 0;
      int tl = 0;
#ifdef PNET_TIMING
      long long m_wfull = 0, m_wtmem = 0, m_total = clock64();
#endif
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
        const int buf = tl & 1;
#ifdef PNET_TIMING
        const long long tq0 = clock64();
#endif
        mbar_wait(&tempty_bar[buf], ((uint32_t)(tl >> 1) & 1u) ^ 1u);       // the epilogue has drained this accumulator
#ifdef PNET_TIMING
        m_wtmem += clock64() - tq0;
#endif
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)buf * acc_cols;
        for (int it = 0; it < k_iters; ++it) {
#ifdef PNET_TIMING
          const long long tq1 = clock64();
#endif
          mbar_wait(&full_bar[stage], phase);
#ifdef PNET_TIMING
          m_wfull += clock64() - tq1;
#endif
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_u32(smem_a + stage * a_bytes), sbo, layout);
          const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + stage * b_bytes), sbo, layout);
          for (int kk = 0; kk < BK / 16; ++kk)
            mma_bf16(d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (it > 0 || kk > 0) ? 1u : 0u);
          mma_commit(&empty_bar[stage]);          // frees this smem stage when the MMAs above retire
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        mma_commit(&tfull_bar[buf]);
      }
#ifdef PNET_TIMING
      atomicAdd(&g_umma_phase[2], (unsigned long long)(clock64() - m_total));
      atomicAdd(&g_umma_phase[3], (unsigned long long)m_wfull);
      atomicAdd(&g_umma_phase[4], (unsigned long long)m_wtmem);
      atomicAdd(&g_umma_phase[5], (unsigned long long)tl);
#endif
    }
  } else if (warp >= 2 && warp <= 9) {
    // ===== epilogue (8 warps): warp w owns TMEM lanes 32*(w%4) .. +31 and one half of the tile's columns; thread <->
    // one output pixel (tile row).  No global loads on this path: bias comes from shared memory, the residual tile was
    // fetched by the producer's TMA.  All accumulator columns of the thread are loaded first (one wait) and the
    // accumulator is handed back to the MMA warp before the bias / residual / ReLU / store work starts.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int BNh = BN >> 1;                      // columns of this thread: [half * BNh, half * BNh + BNh), BNh in {16..96}
    int tl = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
      const TileCoord tc = tile_coord(op, t, n_ntiles, tiles_w, tiles_h);
      const int ox0 = tc.ox0, oy0 = tc.oy0, n0 = tc.n0, n_off = tc.n_off;
      const int buf = tl & 1;
      bool valid;
      long long m;
      if (op.flat) {
        m = (long long)ox0 + r;
        valid = m < (long long)n_img * op.Hout * op.Wout;
      } else {
        const int per_img = op.box_w * op.box_h;
        const int ni = r / per_img, rem = r - ni * per_img;
        const int hy = rem / op.box_w, wx = rem - hy * op.box_w;
        const int n = n0 + ni, oy = oy0 + hy, ox = ox0 + wx;
        valid = (ni < op.box_n) && (n < n_img) && (oy < op.Hout) && (ox < op.Wout);
        m = ((long long)n * op.Hout + oy) * op.Wout + ox;
      }
      const int cbase = half * BNh;
      mbar_wait(&tfull_bar[buf], (uint32_t)(tl >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * acc_cols + (uint32_t)cbase;
      uint32_t v[6][16];
#pragma unroll
      for (int cc = 0; cc < 6; ++cc)
        if (cc * 16 < BNh) tmem_ld16_async(taddr_row + (uint32_t)(cc * 16), v[cc]);
      tmem_ld_wait();
      tc_fence_before();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty_bar[buf])) : "memory");
      if (has_resid) mbar_wait(&rfull_bar[buf], (uint32_t)(tl >> 1) & 1u);
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        if (cc * 16 >= BNh) break;
        const int cl = cbase + cc * 16;            // column within the tile
        const int ng = n_off + cl;
        float f[16];
        const float4* b4 = reinterpret_cast<const float4*>(sbias + ng);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 bv = b4[j];
          f[4 * j + 0] = __uint_as_float(v[cc][4 * j + 0]) + bv.x;
          f[4 * j + 1] = __uint_as_float(v[cc][4 * j + 1]) + bv.y;
          f[4 * j + 2] = __uint_as_float(v[cc][4 * j + 2]) + bv.z;
          f[4 * j + 3] = __uint_as_float(v[cc][4 * j + 3]) + bv.w;
        }
        if (has_resid) {
          // 128-byte swizzle: 16-byte chunk c of row r sits at chunk c ^ (r & 7)
          const uint8_t* box = rt + buf * rt_bytes + (cl >> 6) * 16384 + r * 128;
          const int ch = (cl & 63) >> 3;             // first of the two 16-byte chunks
          const uint4 x0 = *reinterpret_cast<const uint4*>(box + ((ch ^ (r & 7)) << 4));
          const uint4 x1 = *reinterpret_cast<const uint4*>(box + (((ch + 1) ^ (r & 7)) << 4));
          const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&x0);
          const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&x1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 a = __bfloat1622float2(h0[j]), b = __bfloat1622float2(h1[j]);
            f[2 * j] = fmaf(op.scale, f[2 * j], a.x);
            f[2 * j + 1] = fmaf(op.scale, f[2 * j + 1], a.y);
            f[8 + 2 * j] = fmaf(op.scale, f[8 + 2 * j], b.x);
            f[8 + 2 * j + 1] = fmaf(op.scale, f[8 + 2 * j + 1], b.y);
          }
        }
        if (op.epi != EPI_RESID) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (valid) {
          uint4 o0, o1;
          o0.x = pack_bf16(f[0], f[1]); o0.y = pack_bf16(f[2], f[3]); o0.z = pack_bf16(f[4], f[5]); o0.w = pack_bf16(f[6], f[7]);
          o1.x = pack_bf16(f[8], f[9]); o1.y = pack_bf16(f[10], f[11]); o1.z = pack_bf16(f[12], f[13]); o1.w = pack_bf16(f[14], f[15]);
          for (int sidx = 0; sidx < op.nseg; ++sidx) {
            const OutSeg& sg = op.seg[sidx];
            if (ng >= sg.n_begin && ng < sg.n_end) {
              uint4* d = reinterpret_cast<uint4*>(sg.dst + (size_t)m * sg.dst_ctot + sg.dst_coff + (ng - sg.n_begin));
              d[0] = o0;
              d[1] = o1;
            }
          }
        }
      }
      if (has_resid) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&rempty_bar[buf])) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * acc_cols) : "memory");
  }
}


}  // namespace umma
#ifdef PNET_TIMING
extern "C" void trl_debug_umma_timing(unsigned long long* out) {
  cudaMemcpyFromSymbol(out, umma::g_umma_phase, sizeof(unsigned long long) * 8);
  unsigned long long z[8] = {0};
  cudaMemcpyToSymbol(umma::g_umma_phase, z, sizeof(z));
}
#endif

// ----------------------------------------------------------------------------- host side

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

int umma_init(trl_ctx* c) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    TRL_CUDA(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) TRL_FAIL(c, TRL_E_CUDA, "cuTensorMapEncodeTiled unavailable");
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  TRL_CUDA(c, cudaFuncSetAttribute(umma::conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return TRL_OK;
}

// fp32 tiled tensor map for other kernels of the library (P-Net input staging); `map` = CUtensorMap storage (128 B, 64-byte
// aligned); zero fill out of bounds, no swizzle, no interleave
int tma_encode_tiled_f32(trl_ctx* c, void* map, const void* base, int rank, const unsigned long long* dims,
                         const unsigned long long* strides_bytes, const unsigned* box) {
  int rc = umma_init(c);
  if (rc != TRL_OK) return rc;
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), d,
                        st, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) TRL_FAIL(c, TRL_E_CUDA, "cuTensorMapEncodeTiled(f32) failed with %d", (int)r);
  return TRL_OK;
}

// Choose the 128-row tile shape (box_w, box_h, box_n) that minimises the number of tiles for a nominal batch.
static void choose_box(const ConvOp& op, int* bw, int* bh, int* bn) {
  const int NB = 240;   // nominal batch
  long long best = -1;
  const int lim = 256 / op.stride;      // TMA boxDim <= 256 along the traversed extent
  for (int w = 1; w <= op.Wout && w <= 128 && w <= lim; ++w)
    for (int h = 1; h <= op.Hout && w * h <= 128 && h <= lim; ++h) {
      const int nn = 128 / (w * h);
      const long long tiles = (long long)ceil_div(op.Wout, w) * ceil_div(op.Hout, h) * ceil_div(NB, nn);
      if (best < 0 || tiles < best) { best = tiles; *bw = w; *bh = h; *bn = nn; }
    }
}

int umma_encode_maps(trl_ctx* c, ConvOp& op, int n_cap) {
  if (op.Cin % 64 == 0) op.block_k = 64;
  else if (op.Cin % 32 == 0) op.block_k = 32;
  else TRL_FAIL(c, TRL_E_INVALID, "%s: Cin=%d not a multiple of 32", op.name, op.Cin);
  if (op.Cout % 128 == 0) op.block_n = 128;
  else if (op.Cout <= 256 && op.Cout % 16 == 0) op.block_n = op.Cout;
  else if (op.Cout % 96 == 0) op.block_n = 96;
  else if (op.Cout % 64 == 0) op.block_n = 64;
  else TRL_FAIL(c, TRL_E_INVALID, "%s: unsupported Cout=%d", op.name, op.Cout);
  if (op.block_n > 256) TRL_FAIL(c, TRL_E_INVALID, "%s: block_n", op.name);
  op.flat = (op.kh == 1 && op.kw == 1 && op.stride == 1 && op.pad_h == 0 && op.pad_w == 0) ? 1 : 0;
  const CUtensorMapSwizzle swz = (op.block_k == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;

  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], estr[4];
  const cuuint64_t pix = (cuuint64_t)n_cap * op.Hin * op.Win;
  if (op.flat) {
    op.box_w = 128; op.box_h = 1; op.box_n = 1;
    dims[0] = op.Cin; dims[1] = pix; dims[2] = 1; dims[3] = 1;
    strides[0] = (cuuint64_t)op.in_ctot * 2; strides[1] = pix * op.in_ctot * 2; strides[2] = strides[1];
    box[0] = op.block_k; box[1] = 128; box[2] = 1; box[3] = 1;
    estr[0] = estr[1] = estr[2] = estr[3] = 1;
  } else {
    choose_box(op, &op.box_w, &op.box_h, &op.box_n);
    dims[0] = op.Cin; dims[1] = op.Win; dims[2] = op.Hin; dims[3] = n_cap;
    strides[0] = (cuuint64_t)op.in_ctot * 2;
    strides[1] = (cuuint64_t)op.Win * op.in_ctot * 2;
    strides[2] = (cuuint64_t)op.Hin * op.Win * op.in_ctot * 2;
    box[0] = op.block_k; box[1] = op.box_w * op.stride; box[2] = op.box_h * op.stride; box[3] = op.box_n;
    estr[0] = 1; estr[1] = op.stride; estr[2] = op.stride; estr[3] = 1;
  }
  CUresult r = g_encode(&op.tmap_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)(op.in + op.in_coff), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) TRL_FAIL(c, TRL_E_CUDA, "%s: cuTensorMapEncodeTiled(A) failed with %d", op.name, (int)r);
  const cuuint64_t ktot = (cuuint64_t)op.kh * op.kw * op.Cin;
  cuuint64_t wd[2] = {ktot, (cuuint64_t)op.Cout};
  cuuint64_t ws[1] = {ktot * 2};
  cuuint32_t wb[2] = {(cuuint32_t)op.block_k, (cuuint32_t)op.block_n};
  cuuint32_t we[2] = {1, 1};
  r = g_encode(&op.tmap_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)op.w, wd, ws, wb, we, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) TRL_FAIL(c, TRL_E_CUDA, "%s: cuTensorMapEncodeTiled(W) failed with %d", op.name, (int)r);
  if (op.epi != EPI_RELU) {
    // residual operand [pixels][Cout] of the (flat 1x1) up-projections: 128 rows x 64 channels per box, 128-byte swizzle
    if (!op.flat || op.block_n % 64 != 0) TRL_FAIL(c, TRL_E_INVALID, "%s: residual epilogue needs a flat 1x1 layer with block_n %% 64 == 0", op.name);
    const cuuint64_t opix = (cuuint64_t)n_cap * op.Hout * op.Wout;
    cuuint64_t rd[2] = {(cuuint64_t)op.Cout, opix};
    cuuint64_t rs[1] = {(cuuint64_t)op.Cout * 2};
    cuuint32_t rb[2] = {64, 128};
    cuuint32_t re[2] = {1, 1};
    r = g_encode(&op.tmap_r, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)op.resid, rd, rs, rb, re, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) TRL_FAIL(c, TRL_E_CUDA, "%s: cuTensorMapEncodeTiled(R) failed with %d", op.name, (int)r);
  }
  return TRL_OK;
}

int launch_conv_umma(trl_ctx* c, const ConvOp& op, int n, cudaStream_t s, const int* d_n) {
  using namespace umma;
  if (n <= 0) return TRL_OK;
  int tiles_w, tiles_h, tiles_n;
  if (op.flat) {
    tiles_w = (int)(((long long)n * op.Hout * op.Wout + 127) / 128);
    tiles_h = 1; tiles_n = 1;
  } else {
    tiles_w = ceil_div(op.Wout, op.box_w);
    tiles_h = ceil_div(op.Hout, op.box_h);
    tiles_n = ceil_div(n, op.box_n);
  }
  const int stage_bytes = (BLOCK_M + op.block_n) * op.block_k * 2;
  // persistent CTA, one per SM: as many ring stages as fit (the producer prefetches across tile boundaries); the
  // footprint is kept above half of the SM's shared memory so that two CTAs (2 x 512 TMEM columns) never co-reside
  const bool has_resid = op.epi != EPI_RELU;
  const size_t epi_bytes = (size_t)op.Cout * 4 + (has_resid ? 1024 + 2 * (size_t)(op.block_n / 64) * 16384 : 1024);
  const size_t fixed = 1024 /*alignment*/ + 512 /*barriers: 2 x 20 stages + 8*/ + 64 + epi_bytes;
  int stages = (int)((196 * 1024 - fixed) / stage_bytes);
  if (stages > 20) stages = 20;      // small-K / narrow layers are TMA-latency bound: keep many loads in flight
  if (stages < 2) stages = 2;
  size_t smem = (size_t)stages * stage_bytes + fixed;
  if (smem < 120 * 1024) smem = 120 * 1024;
  const uint32_t acc_cols = op.block_n <= 32 ? 32u : op.block_n <= 64 ? 64u : op.block_n <= 128 ? 128u : 256u;
  const int n_ntiles = op.Cout / op.block_n;
  const long long total = (long long)tiles_w * tiles_h * tiles_n * n_ntiles;
  const int grid = (int)(total < c->num_sms ? total : c->num_sms);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // the ~100 back-to-back layers overlap their prologues
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TRL_CUDA(c, cudaLaunchKernelEx(&cfg, conv_umma_kernel, op, n, tiles_w, tiles_h, n_ntiles, (int)total, stages, acc_cols, d_n));
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}
