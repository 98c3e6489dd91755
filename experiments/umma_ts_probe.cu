// Probe for the P-Net tcgen05 design, part 2 (not product code): A operand in TMEM ("TS" form of tcgen05.mma).
//  1. semantics: A[128 x K] fp16 written with tcgen05.st.32x32b (lane = row, 32-bit column c = elements 2c, 2c+1),
//     B from a no-swizzle K-major shared-memory descriptor, D fp32 in TMEM;
//  2. throughput of small-N TS MMAs (N = 16 / 32 / 64) against the SS form;
//  3. cost of the im2col fill (LDS.128 -> tcgen05.st) running beside the MMAs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_ts_probe umma_ts_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) { if (clock64() - t0 > 2000000000LL) __trap(); }
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred px;\nelect.sync _|px, %1;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- 1. semantics.  A: [128][K] halves (global, row major), B: [N][K] halves.  K = 16 * KS.
__global__ void __launch_bounds__(128) sem_kernel(const __half* A, const __half* B, float* out, int N, int KS) {
  extern __shared__ __align__(128) uint8_t smem[];
  // B image: [k-half (2 KS)][n][8 halves = 16 B]
  __half* sb = reinterpret_cast<__half*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * KS * N * 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = 16 * KS;
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i - n * K;
    sb[((k >> 3) * N + n) * 8 + (k & 7)] = B[i];
  }
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_a = tmem + 128;                 // A at columns 128.., D at columns 0..N-1
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  // thread = row: 8 columns (16 halves) per k-step
  for (int ks = 0; ks < KS; ++ks) {
    uint32_t v[8];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(A + (size_t)tid * K + ks * 16);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = src[j];
    tmem_st8(tmem_a + lane_base + ks * 8, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (tid == 0) {
    for (int ks = 0; ks < KS; ++ks) {
      const uint64_t bd = make_desc(smem_u32(sb) + (uint32_t)(2 * ks) * N * 16, N * 16, 128);
      mma_ts(tmem, tmem_a + ks * 8, bd, idesc, ks > 0);
    }
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + lane_base + (uint32_t)c0, v);
    for (int j = 0; j < 16; ++j) out[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

static void run_sem(int N, int KS) {
  const int K = 16 * KS;
  std::vector<__half> A(128 * K), B(N * K);
  std::vector<float> Af(128 * K), Bf(N * K);
  srand(7);
  for (int i = 0; i < 128 * K; ++i) { Af[i] = (float)((rand() % 17) - 8) / 8.f; A[i] = __float2half(Af[i]); }
  for (int i = 0; i < N * K; ++i) { Bf[i] = (float)((rand() % 9) - 4) / 4.f; B[i] = __float2half(Bf[i]); }
  __half *dA, *dB; float* dO;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dO, 128 * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(sem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  sem_kernel<<<1, 128, 2 * KS * N * 16 + 64>>>(dA, dB, dO, N, KS);
  CK(cudaDeviceSynchronize());
  std::vector<float> out(128 * N);
  CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)Af[m * K + k] * Bf[n * K + k];
      maxerr = fmax(maxerr, fabs(ref - out[m * N + n]));
    }
  printf("TS semantics N=%d K=%d: max|err| = %.3e %s\n", N, K, maxerr, maxerr == 0 ? "(exact: layout confirmed)" : "(MISMATCH)");
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
}

// ---- 2/3. throughput.  160 threads: warps 0-3 optionally run the fill loop, warp 4 lane 0 issues MMAs.
//  ts = 1: A from TMEM, 0: A from smem (no-swizzle).  fill: 0 none, 1 = LDS.128 x4 + tcgen05.st x2 per (pixel, tap)
__global__ void __launch_bounds__(256) tput_kernel(long long* cycles, int ts, int N, int n_mma, int fill, int n_fill, int ndst, int n_iss, int M) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem;                  // 16 KB of A image / fill source
  uint8_t* sb = smem + 16384;          // B: up to 256 rows x 32 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 8192) / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(bar, n_iss); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (((uint32_t)M >> 4) << 24);
  if (warp < 4 && fill) {
    const long long t0 = clock64();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t sink = 0;
    for (int i = 0; i < n_fill; ++i) {
      // one (pixel, tap): 16 channels hi + 16 channels lo from two planes at 16 B pixel stride
      const int shift = (i % 9) * 16;
      const uint4 h0 = *reinterpret_cast<const uint4*>(sa + tid * 16 + shift);
      const uint4 h1 = *reinterpret_cast<const uint4*>(sa + 4096 + tid * 16 + shift);
      const uint4 l0 = *reinterpret_cast<const uint4*>(sa + 8192 + tid * 16 + shift);
      const uint4 l1 = *reinterpret_cast<const uint4*>(sa + 12288 - 256 + tid * 16 + shift);
      const uint32_t vh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
      const uint32_t vl[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
      tmem_st8(tmem + lane_base + 64 + (i % 9) * 8, vh);
      tmem_st8(tmem + lane_base + 64 + 72 + (i % 9) * 8, vl);
      sink += h0.x;
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    if (tid == 0) { cycles[3] = t1 - t0; cycles[4] = sink; }
  }
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (warp_u >= 4 && warp_u < 4 + n_iss) {
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t sbu = __shfl_sync(0xffffffffu, smem_u32(sb), 0);
    const uint32_t sau = __shfl_sync(0xffffffffu, smem_u32(sa), 0);
    const uint64_t bhi = ((uint64_t)(((uint32_t)N * 16u >> 4) & 0x3FFFu) << 16) | ((uint64_t)((128u >> 4) & 0x3FFFu) << 32) | (1ull << 46);
    const uint64_t ahi = ((uint64_t)((4096u >> 4) & 0x3FFFu) << 16) | ((uint64_t)((128u >> 4) & 0x3FFFu) << 32) | (1ull << 46);
    const uint32_t dcol = tm + 256 + (uint32_t)(warp_u - 4) * 64;
    if (elect_one()) {
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; i += 27) {
#pragma unroll
        for (int term = 0; term < 3; ++term)
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t d = dcol + (ndst > 1 ? (uint32_t)(tap & 1) * 32u : 0u);
            if (ts) mma_ts(d, tm + 64 + (term == 1 ? 72 : 0) + tap * 8, bhi | (uint64_t)(((sbu + (term == 2 ? 512u : 0u)) >> 4) & 0x3FFFu), idesc, 1u);
            else mma_ss(d, ahi | (uint64_t)(((sau + (term == 1 ? 8192u : 0u) + tap * 16u) >> 4) & 0x3FFFu),
                        bhi | (uint64_t)(((sbu + (term == 2 ? 512u : 0u)) >> 4) & 0x3FFFu), idesc, 1u);
          }
      }
      const long long t1 = clock64();
      mma_commit(bar);
      if (warp_u == 4) { cycles[0] = t1 - t0; cycles[2] = t0; }
    }
  }
  __syncwarp();
  if (n_mma > 0) mbar_wait(bar, 0);
  __syncthreads();
  if (tid == 128) cycles[1] = clock64() - cycles[2];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static void run_tput(int ts, int N, int n_mma, int fill, int n_fill, int ndst, int ctas = 1, int n_iss = 1, int M = 128) {
  long long* d_cyc; CK(cudaMalloc(&d_cyc, 64)); CK(cudaMemset(d_cyc, 0, 64));
  CK(cudaFuncSetAttribute(tput_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  tput_kernel<<<ctas, 256, 16384 + 8192 + 64>>>(d_cyc, ts, N, n_mma, fill, n_fill, ndst, n_iss, M);
  CK(cudaDeviceSynchronize());
  long long cyc[5]; CK(cudaMemcpy(cyc, d_cyc, 40, cudaMemcpyDeviceToHost));
  printf("tput %s M=%d N=%3d ndst=%d ctas=%d iss=%d mmas=%d fill=%d: mma issue %.1f total %.1f cyc/mma(per issuer)", ts ? "TS" : "SS", M, N, ndst, ctas, n_iss, n_mma, fill,
         n_mma ? (double)cyc[0] / n_mma : 0.0, n_mma ? (double)cyc[1] / n_mma : 0.0);
  if (fill) printf(" | fill %.1f cyc per (128 px, tap) = %.0f per 9-tap tile", (double)cyc[3] / n_fill, 9.0 * cyc[3] / n_fill);
  printf("\n");
  cudaFree(d_cyc);
}

int main() {
  run_sem(32, 1);
  run_sem(32, 9);
  run_sem(16, 6);
  run_sem(64, 2);
  for (int N : {16, 32}) for (int nd : {1, 2}) for (int ni : {1, 2}) { run_tput(1, N, 3645, 0, 0, nd, 1, ni); run_tput(0, N, 3645, 0, 0, nd, 1, ni); }
  run_tput(1, 64, 3645, 0, 0, 1, 1, 1); run_tput(0, 64, 3645, 0, 0, 1, 1, 1);
  run_tput(1, 32, 0, 1, 3600, 2);          // fill alone
  run_tput(1, 32, 10800, 1, 3600, 2); run_tput(1, 32, 10800, 1, 3600, 1);      // fill + TS MMAs (27 MMAs per 9 fills, as conv3 with the 3-term split)
  run_tput(1, 32, 10800, 1, 3600, 2, 2);   // two CTAs (different SMs: no contention expected; sanity)
  return 0;
}
