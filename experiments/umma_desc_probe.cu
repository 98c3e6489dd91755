// Probe (not product code): address semantics and cost of no-swizzle K-major tcgen05 shared-memory descriptors with
// arbitrary LBO / SBO -- the two tricks the all-tensor-pipe P-Net (pnet2.cu) is built on:
//   * LBO = 16 B: the second 8-element K chunk of row m is the first chunk of row m + 1 (two horizontally adjacent pixels
//     form one K = 16 operand row without an im2col copy),
//   * SBO = any multiple of 16 B: the sixteen 8-row groups of an M = 128 tile may be 16 image rows of the same parity
//     (a 2-D patch as one M tile).
// Method: the A region holds fp16 values that encode their own byte offset (two passes: low 10 bits, high bits), B is
// the 16 x 16 identity, so D[m][n] = the half-word the hardware fetched for (row m, k = n).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_desc_probe umma_desc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (clock64() - t0 > 2000000000LL) __trap();
  }
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int A_BYTES = 64 * 1024;

// mode 0: semantics (one MMA, N = 16, identity B; pass selects the encoding); mode 1: timing (n_mma MMAs of width N)
__global__ void __launch_bounds__(128) probe(float* out, long long* cycles, int mode, int pass, uint32_t start, uint32_t lbo, uint32_t sbo,
                                             int N, int n_mma, int ndst) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __half* a = reinterpret_cast<__half*>(smem);
  __half* b = reinterpret_cast<__half*>(smem + A_BYTES);            // [2 k-chunks][256 rows][8]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + A_BYTES + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < A_BYTES / 2; i += 128) a[i] = __float2half_rn(mode == 0 ? (float)(pass == 0 ? (i & 1023) : (i >> 10)) : 0.f);
  for (int i = tid; i < 2 * 256 * 8; i += 128) {
    const int kc = i / (256 * 8), n = (i / 8) % 256, j = i % 8;
    b[i] = __float2half_rn((mode == 0 && n == kc * 8 + j) ? 1.f : 0.f);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);      // f32 acc, f16 x f16, K-major, M = 128
  long long t0 = 0;
  bool leader = false;
  if (warp == 0) {
    uint32_t pred = 0;
    asm volatile("{\n.reg .pred px;\nelect.sync _|px, %1;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xFFFFFFFFu));
    leader = pred != 0;
  }
  if (leader) {
    const uint64_t bd = make_desc(smem_u32(b), 256u * 16u, 128u);
    t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      // timing: walk the start address so consecutive MMAs read different rows, as the conv taps do
      const uint64_t ad = make_desc(smem_u32(a) + start + (mode == 1 ? (uint32_t)(i & 7) * 16u : 0u), lbo, sbo);
      mma_f16(tmem + (uint32_t)((i & (ndst - 1)) * N), ad, bd, idesc, mode == 1 ? 1u : 0u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  mbar_wait(bar, 0);
  if (leader) cycles[0] = clock64() - t0;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (mode == 0) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    for (int j = 0; j < 16; ++j) out[tid * 16 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, 128 * 16 * 4)); CK(cudaMalloc(&d_cyc, 8));
  const size_t smem = A_BYTES + 8192 + 64;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  struct Case { uint32_t start, lbo, sbo; const char* what; };
  const Case cases[] = {
      {0, 2048, 128, "baseline: plane stride LBO, contiguous 8-row groups"},
      {160, 2048, 128, "shifted start"},
      {0, 16, 128, "LBO = 16: k-chunk 1 of row m = chunk 0 of row m + 1"},
      {48, 16, 1056, "LBO = 16, SBO = 1056 (two image rows of 33 pairs)"},
      {0, 528, 1088, "LBO = 528, SBO = 1088"},
      {32, 4096 - 32, 128, "LBO = plane - 32"},
      {0, 16, 2112, "LBO = 16, SBO = 2112"},
  };
  std::vector<float> o0(128 * 16), o1(128 * 16);
  for (const Case& c : cases) {
    probe<<<1, 128, smem>>>(d_out, d_cyc, 0, 0, c.start, c.lbo, c.sbo, 16, 1, 1);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(o0.data(), d_out, o0.size() * 4, cudaMemcpyDeviceToHost));
    probe<<<1, 128, smem>>>(d_out, d_cyc, 0, 1, c.start, c.lbo, c.sbo, 16, 1, 1);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(o1.data(), d_out, o1.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0, first_m = -1, first_n = -1; long long got0 = 0, exp0 = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 16; ++n) {
        const long long got = 2 * ((long long)o0[m * 16 + n] + 1024 * (long long)o1[m * 16 + n]);
        const long long exp = (long long)c.start + (m % 8) * 16 + (long long)(m / 8) * c.sbo + (long long)(n / 8) * c.lbo + (n % 8) * 2;
        if (got != exp && exp < A_BYTES) { if (!bad) { first_m = m; first_n = n; got0 = got; exp0 = exp; } ++bad; }
      }
    printf("%-58s start=%4u lbo=%5u sbo=%5u : %s", c.what, c.start, c.lbo, c.sbo, bad ? "MISMATCH" : "address formula holds");
    if (bad) printf("  (%d of 2048; first at m=%d k=%d: fetched byte %lld, expected %lld)", bad, first_m, first_n, got0, exp0);
    printf("\n");
  }
  // cost per MMA (issue to commit, 1024 MMAs round-robin over 4 (N = 256: 2) accumulators; index arithmetic by masks -- a runtime modulo in the issue loop costs more than the MMA)
  struct T { uint32_t lbo, sbo; int N; };
  const T ts[] = {{2048, 128, 16}, {2048, 128, 32}, {2048, 128, 64}, {16, 128, 16}, {16, 128, 32}, {16, 1056, 32}, {16, 1056, 64},
                  {16, 2112, 32}, {2048, 128, 128}, {2048, 128, 256}, {3072 - 32, 128, 16}};
  // dependent accumulation: consecutive MMAs into the SAME accumulator against round-robin over 2 / 4 / 8 accumulators
  for (int nd = 1; nd <= 8; nd *= 2)
    for (int N : {16, 32, 64}) {
      probe<<<1, 128, smem>>>(d_out, d_cyc, 1, 0, 0, 16, 1056, N, 1024, nd);
      CK(cudaDeviceSynchronize());
      long long cyc;
      CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
      printf("chain  M=128 N=%3d K=16 accumulators in rotation %d : %.1f cycles per MMA\n", N, nd, (double)cyc / 1024);
    }
  for (const T& t : ts) {
    const int n_mma = 1024, ndst = t.N >= 256 ? 2 : 4;
    probe<<<1, 128, smem>>>(d_out, d_cyc, 1, 0, 0, t.lbo, t.sbo, t.N, n_mma, ndst);
    CK(cudaDeviceSynchronize());
    long long cyc;
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("timing M=128 N=%3d K=16 lbo=%5u sbo=%5u : %.1f cycles per MMA\n", t.N, t.lbo, t.sbo, (double)cyc / n_mma);
  }
  return 0;
}
