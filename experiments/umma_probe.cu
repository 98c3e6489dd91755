// Probe for the P-Net tensor-core design (not product code):
//  1. semantics of the no-swizzle K-major shared-memory descriptor with kind::tf32 (A = [chunk][pixel][4] image,
//     filter taps = start-address shifts, tap/chunk pairing through LBO),
//  2. error of the 3xTF32 split against fp64,
//  3. issue/throughput cost of many small-N MMAs from one thread.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

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
// no-swizzle K-major descriptor: start>>4 | LBO>>4 @16 | SBO>>4 @32 | version 1 @46 | layout 0 @61
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// conv as shifted GEMM.  image: [CC chunks][NPIX][4] floats (hi and lo copies), weights [taps][CC][N][4] (hi, lo)
// D[m][n] = sum_{tap, c, j} img[c][m + shift[tap]][j] * w[tap][c][n][j]       m in [0,128)
struct ProbeParams {
  int CC, NPIX, N, taps, split;   // split: 1 = plain tf32 (hi only), 3 = 3xTF32
  int shift[9];
  int reps;                       // timing: repeat the whole MMA sequence reps times (accumulating)
};

__global__ void __launch_bounds__(128) probe_kernel(const float* img_hi, const float* img_lo, const float* w_hi,
                                                    const float* w_lo, float* out, long long* cycles, ProbeParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int img_floats = p.CC * p.NPIX * 4;
  const int w_floats = p.taps * p.CC * p.N * 4;
  float* s_img_hi = reinterpret_cast<float*>(smem);
  float* s_img_lo = s_img_hi + img_floats;
  float* s_w_hi = s_img_lo + img_floats;
  float* s_w_lo = s_w_hi + w_floats;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_w_lo + w_floats);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < img_floats; i += 128) { s_img_hi[i] = img_hi[i]; s_img_lo[i] = img_lo[i]; }
  for (int i = tid; i < w_floats; i += 128) { s_w_hi[i] = w_hi[i]; s_w_lo[i] = w_lo[i]; }
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA) reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  // idesc: c=f32 (1<<4), a=b=tf32 (2<<7, 2<<10), K-major both, N>>3 @17, M>>4 @24
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
  long long t0 = 0, t1 = 0, t2 = 0;
  if (tid == 0) {
    t0 = clock64();
    uint32_t acc = 0;
    const uint32_t plane = (uint32_t)p.NPIX * 16u;        // bytes between channel chunks of the image
    const uint32_t wplane = (uint32_t)p.N * 16u;          // bytes between k-chunks of the weights
    for (int r = 0; r < p.reps; ++r)
      for (int s = 0; s < p.split; ++s) {
        const float* ai = (s == 1) ? s_img_lo : s_img_hi;   // s=0: hi*hi, s=1: lo*hi, s=2: hi*lo
        const float* wi = (s == 2) ? s_w_lo : s_w_hi;
        for (int t = 0; t < p.taps; ++t)
          for (int c = 0; c < p.CC; c += 2) {
            const uint64_t ad = make_desc(smem_u32(ai) + (uint32_t)p.shift[t] * 16u + (uint32_t)c * plane, plane, 128u);
            const uint64_t bd = make_desc(smem_u32(wi) + (uint32_t)(t * p.CC + c) * wplane, wplane, 128u);
            mma_tf32(tmem, ad, bd, idesc, acc);
            acc = 1;
          }
      }
    t1 = clock64();
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  if (tid == 0) { t2 = clock64(); cycles[0] = t1 - t0; cycles[1] = t2 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    for (int j = 0; j < 16; ++j) out[(size_t)tid * p.N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

static float tf32_rn(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  float r; memcpy(&r, &u, 4); return r;
}

static void run_case(const char* name, int CC, int N, int taps, int pitch, int split, int reps, bool exact_inputs) {
  ProbeParams p{};
  p.CC = CC; p.N = N; p.taps = taps; p.split = split; p.reps = reps;
  int maxshift = 0;
  for (int t = 0; t < taps; ++t) { p.shift[t] = (taps == 1) ? 0 : (t / 3) * pitch + (t % 3); if (p.shift[t] > maxshift) maxshift = p.shift[t]; }
  p.NPIX = 128 + maxshift;
  const int img_n = CC * p.NPIX * 4, w_n = taps * CC * N * 4;
  std::vector<float> img(img_n), w(w_n), ih(img_n), il(img_n), wh(w_n), wl(w_n);
  srand(1234);
  for (auto& v : img) v = exact_inputs ? (float)((rand() % 17) - 8) / 8.f : ((float)rand() / RAND_MAX * 2.f - 1.f);
  for (auto& v : w) v = exact_inputs ? (float)((rand() % 9) - 4) / 4.f : ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.3f;
  for (int i = 0; i < img_n; ++i) { ih[i] = tf32_rn(img[i]); il[i] = img[i] - ih[i]; }
  for (int i = 0; i < w_n; ++i) { wh[i] = tf32_rn(w[i]); wl[i] = w[i] - wh[i]; }
  float *d_ih, *d_il, *d_wh, *d_wl, *d_out; long long* d_cyc;
  CK(cudaMalloc(&d_ih, img_n * 4)); CK(cudaMalloc(&d_il, img_n * 4)); CK(cudaMalloc(&d_wh, w_n * 4)); CK(cudaMalloc(&d_wl, w_n * 4));
  CK(cudaMalloc(&d_out, 128 * N * 4)); CK(cudaMalloc(&d_cyc, 16));
  CK(cudaMemcpy(d_ih, ih.data(), img_n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_il, il.data(), img_n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_wh, wh.data(), w_n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_wl, wl.data(), w_n * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(2 * img_n + 2 * w_n) * 4 + 64;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  probe_kernel<<<1, 128, smem>>>(d_ih, d_il, d_wh, d_wl, d_out, d_cyc, p);
  CK(cudaDeviceSynchronize());
  std::vector<float> out(128 * N); long long cyc[2];
  CK(cudaMemcpy(out.data(), d_out, 128 * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int t = 0; t < taps; ++t)
        for (int c = 0; c < CC; ++c)
          for (int j = 0; j < 4; ++j) {
            const double a = (split == 1 ? ih : img)[(c * p.NPIX + m + p.shift[t]) * 4 + j];
            const double b = (split == 1 ? wh : w)[((t * CC + c) * N + n) * 4 + j];
            ref += a * b;
          }
      ref *= reps;
      maxerr = fmax(maxerr, fabs(ref - out[m * N + n])); maxref = fmax(maxref, fabs(ref));
    }
  const int n_mma = reps * split * taps * (CC / 2);
  printf("%-34s N=%3d K=%4d split=%d mmas=%5d  max|err|=%.3e (max|ref|=%.2f)  issue=%.1f cyc/mma  total=%.1f cyc/mma\n", name, N,
         taps * CC * 4, split, n_mma, maxerr, maxref, (double)cyc[0] / n_mma, (double)cyc[1] / n_mma);
  cudaFree(d_ih); cudaFree(d_il); cudaFree(d_wh); cudaFree(d_wl); cudaFree(d_out); cudaFree(d_cyc);
}


__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0, lane = 0;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, %2;\n@px mov.s32 %1, 1;\nmov.s32 %0, rx;\n}\n"
               : "+r"(lane), "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}
// ---- throughput-only kernel: dense [128 x KB bytes] A tile and [N x KB] B tile, layout selectable
//  mode 0: no-swizzle, tf32   (A = [chunk][row][16B], LBO = plane, SBO = 128)
//  mode 1: no-swizzle, bf16 kind::f16
//  mode 2: SWIZZLE_128B tf32  (rows of 128 B, SBO = 1024)   (data garbage, timing only)
//  mode 3: SWIZZLE_128B bf16
template <bool ELECT>
__global__ void __launch_bounds__(128) tput_kernel(long long* cycles, int mode, int N, int n_mma, int ndst, int M, int use_elect) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = base;                 // 128 rows x 128 B = 16 KB
  uint8_t* sb = base + 16384;         // up to 256 rows x 128 B = 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 16384 + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0;
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const bool is16 = (mode == 1 || mode == 3);
  const uint32_t fmt = is16 ? 1u : 2u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (((uint32_t)M >> 4) << 24);
  uint64_t ad[4], bd[4];
  for (int k = 0; k < 4; ++k) {
    if (mode < 2) {
      ad[k] = make_desc(smem_u32(sa) + k * 2 * 2048, 2048, 128);          // chunk plane = 128 rows * 16 B
      bd[k] = make_desc(smem_u32(sb) + k * 2 * (N * 16), N * 16, 128);
    } else {
      ad[k] = (make_desc(smem_u32(sa) + k * 32, 16, 1024)) | (2ull << 61);
      bd[k] = (make_desc(smem_u32(sb) + k * 32, 16, 1024)) | (2ull << 61);
    }
  }
  if (ELECT ? (warp == 0 && elect_one()) : (tid == 0)) {
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
      const uint32_t d = tmem + (uint32_t)(((i >> 2) % ndst) * N);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (is16)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                       ::"r"(d), "l"(ad[k]), "l"(bd[k]), "r"(idesc), "r"(1u) : "memory");
        else
          mma_tf32(d, ad[k], bd[k], idesc, 1u);
      }
    }
    const long long t1 = clock64();
    mma_commit(bar);
    cycles[0] = t1 - t0;
    cycles[2] = t0;
  }
  __syncwarp();
  mbar_wait(bar, 0);
  if (tid == 0) cycles[1] = clock64() - cycles[2];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static void run_tput(int mode, int N, int ndst, int M = 128, int ctas = 1, int use_elect = 1) {
  long long* d_cyc; CK(cudaMalloc(&d_cyc, 32));
  CK(cudaFuncSetAttribute(tput_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaFuncSetAttribute(tput_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int n_mma = 4000;
  if (use_elect) tput_kernel<true><<<ctas, 128, 16384 + 32768 + 1024 + 64>>>(d_cyc, mode, N, n_mma, ndst, M, use_elect);
  else tput_kernel<false><<<ctas, 128, 16384 + 32768 + 1024 + 64>>>(d_cyc, mode, N, n_mma, ndst, M, use_elect);
  CK(cudaDeviceSynchronize());
  long long cyc[3]; CK(cudaMemcpy(cyc, d_cyc, 24, cudaMemcpyDeviceToHost));
  const char* names[4] = {"noswz tf32", "noswz bf16", "sw128 tf32", "sw128 bf16"};
  printf("tput elect=%d %-10s M=%3d N=%3d ndst=%d ctas=%d: issue %.1f  total %.1f cyc/mma (floor N/2=%d, A-read 32)\n", use_elect, names[mode], M, N, ndst, ctas,
         (double)cyc[0] / n_mma, (double)cyc[1] / n_mma, N / 2);
  cudaFree(d_cyc);
}

int main() {
  for (int mode = 0; mode < 4; ++mode)
    for (int N : {16, 32, 128, 256}) run_tput(mode, N, 1);
  run_tput(0, 32, 4); run_tput(2, 32, 4); run_tput(3, 256, 2); run_tput(0, 32, 1, 64); run_tput(2, 32, 1, 64);
  run_tput(0, 32, 1, 128, 2); run_tput(0, 32, 1, 128, 1, 0);
  return 0;

  // 1. semantics with exactly representable inputs (error must be 0)
  run_case("gemm exact", 4, 16, 1, 0, 1, 1, true);
  run_case("gemm exact N32", 8, 32, 1, 0, 1, 1, true);
  run_case("conv3x3 exact (pitch 64)", 4, 32, 9, 64, 1, 1, true);
  run_case("conv3x3 exact (pitch 34)", 4, 16, 9, 34, 1, 1, true);
  // 2. precision on random fp32 data
  run_case("conv3x3 random tf32x1", 4, 32, 9, 64, 1, 1, false);
  run_case("conv3x3 random tf32x3", 4, 32, 9, 64, 3, 1, false);
  // 3. throughput (results accumulate reps times; error columns are still meaningful for split=3)
  run_case("throughput N16", 4, 16, 9, 64, 3, 40, false);
  run_case("throughput N32", 4, 32, 9, 64, 3, 40, false);
  run_case("throughput N64", 4, 64, 9, 64, 3, 40, false);
  run_case("throughput N128", 4, 128, 9, 64, 1, 40, false);
  return 0;
}
