#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred px;\nelect.sync _|px, %1;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__global__ void __launch_bounds__(256) k(uint32_t* slot_g, int n_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp == 0) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
  __syncthreads();
  if (warp == 4) {
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const uint32_t sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint32_t idesc = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t dbase = ((uint64_t)((512u >> 4) & 0x3FFFu) << 16) | ((uint64_t)((128u >> 4) & 0x3FFFu) << 32) | (1ull << 46);
    if (elect_one()) {
      for (int t = 0; t < n_tiles; ++t) {
#pragma unroll
        for (int term = 0; term < 3; ++term)
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a = tmem + 64 + (term == 1 ? 72 : 0) + tap * 8;
            const uint32_t baddr = sb + (term == 2 ? 9 * 1024 : 0) + tap * 1024;
            mma_ts(tmem, a, dbase | (uint64_t)((baddr >> 4) & 0x3FFFu), idesc, (term | tap) != 0);
          }
        mma_commit(&bar);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) slot_g[0] = tmem_slot;
}
