// K12: cosine similarity of each face-bearing frame's embedding with the previous face-bearing one, and the
// `sim < threshold` test (reference server/model.py:60-62).  Frames without a face are skipped, so the comparison
// spans the gap; the embedding that precedes this range (previous batch, or previous rank's
// halo) is passed in as `halo`.  One warp per frame.
#include "common.cuh"

__global__ void __launch_bounds__(256) consistency_kernel(const float* __restrict__ emb, const uint8_t* __restrict__ valid, int B,
                                                         const float* __restrict__ halo, const uint8_t* __restrict__ halo_valid,
                                                         float thr, float* __restrict__ sim, uint8_t* __restrict__ below,
                                                         uint8_t* __restrict__ has_sim) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const int i = warp;
  float out_sim = __int_as_float(0x7fc00000);   // NaN = no comparison
  uint8_t out_below = 0, out_has = 0;
  if (valid[i]) {
    int j = i - 1;
    while (j >= 0 && !valid[j]) --j;
    const float* prev = nullptr;
    if (j >= 0) prev = emb + (size_t)j * TRL_EMB_DIM;
    else if (halo != nullptr && (halo_valid == nullptr || *halo_valid)) prev = halo;
    if (prev) {
      const float* cur = emb + (size_t)i * TRL_EMB_DIM;
      float dot = 0.f, na = 0.f, nb = 0.f;
      for (int k = lane; k < TRL_EMB_DIM; k += 32) {
        const float a = cur[k], b = prev[k];
        dot = fmaf(a, b, dot);
        na = fmaf(a, a, na);
        nb = fmaf(b, b, nb);
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        dot += __shfl_xor_sync(0xffffffffu, dot, d);
        na += __shfl_xor_sync(0xffffffffu, na, d);
        nb += __shfl_xor_sync(0xffffffffu, nb, d);
      }
      out_sim = dot / (sqrtf(na) * sqrtf(nb));    // np.dot / (np.linalg.norm * np.linalg.norm)
      out_below = out_sim < thr ? 1 : 0;
      out_has = 1;
    }
  }
  if (lane == 0) { sim[i] = out_sim; below[i] = out_below; has_sim[i] = out_has; }
}

// last face-bearing embedding of the range (or the incoming halo if the range has none) -> next range's halo
__global__ void __launch_bounds__(128) last_valid_kernel(const float* __restrict__ emb, const uint8_t* __restrict__ valid, int B,
                                                        const float* __restrict__ halo, const uint8_t* __restrict__ halo_valid,
                                                        float* __restrict__ last_emb, uint8_t* __restrict__ last_valid) {
  __shared__ int sj;
  if (threadIdx.x == 0) {
    int j = B - 1;
    while (j >= 0 && !valid[j]) --j;
    sj = j;
  }
  __syncthreads();
  const int j = sj;
  const bool from_halo = j < 0 && halo != nullptr && (halo_valid == nullptr || *halo_valid);
  if (last_emb) {
    for (int k = threadIdx.x; k < TRL_EMB_DIM; k += blockDim.x)
      last_emb[k] = j >= 0 ? emb[(size_t)j * TRL_EMB_DIM + k] : (from_halo ? halo[k] : 0.f);
  }
  if (last_valid && threadIdx.x == 0) *last_valid = (j >= 0 || from_halo) ? 1 : 0;
}

int launch_consistency(trl_ctx* c, const float* d_emb, const uint8_t* d_valid, int B, const float* d_halo,
                       const uint8_t* d_halo_valid, float thr,
                       float* d_sim, uint8_t* d_below, uint8_t* d_has_sim, float* d_last_emb, uint8_t* d_last_valid,
                       cudaStream_t s) {
  if (B <= 0) return TRL_OK;
  consistency_kernel<<<ceil_div(B * 32, 256), 256, 0, s>>>(d_emb, d_valid, B, d_halo, d_halo_valid, thr, d_sim, d_below, d_has_sim);
  TRL_LAUNCH_CHECK(c);
  if (d_last_emb || d_last_valid) {
    last_valid_kernel<<<1, 128, 0, s>>>(d_emb, d_valid, B, d_halo, d_halo_valid, d_last_emb, d_last_valid);
    TRL_LAUNCH_CHECK(c);
  }
  return TRL_OK;
}
