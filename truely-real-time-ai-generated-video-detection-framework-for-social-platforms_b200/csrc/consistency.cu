// K12: cosine similarity of each face-bearing frame's embedding with the previous face-bearing one, and the
// `sim < threshold` test (reference server/model.py:60-62).  Frames without a face are skipped, so the comparison
// spans the gap; the embedding that precedes this range (previous batch, or previous rank's
// halo) is passed in as `halo`.  One warp per frame.
//
// Many-clip batches (BASELINE.json configs[4], SURVEY.md 8d "Config 5"): `previous_face_encoding` and the run-length
// counter are locals of one run() call (server/model.py:37-39), so the chain is cut at every clip boundary.
// `clip_start[i] != 0` marks the first processed frame of a clip: a frame is compared only with a face-bearing frame of
// its own clip, and a halo is used only if no clip starts between the beginning of the range and the frame.
//
// Frame-range sharding: shard_pack_kernel condenses one rank's results into a fixed-size record (per
// frame flags + the embeddings at both ends of the range), ONE all-gather moves the records, shard_resolve_kernel then
// finishes the comparisons that cross shard boundaries on every rank's copy of the gathered buffer -- the only flag of
// a shard that depends on another shard is the one of its first face-bearing frame.
#include "common.cuh"

// cos(a, b) exactly as the reference evaluates it: np.dot / (np.linalg.norm * np.linalg.norm), fp32, fixed reduction
// order (lane-strided partial sums, xor butterfly) so that every kernel that compares a pair gets the same bits.
__device__ __forceinline__ float warp_cosine(const float* __restrict__ cur, const float* __restrict__ prev, int lane) {
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
  return dot / (sqrtf(na) * sqrtf(nb));
}

__global__ void __launch_bounds__(256) consistency_kernel(const float* __restrict__ emb, const uint8_t* __restrict__ valid, int B,
                                                         const uint8_t* __restrict__ clip_start,
                                                         const float* __restrict__ halo, const uint8_t* __restrict__ halo_valid,
                                                         float thr, float* __restrict__ sim, uint8_t* __restrict__ below,
                                                         uint8_t* __restrict__ has_sim) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const int i = warp;
  float out_sim = __int_as_float(0x7fc00000);   // NaN = no comparison
  uint8_t out_below = 0, out_has = 0;
  if (valid[i]) {
    // previous face-bearing frame of the same clip: walk back until one is found or a clip boundary is crossed
    const float* prev = nullptr;
    int j = i;
    bool cut = false;
    while (true) {
      if (clip_start && clip_start[j]) { cut = true; break; }     // frame j opens its clip: nothing before it counts
      if (--j < 0) break;
      if (valid[j]) { prev = emb + (size_t)j * TRL_EMB_DIM; break; }
    }
    if (!prev && !cut && halo != nullptr && (halo_valid == nullptr || *halo_valid)) prev = halo;
    if (prev) {
      out_sim = warp_cosine(emb + (size_t)i * TRL_EMB_DIM, prev, lane);
      out_below = out_sim < thr ? 1 : 0;
      out_has = 1;
    }
  }
  if (lane == 0) { sim[i] = out_sim; below[i] = out_below; has_sim[i] = out_has; }
}

// last face-bearing embedding of the range's last clip (or the incoming halo if the range has neither a face nor a clip
// start) -> next range's halo
__global__ void __launch_bounds__(128) last_valid_kernel(const float* __restrict__ emb, const uint8_t* __restrict__ valid, int B,
                                                        const uint8_t* __restrict__ clip_start,
                                                        const float* __restrict__ halo, const uint8_t* __restrict__ halo_valid,
                                                        float* __restrict__ last_emb, uint8_t* __restrict__ last_valid) {
  __shared__ int sj, scut;
  if (threadIdx.x == 0) {
    int j = B - 1, cut = 0;
    while (j >= 0 && !valid[j]) {
      if (clip_start && clip_start[j]) { cut = 1; break; }
      --j;
    }
    sj = cut ? -1 : j;
    scut = cut;
  }
  __syncthreads();
  const int j = sj;
  const bool from_halo = j < 0 && !scut && halo != nullptr && (halo_valid == nullptr || *halo_valid);
  if (last_emb) {
    for (int k = threadIdx.x; k < TRL_EMB_DIM; k += blockDim.x)
      last_emb[k] = j >= 0 ? emb[(size_t)j * TRL_EMB_DIM + k] : (from_halo ? halo[k] : 0.f);
  }
  if (last_valid && threadIdx.x == 0) *last_valid = (j >= 0 || from_halo) ? 1 : 0;
}

int launch_consistency(trl_ctx* c, const float* d_emb, const uint8_t* d_valid, int B, const uint8_t* d_clip_start,
                       const float* d_halo, const uint8_t* d_halo_valid, float thr,
                       float* d_sim, uint8_t* d_below, uint8_t* d_has_sim, float* d_last_emb, uint8_t* d_last_valid,
                       cudaStream_t s) {
  if (B <= 0) return TRL_OK;
  consistency_kernel<<<ceil_div(B * 32, 256), 256, 0, s>>>(d_emb, d_valid, B, d_clip_start, d_halo, d_halo_valid, thr, d_sim,
                                                           d_below, d_has_sim);
  TRL_LAUNCH_CHECK(c);
  if (d_last_emb || d_last_valid) {
    last_valid_kernel<<<1, 128, 0, s>>>(d_emb, d_valid, B, d_clip_start, d_halo, d_halo_valid, d_last_emb, d_last_valid);
    TRL_LAUNCH_CHECK(c);
  }
  return TRL_OK;
}

// ----------------------------------------------------------------------------- shard records (multi-GPU exchange)
// Record of one rank, `trl_shard_record_bytes(n_max)` bytes, 16-byte aligned fields:
//   int32 hdr[8]   : n_local, first_idx, last_has, blocked, reserved x4
//   float first_emb[512] : embedding of frame first_idx
//   float last_emb[512]  : the range's outgoing halo
//   uint8 flags[3][n_pad]: valid, has_sim, below of the local frames (n_pad = n_max rounded up to 16)
// first_idx = the first face-bearing frame of the range IF it is still waiting for a predecessor from an earlier range
//   (no clip starts at or before it inside the range), else -1;
// last_has  = the range ends with a face-bearing frame of its last clip (last_emb is meaningful);
// blocked   = the range contains a clip start with no face-bearing frame after it: nothing from this range or from any
//             earlier one may be handed on to later ranges.
#define SHARD_HDR_INTS 8
__host__ __device__ inline size_t shard_pad(int n_max) { return ((size_t)n_max + 15) & ~(size_t)15; }
__host__ __device__ inline size_t shard_record_bytes(int n_max) {
  return SHARD_HDR_INTS * sizeof(int) + 2 * TRL_EMB_DIM * sizeof(float) + 3 * shard_pad(n_max);
}

__global__ void __launch_bounds__(256) shard_pack_kernel(const float* __restrict__ emb, const uint8_t* __restrict__ valid,
                                                        const uint8_t* __restrict__ has_sim, const uint8_t* __restrict__ below,
                                                        const uint8_t* __restrict__ clip_start, int n_local, int n_max,
                                                        unsigned char* __restrict__ rec) {
  __shared__ int s_first, s_last, s_blocked;
  int* hdr = reinterpret_cast<int*>(rec);
  float* first_emb = reinterpret_cast<float*>(rec + SHARD_HDR_INTS * sizeof(int));
  float* last_emb = first_emb + TRL_EMB_DIM;
  unsigned char* flags = reinterpret_cast<unsigned char*>(last_emb + TRL_EMB_DIM);
  const size_t np = shard_pad(n_max);
  if (threadIdx.x == 0) {
    int first = -1;
    for (int i = 0; i < n_local; ++i) {
      if (clip_start && clip_start[i]) break;          // the clip that reaches in from the previous range ends here
      if (valid[i]) { first = i; break; }
    }
    int j = n_local - 1, blocked = 0;
    while (j >= 0 && !valid[j]) {
      if (clip_start && clip_start[j]) { blocked = 1; break; }
      --j;
    }
    if (blocked) j = -1;
    // a face-bearing frame that itself opens a clip hands its embedding on, but blocks everything before it
    s_first = first; s_last = j; s_blocked = blocked;
    hdr[0] = n_local; hdr[1] = first; hdr[2] = j >= 0 ? 1 : 0;
    // `blocked` for the chain search of later ranks: a clip start anywhere after the last face-bearing frame -- or, when
    // that frame exists, nothing (it supplies the halo itself)
    hdr[3] = blocked; hdr[4] = hdr[5] = hdr[6] = hdr[7] = 0;
  }
  __syncthreads();
  const int first = s_first, last = s_last;
  for (int k = threadIdx.x; k < TRL_EMB_DIM; k += blockDim.x) {
    first_emb[k] = first >= 0 ? emb[(size_t)first * TRL_EMB_DIM + k] : 0.f;
    last_emb[k] = last >= 0 ? emb[(size_t)last * TRL_EMB_DIM + k] : 0.f;
  }
  for (int i = threadIdx.x; i < (int)np; i += blockDim.x) {
    const bool in = i < n_local;
    flags[i] = in ? valid[i] : 0;
    flags[np + i] = in ? has_sim[i] : 0;
    flags[2 * np + i] = in ? below[i] : 0;
  }
}

// One warp per rank q >= 1: find the nearest earlier rank whose range ends with a usable embedding (stopping at a blocked
// range), compare it with rank q's first pending frame and patch that frame's flags in the gathered buffer; the warp of
// q == my_rank also patches the local per-frame outputs.
__global__ void __launch_bounds__(32) shard_resolve_kernel(unsigned char* __restrict__ all, int world, int my_rank, int n_max,
                                                          float thr, float* __restrict__ sim_local,
                                                          uint8_t* __restrict__ below_local, uint8_t* __restrict__ has_local) {
  const int q = blockIdx.x + 1, lane = threadIdx.x;
  if (q >= world) return;
  const size_t rb = shard_record_bytes(n_max), np = shard_pad(n_max);
  unsigned char* rec = all + (size_t)q * rb;
  const int* hdr = reinterpret_cast<const int*>(rec);
  const int first = hdr[1];
  if (first < 0) return;
  int src = -1;
  for (int r = q - 1; r >= 0; --r) {
    const int* h = reinterpret_cast<const int*>(all + (size_t)r * rb);
    if (h[2]) { src = r; break; }
    if (h[3]) break;                   // a clip boundary with no face after it: the chain does not reach further back
  }
  if (src < 0) return;
  const float* cur = reinterpret_cast<const float*>(rec + SHARD_HDR_INTS * sizeof(int));
  const float* prev = reinterpret_cast<const float*>(all + (size_t)src * rb + SHARD_HDR_INTS * sizeof(int)) + TRL_EMB_DIM;
  const float s = warp_cosine(cur, prev, lane);
  if (lane == 0) {
    unsigned char* flags = rec + SHARD_HDR_INTS * sizeof(int) + 2 * TRL_EMB_DIM * sizeof(float);
    const uint8_t b = s < thr ? 1 : 0;
    flags[np + first] = 1;
    flags[2 * np + first] = b;
    if (q == my_rank) {
      if (sim_local) sim_local[first] = s;
      if (below_local) below_local[first] = b;
      if (has_local) has_local[first] = 1;
    }
  }
}

size_t shard_record_size(int n_max) { return shard_record_bytes(n_max); }

int launch_shard_pack(trl_ctx* c, const float* d_emb, const uint8_t* d_valid, const uint8_t* d_has_sim, const uint8_t* d_below,
                      const uint8_t* d_clip_start, int n_local, int n_max, unsigned char* d_record, cudaStream_t s) {
  shard_pack_kernel<<<1, 256, 0, s>>>(d_emb, d_valid, d_has_sim, d_below, d_clip_start, n_local, n_max, d_record);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

int launch_shard_resolve(trl_ctx* c, unsigned char* d_all, int world, int my_rank, int n_max, float thr, float* d_sim,
                         uint8_t* d_below, uint8_t* d_has_sim, cudaStream_t s) {
  if (world <= 1) return TRL_OK;
  shard_resolve_kernel<<<world - 1, 32, 0, s>>>(d_all, world, my_rank, n_max, thr, d_sim, d_below, d_has_sim);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}
