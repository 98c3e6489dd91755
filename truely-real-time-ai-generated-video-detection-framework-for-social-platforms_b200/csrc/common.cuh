// Shared declarations of libtruely_b200 (sm_100a).  See include/truely_b200.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/truely_b200.h"


// ----------------------------------------------------------------------------- candidates
// One detection candidate as it travels through the cascade (40 bytes).
struct __align__(8) Cand {
  float x1, y1, x2, y2;   // box
  float score;
  float r0, r1, r2, r3;   // regression offsets of the producing net
  uint32_t key;           // tie-break key: position in the upstream list at this stage
};

// One P-Net cell kept by the single-pass screen for exact re-evaluation (pnet_refine.cu)
struct ScreenEntry {
  int group;   // frame * n_levels + level
  int cell;    // oy * ow + ox
};

// capacity overflow record, written by kernels into pinned mapped host memory
struct CapFlag {
  int overflow;   // 0 = ok
  int stage;      // 1 = pnet per-scale, 2 = stage-1 per-frame, 3 = rnet per-frame, 4 = final per-frame
  int frame;
  int count;
  int capacity;
};

struct PyramidGeom {
  int n;
  double scale[TRL_MAX_SCALES];
  float scale_f[TRL_MAX_SCALES];
  int hs[TRL_MAX_SCALES], ws[TRL_MAX_SCALES];   // resampled image size
  int pitch[TRL_MAX_SCALES];                    // row pitch (floats) of level k in the cascade's pyramid buffer: ws rounded up to 4
  int oh[TRL_MAX_SCALES], ow[TRL_MAX_SCALES];   // P-Net output map size
  long long off[TRL_MAX_SCALES];                // float offset of level k in the pyramid buffer, per frame count B: off*B
  int pitch2[TRL_MAX_SCALES];                   // fp16 pair images (pnet2.cu): pixel PAIRS (16 bytes: 2 x BGR0 halves) per row = ceil(ws / 2)
  long long off2[TRL_MAX_SCALES];               // pair offset of level k in the hi / lo pair images, per frame count B: off2*B
  long long pairs_total;                        // pairs per frame of one pair image: sum hs*pitch2
  long long px_total;                           // sum hs*ws
  long long floats_total;                       // floats per frame of the padded pyramid buffer: sum 3*hs*pitch
};

// kernel-parameter view of the pyramid (passed by value)
struct PyrParams {
  int n;
  int hs[TRL_MAX_SCALES], ws[TRL_MAX_SCALES];
  int oh[TRL_MAX_SCALES], ow[TRL_MAX_SCALES];
  float scale[TRL_MAX_SCALES];
  long long off[TRL_MAX_SCALES];       // float offset (already multiplied by B) of level k
  int blk_start[TRL_MAX_SCALES + 1];   // prefix of work blocks per level (kernel specific)
  int pitch[TRL_MAX_SCALES];           // output row pitch in floats (pyramid kernel)
  int rows[TRL_MAX_SCALES];            // output rows per CTA (pyramid kernel)
  int fastdiv[TRL_MAX_SCALES];         // 1: the verified 3-instruction division may be used for this level
  int tab_off[TRL_MAX_SCALES];         // offset of level k's window tables
  int kwmin[TRL_MAX_SCALES];           // narrowest horizontal window of level k; > 0: every window is kwmin or kwmin + 1 columns
                                       // wide and kwmin <= 5 (branch-free pass 2), 0: general loop
  unsigned magic[TRL_MAX_SCALES];      // floor(2^32 / ws) + 1: flat pixel index of a CTA -> row by one multiply-high
};


#ifdef __CUDACC__
// sm_100 packed fp32: two independent IEEE fmas per instruction (SASS FFMA2; a (v, v) pair becomes a scalar broadcast operand)
__device__ __forceinline__ unsigned long long pack_f32x2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ void fadd2(unsigned long long& d, unsigned long long a) {
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(d) : "l"(a));
}
#endif

int tma_encode_tiled_f32(trl_ctx* c, void* map, const void* base, int rank, const unsigned long long* dims,
                         const unsigned long long* strides_bytes, const unsigned* box);   // facenet_umma.cu

// ----------------------------------------------------------------------------- compact work lists
// The cascade keeps per-frame candidate lists at fixed strides (slot = frame * cap + i, count[frame] live entries).
// Kernels that do one CTA of work per live slot run as grid-stride loops over the *compacted* index space instead of
// launching one (mostly empty) CTA per slot: every CTA scans the per-frame counts once (B <= SLOTMAP_MAX_FRAMES ints)
// and maps work item j to its slot by binary search.
#define SLOTMAP_MAX_FRAMES 1024
#ifdef __CUDACC__
// all threads of the CTA call this; s_pref holds n_frames + 1 ints.  Returns the number of live slots.
__device__ __forceinline__ int slotmap_init(const int* __restrict__ d_count, int n_frames, int cap, int* s_pref) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int carry = 0;
    if (lane == 0) s_pref[0] = 0;
    for (int base = 0; base < n_frames; base += 32) {
      const int i = base + lane;
      int v = i < n_frames ? min(max(d_count[i], 0), cap) : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      if (i < n_frames) s_pref[i + 1] = carry + v;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  return s_pref[n_frames];
}
// work item j (0 <= j < total) -> slot = frame * cap + index
__device__ __forceinline__ int slotmap_slot(const int* s_pref, int n_frames, int cap, int j) {
  int lo = 0, hi = n_frames;           // invariant: s_pref[lo] <= j < s_pref[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (s_pref[mid] <= j) lo = mid; else hi = mid;
  }
  return lo * cap + (j - s_pref[lo]);
}
#endif

namespace nms {
// parameters of one cascade NMS stage (nms.cu)
struct StageParams {
  int n_levels;          // stage 1
  int cap_in, cap_out;
  float thr_nms, thr_score;
  int W, H;
  const Cand* in;        // stage1: [B][L][cap_in]; others: [B][cap_in]
  const int* cnt_in;
  const float* prob;     // stage 3/4: net outputs [B][cap_in]
  const float* reg;      //            [B][cap_in][4]
  const int* pad_in;     // stage 3/4: pad of the inputs (degenerate crops are dropped)
  Cand* out;             // [B][cap_out]
  int* cnt_out;          // [B]
  int* pad_out;          // [B][cap_out][4]
  float* boxes_out;      // stage 4: [B][cap_out][5]
  CapFlag* capflag;
};
}  // namespace nms

struct FaceNetEngine;   // facenet_plan.cu

struct trl_ctx {
  int device = 0;
  int num_sms = 0;               // cudaDeviceProp::multiProcessorCount of `device` (148 on B200): persistent-kernel grid size
  trl_config_t cfg{};
  std::string err;
  long long launches = 0;

  // weights (device)
  float h_pnet_head[32 * 8 + 8 + 1 + 20] = {0};   // P-Net head / conv3 epilogue constants, copied into the kernel parameters
  float* d_pnet_packed = nullptr;   // smem image of P-Net (pnet.cu layout)
  int pnet2_range_ok = 0;           // host-side bound of the P-Net activations fits fp16 (pnet2.cu)
  float h_pnet2_epi[152] = {0};     // pnet2.cu epilogue constants (biases, slopes, conv4_1 logit-difference weights)
  float* d_pnet_refine = nullptr;   // fp32 weight image of the exact per-cell kernel (pnet_refine.cu)
  void* d_pnet2_tiles = nullptr;    // pnet2.cu: tile table of the current frame geometry (int4 per tile of one frame)
  int pnet2_tiles_n = 0; long long pnet2_tiles_key = 0;
  uint32_t* d_pnet2_packed = nullptr;   // smem image of the all-tensor-pipe screening P-Net (pnet2.cu)
  float* d_rnet = nullptr;          // packed R-Net (mtcnn_ro.cu layout)
  float* d_onet = nullptr;
  FaceNetEngine* facenet = nullptr;
  void* overlay = nullptr;          // overlay::Stamps (overlay.cu)

  // workspace for trl_detect / trl_process, sized for (ws_B, ws_H, ws_W)
  int ws_B = 0, ws_H = 0, ws_W = 0;
  PyramidGeom geom{};
  float* d_pyr = nullptr;        // fp32 planar pyramid (pnet_precision 0-2)
  uint4* d_pyr_hi = nullptr;     // pnet_precision 3: fp16 hi / lo pair images, [level][B][hs][pitch2] x 16 bytes
  uint4* d_pyr_lo = nullptr;
  ScreenEntry* d_screen = nullptr;   // cells kept by the single-pass screen (pnet_precision 2, 3)
  int* d_screen_cnt = nullptr;
  int screen_cap = 0;
  size_t pyr_smem_set = 0;       // dynamic shared memory the pyramid kernel has been opted in to
  int* d_pyr_tab = nullptr;      // adaptive-average window tables of the current frame shape
  void* d_pyr_blk = nullptr;     // pyramid kernel: (level, first output row) of every CTA of one frame
  long long pyr_blk_key = 0;
  int pyr_tab_H = 0, pyr_tab_W = 0;
  int pyr_tab_off[TRL_MAX_SCALES] = {0};
  int pyr_fastdiv[TRL_MAX_SCALES] = {0};
  int pyr_kwmin[TRL_MAX_SCALES] = {0};
  Cand* d_cand1 = nullptr;       // [B][n_scales][cand_cap_scale]   P-Net candidates
  int* d_cnt1 = nullptr;         // [B][n_scales]
  Cand* d_cand2 = nullptr;       // [B][cand_cap_frame]             after per-scale NMS
  int* d_cnt2 = nullptr;         // [B]
  Cand* d_cand3 = nullptr;       // [B][cand_cap_frame]             R-Net inputs (after cross-scale NMS)
  int* d_cnt3 = nullptr;         // [B]
  int* d_pad3 = nullptr;         // [B][cand_cap_frame][4]
  float* d_rin = nullptr;        // [B*cand_cap_frame][3][24][24]
  Cand* d_cand4 = nullptr;       // [B][box_cap_frame]              O-Net inputs
  int* d_cnt4 = nullptr;         // [B]
  int* d_pad4 = nullptr;
  float* d_oin = nullptr;        // [B*box_cap_frame][3][48][48]
  float* d_rprob = nullptr;      // [B][cand_cap_frame]   R-Net outputs
  float* d_rreg = nullptr;       // [B][cand_cap_frame][4]
  float* d_oprob = nullptr;      // [B][box_cap_frame]    O-Net outputs
  float* d_oreg = nullptr;
  // process scratch
  float* d_boxes = nullptr;      // [B][box_cap_frame][5]
  int* d_nfaces = nullptr;
  uint8_t* d_crops = nullptr;    // [B][S][S][3]

  CapFlag* h_cap = nullptr;      // pinned, mapped
  CapFlag* d_cap = nullptr;      // device alias of h_cap

  // pipelined cascade (trl_detect_align_async): the latency-bound tail of chunk k (NMS, crops, R-Net, O-Net, crop-align)
  // runs on an internal high-priority stream under the pyramid of chunk k+1
  cudaStream_t tail_stream = nullptr;
  cudaEvent_t ev_head = nullptr, ev_tail = nullptr;
  bool tail_pending = false;     // ev_tail has been recorded and no caller stream has waited for it yet

  bool profiling = false;
  struct ProfRec { int stage; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof_events;

  // scratch for the stand-alone trl_nms entry point
  Cand* d_nms_tmp = nullptr; int nms_tmp_cap = 0;
  // global-memory scratch of cascade_nms_big_kernel (allocated with the workspace only when a capacity exceeds 2048)
  unsigned char* d_nms_big = nullptr;
};

#define TRL_FAIL(ctx, code, ...)                                   \
  do {                                                             \
    char _b[512];                                                  \
    snprintf(_b, sizeof(_b), __VA_ARGS__);                         \
    (ctx)->err = _b;                                               \
    return (code);                                                 \
  } while (0)

#define TRL_CUDA(ctx, expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) TRL_FAIL(ctx, TRL_E_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define TRL_LAUNCH_CHECK(ctx)                                                                 \
  do {                                                                                        \
    (ctx)->launches++;                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) TRL_FAIL(ctx, TRL_E_CUDA, "kernel launch: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------- stage launchers (internal)
int compute_geometry(const trl_config_t& cfg, int H, int W, PyramidGeom* g);

int launch_pyramid(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const PyramidGeom& g, float* d_out,
                   bool padded, cudaStream_t s);
int launch_crop_resample(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const int* d_pad, const int* d_img,
                         const int* d_count, int n_max, int size, float* d_out, cudaStream_t s);
int launch_crop_align(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const float* d_boxes, int box_stride,
                      const int* d_nfaces, int S, int* d_box_int, uint8_t* d_valid, uint8_t* d_crops, cudaStream_t s);

int launch_extract_face(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const float* d_boxes, int box_stride,
                        const int* d_nfaces, int S, int margin, int* d_box_int, uint8_t* d_valid, uint8_t* d_crops,
                        cudaStream_t s);
int launch_extract_faces_all(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const float* d_boxes, int box_stride,
                             int box_cap, const int* d_nfaces, int S, int margin, int max_faces, int* d_face_off,
                             int* d_face_frame, int* d_box_int, uint8_t* d_valid, uint8_t* d_crops, cudaStream_t s);

int pnet_pack_weights(trl_ctx* c, const float* h_pnet, size_t len);
// maps mode: d_prob/d_reg non-null, one level.  candidate mode: all levels, thresholded append.
int launch_pnet_maps(trl_ctx* c, const float* d_in, int B, int hs, int ws, float* d_prob, float* d_reg, cudaStream_t s);
int launch_pnet_candidates(trl_ctx* c, const float* d_pyr, int B, const PyramidGeom& g, float thr, Cand* d_cand,
                           int* d_cnt, int cap, cudaStream_t s);

// hybrid P-Net: single-pass screen (prob >= thr - TRL_SCREEN_MARGIN) + exact fp32 re-evaluation of the screened cells
#define TRL_SCREEN_MARGIN 0.05f
int launch_pnet_screen_v1(trl_ctx* c, const float* d_pyr, int B, const PyramidGeom& g, float thr_lo, ScreenEntry* d_screen,
                          int* d_screen_cnt, int screen_cap, cudaStream_t s);
int pnet_refine_pack_weights(trl_ctx* c, const float* h_pnet, size_t len);
int launch_pnet_refine(trl_ctx* c, int fmt, const void* d_pyr, const void* d_pyr_lo, int B, const PyramidGeom& g, float thr,
                       const ScreenEntry* d_screen, const int* d_screen_cnt, int screen_cap, Cand* d_cand, int* d_cnt, int cap,
                       cudaStream_t s);
int pnet2_pack_weights(trl_ctx* c, const float* h_pnet, size_t len);
int launch_pnet2(trl_ctx* c, const uint4* d_pyr_hi, int B, const PyramidGeom& g, float thr_lo, ScreenEntry* d_screen,
                 int* d_screen_cnt, int screen_cap, float* d_logit, cudaStream_t s);
int launch_pnet2_screen(trl_ctx* c, const uint4* d_pyr_hi, int B, const PyramidGeom& g, float thr_lo, ScreenEntry* d_screen,
                        int* d_screen_cnt, int screen_cap, cudaStream_t s);
// fp16 hi / lo pair images of every level (the cascade's pyramid when pnet_precision == 3)
int launch_pyramid_pairs(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const PyramidGeom& g, uint4* d_hi, uint4* d_lo,
                         cudaStream_t s);

int ro_pack_weights(trl_ctx* c, const float* h_rnet, size_t rlen, const float* h_onet, size_t olen);
int launch_rnet(trl_ctx* c, const float* d_in, int n_max, const int* d_count, float* d_prob, float* d_reg, cudaStream_t s);
int launch_onet(trl_ctx* c, const float* d_in, int n_max, const int* d_count, float* d_prob, float* d_reg, cudaStream_t s);

int facenet_create(trl_ctx* c, const float* h_blob, size_t len);
void facenet_destroy(trl_ctx* c);
// norm: 0 = F.to_tensor (x / 255, the reference, server/model.py:58); 1 = fixed_image_standardization ((x - 127.5) / 128)
int facenet_forward(trl_ctx* c, const uint8_t* d_crops, int n, int S, int norm, float* d_emb, cudaStream_t s,
                    const int* d_n = nullptr);
// only the crops with d_valid[i] != 0 are embedded (packed on the device, no host synchronisation); other rows of d_emb = 0
int facenet_forward_valid(trl_ctx* c, const uint8_t* d_crops, const uint8_t* d_valid, int n, int S, int norm, float* d_emb,
                          cudaStream_t s);

void overlay_destroy(trl_ctx* c);                                                                  // overlay.cu
int overlay_set_stamps(trl_ctx* c, const uint8_t* h_lut, int n_lut, const trl_stamp_t* h_stamps, int n_stamps,
                       const uint16_t* h_idx, long long n_idx, int digit_advance);
int launch_overlay(trl_ctx* c, uint8_t* d_frames, int B, int H, int W, const int* d_box, const uint8_t* d_state,
                   const int* d_frame_index, uint8_t* d_text_pending, cudaStream_t s);

int nms_init(trl_ctx* c);
int nms_max_n();
int nms_big_max_n();
size_t nms_big_scratch_bytes(int groups, int cap);
int launch_plain_nms(trl_ctx* c, const float* d_boxes, const float* d_scores, int n, float thr, int mode, int* d_keep,
                     int* d_nkeep, cudaStream_t s);
int launch_cascade_stage(trl_ctx* c, int stage, const nms::StageParams& p, int B, cudaStream_t s);
int launch_crop_resample_ex(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const int* d_pad, const int* d_img,
                            const int* d_count, int per_frame_cap, int n_slots, int size, float* d_out, cudaStream_t s);
int launch_rnet_ex(trl_ctx* c, const float* d_in, int n_slots, const int* d_count, int per_frame_cap, float* d_prob,
                   float* d_reg, cudaStream_t s);
int launch_onet_ex(trl_ctx* c, const float* d_in, int n_slots, const int* d_count, int per_frame_cap, float* d_prob,
                   float* d_reg, cudaStream_t s);

int launch_consistency(trl_ctx* c, const float* d_emb, const uint8_t* d_valid, int B, const uint8_t* d_clip_start,
                       const float* d_halo, const uint8_t* d_halo_valid, float thr,
                       float* d_sim, uint8_t* d_below, uint8_t* d_has_sim, float* d_last_emb, uint8_t* d_last_valid,
                       cudaStream_t s);
size_t shard_record_size(int n_max);
int launch_shard_pack(trl_ctx* c, const float* d_emb, const uint8_t* d_valid, const uint8_t* d_has_sim, const uint8_t* d_below,
                      const uint8_t* d_clip_start, int n_local, int n_max, unsigned char* d_record, cudaStream_t s);
int launch_shard_resolve(trl_ctx* c, unsigned char* d_all, int world, int my_rank, int n_max, float thr, float* d_sim,
                         uint8_t* d_below, uint8_t* d_has_sim, cudaStream_t s);
