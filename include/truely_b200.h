/*
 * truely_b200.h — C ABI of libtruely_b200.so: the B200 (sm_100a) implementation of the
 * Truely visual-analysis hot path, i.e. everything reference server/model.py::run does per
 * processed frame between cv2.VideoCapture.read (server/model.py:43) and the run-length
 * counter (server/model.py:62).
 *
 * The reference has no FFI of its own (it is pure Python over facenet_pytorch / torchvision /
 * OpenCV, all on CPU); each entry point below therefore cites the Python call it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - return 0 (TRL_OK) on success, a negative TRL_E_* code otherwise; no exceptions cross the ABI;
 *     trl_last_error() gives the message of the last failure on that context.
 *   - every pointer named d_* is DEVICE memory owned by the caller; h_* is host memory.
 *   - every call is asynchronous on the CUDA stream passed as `stream` (a cudaStream_t cast to
 *     void*; NULL = legacy default stream).  The library never calls cudaDeviceSynchronize.
 *     Workspace is (re)allocated with cudaMalloc only when a call needs more than the context holds.
 *   - one context per (process, device); a context is not thread safe.
 *   - frames are BGR uint8, layout [B, H, W, 3] (what cv2.VideoCapture.read returns).
 *   - there is no CPU fallback: without a CUDA device trl_create fails with TRL_E_CUDA.
 */
#ifndef TRUELY_B200_H
#define TRUELY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRL_OK 0
#define TRL_E_INVALID (-1)   /* bad argument */
#define TRL_E_CUDA (-2)      /* CUDA runtime / driver error */
#define TRL_E_CAPACITY (-3)  /* a candidate buffer overflowed (reported, never silently truncated) */
#define TRL_E_NOMEM (-4)
#define TRL_E_STATE (-5)     /* call order / missing weights */

#define TRL_EMB_DIM 512
#define TRL_NUM_STAGES 13
#define TRL_MAX_SCALES 24

typedef struct trl_ctx trl_ctx_t;

/* Flat float32 weight blobs (host memory, copied during trl_create).
 *   pnet / rnet / onet: tensors of the upstream state dict concatenated in the order of
 *     SURVEY.md Appendix C (conv1.weight, conv1.bias, prelu1.weight, conv2.weight, ...), PyTorch layouts.
 *   facenet: InceptionResnetV1 with BatchNorm folded (fp32): for every BasicConv2d in execution order
 *     W[cout][kh][kw][cin] then bias[cout]; then for every residual up-projection W[cout][cin], bias[cout];
 *     then last_linear folded with last_bn: W[512][1792], bias[512].  (weights.py::pack_facenet) */
typedef struct {
  const float* h_pnet;    size_t pnet_len;     /* 6,632 floats   */
  const float* h_rnet;    size_t rnet_len;     /* 100,178 floats */
  const float* h_onet;    size_t onet_len;     /* 389,040 floats (dense6_3 landmarks head is accepted and unused) */
  const float* h_facenet; size_t facenet_len;  /* trl_facenet_blob_len() floats */
} trl_weights_t;

/* Replaces the constructor arguments of MTCNN() (server/model.py:18, upstream defaults) and the
 * literals of server/model.py:16,41. */
typedef struct {
  int min_face_size;      /* 20 */
  float thresholds[3];    /* 0.6, 0.7, 0.7 */
  double factor;          /* 0.709 */
  int crop_size;          /* 80: side of the FaceNet input (server/model.py:41) */
  int cand_cap_scale;     /* capacity: P-Net candidates per (frame, scale); default 2048, at most 16384 */
  int cand_cap_frame;     /* capacity: R-Net inputs per frame (candidates surviving stage-1 NMS); default 1024, at most 16384 */
  int box_cap_frame;      /* capacity: O-Net inputs / final boxes per frame; default 128, at most 2048 */
  int facenet_impl;       /* 0 = tcgen05 implicit-GEMM path (product); 1 = SIMT direct-conv kernels (validation) */
  int pnet_precision;     /* how the cascade evaluates P-Net (the trl_pnet stage entry point always uses 0):
                             0 = one kernel, 3-term fp16 operand split on the tensor pipe (maps within 2e-5 of the fp32 reference);
                             1 = one kernel, single-pass fp16 (experiment: maps within ~2e-3, candidates may differ);
                             2 = hybrid: the single-pass kernel of (1) only screens -- cells with prob >= thresholds[0] - 0.05
                                 are re-evaluated exactly in fp32 (12x12 receptive field each) and thresholded there, so the
                                 candidate set and its scores / regressions are those of an fp32 P-Net;
                             3 = hybrid with the all-tcgen05 screening kernel on fp16 pixel-pair pyramid images (the default
                                 of the Python host; fastest) */
  int mode;               /* 0 = the reference's crop path (server/model.py:49-58): truncated + clamped box, cv2.resize
                             INTER_LINEAR to crop_size, F.to_tensor (/255).  1 = "mode B", upstream facenet_pytorch's own
                             face crop as the north star words it: extract_face (margin-adjusted box, cv2.resize INTER_AREA
                             to crop_size, normally 160) + fixed_image_standardization ((x - 127.5) / 128) */
  int margin;             /* mode B: extract_face margin in pixels of the crop_size image (upstream default 0) */
} trl_config_t;

void trl_default_config(trl_config_t* cfg);
/* The P-Net mode the context actually runs (trl_create falls back from 2 / 3 to 0 when a host-side bound of the P-Net
 * activations computed from the weights does not fit the fp16 operands of the screen). */
int trl_pnet_precision(const trl_ctx_t* ctx);
size_t trl_facenet_blob_len(void);

/* Replaces MTCNN() + InceptionResnetV1(pretrained="vggface2").eval()   (server/model.py:18-19). */
int trl_create(int device, const trl_weights_t* w, const trl_config_t* cfg, trl_ctx_t** out);
void trl_destroy(trl_ctx_t* ctx);
const char* trl_last_error(const trl_ctx_t* ctx);

/* Pyramid geometry of detect_face (upstream detect_face.py: scale loop).
 * Fills up to TRL_MAX_SCALES entries; returns the number of scales (>= 0) or a TRL_E_* code. */
int trl_pyramid_geometry(const trl_ctx_t* ctx, int H, int W, double* scales, int* hs, int* ws, int* oh, int* ow);

/* ---- stage entry points (each is one stage of detect_face / run, exposed for stage-isolated parity) ---- */

/* K1: imresample(mode="area") + (x-127.5)*0.0078125 for every scale.
 * d_out: concatenation over scales of float32 [B,3,hs,ws]. */
int trl_pyramid(trl_ctx_t* ctx, const uint8_t* d_frames, int B, int H, int W, float* d_out, void* stream);

/* K1 in the layout of the all-tensor-pipe P-Net (pnet_precision 3): every level as two images of pixel PAIRS, 16 bytes per pair
 * = 2 x (B, G, R, 0) fp16; d_hi holds fp16(v), d_lo holds fp16(v - hi) (hi + lo restores v to 2^-23 relative).  Level k of a
 * batch of B frames starts at pair offset level_off[k] * B and is [B, hs, pitch_pairs[k]] pairs (trl_pyramid_pairs_size, which
 * returns the number of levels); the pad pixel of an odd-width level is not written. */
int trl_pyramid_pairs_size(const trl_ctx_t* ctx, int H, int W, long long* pairs_per_frame, long long* level_off, int* pitch_pairs);
int trl_pyramid_pairs(trl_ctx_t* ctx, const uint8_t* d_frames, int B, int H, int W, void* d_hi, void* d_lo, void* stream);

/* K2 (screen): the single-pass tcgen05 P-Net on the hi image of trl_pyramid_pairs; d_logit receives, levels concatenated,
 * float32 [B, oh, ow] = conv4_1 logit of class 1 minus class 0 (sigmoid of it approximates PNet's probability map to ~1e-3). */
int trl_pnet_screen_maps(trl_ctx_t* ctx, const void* d_hi, int B, int H, int W, float* d_logit, void* stream);

/* K2: PNet.forward on one pyramid level.  d_in float32 [B,3,hs,ws] -> d_prob [B,oh,ow] (softmax class 1),
 * d_reg [B,4,oh,ow]. */
int trl_pnet(trl_ctx_t* ctx, const float* d_in, int B, int hs, int ws, float* d_prob, float* d_reg, void* stream);

/* K4/K5: greedy NMS of one group.  mode 0 = torchvision.ops.nms (area (x2-x1)(y2-y1), suppress IoU > thr,
 * ties by index ascending); mode 1 = upstream nms_numpy(..., 'Min') (+1 areas, keep o <= thr, ties by index
 * descending).  d_keep receives indices in pick order, *d_nkeep their count. */
int trl_nms(trl_ctx_t* ctx, const float* d_boxes, const float* d_scores, int n, float thr, int mode,
            int* d_keep, int* d_nkeep, void* stream);

/* K7: crop (1-based inclusive y..ey, x..ex as returned by upstream pad()) -> area resample to size x size
 * -> normalise.  d_pad int32 [N,4] = (y, ey, x, ex); d_img int32 [N]; d_out float32 [N,3,size,size]. */
int trl_crop_resample(trl_ctx_t* ctx, const uint8_t* d_frames, int B, int H, int W, const int* d_pad,
                      const int* d_img, int n, int size, float* d_out, void* stream);

/* K8 / K9: RNet.forward / ONet.forward.  d_in float32 [N,3,24,24] / [N,3,48,48] -> d_prob [N], d_reg [N,4]. */
int trl_rnet(trl_ctx_t* ctx, const float* d_in, int n, float* d_prob, float* d_reg, void* stream);
int trl_onet(trl_ctx_t* ctx, const float* d_in, int n, float* d_prob, float* d_reg, void* stream);

/* MTCNN.detect for a batch of frames (server/model.py:47): the whole cascade K1..K9 on device.
 * d_nfaces int32 [B]; d_boxes float32 [B, box_cap_frame, 5] (x1,y1,x2,y2,score), per frame sorted
 * largest-area first (select_largest=True); d_counts (optional, may be NULL) int32 [B,4] =
 * (#P-Net candidates, #R-Net inputs, #O-Net inputs, #final boxes).
 * Capacity overflow is recorded on device; trl_check_capacity() reports it after a stream sync. */
int trl_detect(trl_ctx_t* ctx, const uint8_t* d_frames, int B, int H, int W, int* d_nfaces, float* d_boxes,
               int* d_counts, void* stream);

/* K10: box truncation + clamp (server/model.py:49-53), crop, cv2.resize(.., (S,S)) INTER_LINEAR uint8
 * (server/model.py:55-57), bit exact.  d_boxes float32 [B, box_stride] (first 4 floats of each row used),
 * d_nfaces int32 [B].  d_box_int int32 [B,4]; d_valid uint8 [B] (1 = a face was cropped);
 * d_crops uint8 [B,S,S,3]. */
int trl_crop_align(trl_ctx_t* ctx, const uint8_t* d_frames, int B, int H, int W, const float* d_boxes,
                   int box_stride, const int* d_nfaces, int* d_box_int, uint8_t* d_valid, uint8_t* d_crops,
                   void* stream);

/* K11: F.to_tensor (/255) + InceptionResnetV1.forward (server/model.py:58-59).
 * d_crops uint8 [N,S,S,3] BGR -> d_emb float32 [N,512], unit norm. */
int trl_facenet(trl_ctx_t* ctx, const uint8_t* d_crops, int n, int S, float* d_emb, void* stream);

/* Mode B building blocks (upstream facenet_pytorch models/utils/detect_face.py, SURVEY.md Appendix A "Mode-B extras").
 * trl_extract_face: extract_face(img, boxes[0], image_size, margin) per frame -- box arithmetic in fp32, int() truncation,
 *   crop, cv2.resize(.., (image_size, image_size), interpolation=cv2.INTER_AREA) on uint8, bit exact with OpenCV
 *   (integer-ratio, general-area and enlarging code paths).  Same argument meaning as trl_crop_align.
 * trl_extract_faces_all: keep_all -- every box of every frame (d_boxes float32 [B, box_stride], rows of 5 floats per box,
 *   at most box_cap boxes per frame).  Faces are numbered frame by frame: d_face_off int32 [B+1] receives the prefix
 *   (face i of frame b is face d_face_off[b] + i of the batch, d_face_off[B] = total), d_face_frame int32 [max_faces]
 *   (optional) the frame of each face; d_box_int [max_faces,4], d_valid [max_faces], d_crops [max_faces,S,S,3].  More than
 *   max_faces faces is reported through trl_check_capacity (stage 6), never silently dropped.
 * trl_facenet_norm: trl_facenet with the input normalisation chosen per call: 0 = F.to_tensor (x / 255, the reference),
 *   1 = fixed_image_standardization ((x - 127.5) / 128, what MTCNN.forward applies with post_process=True). */
int trl_extract_face(trl_ctx_t* ctx, const uint8_t* d_frames, int B, int H, int W, const float* d_boxes, int box_stride,
                     const int* d_nfaces, int image_size, int margin, int* d_box_int, uint8_t* d_valid, uint8_t* d_crops,
                     void* stream);
int trl_extract_faces_all(trl_ctx_t* ctx, const uint8_t* d_frames, int B, int H, int W, const float* d_boxes, int box_stride,
                          int box_cap, const int* d_nfaces, int image_size, int margin, int max_faces, int* d_face_off,
                          int* d_face_frame, int* d_box_int, uint8_t* d_valid, uint8_t* d_crops, void* stream);
int trl_facenet_norm(trl_ctx_t* ctx, const uint8_t* d_crops, int n, int S, int norm, float* d_emb, void* stream);

/* trl_facenet restricted to the face-bearing crops: the reference never embeds a frame without a face
 * (server/model.py:48 `continue`s first).  Crops with d_valid[i] != 0 are packed on the device, embedded as one batch
 * whose size never travels to the host (every kernel of the pass reads it from device memory), and scattered back;
 * rows of d_emb that belong to frames without a face are set to zero.  Normalisation follows trl_config_t.mode. */
int trl_facenet_valid(trl_ctx_t* ctx, const uint8_t* d_crops, const uint8_t* d_valid, int n, int S, float* d_emb, void* stream);

/* K12: cosine similarity against the previous face-bearing frame and the 0.99 test (server/model.py:60-62).
 * d_emb [B,512], d_valid [B]; d_halo_emb [512] (or NULL) is the last face-bearing embedding before this
 * range (previous batch or previous rank), d_halo_valid uint8[1] (or NULL = valid) says on the device whether
 * it holds one, so consecutive batches chain without a host sync.  Outputs: d_sim float32 [B] (NaN where no comparison),
 * d_below uint8 [B] (1 = sim < thr), d_has_sim uint8 [B];  d_last_emb [512] + d_last_valid uint8[1]
 * (either may be NULL): the halo to hand to the next range. */
int trl_consistency(trl_ctx_t* ctx, const float* d_emb, const uint8_t* d_valid, int B, const float* d_halo_emb,
                    const uint8_t* d_halo_valid, float thr, float* d_sim, uint8_t* d_below, uint8_t* d_has_sim, float* d_last_emb,
                    uint8_t* d_last_valid, void* stream);

/* K12 for a batch that holds several clips back to back (BASELINE.json configs[4]): d_clip_start uint8 [B] (or NULL =
 * one clip) marks the first processed frame of every clip.  `previous_face_encoding` is a local of one run() call
 * (server/model.py:37), so a frame is only compared with a face-bearing frame of its own clip, the incoming halo is used
 * only for frames of the clip that reaches in from the previous range, and d_last_emb / d_last_valid describe the last
 * clip of the range.  With d_clip_start == NULL identical to trl_consistency. */
int trl_consistency_clips(trl_ctx_t* ctx, const float* d_emb, const uint8_t* d_valid, int B, const uint8_t* d_clip_start,
                          const float* d_halo_emb, const uint8_t* d_halo_valid, float thr, float* d_sim, uint8_t* d_below,
                          uint8_t* d_has_sim, float* d_last_emb, uint8_t* d_last_valid, void* stream);

/* Frame-range sharding over the GPUs of a box (no counterpart in the single-process reference).
 * Rank r runs trl_consistency[_clips] on its range WITHOUT a halo, condenses the result into a fixed-size record
 * (trl_shard_pack: per-frame flags valid / has_sim / below, the embedding of the first frame still waiting for a
 * predecessor and the outgoing halo), the host all-gathers the records of all ranks with ONE collective
 * (world * trl_shard_record_bytes(n_max) bytes, n_max = the largest local range), and trl_shard_resolve finishes the
 * comparisons that cross shard boundaries on the gathered buffer -- exact across shards without a face and across clip
 * boundaries -- patching the per-frame flags of every rank's section and this rank's local d_sim / d_below / d_has_sim
 * (any of which may be NULL).  Layout of a record: int32 hdr[8] = {n_local, first_idx, last_has, blocked, 0...},
 * float first_emb[512], float last_emb[512], uint8 flags[3][n_pad] (valid, has_sim, below; n_pad = n_max rounded up to
 * 16).  d_record / d_all_records must be 16-byte aligned device memory. */
size_t trl_shard_record_bytes(int n_max);
int trl_shard_pack(trl_ctx_t* ctx, const float* d_emb, const uint8_t* d_valid, const uint8_t* d_has_sim, const uint8_t* d_below,
                   const uint8_t* d_clip_start, int n_local, int n_max, void* d_record, void* stream);
int trl_shard_resolve(trl_ctx_t* ctx, void* d_all_records, int world, int rank, int n_max, float thr, float* d_sim,
                      uint8_t* d_below, uint8_t* d_has_sim, void* stream);

/* The fused per-batch hot path: detect -> crop-align -> facenet (valid frames only) -> consistency.
 * Equivalent to the body of the reference loop (server/model.py:47-65) for B processed frames. */
int trl_process(trl_ctx_t* ctx, const uint8_t* d_frames, int B, int H, int W, const float* d_halo_emb,
                const uint8_t* d_halo_valid, float thr,
                int* d_box_int, uint8_t* d_valid, float* d_emb, float* d_sim, uint8_t* d_below, uint8_t* d_has_sim,
                int* d_nfaces, float* d_last_emb, uint8_t* d_last_valid, void* stream);

/* Device-side annotation of processed frames, in place (replaces cv2.rectangle + cv2.putText, server/model.py:67-74;
 * SURVEY.md 8f).  d_state uint8 [B]: 0 = leave the frame alone, 1 = green box + "Real Frame" at (x1, y1 - 10),
 * 2 = red box + "AI Detected - Frame n" at (10, 30) with n = d_frame_index[b]; d_box int32 [B,4] = (x1, y1, x2, y2) as
 * written by trl_process.  Bit exact with OpenCV: the thickness-2 rectangle by rule, the anti-aliased text through stamps
 * of per-pixel look-up tables that the host builds once with OpenCV's own rasteriser (overlay.py::build_stamps) and
 * registers with trl_overlay_set_stamps -- h_lut uint8 [n_lut][2][256] (blend towards 0 / towards 255 as a function of the
 * background value), twelve stamps ("Real Frame", the "AI Detected - Frame " prefix, the digits 0-9 in the first digit
 * slot; box relative to the text origin, index map uint16 [h][w] at h_idx + idx_off, 0 = untouched, k = table k - 1) and
 * the font's digit advance in pixels.  Text that would leave the frame is not drawn (OpenCV's clipped rasterisation is not
 * translation invariant): d_text_pending uint8 [B] (may be NULL) is set to 1 for such frames and the caller draws their
 * text on the host; the rectangle is always drawn here. */
typedef struct { int ox, oy, w, h; long long idx_off; } trl_stamp_t;
int trl_overlay_set_stamps(trl_ctx_t* ctx, const uint8_t* h_lut, int n_lut, const trl_stamp_t* h_stamps, int n_stamps,
                           const uint16_t* h_idx, long long n_idx, int digit_advance);
int trl_overlay(trl_ctx_t* ctx, uint8_t* d_frames, int B, int H, int W, const int* d_box, const uint8_t* d_state,
                const int* d_frame_index, uint8_t* d_text_pending, void* stream);

/* First half of trl_process: detect + crop-align only (server/model.py:47-57), writing the crops of this batch into a
 * caller-owned buffer d_crops uint8 [B,S,S,3].  Lets a host that holds a whole clip run the cascade chunk by chunk
 * and then embed every crop with ONE trl_facenet call (large-M GEMMs) followed by one trl_consistency call. */
int trl_detect_align(trl_ctx_t* ctx, const uint8_t* d_frames, int B, int H, int W, int* d_box_int, uint8_t* d_valid,
                     int* d_nfaces, uint8_t* d_crops, void* stream);

/* Pipelined form of trl_detect_align for a host that feeds a clip chunk by chunk: the two throughput stages (pyramid,
 * P-Net) are enqueued on `stream`, the latency-bound rest of the cascade (NMS, crops, R-Net, O-Net, crop-align) on an
 * internal high-priority stream, so that it runs under the pyramid of the NEXT chunk.  The outputs of a call
 * (d_box_int, d_valid, d_nfaces, d_crops) and its reads of d_frames are complete, in `stream` order, only after
 * the next trl_detect_align_async call has returned (it orders them before its own P-Net), after trl_pipeline_join,
 * or after any other stream-taking call on this context (each of them joins first).  Same results as
 * trl_detect_align, bit for bit. */
int trl_detect_align_async(trl_ctx_t* ctx, const uint8_t* d_frames, int B, int H, int W, int* d_box_int, uint8_t* d_valid,
                           int* d_nfaces, uint8_t* d_crops, void* stream);
int trl_pipeline_join(trl_ctx_t* ctx, void* stream);

/* After the stream has been synchronised by the caller: TRL_E_CAPACITY if any candidate buffer overflowed
 * since the last check (h_detail, optional int32[4]: which stage, frame, count, capacity). */
int trl_check_capacity(trl_ctx_t* ctx, int* h_detail);

/* Changes the candidate capacities of a live context (the workspace is released and re-allocated by the next call;
 * this synchronises the device: slow path).  cand_cap_scale / cand_cap_frame up to 16384 -- groups above 2048 candidates
 * then run through a global-memory NMS instead of the shared-memory one -- box_cap_frame up to 2048.  Upstream
 * detect_face has no cap at all: a host that sees TRL_E_CAPACITY raises the capacities, re-runs the affected frames one
 * at a time and restores the fast-path capacities (model.analyze_stream does this). */
int trl_set_capacity(trl_ctx_t* ctx, int cand_cap_scale, int cand_cap_frame, int box_cap_frame);

/* Per-stage device timing of trl_process (measurement only; off by default).  With profiling on, CUDA events are
 * recorded on the launching stream around each stage; trl_read_stage_times (after the caller synchronised the
 * stream) returns in h_ms[TRL_NUM_STAGES] the summed milliseconds per stage since the last read, and as return
 * value the number of timed stage launches accumulated.  trl_stage_name gives the stage labels. */
int trl_set_profiling(trl_ctx_t* ctx, int on);
int trl_read_stage_times(trl_ctx_t* ctx, float* h_ms);
int trl_stage_name(int stage, char* buf, int len);

/* Host staging buffers for the frames on their way to the device (what replaces the pageable numpy frame that
 * cv2.VideoCapture.read returns, server/model.py:43, as the source of the host->device copy).  Page-locked;
 * with write_combined != 0 the pages are also write-combined (cudaHostAllocWriteCombined): the host only ever
 * WRITES decoded frames into a staging buffer and the copy engine reads it without snooping the CPU caches, which
 * matters when eight GPUs of one box pull frames at the same time.  Reading such a buffer from the CPU is slow.
 * No context needed; returns TRL_E_NOMEM / TRL_E_CUDA on failure. */
int trl_host_alloc(size_t bytes, int write_combined, void** out);
int trl_host_free(void* p);

/* Number of kernels launched by this context so far (bench.py reports it as gpu_launches). */
long long trl_launch_count(const trl_ctx_t* ctx);

#ifdef __cplusplus
}
#endif
#endif /* TRUELY_B200_H */
