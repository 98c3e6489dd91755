// C ABI of libtruely_b200.so (include/truely_b200.h): context, workspace and the device-side MTCNN cascade driver.
#include <string.h>

#include "common.cuh"

static thread_local std::string g_create_err;

// ---- optional per-stage timing (trl_set_profiling): CUDA events recorded on the launching stream around each stage
static const char* kStageNames[TRL_NUM_STAGES] = {"pyramid", "pnet", "nms_scale", "nms_frame", "crop24", "rnet", "nms_rnet",
                                                  "crop48", "onet", "nms_final", "crop_align", "facenet", "consistency"};
// scoped timer: records an event pair on the stream around one stage when profiling is on
struct StageScope {
  trl_ctx* c;
  cudaStream_t s;
  int stage;
  cudaEvent_t e0 = nullptr;
  StageScope(trl_ctx* c_, cudaStream_t s_, int stage_) : c(c_), s(s_), stage(stage_) {
    if (c->profiling) {
      cudaEventCreate(&e0);
      cudaEventRecord(e0, s);
    }
  }
  ~StageScope() {
    if (e0) {
      cudaEvent_t e1;
      cudaEventCreate(&e1);
      cudaEventRecord(e1, s);
      c->prof_events.push_back({stage, e0, e1});
    }
  }
};
#define TIMED(stage_id, expr)                      \
  do {                                             \
    StageScope _sc(c, s, stage_id);                \
    if ((rc = (expr)) != TRL_OK) return rc;        \
  } while (0)

extern "C" {

void trl_default_config(trl_config_t* cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->min_face_size = 20;
  cfg->thresholds[0] = 0.6f; cfg->thresholds[1] = 0.7f; cfg->thresholds[2] = 0.7f;
  cfg->factor = 0.709;
  cfg->crop_size = 80;
  cfg->cand_cap_scale = 2048;
  cfg->cand_cap_frame = 1024;
  cfg->box_cap_frame = 128;
  cfg->facenet_impl = 0;
  cfg->pnet_precision = 3;      // hybrid P-Net: all-tcgen05 screen (pnet2.cu) + exact fp32 re-evaluation (pnet_refine.cu)
  cfg->mode = 0;
  cfg->margin = 0;
}

const char* trl_last_error(const trl_ctx_t* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

long long trl_launch_count(const trl_ctx_t* ctx) { return ctx ? ctx->launches : 0; }

int trl_host_alloc(size_t bytes, int write_combined, void** out) {
  if (!out || bytes == 0) return TRL_E_INVALID;
  *out = nullptr;
  const cudaError_t e = cudaHostAlloc(out, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
  if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return TRL_E_NOMEM; }
  if (e != cudaSuccess) { cudaGetLastError(); return TRL_E_CUDA; }
  return TRL_OK;
}
int trl_host_free(void* p) {
  if (!p) return TRL_OK;
  return cudaFreeHost(p) == cudaSuccess ? TRL_OK : TRL_E_CUDA;
}

// P-Net / R-Net candidate lists may be raised up to the global-memory NMS limit (16384 per group); the final box list
// stays within the shared-memory limit (its epilogue ranks the boxes by area in O(n^2)).
static bool capacities_ok(int cap_scale, int cap_frame, int box_cap) {
  return cap_scale >= 1 && cap_frame >= 1 && box_cap >= 1 && cap_scale <= nms_big_max_n() && cap_frame <= nms_big_max_n() &&
         box_cap <= nms_max_n();
}

static void free_workspace(trl_ctx* c) {
  void* ptrs[] = {c->d_pyr, c->d_pyr_hi, c->d_pyr_lo, c->d_screen, c->d_screen_cnt, c->d_cand1, c->d_cnt1, c->d_cand2, c->d_cnt2, c->d_cand3, c->d_cnt3, c->d_pad3, c->d_rin,
                  c->d_cand4, c->d_cnt4, c->d_pad4, c->d_oin, c->d_rprob, c->d_rreg, c->d_oprob, c->d_oreg,
                  c->d_boxes, c->d_nfaces, c->d_crops, c->d_nms_big};
  for (void* p : ptrs) if (p) cudaFree(p);
  c->d_pyr_hi = nullptr; c->d_pyr_lo = nullptr; c->d_screen = nullptr; c->d_screen_cnt = nullptr; c->screen_cap = 0;
  c->d_pyr = nullptr; c->d_cand1 = nullptr; c->d_cnt1 = nullptr; c->d_cand2 = nullptr; c->d_cnt2 = nullptr;
  c->d_cand3 = nullptr; c->d_cnt3 = nullptr; c->d_pad3 = nullptr; c->d_rin = nullptr; c->d_cand4 = nullptr;
  c->d_cnt4 = nullptr; c->d_pad4 = nullptr; c->d_oin = nullptr; c->d_rprob = nullptr; c->d_rreg = nullptr;
  c->d_oprob = nullptr; c->d_oreg = nullptr; c->d_boxes = nullptr; c->d_nfaces = nullptr; c->d_crops = nullptr;
  c->d_nms_big = nullptr;
  c->ws_B = c->ws_H = c->ws_W = 0;
}

void trl_destroy(trl_ctx_t* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->tail_stream) { cudaStreamSynchronize(c->tail_stream); cudaStreamDestroy(c->tail_stream); }
  if (c->ev_head) cudaEventDestroy(c->ev_head);
  if (c->ev_tail) cudaEventDestroy(c->ev_tail);
  free_workspace(c);
  facenet_destroy(c);
  if (c->d_pnet_packed) cudaFree(c->d_pnet_packed);
  if (c->d_pnet_refine) cudaFree(c->d_pnet_refine);
  if (c->d_pnet2_packed) cudaFree(c->d_pnet2_packed);
  if (c->d_pnet2_tiles) cudaFree(c->d_pnet2_tiles);
  if (c->d_rnet) cudaFree(c->d_rnet);
  if (c->d_onet) cudaFree(c->d_onet);
  if (c->d_nms_tmp) cudaFree(c->d_nms_tmp);
  if (c->d_pyr_tab) cudaFree(c->d_pyr_tab);
  if (c->d_pyr_blk) cudaFree(c->d_pyr_blk);
  overlay_destroy(c);
  if (c->h_cap) cudaFreeHost(c->h_cap);
  delete c;
}

int trl_create(int device, const trl_weights_t* w, const trl_config_t* cfg, trl_ctx_t** out) {
  if (!out || !w) { g_create_err = "trl_create: null argument"; return TRL_E_INVALID; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("trl_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
    return TRL_E_CUDA;
  }
  if (device < 0 || device >= ndev) { g_create_err = "trl_create: bad device index"; return TRL_E_INVALID; }
  trl_ctx* c = new trl_ctx();
  c->device = device;
  if (cfg) c->cfg = *cfg; else trl_default_config(&c->cfg);
  auto fail = [&](int rc) { g_create_err = c->err; trl_destroy(c); return rc; };
  if (!capacities_ok(c->cfg.cand_cap_scale, c->cfg.cand_cap_frame, c->cfg.box_cap_frame)) {
    c->err = "trl_create: candidate capacities must be in [1, " + std::to_string(nms_big_max_n()) + "], box_cap_frame in [1, " +
             std::to_string(nms_max_n()) + "]";
    return fail(TRL_E_INVALID);
  }
  if (c->cfg.pnet_precision < 0 || c->cfg.pnet_precision > 3) { c->err = "trl_create: pnet_precision must be 0..3"; return fail(TRL_E_INVALID); }
  if (c->cfg.mode != 0 && c->cfg.mode != 1) { c->err = "trl_create: mode must be 0 (reference) or 1 (mode B)"; return fail(TRL_E_INVALID); }
  if (c->cfg.margin < 0 || c->cfg.margin >= c->cfg.crop_size) { c->err = "trl_create: margin must be in [0, crop_size)"; return fail(TRL_E_INVALID); }
  if (cudaSetDevice(device) != cudaSuccess) { c->err = "cudaSetDevice failed"; return fail(TRL_E_CUDA); }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { c->err = "cudaGetDeviceProperties failed"; return fail(TRL_E_CUDA); }
  c->num_sms = prop.multiProcessorCount;
  if (prop.major != 10) {
    c->err = "trl_create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ", this library is built for sm_100a only";
    return fail(TRL_E_CUDA);
  }
  if (cudaHostAlloc(&c->h_cap, sizeof(CapFlag), cudaHostAllocMapped) != cudaSuccess) { c->err = "cudaHostAlloc failed"; return fail(TRL_E_NOMEM); }
  memset(c->h_cap, 0, sizeof(CapFlag));
  if (cudaHostGetDevicePointer(&c->d_cap, c->h_cap, 0) != cudaSuccess) { c->err = "cudaHostGetDevicePointer failed"; return fail(TRL_E_CUDA); }
  int rc;
  if ((rc = nms_init(c)) != TRL_OK) return fail(rc);
  if (w->h_pnet && (rc = pnet_pack_weights(c, w->h_pnet, w->pnet_len)) != TRL_OK) return fail(rc);
  if (w->h_pnet && (rc = pnet_refine_pack_weights(c, w->h_pnet, w->pnet_len)) != TRL_OK) return fail(rc);
  if (w->h_pnet && (rc = pnet2_pack_weights(c, w->h_pnet, w->pnet_len)) != TRL_OK) return fail(rc);
  // the fp16 screen needs activations inside fp16's range (bounded on the host from the weights): otherwise the 3-term kernel
  if (w->h_pnet && c->cfg.pnet_precision >= 2 && !c->pnet2_range_ok) c->cfg.pnet_precision = 0;
  if (w->h_rnet && w->h_onet && (rc = ro_pack_weights(c, w->h_rnet, w->rnet_len, w->h_onet, w->onet_len)) != TRL_OK) return fail(rc);
  if (w->h_facenet && (rc = facenet_create(c, w->h_facenet, w->facenet_len)) != TRL_OK) return fail(rc);
  *out = c;
  return TRL_OK;
}

int trl_overlay_set_stamps(trl_ctx_t* c, const uint8_t* h_lut, int n_lut, const trl_stamp_t* h_stamps, int n_stamps,
                           const uint16_t* h_idx, long long n_idx, int digit_advance) {
  if (!c) return TRL_E_INVALID;
  return overlay_set_stamps(c, h_lut, n_lut, h_stamps, n_stamps, h_idx, n_idx, digit_advance);
}

int trl_overlay(trl_ctx_t* c, uint8_t* d_frames, int B, int H, int W, const int* d_box, const uint8_t* d_state,
                const int* d_frame_index, uint8_t* d_text_pending, void* stream) {
  if (!c) return TRL_E_INVALID;
  if (!d_frames || !d_box || !d_state || !d_frame_index || B < 0 || H <= 0 || W <= 0) TRL_FAIL(c, TRL_E_INVALID, "trl_overlay: bad argument");
  return launch_overlay(c, d_frames, B, H, W, d_box, d_state, d_frame_index, d_text_pending, (cudaStream_t)stream);
}

int trl_pyramid_geometry(const trl_ctx_t* ctx, int H, int W, double* scales, int* hs, int* ws, int* oh, int* ow) {
  if (!ctx) return TRL_E_INVALID;
  PyramidGeom g;
  int rc = compute_geometry(ctx->cfg, H, W, &g);
  if (rc != TRL_OK) return rc;
  for (int k = 0; k < g.n; ++k) {
    if (scales) scales[k] = g.scale[k];
    if (hs) hs[k] = g.hs[k];
    if (ws) ws[k] = g.ws[k];
    if (oh) oh[k] = g.oh[k];
    if (ow) ow[k] = g.ow[k];
  }
  return g.n;
}

int trl_pyramid(trl_ctx_t* c, const uint8_t* d_frames, int B, int H, int W, float* d_out, void* stream) {
  if (!c || !d_frames || !d_out || B < 0) return TRL_E_INVALID;
  PyramidGeom g;
  int rc = compute_geometry(c->cfg, H, W, &g);
  if (rc != TRL_OK) TRL_FAIL(c, rc, "trl_pyramid: bad geometry %dx%d", H, W);
  return launch_pyramid(c, d_frames, B, H, W, g, d_out, false, (cudaStream_t)stream);
}

int trl_pnet_precision(const trl_ctx_t* c) { return c ? c->cfg.pnet_precision : TRL_E_INVALID; }

int trl_pyramid_pairs(trl_ctx_t* c, const uint8_t* d_frames, int B, int H, int W, void* d_hi, void* d_lo, void* stream) {
  if (!c || !d_frames || !d_hi || !d_lo || B < 0) return TRL_E_INVALID;
  PyramidGeom g;
  int rc = compute_geometry(c->cfg, H, W, &g);
  if (rc != TRL_OK) TRL_FAIL(c, rc, "trl_pyramid_pairs: bad geometry %dx%d", H, W);
  return launch_pyramid_pairs(c, d_frames, B, H, W, g, reinterpret_cast<uint4*>(d_hi), reinterpret_cast<uint4*>(d_lo), (cudaStream_t)stream);
}

int trl_pyramid_pairs_size(const trl_ctx_t* c, int H, int W, long long* pairs_per_frame, long long* level_off, int* pitch_pairs) {
  if (!c) return TRL_E_INVALID;
  PyramidGeom g;
  int rc = compute_geometry(c->cfg, H, W, &g);
  if (rc != TRL_OK) return rc;
  if (pairs_per_frame) *pairs_per_frame = g.pairs_total;
  for (int k = 0; k < g.n; ++k) {
    if (level_off) level_off[k] = g.off2[k];
    if (pitch_pairs) pitch_pairs[k] = g.pitch2[k];
  }
  return g.n;
}

int trl_pnet_screen_maps(trl_ctx_t* c, const void* d_hi, int B, int H, int W, float* d_logit, void* stream) {
  if (!c || !d_hi || !d_logit || B < 0) return TRL_E_INVALID;
  PyramidGeom g;
  int rc = compute_geometry(c->cfg, H, W, &g);
  if (rc != TRL_OK) TRL_FAIL(c, rc, "trl_pnet_screen_maps: bad geometry %dx%d", H, W);
  return launch_pnet2(c, reinterpret_cast<const uint4*>(d_hi), B, g, 0.5f, nullptr, nullptr, 0, d_logit, (cudaStream_t)stream);
}

int trl_pnet(trl_ctx_t* c, const float* d_in, int B, int hs, int ws, float* d_prob, float* d_reg, void* stream) {
  if (!c || !d_in || !d_prob || !d_reg) return TRL_E_INVALID;
  if (!c->d_pnet_packed) TRL_FAIL(c, TRL_E_STATE, "P-Net weights not loaded");
  return launch_pnet_maps(c, d_in, B, hs, ws, d_prob, d_reg, (cudaStream_t)stream);
}

int trl_nms(trl_ctx_t* c, const float* d_boxes, const float* d_scores, int n, float thr, int mode, int* d_keep, int* d_nkeep,
            void* stream) {
  if (!c || n < 0 || (mode != 0 && mode != 1)) return TRL_E_INVALID;
  return launch_plain_nms(c, d_boxes, d_scores, n, thr, mode, d_keep, d_nkeep, (cudaStream_t)stream);
}

int trl_crop_resample(trl_ctx_t* c, const uint8_t* d_frames, int B, int H, int W, const int* d_pad, const int* d_img, int n,
                      int size, float* d_out, void* stream) {
  if (!c || !d_frames || !d_pad || !d_img || !d_out || (size != 24 && size != 48)) return TRL_E_INVALID;
  return launch_crop_resample(c, d_frames, B, H, W, d_pad, d_img, nullptr, n, size, d_out, (cudaStream_t)stream);
}

int trl_rnet(trl_ctx_t* c, const float* d_in, int n, float* d_prob, float* d_reg, void* stream) {
  if (!c || !d_in || !d_prob || !d_reg) return TRL_E_INVALID;
  if (!c->d_rnet) TRL_FAIL(c, TRL_E_STATE, "R-Net weights not loaded");
  return launch_rnet(c, d_in, n, nullptr, d_prob, d_reg, (cudaStream_t)stream);
}

int trl_onet(trl_ctx_t* c, const float* d_in, int n, float* d_prob, float* d_reg, void* stream) {
  if (!c || !d_in || !d_prob || !d_reg) return TRL_E_INVALID;
  if (!c->d_onet) TRL_FAIL(c, TRL_E_STATE, "O-Net weights not loaded");
  return launch_onet(c, d_in, n, nullptr, d_prob, d_reg, (cudaStream_t)stream);
}

// ---- workspace

#define WS_ALLOC(ptr, bytes)                                                       \
  do {                                                                             \
    cudaError_t _e = cudaMalloc((void**)&(ptr), (bytes));                          \
    if (_e != cudaSuccess) { free_workspace(c); TRL_FAIL(c, TRL_E_NOMEM, "workspace cudaMalloc(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(_e)); } \
  } while (0)

static int ensure_workspace(trl_ctx* c, int B, int H, int W) {
  if (c->ws_H == H && c->ws_W == W && c->ws_B >= B) return TRL_OK;
  PyramidGeom geom;
  int rc = compute_geometry(c->cfg, H, W, &geom);
  if (rc != TRL_OK) TRL_FAIL(c, rc, "bad frame geometry %dx%d", H, W);
  // upstream detect_face ends in torch.cat([]) -> RuntimeError when the scale list is empty; the reference's run() then
  // raises (HTTP 500 in server.py).  Same error behaviour here, reported before anything is launched or freed.
  if (geom.n == 0)
    TRL_FAIL(c, TRL_E_INVALID, "frame %dx%d is too small for a single pyramid level at min_face_size %d (upstream detect_face raises on it)",
             H, W, c->cfg.min_face_size);
  const int Bc = (c->ws_H == H && c->ws_W == W && c->ws_B > B) ? c->ws_B : B;
  if (c->tail_stream) TRL_CUDA(c, cudaStreamSynchronize(c->tail_stream));     // a pipelined tail may still read the old workspace
  free_workspace(c);
  c->geom = geom;
  const PyramidGeom& g = c->geom;
  const size_t c1 = c->cfg.cand_cap_scale, c2 = c->cfg.cand_cap_frame, c4 = c->cfg.box_cap_frame;
  const int S = c->cfg.crop_size;
  if (c->cfg.pnet_precision == 3) {
    // fp16 hi / lo pair images; the pad pixel of an odd-width level's last pair is never written again: zero it once
    const size_t bytes = (size_t)Bc * g.pairs_total * sizeof(uint4);
    WS_ALLOC(c->d_pyr_hi, bytes);
    WS_ALLOC(c->d_pyr_lo, bytes);
    TRL_CUDA(c, cudaMemset(c->d_pyr_hi, 0, bytes));
    TRL_CUDA(c, cudaMemset(c->d_pyr_lo, 0, bytes));
  } else {
    WS_ALLOC(c->d_pyr, (size_t)Bc * g.floats_total * sizeof(float));
  }
  if (c->cfg.pnet_precision >= 2) {
    long long cells = 0;
    for (int k = 0; k < g.n; ++k) cells += (long long)(g.oh[k] > 0 ? g.oh[k] : 0) * (g.ow[k] > 0 ? g.ow[k] : 0);
    const long long per_frame = std::min<long long>((long long)g.n * (long long)c1, cells);
    c->screen_cap = (int)std::min<long long>((long long)Bc * std::max<long long>(per_frame, 1), 0x7fffffffLL / 16);
    WS_ALLOC(c->d_screen, (size_t)c->screen_cap * sizeof(ScreenEntry));
    WS_ALLOC(c->d_screen_cnt, sizeof(int));
  }
  WS_ALLOC(c->d_cand1, (size_t)Bc * g.n * c1 * sizeof(Cand));
  WS_ALLOC(c->d_cnt1, (size_t)Bc * (g.n + 3) * sizeof(int));     // cnt1 [B][n] then cnt2 [B], cnt3 [B], cnt4 [B]
  c->d_cnt2 = nullptr;
  WS_ALLOC(c->d_cand2, (size_t)Bc * c2 * sizeof(Cand));
  WS_ALLOC(c->d_cand3, (size_t)Bc * c2 * sizeof(Cand));
  WS_ALLOC(c->d_pad3, (size_t)Bc * c2 * 4 * sizeof(int));
  WS_ALLOC(c->d_rin, (size_t)Bc * c2 * 3 * 24 * 24 * sizeof(float));
  WS_ALLOC(c->d_rprob, (size_t)Bc * c2 * sizeof(float));
  WS_ALLOC(c->d_rreg, (size_t)Bc * c2 * 4 * sizeof(float));
  WS_ALLOC(c->d_cand4, (size_t)Bc * c4 * sizeof(Cand));
  WS_ALLOC(c->d_pad4, (size_t)Bc * c4 * 4 * sizeof(int));
  WS_ALLOC(c->d_oin, (size_t)Bc * c4 * 3 * 48 * 48 * sizeof(float));
  WS_ALLOC(c->d_oprob, (size_t)Bc * c4 * sizeof(float));
  WS_ALLOC(c->d_oreg, (size_t)Bc * c4 * 4 * sizeof(float));
  WS_ALLOC(c->d_boxes, (size_t)Bc * c4 * 5 * sizeof(float));
  WS_ALLOC(c->d_nfaces, (size_t)Bc * sizeof(int));
  WS_ALLOC(c->d_crops, (size_t)Bc * S * S * 3 + 256);
  {
    const int cmax = (int)(c1 > c2 ? c1 : c2);
    const size_t big = nms_big_scratch_bytes(Bc * g.n, cmax);
    if (big) WS_ALLOC(c->d_nms_big, big);
  }
  c->ws_B = Bc; c->ws_H = H; c->ws_W = W;
  return TRL_OK;
}

// Makes `s` wait for the tail of the last pipelined cascade call (no-op when none is pending).
static int join_tail(trl_ctx* c, cudaStream_t s) {
  if (!c->tail_pending) return TRL_OK;
  TRL_CUDA(c, cudaStreamWaitEvent(s, c->ev_tail, 0));
  c->tail_pending = false;
  return TRL_OK;
}

// The cascade in two parts.  Head = the two throughput stages (pyramid, P-Net: every SM busy); tail = the latency-bound
// rest (NMS x4, crops, R-Net, O-Net, crop-align: a few hundred candidates).  `ts` is the stream of the tail: the
// caller's stream (serial), or the context's internal stream (pipelined: the tail of chunk k runs under the pyramid of
// chunk k+1; P-Net of chunk k+1 is held back until that tail is done, because it shares the candidate workspace and
// because a persistent P-Net CTA that starts late on an SM still held by a tail kernel would finish late).
static int detect_head(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, cudaStream_t s) {
  const PyramidGeom& g = c->geom;
  int rc;
  const int prec = c->cfg.pnet_precision;
  if (prec == 3) { TIMED(0, launch_pyramid_pairs(c, d_frames, B, H, W, g, c->d_pyr_hi, c->d_pyr_lo, s)); }
  else { TIMED(0, launch_pyramid(c, d_frames, B, H, W, g, c->d_pyr, true, s)); }
  if ((rc = join_tail(c, s)) != TRL_OK) return rc;          // previous tail still reads the candidate workspace
  TRL_CUDA(c, cudaMemsetAsync(c->d_cnt1, 0, (size_t)B * (g.n + 3) * sizeof(int), s));
  if (prec < 2) {
    TIMED(1, launch_pnet_candidates(c, c->d_pyr, B, g, c->cfg.thresholds[0], c->d_cand1, c->d_cnt1, c->cfg.cand_cap_scale, s));
    return TRL_OK;
  }
  // hybrid: single-pass tensor-core screen at thr - margin, then the exact fp32 re-evaluation of the screened cells
  const float thr = c->cfg.thresholds[0], thr_lo = thr - TRL_SCREEN_MARGIN;
  TRL_CUDA(c, cudaMemsetAsync(c->d_screen_cnt, 0, sizeof(int), s));
  {
    StageScope _sc(c, s, 1);      // one stage record ("pnet") for screen + refine
    if (prec == 3) rc = launch_pnet2_screen(c, c->d_pyr_hi, B, g, thr_lo, c->d_screen, c->d_screen_cnt, c->screen_cap, s);
    else rc = launch_pnet_screen_v1(c, c->d_pyr, B, g, thr_lo, c->d_screen, c->d_screen_cnt, c->screen_cap, s);
    if (rc != TRL_OK) return rc;
    rc = launch_pnet_refine(c, prec == 3 ? 1 : 0, prec == 3 ? (const void*)c->d_pyr_hi : (const void*)c->d_pyr, c->d_pyr_lo, B, g, thr,
                            c->d_screen, c->d_screen_cnt, c->screen_cap, c->d_cand1, c->d_cnt1, c->cfg.cand_cap_scale, s);
    if (rc != TRL_OK) return rc;
  }
  return TRL_OK;
}

static int detect_tail(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, int* d_nfaces, float* d_boxes, int* d_counts,
                       cudaStream_t s) {
  const PyramidGeom& g = c->geom;
  const int c1 = c->cfg.cand_cap_scale, c2 = c->cfg.cand_cap_frame, c4 = c->cfg.box_cap_frame;
  int rc;
  // counters are laid out for the *current* B: cnt1 [B][n], cnt2 [B], cnt3 [B], cnt4 [B]
  int* cnt1 = c->d_cnt1;
  int* cnt2 = cnt1 + (size_t)B * g.n;
  int* cnt3 = cnt2 + B;
  int* cnt4 = cnt3 + B;
  nms::StageParams p{};
  p.W = W; p.H = H; p.capflag = c->d_cap;
  // stage 1: per (frame, level) NMS 0.5
  p.n_levels = g.n; p.cap_in = c1; p.cap_out = c2; p.thr_nms = 0.5f; p.thr_score = 0.f;
  p.in = c->d_cand1; p.cnt_in = cnt1; p.out = c->d_cand2; p.cnt_out = cnt2;
  TIMED(2, launch_cascade_stage(c, 1, p, B, s));
  // stage 2: per frame NMS 0.7 + regression + rerec + pad -> R-Net inputs
  p.cap_in = c2; p.cap_out = c2; p.thr_nms = 0.7f;
  p.in = c->d_cand2; p.cnt_in = cnt2; p.out = c->d_cand3; p.cnt_out = cnt3; p.pad_out = c->d_pad3;
  TIMED(3, launch_cascade_stage(c, 2, p, B, s));
  TIMED(4, launch_crop_resample_ex(c, d_frames, B, H, W, c->d_pad3, nullptr, cnt3, c2, B * c2, 24, c->d_rin, s));
  TIMED(5, launch_rnet_ex(c, c->d_rin, B * c2, cnt3, c2, c->d_rprob, c->d_rreg, s));
  // stage 3: R-Net score > thr, NMS 0.7, bbreg, rerec, pad -> O-Net inputs
  p.cap_in = c2; p.cap_out = c4; p.thr_nms = 0.7f; p.thr_score = c->cfg.thresholds[1];
  p.in = c->d_cand3; p.cnt_in = cnt3; p.prob = c->d_rprob; p.reg = c->d_rreg; p.pad_in = c->d_pad3;
  p.out = c->d_cand4; p.cnt_out = cnt4; p.pad_out = c->d_pad4;
  TIMED(6, launch_cascade_stage(c, 3, p, B, s));
  TIMED(7, launch_crop_resample_ex(c, d_frames, B, H, W, c->d_pad4, nullptr, cnt4, c4, B * c4, 48, c->d_oin, s));
  TIMED(8, launch_onet_ex(c, c->d_oin, B * c4, cnt4, c4, c->d_oprob, c->d_oreg, s));
  // stage 4: O-Net score > thr, bbreg, 'Min' NMS 0.7, largest-first
  p.cap_in = c4; p.cap_out = c4; p.thr_nms = 0.7f; p.thr_score = c->cfg.thresholds[2];
  p.in = c->d_cand4; p.cnt_in = cnt4; p.prob = c->d_oprob; p.reg = c->d_oreg; p.pad_in = c->d_pad4;
  p.out = nullptr; p.cnt_out = d_nfaces; p.pad_out = nullptr; p.boxes_out = d_boxes;
  TIMED(9, launch_cascade_stage(c, 4, p, B, s));
  if (d_counts) {
    // (#P-Net candidates summed over levels is not needed on the hot path; report per-stage list sizes)
    TRL_CUDA(c, cudaMemcpy2DAsync(d_counts + 1, 4 * sizeof(int), cnt3, sizeof(int), sizeof(int), B, cudaMemcpyDeviceToDevice, s));
    TRL_CUDA(c, cudaMemcpy2DAsync(d_counts + 2, 4 * sizeof(int), cnt4, sizeof(int), sizeof(int), B, cudaMemcpyDeviceToDevice, s));
    TRL_CUDA(c, cudaMemcpy2DAsync(d_counts + 3, 4 * sizeof(int), d_nfaces, sizeof(int), sizeof(int), B, cudaMemcpyDeviceToDevice, s));
    TRL_CUDA(c, cudaMemcpy2DAsync(d_counts + 0, 4 * sizeof(int), cnt2, sizeof(int), sizeof(int), B, cudaMemcpyDeviceToDevice, s));
  }
  return TRL_OK;
}

static int detect_impl(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, int* d_nfaces, float* d_boxes, int* d_counts,
                       cudaStream_t s) {
  if (!c->d_pnet_packed || !c->d_rnet || !c->d_onet) TRL_FAIL(c, TRL_E_STATE, "MTCNN weights not loaded");
  int rc = ensure_workspace(c, B, H, W);
  if (rc != TRL_OK) return rc;
  if ((rc = detect_head(c, d_frames, B, H, W, s)) != TRL_OK) return rc;
  return detect_tail(c, d_frames, B, H, W, d_nfaces, d_boxes, d_counts, s);
}

int trl_detect(trl_ctx_t* c, const uint8_t* d_frames, int B, int H, int W, int* d_nfaces, float* d_boxes, int* d_counts,
               void* stream) {
  if (!c || !d_frames || !d_nfaces || !d_boxes || B <= 0) return TRL_E_INVALID;
  int rc = join_tail(c, (cudaStream_t)stream);
  if (rc != TRL_OK) return rc;
  return detect_impl(c, d_frames, B, H, W, d_nfaces, d_boxes, d_counts, (cudaStream_t)stream);
}

int trl_crop_align(trl_ctx_t* c, const uint8_t* d_frames, int B, int H, int W, const float* d_boxes, int box_stride,
                   const int* d_nfaces, int* d_box_int, uint8_t* d_valid, uint8_t* d_crops, void* stream) {
  if (!c || !d_frames || !d_boxes || !d_nfaces || !d_box_int || !d_valid || !d_crops) return TRL_E_INVALID;
  int rc = join_tail(c, (cudaStream_t)stream);
  if (rc != TRL_OK) return rc;
  return launch_crop_align(c, d_frames, B, H, W, d_boxes, box_stride, d_nfaces, c->cfg.crop_size, d_box_int, d_valid, d_crops,
                           (cudaStream_t)stream);
}

int trl_facenet(trl_ctx_t* c, const uint8_t* d_crops, int n, int S, float* d_emb, void* stream) {
  if (!c || !d_crops || !d_emb || n < 0) return TRL_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if ((rc = join_tail(c, s)) != TRL_OK) return rc;
  TIMED(11, facenet_forward(c, d_crops, n, S, c->cfg.mode == 1 ? 1 : 0, d_emb, s));
  return TRL_OK;
}

int trl_facenet_valid(trl_ctx_t* c, const uint8_t* d_crops, const uint8_t* d_valid, int n, int S, float* d_emb, void* stream) {
  if (!c || !d_crops || !d_valid || !d_emb || n < 0) return TRL_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if ((rc = join_tail(c, s)) != TRL_OK) return rc;
  TIMED(11, facenet_forward_valid(c, d_crops, d_valid, n, S, c->cfg.mode == 1 ? 1 : 0, d_emb, s));
  return TRL_OK;
}

int trl_facenet_norm(trl_ctx_t* c, const uint8_t* d_crops, int n, int S, int norm, float* d_emb, void* stream) {
  if (!c || !d_crops || !d_emb || n < 0 || (norm != 0 && norm != 1)) return TRL_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if ((rc = join_tail(c, s)) != TRL_OK) return rc;
  TIMED(11, facenet_forward(c, d_crops, n, S, norm, d_emb, s));
  return TRL_OK;
}

int trl_extract_face(trl_ctx_t* c, const uint8_t* d_frames, int B, int H, int W, const float* d_boxes, int box_stride,
                     const int* d_nfaces, int image_size, int margin, int* d_box_int, uint8_t* d_valid, uint8_t* d_crops,
                     void* stream) {
  if (!c || !d_frames || !d_boxes || !d_nfaces || !d_box_int || !d_valid || !d_crops) return TRL_E_INVALID;
  int rc = join_tail(c, (cudaStream_t)stream);
  if (rc != TRL_OK) return rc;
  return launch_extract_face(c, d_frames, B, H, W, d_boxes, box_stride, d_nfaces, image_size, margin, d_box_int, d_valid, d_crops,
                             (cudaStream_t)stream);
}

int trl_extract_faces_all(trl_ctx_t* c, const uint8_t* d_frames, int B, int H, int W, const float* d_boxes, int box_stride,
                          int box_cap, const int* d_nfaces, int image_size, int margin, int max_faces, int* d_face_off,
                          int* d_face_frame, int* d_box_int, uint8_t* d_valid, uint8_t* d_crops, void* stream) {
  if (!c || !d_frames || !d_boxes || !d_nfaces || !d_face_off || !d_box_int || !d_valid || !d_crops || box_stride < 5 * box_cap)
    return TRL_E_INVALID;
  int rc = join_tail(c, (cudaStream_t)stream);
  if (rc != TRL_OK) return rc;
  return launch_extract_faces_all(c, d_frames, B, H, W, d_boxes, box_stride, box_cap, d_nfaces, image_size, margin, max_faces,
                                  d_face_off, d_face_frame, d_box_int, d_valid, d_crops, (cudaStream_t)stream);
}

int trl_consistency(trl_ctx_t* c, const float* d_emb, const uint8_t* d_valid, int B, const float* d_halo_emb,
                    const uint8_t* d_halo_valid, float thr,
                    float* d_sim, uint8_t* d_below, uint8_t* d_has_sim, float* d_last_emb, uint8_t* d_last_valid, void* stream) {
  if (!c || !d_emb || !d_valid || !d_sim || !d_below || !d_has_sim) return TRL_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if ((rc = join_tail(c, s)) != TRL_OK) return rc;
  TIMED(12, launch_consistency(c, d_emb, d_valid, B, nullptr, d_halo_emb, d_halo_valid, thr, d_sim, d_below, d_has_sim, d_last_emb,
                               d_last_valid, s));
  return TRL_OK;
}

int trl_consistency_clips(trl_ctx_t* c, const float* d_emb, const uint8_t* d_valid, int B, const uint8_t* d_clip_start,
                          const float* d_halo_emb, const uint8_t* d_halo_valid, float thr, float* d_sim, uint8_t* d_below,
                          uint8_t* d_has_sim, float* d_last_emb, uint8_t* d_last_valid, void* stream) {
  if (!c || !d_emb || !d_valid || !d_sim || !d_below || !d_has_sim) return TRL_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if ((rc = join_tail(c, s)) != TRL_OK) return rc;
  TIMED(12, launch_consistency(c, d_emb, d_valid, B, d_clip_start, d_halo_emb, d_halo_valid, thr, d_sim, d_below, d_has_sim,
                               d_last_emb, d_last_valid, s));
  return TRL_OK;
}

size_t trl_shard_record_bytes(int n_max) { return n_max < 0 ? 0 : shard_record_size(n_max); }

int trl_shard_pack(trl_ctx_t* c, const float* d_emb, const uint8_t* d_valid, const uint8_t* d_has_sim, const uint8_t* d_below,
                   const uint8_t* d_clip_start, int n_local, int n_max, void* d_record, void* stream) {
  if (!c || !d_record || n_local < 0 || n_local > n_max) return TRL_E_INVALID;
  if (n_local > 0 && (!d_emb || !d_valid || !d_has_sim || !d_below)) return TRL_E_INVALID;
  if ((reinterpret_cast<uintptr_t>(d_record) & 15) != 0) TRL_FAIL(c, TRL_E_INVALID, "trl_shard_pack: record must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = join_tail(c, s);
  if (rc != TRL_OK) return rc;
  return launch_shard_pack(c, d_emb, d_valid, d_has_sim, d_below, d_clip_start, n_local, n_max, (unsigned char*)d_record, s);
}

int trl_shard_resolve(trl_ctx_t* c, void* d_all_records, int world, int rank, int n_max, float thr, float* d_sim,
                      uint8_t* d_below, uint8_t* d_has_sim, void* stream) {
  if (!c || !d_all_records || world < 1 || rank < 0 || rank >= world || n_max < 0) return TRL_E_INVALID;
  return launch_shard_resolve(c, (unsigned char*)d_all_records, world, rank, n_max, thr, d_sim, d_below, d_has_sim,
                              (cudaStream_t)stream);
}

// crop of the largest face of every frame: the reference's (server/model.py:49-57) or, in mode B, upstream extract_face
static int crop_stage(trl_ctx* c, const uint8_t* d_frames, int B, int H, int W, const int* nf, int* d_box_int, uint8_t* d_valid,
                      uint8_t* d_crops, cudaStream_t s) {
  if (c->cfg.mode == 1)
    return launch_extract_face(c, d_frames, B, H, W, c->d_boxes, c->cfg.box_cap_frame * 5, nf, c->cfg.crop_size, c->cfg.margin,
                               d_box_int, d_valid, d_crops, s);
  return launch_crop_align(c, d_frames, B, H, W, c->d_boxes, c->cfg.box_cap_frame * 5, nf, c->cfg.crop_size, d_box_int, d_valid,
                           d_crops, s);
}

int trl_detect_align(trl_ctx_t* c, const uint8_t* d_frames, int B, int H, int W, int* d_box_int, uint8_t* d_valid,
                     int* d_nfaces, uint8_t* d_crops, void* stream) {
  if (!c || !d_frames || !d_box_int || !d_valid || !d_crops || B <= 0) return TRL_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = join_tail(c, s);
  if (rc != TRL_OK) return rc;
  if ((rc = ensure_workspace(c, B, H, W)) != TRL_OK) return rc;
  int* nf = d_nfaces ? d_nfaces : c->d_nfaces;
  if ((rc = detect_impl(c, d_frames, B, H, W, nf, c->d_boxes, nullptr, s)) != TRL_OK) return rc;
  TIMED(10, crop_stage(c, d_frames, B, H, W, nf, d_box_int, d_valid, d_crops, s));
  return TRL_OK;
}

int trl_detect_align_async(trl_ctx_t* c, const uint8_t* d_frames, int B, int H, int W, int* d_box_int, uint8_t* d_valid,
                           int* d_nfaces, uint8_t* d_crops, void* stream) {
  if (!c || !d_frames || !d_box_int || !d_valid || !d_crops || B <= 0) return TRL_E_INVALID;
  if (c->profiling) return trl_detect_align(c, d_frames, B, H, W, d_box_int, d_valid, d_nfaces, d_crops, stream);   // stage times need a serial schedule
  if (!c->d_pnet_packed || !c->d_rnet || !c->d_onet) TRL_FAIL(c, TRL_E_STATE, "MTCNN weights not loaded");
  cudaStream_t s = (cudaStream_t)stream;
  if (!c->tail_stream) {
    int lo = 0, hi = 0;
    TRL_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    TRL_CUDA(c, cudaStreamCreateWithPriority(&c->tail_stream, cudaStreamNonBlocking, hi));     // tail CTAs go first when an SM frees up
    TRL_CUDA(c, cudaEventCreateWithFlags(&c->ev_head, cudaEventDisableTiming));
    TRL_CUDA(c, cudaEventCreateWithFlags(&c->ev_tail, cudaEventDisableTiming));
  }
  int rc = ensure_workspace(c, B, H, W);
  if (rc != TRL_OK) return rc;
  int* nf = d_nfaces ? d_nfaces : c->d_nfaces;
  if ((rc = detect_head(c, d_frames, B, H, W, s)) != TRL_OK) return rc;       // joins the previous tail before P-Net
  TRL_CUDA(c, cudaEventRecord(c->ev_head, s));
  cudaStream_t ts = c->tail_stream;
  TRL_CUDA(c, cudaStreamWaitEvent(ts, c->ev_head, 0));
  if ((rc = detect_tail(c, d_frames, B, H, W, nf, c->d_boxes, nullptr, ts)) != TRL_OK) return rc;
  if ((rc = crop_stage(c, d_frames, B, H, W, nf, d_box_int, d_valid, d_crops, ts)) != TRL_OK) return rc;
  TRL_CUDA(c, cudaEventRecord(c->ev_tail, ts));
  c->tail_pending = true;
  return TRL_OK;
}

int trl_set_capacity(trl_ctx_t* c, int cand_cap_scale, int cand_cap_frame, int box_cap_frame) {
  if (!c) return TRL_E_INVALID;
  if (!capacities_ok(cand_cap_scale, cand_cap_frame, box_cap_frame))
    TRL_FAIL(c, TRL_E_INVALID, "trl_set_capacity: capacities must be in [1, %d] (box_cap_frame: [1, %d])", nms_big_max_n(), nms_max_n());
  if (c->tail_stream) TRL_CUDA(c, cudaStreamSynchronize(c->tail_stream));
  free_workspace(c);                        // cudaFree waits for outstanding work that may still use the old workspace (slow path)
  c->tail_pending = false;
  c->cfg.cand_cap_scale = cand_cap_scale; c->cfg.cand_cap_frame = cand_cap_frame; c->cfg.box_cap_frame = box_cap_frame;
  return TRL_OK;
}

int trl_pipeline_join(trl_ctx_t* c, void* stream) {
  if (!c) return TRL_E_INVALID;
  return join_tail(c, (cudaStream_t)stream);
}

int trl_process(trl_ctx_t* c, const uint8_t* d_frames, int B, int H, int W, const float* d_halo_emb,
                const uint8_t* d_halo_valid, float thr, int* d_box_int,
                uint8_t* d_valid, float* d_emb, float* d_sim, uint8_t* d_below, uint8_t* d_has_sim, int* d_nfaces,
                float* d_last_emb, uint8_t* d_last_valid, void* stream) {
  if (!c || !d_frames || !d_box_int || !d_valid || !d_emb || !d_sim || !d_below || !d_has_sim || B <= 0) return TRL_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = ensure_workspace(c, B, H, W);       // before c->d_crops is read: the workspace may be (re)allocated here
  if (rc != TRL_OK) return rc;
  if ((rc = trl_detect_align(c, d_frames, B, H, W, d_box_int, d_valid, d_nfaces, c->d_crops, stream)) != TRL_OK) return rc;
  // only face-bearing frames are embedded (server/model.py:48 skips the others): the crops are packed on the device and
  // the live batch size stays there, so the batch remains free of host synchronisation
  TIMED(11, facenet_forward_valid(c, c->d_crops, d_valid, B, c->cfg.crop_size, c->cfg.mode == 1 ? 1 : 0, d_emb, s));
  TIMED(12, launch_consistency(c, d_emb, d_valid, B, nullptr, d_halo_emb, d_halo_valid, thr, d_sim, d_below, d_has_sim, d_last_emb,
                               d_last_valid, s));
  return TRL_OK;
}

int trl_set_profiling(trl_ctx_t* c, int on) {
  if (!c) return TRL_E_INVALID;
  c->profiling = on != 0;
  return TRL_OK;
}

int trl_stage_name(int stage, char* buf, int len) {
  if (stage < 0 || stage >= TRL_NUM_STAGES || !buf) return TRL_E_INVALID;
  snprintf(buf, len, "%s", kStageNames[stage]);
  return TRL_OK;
}

/* Sum of the per-stage device times (ms) since the last read; the caller must have synchronised the stream.
 * h_ms: float[TRL_NUM_STAGES].  Returns the number of timed stage launches accumulated. */
int trl_read_stage_times(trl_ctx_t* c, float* h_ms) {
  if (!c || !h_ms) return TRL_E_INVALID;
  for (int i = 0; i < TRL_NUM_STAGES; ++i) h_ms[i] = 0.f;
  int n = 0;
  for (auto& r : c->prof_events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess && r.stage >= 0 && r.stage < TRL_NUM_STAGES) h_ms[r.stage] += ms;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
    ++n;
  }
  c->prof_events.clear();
  return n;
}

int trl_check_capacity(trl_ctx_t* c, int* h_detail) {
  if (!c) return TRL_E_INVALID;
  if (c->h_cap->overflow) {
    if (h_detail) { h_detail[0] = c->h_cap->stage; h_detail[1] = c->h_cap->frame; h_detail[2] = c->h_cap->count; h_detail[3] = c->h_cap->capacity; }
    if (c->h_cap->stage == 5)
      c->err = "P-Net activation beyond +-" + std::to_string(c->h_cap->capacity) + " (frame " + std::to_string(c->h_cap->frame) +
               "): outside the range of the scaled fp16 operand split (pnet.cu)";
    else
      c->err = "candidate buffer overflow at stage " + std::to_string(c->h_cap->stage) + " (frame " + std::to_string(c->h_cap->frame) +
               ": " + std::to_string(c->h_cap->count) + " > capacity " + std::to_string(c->h_cap->capacity) + ")";
    memset(c->h_cap, 0, sizeof(CapFlag));
    return TRL_E_CAPACITY;
  }
  return TRL_OK;
}

}  // extern "C"
