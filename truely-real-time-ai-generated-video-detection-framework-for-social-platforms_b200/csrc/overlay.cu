// Device-side annotation of the processed frames (SURVEY.md 8f "device overlay"; reference server/model.py:66-74):
// cv2.rectangle(frame, (x1, y1), (x2, y2), colour, 2) followed by cv2.putText(..., LINE_AA) -- bit exact with OpenCV.
//
//  * rectangle, thickness 2, LINE_8: OpenCV paints the pixels within one pixel of the outline, minus the four outer corner
//    pixels, clipped to the image (the rule is pinned against cv2.rectangle on thousands of random boxes, clipped ones
//    included, in tests/test_host.py).
//  * anti-aliased text: every stroke blends  p += ((c - p) * a + 127) >> 8  into the frame, strokes overlap, so the result per
//    pixel and channel is a function f(p) of the background value p alone.  The host builds these functions once with
//    OpenCV's own rasteriser (overlay.py: the string on the 256 constant backgrounds) as 256-entry look-up tables per touched
//    pixel, de-duplicated; a "stamp" is a bounding box of indices into that table.  OpenCV's fixed-point text geometry is
//    translation invariant for integer origins while the text stays inside the image, so a stamp applies at any origin; text
//    that would be clipped is left to the host (d_text_pending[b] = 1; the rectangle is still drawn here).  The frame number
//    of "AI Detected - Frame n" is composed from ten digit stamps at the font's fixed digit advance, applied in drawing order
//    (sequential blending is exactly what applying the tables one after the other computes).
// One CTA per frame: the rectangle, a barrier, then the stamps in order (the text may cross the rectangle).
#include "common.cuh"

namespace overlay {

struct Stamps {
  uint8_t* d_lut = nullptr;        // [n_lut][2][256]: towards 0 / towards 255
  trl_stamp_t* d_stamps = nullptr; // 0: "Real Frame", 1: "AI Detected - Frame ", 2 + d: digit d in slot 0
  uint16_t* d_idx = nullptr;       // index maps, 0 = untouched, k = table k - 1
  int n_stamps = 0, digit_advance = 0;
  trl_stamp_t h_stamps[12];
};

__device__ __forceinline__ bool stamp_inside(const trl_stamp_t& s, int ox, int oy, int W, int H) {
  // two pixels of slack: OpenCV's clipping changes the rasterisation only when a stroke really leaves the image
  return ox + s.ox - 2 >= 0 && oy + s.oy - 2 >= 0 && ox + s.ox + s.w + 2 <= W && oy + s.oy + s.h + 2 <= H;
}

__device__ __forceinline__ void apply_stamp(uint8_t* __restrict__ frame, int W, const trl_stamp_t& s, int ox, int oy,
                                            const uint16_t* __restrict__ idx, const uint8_t* __restrict__ lut, int t0, int t1, int t2) {
  const uint16_t* m = idx + s.idx_off;
  const int n = s.w * s.h;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int k = m[i];
    if (k == 0) continue;
    const int y = i / s.w, x = i - y * s.w;
    uint8_t* px = frame + ((size_t)(oy + s.oy + y) * W + (ox + s.ox + x)) * 3;
    const uint8_t* l = lut + (size_t)(k - 1) * 512;
    px[0] = l[t0 * 256 + px[0]];
    px[1] = l[t1 * 256 + px[1]];
    px[2] = l[t2 * 256 + px[2]];
  }
}

// state: 0 = untouched, 1 = "Real Frame" (green), 2 = "AI Detected - Frame n" (red)
__global__ void __launch_bounds__(256) overlay_kernel(uint8_t* __restrict__ frames, int H, int W, const int* __restrict__ box,
                                                     const uint8_t* __restrict__ state, const int* __restrict__ frame_index,
                                                     const trl_stamp_t* __restrict__ stamps, const uint16_t* __restrict__ idx,
                                                     const uint8_t* __restrict__ lut, int digit_advance,
                                                     uint8_t* __restrict__ text_pending) {
  const int b = blockIdx.x;
  const int st = state[b];
  if (st == 0) { if (threadIdx.x == 0 && text_pending) text_pending[b] = 0; return; }
  uint8_t* frame = frames + (size_t)b * H * W * 3;
  const int x1 = box[b * 4 + 0], y1 = box[b * 4 + 1], x2 = box[b * 4 + 2], y2 = box[b * 4 + 3];
  const uint8_t c0 = 0, c1 = st == 1 ? 255 : 0, c2 = st == 1 ? 0 : 255;          // BGR

  const int bx0 = max(x1 - 1, 0), bx1 = min(x2 + 1, W - 1), by0 = max(y1 - 1, 0), by1 = min(y2 + 1, H - 1);
  const int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
  if (bw > 0 && bh > 0) {
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) {
      const int y = by0 + i / bw, x = bx0 + i % bw;
      const bool inner = x >= x1 + 2 && x <= x2 - 2 && y >= y1 + 2 && y <= y2 - 2;
      const bool corner = (x == x1 - 1 || x == x2 + 1) && (y == y1 - 1 || y == y2 + 1);
      if (!inner && !corner) {
        uint8_t* px = frame + ((size_t)y * W + x) * 3;
        px[0] = c0; px[1] = c1; px[2] = c2;
      }
    }
  }
  __syncthreads();

  const int t0 = 0, t1 = c1 ? 1 : 0, t2 = c2 ? 1 : 0;
  if (st == 1) {
    const trl_stamp_t s = stamps[0];
    const int ox = x1, oy = y1 - 10;
    const bool ok = stamp_inside(s, ox, oy, W, H);
    if (threadIdx.x == 0 && text_pending) text_pending[b] = ok ? 0 : 1;
    if (ok) apply_stamp(frame, W, s, ox, oy, idx, lut, t0, t1, t2);
    return;
  }
  // "AI Detected - Frame n" at (10, 30)
  int n = frame_index[b];
  if (n < 0) n = 0;
  int digits[12], nd = 0;
  do { digits[nd++] = n % 10; n /= 10; } while (n > 0 && nd < 12);
  const int ox = 10, oy = 30;
  bool ok = stamp_inside(stamps[1], ox, oy, W, H);
  for (int k = 0; k < nd; ++k) ok = ok && stamp_inside(stamps[2 + digits[nd - 1 - k]], ox + k * digit_advance, oy, W, H);
  if (threadIdx.x == 0 && text_pending) text_pending[b] = ok ? 0 : 1;
  if (!ok) return;
  apply_stamp(frame, W, stamps[1], ox, oy, idx, lut, t0, t1, t2);
  for (int k = 0; k < nd; ++k) {
    __syncthreads();                       // neighbouring glyphs may share anti-aliased fringe pixels: drawing order
    apply_stamp(frame, W, stamps[2 + digits[nd - 1 - k]], ox + k * digit_advance, oy, idx, lut, t0, t1, t2);
  }
}

}  // namespace overlay

void overlay_destroy(trl_ctx* c) {
  overlay::Stamps* s = reinterpret_cast<overlay::Stamps*>(c->overlay);
  if (!s) return;
  if (s->d_lut) cudaFree(s->d_lut);
  if (s->d_stamps) cudaFree(s->d_stamps);
  if (s->d_idx) cudaFree(s->d_idx);
  delete s;
  c->overlay = nullptr;
}

int overlay_set_stamps(trl_ctx* c, const uint8_t* h_lut, int n_lut, const trl_stamp_t* h_stamps, int n_stamps,
                       const uint16_t* h_idx, long long n_idx, int digit_advance) {
  if (!h_lut || !h_stamps || !h_idx || n_lut <= 0 || n_lut > 65534 || n_stamps != 12 || n_idx <= 0 || digit_advance <= 0)
    TRL_FAIL(c, TRL_E_INVALID, "trl_overlay_set_stamps: expected 12 stamps (\"Real Frame\", the \"AI Detected\" prefix, ten digits)");
  for (int i = 0; i < n_stamps; ++i) {
    const trl_stamp_t& s = h_stamps[i];
    if (s.w <= 0 || s.h <= 0 || s.idx_off < 0 || s.idx_off + (long long)s.w * s.h > n_idx)
      TRL_FAIL(c, TRL_E_INVALID, "trl_overlay_set_stamps: stamp %d out of range", i);
  }
  for (long long i = 0; i < n_idx; ++i)
    if (h_idx[i] > n_lut) TRL_FAIL(c, TRL_E_INVALID, "trl_overlay_set_stamps: table index out of range");
  overlay_destroy(c);
  overlay::Stamps* s = new overlay::Stamps();
  c->overlay = s;
  TRL_CUDA(c, cudaMalloc(&s->d_lut, (size_t)n_lut * 512));
  TRL_CUDA(c, cudaMalloc(&s->d_stamps, n_stamps * sizeof(trl_stamp_t)));
  TRL_CUDA(c, cudaMalloc(&s->d_idx, (size_t)n_idx * sizeof(uint16_t)));
  TRL_CUDA(c, cudaMemcpy(s->d_lut, h_lut, (size_t)n_lut * 512, cudaMemcpyHostToDevice));
  TRL_CUDA(c, cudaMemcpy(s->d_stamps, h_stamps, n_stamps * sizeof(trl_stamp_t), cudaMemcpyHostToDevice));
  TRL_CUDA(c, cudaMemcpy(s->d_idx, h_idx, (size_t)n_idx * sizeof(uint16_t), cudaMemcpyHostToDevice));
  s->n_stamps = n_stamps;
  s->digit_advance = digit_advance;
  for (int i = 0; i < n_stamps; ++i) s->h_stamps[i] = h_stamps[i];
  return TRL_OK;
}

int launch_overlay(trl_ctx* c, uint8_t* d_frames, int B, int H, int W, const int* d_box, const uint8_t* d_state,
                   const int* d_frame_index, uint8_t* d_text_pending, cudaStream_t s) {
  overlay::Stamps* st = reinterpret_cast<overlay::Stamps*>(c->overlay);
  if (!st) TRL_FAIL(c, TRL_E_INVALID, "trl_overlay: no stamps (trl_overlay_set_stamps)");
  if (B <= 0) return TRL_OK;
  overlay::overlay_kernel<<<B, 256, 0, s>>>(d_frames, H, W, d_box, d_state, d_frame_index, st->d_stamps, st->d_idx, st->d_lut,
                                            st->digit_advance, d_text_pending);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}
