// InceptionResnetV1 layer program (upstream models/inception_resnet_v1.py), built once from the
// folded-BN weight blob, instantiated for a crop size S and a batch capacity.
//
// Design: activations are NHWC bf16; the 1x1 convs that open the branches of a block are fused
// into one GEMM along N (Block35 3x32, Block17 2x128, Mixed_7a 3x256, Block8 2x192); every branch writes straight
// into its channel slice of the block's concat buffer; the block's 1x1 up-projection applies bias, `x + scale*y`
// and ReLU in its epilogue, in place on the trunk.  103 GEMM launches + 3 max-pools + stem + head per batch.
#include <string.h>

#include <map>

#include "facenet.cuh"

namespace {

struct HostT {
  std::vector<float> w;   // [cout][kh][kw][cin]
  std::vector<float> b;
  int cin, cout, kh, kw;
};

enum Lvl { L1 = 0, L2, L4, L6, L7, L8, L9, NLVL };   // spatial sizes s1 (39|79), s2 (37|77), s4 (18|38), s6 (16|36), s7 (7|17), s8 (3|8), s9 (1|3)

struct BufSpec { int lvl; int C; };

struct SegSpec { int n_begin, n_end, buf, coff; };

struct LayerSpec {
  std::string name;
  int kind;            // 0 conv, 1 maxpool
  int src, src_coff, Cin;
  int Cout, kh, kw, stride, ph, pw;
  int epi; float scale; int resid;
  int nseg; SegSpec seg[4];
  size_t w_off, b_off;   // element offsets into the bf16 weight arena / fp32 bias arena
};

}  // namespace

struct FaceNetEngine {
  std::vector<BufSpec> bufs;
  std::vector<LayerSpec> layers;
  bf16* d_w = nullptr;
  float* d_bias = nullptr;
  float* d_stem_w = nullptr;   // [27][32]  (already divided by 255: F.to_tensor folded in)
  float* d_stem_b = nullptr;
  // compaction scratch of facenet_forward_valid
  uint8_t* d_cmp_crops = nullptr; float* d_cmp_emb = nullptr; int* d_cmp_rank = nullptr; int cmp_cap = 0, cmp_S = 0;
  float* d_stem_w_std = nullptr;   // mode B: fixed_image_standardization folded in: w / 128, bias - (127.5 / 128) sum(w)
  float* d_stem_b_std = nullptr;
  float* d_head_w = nullptr;   // [1792][512]
  float* d_head_b = nullptr;
  int buf_stem = -1, buf_final = -1;
  // instantiated
  int S = 0, cap = 0;
  int sz[NLVL] = {0};
  std::vector<bf16*> d_act;
  std::vector<size_t> act_elems;
  struct Step { int kind; ConvOp conv; PoolOp pool; };
  std::vector<Step> steps;
};

size_t trl_facenet_blob_len(void) {
  // must equal weights.py::facenet_blob_size()
  size_t n = 0;
  auto C = [&](int cin, int cout, int kh, int kw) { n += (size_t)cout * cin * kh * kw + cout; };
  C(3, 32, 3, 3); C(32, 32, 3, 3); C(32, 64, 3, 3); C(64, 80, 1, 1); C(80, 192, 3, 3); C(192, 256, 3, 3);
  for (int i = 0; i < 5; ++i) { C(256, 32, 1, 1); C(256, 32, 1, 1); C(32, 32, 3, 3); C(256, 32, 1, 1); C(32, 32, 3, 3); C(32, 32, 3, 3); }
  C(256, 384, 3, 3); C(256, 192, 1, 1); C(192, 192, 3, 3); C(192, 256, 3, 3);
  for (int i = 0; i < 10; ++i) { C(896, 128, 1, 1); C(896, 128, 1, 1); C(128, 128, 1, 7); C(128, 128, 7, 1); }
  C(896, 256, 1, 1); C(256, 384, 3, 3); C(896, 256, 1, 1); C(256, 256, 3, 3); C(896, 256, 1, 1); C(256, 256, 3, 3); C(256, 256, 3, 3);
  for (int i = 0; i < 6; ++i) { C(1792, 192, 1, 1); C(1792, 192, 1, 1); C(192, 192, 1, 3); C(192, 192, 3, 1); }
  for (int i = 0; i < 5; ++i) C(96, 256, 1, 1);
  for (int i = 0; i < 10; ++i) C(256, 896, 1, 1);
  for (int i = 0; i < 6; ++i) C(384, 1792, 1, 1);
  n += (size_t)512 * 1792 + 512;
  return n;
}

namespace {

struct Builder {
  FaceNetEngine* e;
  std::vector<HostT> convs;    // BasicConv2d, table order
  std::vector<HostT> resids;   // up-projections, table order
  std::vector<bf16> w_arena;
  std::vector<float> b_arena;
  size_t ci = 0, ri = 0;       // cursors

  int add_buf(int lvl, int C) { e->bufs.push_back({lvl, C}); return (int)e->bufs.size() - 1; }

  // append (possibly N-fused, possibly channel-padded) weights to the arenas
  void push_weights(LayerSpec& L, const std::vector<const HostT*>& parts, int cin_pad, int cout_pad) {
    while (w_arena.size() % 128) w_arena.push_back(__float2bfloat16_rn(0.f));   // 256-byte alignment for TMA
    while (b_arena.size() % 4) b_arena.push_back(0.f);
    L.w_off = w_arena.size();
    L.b_off = b_arena.size();
    int cout_total = 0;
    for (const HostT* t : parts) {
      for (int co = 0; co < t->cout; ++co) {
        for (int tap = 0; tap < t->kh * t->kw; ++tap)
          for (int c = 0; c < cin_pad; ++c)
            w_arena.push_back(__float2bfloat16_rn(c < t->cin ? t->w[((size_t)co * t->kh * t->kw + tap) * t->cin + c] : 0.f));
        b_arena.push_back(t->b[co]);
      }
      cout_total += t->cout;
    }
    for (int co = cout_total; co < cout_pad; ++co) {
      for (int q = 0; q < parts[0]->kh * parts[0]->kw * cin_pad; ++q) w_arena.push_back(__float2bfloat16_rn(0.f));
      b_arena.push_back(0.f);
    }
  }

  LayerSpec base(const std::string& name, int src, int src_coff, int Cin, int Cout, int kh, int kw, int stride, int ph, int pw) {
    LayerSpec L{};
    L.name = name; L.kind = 0; L.src = src; L.src_coff = src_coff; L.Cin = Cin; L.Cout = Cout;
    L.kh = kh; L.kw = kw; L.stride = stride; L.ph = ph; L.pw = pw; L.epi = EPI_RELU; L.scale = 1.f; L.resid = -1; L.nseg = 0;
    return L;
  }
  void seg(LayerSpec& L, int nb, int ne, int buf, int coff) { L.seg[L.nseg++] = {nb, ne, buf, coff}; }

  // single BasicConv2d from the table
  void conv(const std::string& name, int src, int src_coff, int dst, int dst_coff, int stride, int ph, int pw,
            int cin_pad = 0, int cout_pad = 0) {
    const HostT& t = convs[ci++];
    const int cin = cin_pad ? cin_pad : t.cin, cout = cout_pad ? cout_pad : t.cout;
    LayerSpec L = base(name, src, src_coff, cin, cout, t.kh, t.kw, stride, ph, pw);
    seg(L, 0, cout, dst, dst_coff);
    push_weights(L, {&t}, cin, cout);
    e->layers.push_back(L);
  }
  void pool(const std::string& name, int src, int C, int dst, int dst_coff) {
    LayerSpec L{};
    L.name = name; L.kind = 1; L.src = src; L.src_coff = 0; L.Cin = C; L.Cout = C; L.nseg = 1;
    L.seg[0] = {0, C, dst, dst_coff};
    e->layers.push_back(L);
  }
  void resid(const std::string& name, int src, int trunk, int epi) {
    const HostT& t = resids[ri++];
    LayerSpec L = base(name, src, 0, t.cin, t.cout, 1, 1, 1, 0, 0);
    L.epi = epi; L.resid = trunk;
    seg(L, 0, t.cout, trunk, 0);
    push_weights(L, {&t}, t.cin, t.cout);
    e->layers.push_back(L);
  }
};

}  // namespace

static const float RESID_SCALE_35 = 0.17f, RESID_SCALE_17 = 0.10f, RESID_SCALE_8 = 0.20f;

int facenet_create(trl_ctx* c, const float* blob, size_t len) {
  if (len != trl_facenet_blob_len())
    TRL_FAIL(c, TRL_E_INVALID, "facenet blob has %zu floats, expected %zu", len, trl_facenet_blob_len());
  FaceNetEngine* e = new FaceNetEngine();
  c->facenet = e;
  Builder B;
  B.e = e;
  const float* p = blob;
  auto take = [&](int cin, int cout, int kh, int kw) {
    HostT t;
    t.cin = cin; t.cout = cout; t.kh = kh; t.kw = kw;
    t.w.assign(p, p + (size_t)cout * cin * kh * kw); p += (size_t)cout * cin * kh * kw;
    t.b.assign(p, p + cout); p += cout;
    return t;
  };
  auto T = [&](int cin, int cout, int kh, int kw) { B.convs.push_back(take(cin, cout, kh, kw)); };
  T(3, 32, 3, 3); T(32, 32, 3, 3); T(32, 64, 3, 3); T(64, 80, 1, 1); T(80, 192, 3, 3); T(192, 256, 3, 3);
  for (int i = 0; i < 5; ++i) { T(256, 32, 1, 1); T(256, 32, 1, 1); T(32, 32, 3, 3); T(256, 32, 1, 1); T(32, 32, 3, 3); T(32, 32, 3, 3); }
  T(256, 384, 3, 3); T(256, 192, 1, 1); T(192, 192, 3, 3); T(192, 256, 3, 3);
  for (int i = 0; i < 10; ++i) { T(896, 128, 1, 1); T(896, 128, 1, 1); T(128, 128, 1, 7); T(128, 128, 7, 1); }
  T(896, 256, 1, 1); T(256, 384, 3, 3); T(896, 256, 1, 1); T(256, 256, 3, 3); T(896, 256, 1, 1); T(256, 256, 3, 3); T(256, 256, 3, 3);
  for (int i = 0; i < 6; ++i) { T(1792, 192, 1, 1); T(1792, 192, 1, 1); T(192, 192, 1, 3); T(192, 192, 3, 1); }
  for (int i = 0; i < 5; ++i) B.resids.push_back(take(96, 256, 1, 1));
  for (int i = 0; i < 10; ++i) B.resids.push_back(take(256, 896, 1, 1));
  for (int i = 0; i < 6; ++i) B.resids.push_back(take(384, 1792, 1, 1));
  const float* head_w = p; p += (size_t)512 * 1792;
  const float* head_b = p; p += 512;

  // ---- stem conv2d_1a: fp32 SIMT kernel on the uint8 crop; fold F.to_tensor's 1/255 into the weights
  {
    const HostT& t = B.convs[B.ci++];
    std::vector<float> w(27 * 32);
    for (int co = 0; co < 32; ++co)
      for (int k = 0; k < 27; ++k) w[k * 32 + co] = (float)((double)t.w[co * 27 + k] / 255.0);
    TRL_CUDA(c, cudaMalloc(&e->d_stem_w, w.size() * 4));
    TRL_CUDA(c, cudaMemcpy(e->d_stem_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    TRL_CUDA(c, cudaMalloc(&e->d_stem_b, 32 * 4));
    TRL_CUDA(c, cudaMemcpy(e->d_stem_b, t.b.data(), 32 * 4, cudaMemcpyHostToDevice));
    // mode B input (x - 127.5) / 128: conv(w, (x - 127.5) / 128) = conv(w / 128, x) - (127.5 / 128) * sum(w); the conv has no
    // padding, so the shift is the same for every output pixel.  w / 128 is exact.
    std::vector<float> ws(27 * 32), bs(32);
    for (int co = 0; co < 32; ++co) {
      double sum = 0.0;
      for (int k = 0; k < 27; ++k) { ws[k * 32 + co] = t.w[co * 27 + k] * 0.0078125f; sum += (double)t.w[co * 27 + k]; }
      bs[co] = (float)((double)t.b[co] - 127.5 / 128.0 * sum);
    }
    TRL_CUDA(c, cudaMalloc(&e->d_stem_w_std, ws.size() * 4));
    TRL_CUDA(c, cudaMemcpy(e->d_stem_w_std, ws.data(), ws.size() * 4, cudaMemcpyHostToDevice));
    TRL_CUDA(c, cudaMalloc(&e->d_stem_b_std, 32 * 4));
    TRL_CUDA(c, cudaMemcpy(e->d_stem_b_std, bs.data(), 32 * 4, cudaMemcpyHostToDevice));
  }
  // ---- buffers
  const int a1 = B.add_buf(L1, 32), a2 = B.add_buf(L2, 32), a3 = B.add_buf(L2, 64), a4 = B.add_buf(L4, 64);
  const int a5 = B.add_buf(L4, 96), a6 = B.add_buf(L6, 192), t35 = B.add_buf(L7, 256);
  const int e35 = B.add_buf(L7, 64), cat35 = B.add_buf(L7, 96), tmp35 = B.add_buf(L7, 32);
  const int m6a = B.add_buf(L7, 192), m6b = B.add_buf(L7, 192), t17 = B.add_buf(L8, 896);
  const int e17 = B.add_buf(L8, 128), e17b = B.add_buf(L8, 128), cat17 = B.add_buf(L8, 256);
  const int m7 = B.add_buf(L8, 768), m7b = B.add_buf(L8, 256), t8 = B.add_buf(L9, 1792);
  const int e8 = B.add_buf(L9, 192), e8b = B.add_buf(L9, 192), cat8 = B.add_buf(L9, 384);
  e->buf_stem = a1;
  e->buf_final = t8;

  // ---- stem
  B.conv("conv2d_2a", a1, 0, a2, 0, 1, 0, 0);
  B.conv("conv2d_2b", a2, 0, a3, 0, 1, 1, 1);
  B.pool("maxpool_3a", a3, 64, a4, 0);
  B.conv("conv2d_3b", a4, 0, a5, 0, 1, 0, 0, 0, 96);        // 80 -> 96 output channels (zero weights) so K blocks by 32
  B.conv("conv2d_4a", a5, 0, a6, 0, 1, 0, 0, 96, 0);
  B.conv("conv2d_4b", a6, 0, t35, 0, 2, 0, 0);
  // ---- 5 x Block35
  for (int i = 0; i < 5; ++i) {
    const std::string n = "repeat_1." + std::to_string(i);
    const HostT &b0 = B.convs[B.ci], &b10 = B.convs[B.ci + 1], &b20 = B.convs[B.ci + 3];
    LayerSpec L = B.base(n + ".entry", t35, 0, 256, 96, 1, 1, 1, 0, 0);
    B.seg(L, 0, 32, cat35, 0);
    B.seg(L, 32, 96, e35, 0);
    B.push_weights(L, {&b0, &b10, &b20}, 256, 96);
    e->layers.push_back(L);
    B.ci += 2;                                                 // -> branch1.1
    B.conv(n + ".branch1.1", e35, 0, cat35, 32, 1, 1, 1);
    B.ci += 1;                                                 // skip branch2.0 (fused) -> branch2.1
    B.conv(n + ".branch2.1", e35, 32, tmp35, 0, 1, 1, 1);
    B.conv(n + ".branch2.2", tmp35, 0, cat35, 64, 1, 1, 1);
    B.resid(n + ".conv2d", cat35, t35, EPI_RESID_RELU);
    e->layers.back().scale = RESID_SCALE_35;
  }
  // ---- Mixed_6a
  B.conv("mixed_6a.branch0", t35, 0, t17, 0, 2, 0, 0);
  B.conv("mixed_6a.branch1.0", t35, 0, m6a, 0, 1, 0, 0);
  B.conv("mixed_6a.branch1.1", m6a, 0, m6b, 0, 1, 1, 1);
  B.conv("mixed_6a.branch1.2", m6b, 0, t17, 384, 2, 0, 0);
  B.pool("mixed_6a.branch2", t35, 256, t17, 640);
  // ---- 10 x Block17
  for (int i = 0; i < 10; ++i) {
    const std::string n = "repeat_2." + std::to_string(i);
    const HostT &b0 = B.convs[B.ci], &b10 = B.convs[B.ci + 1];
    LayerSpec L = B.base(n + ".entry", t17, 0, 896, 256, 1, 1, 1, 0, 0);
    B.seg(L, 0, 128, cat17, 0);
    B.seg(L, 128, 256, e17, 0);
    B.push_weights(L, {&b0, &b10}, 896, 256);
    e->layers.push_back(L);
    B.ci += 2;
    B.conv(n + ".branch1.1", e17, 0, e17b, 0, 1, 0, 3);
    B.conv(n + ".branch1.2", e17b, 0, cat17, 128, 1, 3, 0);
    B.resid(n + ".conv2d", cat17, t17, EPI_RESID_RELU);
    e->layers.back().scale = RESID_SCALE_17;
  }
  // ---- Mixed_7a
  {
    const HostT &b00 = B.convs[B.ci], &b10 = B.convs[B.ci + 2], &b20 = B.convs[B.ci + 4];
    LayerSpec L = B.base("mixed_7a.entry", t17, 0, 896, 768, 1, 1, 1, 0, 0);
    B.seg(L, 0, 768, m7, 0);
    B.push_weights(L, {&b00, &b10, &b20}, 896, 768);
    e->layers.push_back(L);
    B.ci += 1;
    B.conv("mixed_7a.branch0.1", m7, 0, t8, 0, 2, 0, 0);
    B.ci += 1;
    B.conv("mixed_7a.branch1.1", m7, 256, t8, 384, 2, 0, 0);
    B.ci += 1;
    B.conv("mixed_7a.branch2.1", m7, 512, m7b, 0, 1, 1, 1);
    B.conv("mixed_7a.branch2.2", m7b, 0, t8, 640, 2, 0, 0);
    B.pool("mixed_7a.branch3", t17, 896, t8, 896);
  }
  // ---- 5 x Block8 (scale 0.2) + block8 (scale 1, no ReLU)
  for (int i = 0; i < 6; ++i) {
    const std::string n = i < 5 ? "repeat_3." + std::to_string(i) : std::string("block8");
    const HostT &b0 = B.convs[B.ci], &b10 = B.convs[B.ci + 1];
    LayerSpec L = B.base(n + ".entry", t8, 0, 1792, 384, 1, 1, 1, 0, 0);
    B.seg(L, 0, 192, cat8, 0);
    B.seg(L, 192, 384, e8, 0);
    B.push_weights(L, {&b0, &b10}, 1792, 384);
    e->layers.push_back(L);
    B.ci += 2;
    B.conv(n + ".branch1.1", e8, 0, e8b, 0, 1, 0, 1);
    B.conv(n + ".branch1.2", e8b, 0, cat8, 192, 1, 1, 0);
    B.resid(n + ".conv2d", cat8, t8, i < 5 ? EPI_RESID_RELU : EPI_RESID);
    e->layers.back().scale = i < 5 ? RESID_SCALE_8 : 1.0f;
  }
  if (B.ci != B.convs.size() || B.ri != B.resids.size())
    TRL_FAIL(c, TRL_E_STATE, "facenet plan consumed %zu/%zu convs, %zu/%zu resids", B.ci, B.convs.size(), B.ri, B.resids.size());

  // ---- upload arenas
  TRL_CUDA(c, cudaMalloc(&e->d_w, B.w_arena.size() * sizeof(bf16)));
  TRL_CUDA(c, cudaMemcpy(e->d_w, B.w_arena.data(), B.w_arena.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  TRL_CUDA(c, cudaMalloc(&e->d_bias, B.b_arena.size() * 4));
  TRL_CUDA(c, cudaMemcpy(e->d_bias, B.b_arena.data(), B.b_arena.size() * 4, cudaMemcpyHostToDevice));
  {
    std::vector<float> wt((size_t)1792 * 512);
    for (int o = 0; o < 512; ++o)
      for (int k = 0; k < 1792; ++k) wt[(size_t)k * 512 + o] = head_w[(size_t)o * 1792 + k];
    TRL_CUDA(c, cudaMalloc(&e->d_head_w, wt.size() * 4));
    TRL_CUDA(c, cudaMemcpy(e->d_head_w, wt.data(), wt.size() * 4, cudaMemcpyHostToDevice));
    TRL_CUDA(c, cudaMalloc(&e->d_head_b, 512 * 4));
    TRL_CUDA(c, cudaMemcpy(e->d_head_b, head_b, 512 * 4, cudaMemcpyHostToDevice));
  }
  if (c->cfg.facenet_impl == 0) {
    int rc = umma_init(c);
    if (rc != TRL_OK) return rc;
  }
  return TRL_OK;
}

void facenet_destroy(trl_ctx* c) {
  FaceNetEngine* e = c->facenet;
  if (!e) return;
  for (bf16* p : e->d_act) cudaFree(p);
  cudaFree(e->d_w); cudaFree(e->d_bias); cudaFree(e->d_stem_w); cudaFree(e->d_stem_b); cudaFree(e->d_stem_w_std); cudaFree(e->d_stem_b_std);
  cudaFree(e->d_cmp_crops); cudaFree(e->d_cmp_emb); cudaFree(e->d_cmp_rank); cudaFree(e->d_head_w); cudaFree(e->d_head_b);
  delete e;
  c->facenet = nullptr;
}

// (re)instantiate buffers, ops and tensor maps for crop size S and batch capacity cap
static int facenet_prepare(trl_ctx* c, int S, int cap) {
  FaceNetEngine* e = c->facenet;
  if (e->S == S && e->cap >= cap) return TRL_OK;
  if (S < 75) TRL_FAIL(c, TRL_E_INVALID, "crop size %d too small for InceptionResnetV1 (needs >= 75)", S);
  for (bf16* p : e->d_act) cudaFree(p);
  e->d_act.clear();
  e->act_elems.clear();
  e->steps.clear();
  e->S = S;
  e->cap = cap;
  int* z = e->sz;
  z[L1] = (S - 3) / 2 + 1; z[L2] = z[L1] - 2; z[L4] = (z[L2] - 3) / 2 + 1; z[L6] = z[L4] - 2;
  z[L7] = (z[L6] - 3) / 2 + 1; z[L8] = (z[L7] - 3) / 2 + 1; z[L9] = (z[L8] - 3) / 2 + 1;
  for (const BufSpec& b : e->bufs) {
    const size_t elems = (size_t)cap * z[b.lvl] * z[b.lvl] * b.C;
    bf16* p = nullptr;
    TRL_CUDA(c, cudaMalloc(&p, elems * sizeof(bf16) + 256));
    TRL_CUDA(c, cudaMemset(p, 0, elems * sizeof(bf16) + 256));
    e->d_act.push_back(p);
    e->act_elems.push_back(elems);
  }
  for (const LayerSpec& L : e->layers) {
    FaceNetEngine::Step st{};
    st.kind = L.kind;
    const BufSpec& sb = e->bufs[L.src];
    const int hin = z[sb.lvl];
    if (L.kind == 1) {
      PoolOp& po = st.pool;
      const SegSpec& sg = L.seg[0];
      const BufSpec& db = e->bufs[sg.buf];
      po.in = e->d_act[L.src]; po.Hin = hin; po.Win = hin; po.C = L.Cin; po.in_ctot = sb.C; po.in_coff = 0;
      po.out = e->d_act[sg.buf]; po.Hout = z[db.lvl]; po.Wout = z[db.lvl]; po.out_ctot = db.C; po.out_coff = sg.coff;
      if (po.Hout != (hin - 3) / 2 + 1) TRL_FAIL(c, TRL_E_STATE, "%s: pool geometry", L.name.c_str());
    } else {
      ConvOp& op = st.conv;
      memset(&op, 0, sizeof(op));
      snprintf(op.name, sizeof(op.name), "%s", L.name.c_str());
      op.in = e->d_act[L.src]; op.Hin = hin; op.Win = hin; op.Cin = L.Cin; op.in_ctot = sb.C; op.in_coff = L.src_coff;
      op.kh = L.kh; op.kw = L.kw; op.stride = L.stride; op.pad_h = L.ph; op.pad_w = L.pw;
      op.Hout = (hin + 2 * L.ph - L.kh) / L.stride + 1;
      op.Wout = (hin + 2 * L.pw - L.kw) / L.stride + 1;
      op.Cout = L.Cout;
      op.w = e->d_w + L.w_off;
      op.bias = e->d_bias + L.b_off;
      op.epi = L.epi; op.scale = L.scale;
      op.resid = L.resid >= 0 ? e->d_act[L.resid] : nullptr;
      op.nseg = L.nseg;
      for (int s = 0; s < L.nseg; ++s) {
        const BufSpec& db = e->bufs[L.seg[s].buf];
        if (z[db.lvl] != op.Hout || op.Hout != op.Wout) TRL_FAIL(c, TRL_E_STATE, "%s: output geometry %d vs %d", L.name.c_str(), op.Hout, z[db.lvl]);
        op.seg[s] = {L.seg[s].n_begin, L.seg[s].n_end, e->d_act[L.seg[s].buf], db.C, L.seg[s].coff};
      }
      if (c->cfg.facenet_impl == 0) {
        int rc = umma_encode_maps(c, op, cap);
        if (rc != TRL_OK) return rc;
      }
    }
    e->steps.push_back(st);
  }
  return TRL_OK;
}

int facenet_forward(trl_ctx* c, const uint8_t* d_crops, int n, int S, int norm, float* d_emb, cudaStream_t s, const int* d_n) {
  FaceNetEngine* e = c->facenet;
  if (!e) TRL_FAIL(c, TRL_E_STATE, "facenet weights not loaded");
  if (n <= 0) return TRL_OK;
  int cap = e->cap;
  if (e->S != S || cap < n) {
    cap = cap < n ? ((n + 63) / 64) * 64 : cap;
    int rc = facenet_prepare(c, S, cap);
    if (rc != TRL_OK) return rc;
  }
  int rc = launch_stem_conv(c, d_crops, n, S, norm ? e->d_stem_w_std : e->d_stem_w, norm ? e->d_stem_b_std : e->d_stem_b,
                            e->d_act[e->buf_stem], e->sz[L1], s, d_n);
  if (rc != TRL_OK) return rc;
  for (const FaceNetEngine::Step& st : e->steps) {
    if (st.kind == 1) rc = launch_maxpool(c, st.pool, n, s, d_n);
    else rc = (c->cfg.facenet_impl == 0) ? launch_conv_umma(c, st.conv, n, s, d_n) : launch_conv_simt(c, st.conv, n, s, d_n);
    if (rc != TRL_OK) return rc;
  }
  return launch_head(c, e->d_act[e->buf_final], n, e->sz[L9] * e->sz[L9], e->d_head_w, e->d_head_b, d_emb, s, d_n);
}

// ---- face-bearing crops only (server/model.py:48 skips frames without a face before anything is embedded)
// rank[i] = position of crop i among the valid ones (exclusive prefix of valid), *count = number of valid crops.  One CTA.
__global__ void __launch_bounds__(1024) valid_prefix_kernel(const uint8_t* __restrict__ valid, int n, int* __restrict__ rank,
                                                           int* __restrict__ count) {
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < n && valid[i]) ? 1 : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    int add = carry;
    for (int q = 0; q < w; ++q) add += wsum[q];
    if (i < n) rank[i] = add + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = add + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}

// crops of the valid frames, packed front to back (16-byte copies; a crop is S*S*3 bytes, a multiple of 16 for even S)
__global__ void __launch_bounds__(256) gather_crops_kernel(const uint8_t* __restrict__ crops, const uint8_t* __restrict__ valid,
                                                          const int* __restrict__ rank, int bytes, uint8_t* __restrict__ out) {
  const int i = blockIdx.x;
  if (!valid[i]) return;
  const uint8_t* src = crops + (size_t)i * bytes;
  uint8_t* dst = out + (size_t)rank[i] * bytes;
  if ((bytes & 15) == 0) {
    for (int k = threadIdx.x; k < bytes / 16; k += blockDim.x)
      reinterpret_cast<uint4*>(dst)[k] = reinterpret_cast<const uint4*>(src)[k];
  } else {
    for (int k = threadIdx.x; k < bytes; k += blockDim.x) dst[k] = src[k];
  }
}

// emb[i] = packed embedding of frame i, zeros for a frame without a face
__global__ void __launch_bounds__(128) scatter_emb_kernel(const float* __restrict__ packed, const uint8_t* __restrict__ valid,
                                                         const int* __restrict__ rank, float* __restrict__ emb) {
  const int i = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(packed + (size_t)(valid[i] ? rank[i] : 0) * TRL_EMB_DIM);
  float4* dst = reinterpret_cast<float4*>(emb + (size_t)i * TRL_EMB_DIM);
  dst[threadIdx.x] = valid[i] ? src[threadIdx.x] : make_float4(0.f, 0.f, 0.f, 0.f);
}

int facenet_forward_valid(trl_ctx* c, const uint8_t* d_crops, const uint8_t* d_valid, int n, int S, int norm, float* d_emb,
                          cudaStream_t s) {
  FaceNetEngine* e = c->facenet;
  if (!e) TRL_FAIL(c, TRL_E_STATE, "facenet weights not loaded");
  if (n <= 0) return TRL_OK;
  const size_t bytes = (size_t)S * S * 3;
  if (e->cmp_cap < n || e->cmp_S != S) {
    cudaFree(e->d_cmp_crops); cudaFree(e->d_cmp_emb); cudaFree(e->d_cmp_rank);
    e->d_cmp_crops = nullptr; e->d_cmp_emb = nullptr; e->d_cmp_rank = nullptr; e->cmp_cap = 0;
    const int cap = ((n + 63) / 64) * 64;
    TRL_CUDA(c, cudaMalloc(&e->d_cmp_crops, (size_t)cap * bytes + 256));
    TRL_CUDA(c, cudaMalloc(&e->d_cmp_emb, (size_t)cap * TRL_EMB_DIM * sizeof(float)));
    TRL_CUDA(c, cudaMalloc(&e->d_cmp_rank, ((size_t)cap + 4) * sizeof(int)));
    e->cmp_cap = cap; e->cmp_S = S;
  }
  int* d_count = e->d_cmp_rank + e->cmp_cap;
  valid_prefix_kernel<<<1, 1024, 0, s>>>(d_valid, n, e->d_cmp_rank, d_count);
  TRL_LAUNCH_CHECK(c);
  gather_crops_kernel<<<n, 256, 0, s>>>(d_crops, d_valid, e->d_cmp_rank, (int)bytes, e->d_cmp_crops);
  TRL_LAUNCH_CHECK(c);
  int rc = facenet_forward(c, e->d_cmp_crops, n, S, norm, e->d_cmp_emb, s, d_count);
  if (rc != TRL_OK) return rc;
  scatter_emb_kernel<<<n, TRL_EMB_DIM / 4, 0, s>>>(e->d_cmp_emb, d_valid, e->d_cmp_rank, d_emb);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}


// debug / profiling: one forward pass with a CUDA-event pair around every step.  info[i] = {kind, Cin, Cout, kh, kw, Hout,
// block_n, block_k}, ms[i] = device time of step i.  Returns the number of steps (<= max_steps) or a negative error.
extern "C" int trl_debug_facenet_step_times(trl_ctx* c, const uint8_t* d_crops, int n, int S, float* d_emb, int max_steps,
                                            int* info /*[max_steps][8]*/, float* ms /*[max_steps]*/, void* stream) {
  FaceNetEngine* e = c->facenet;
  if (!e || !info || !ms) return TRL_E_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = facenet_forward(c, d_crops, n, S, 0, d_emb, s);      // sizes the engine, warms up
  if (rc != TRL_OK) return rc;
  const int ns = (int)e->steps.size();
  if (ns > max_steps) return TRL_E_INVALID;
  std::vector<cudaEvent_t> ev(ns + 1);
  for (auto& x : ev) cudaEventCreate(&x);
  cudaEventRecord(ev[0], s);
  for (int i = 0; i < ns; ++i) {
    const FaceNetEngine::Step& st = e->steps[i];
    if (st.kind == 1) rc = launch_maxpool(c, st.pool, n, s);
    else rc = (c->cfg.facenet_impl == 0) ? launch_conv_umma(c, st.conv, n, s) : launch_conv_simt(c, st.conv, n, s);
    if (rc != TRL_OK) return rc;
    cudaEventRecord(ev[i + 1], s);
    int* f = info + i * 8;
    if (st.kind == 1) { f[0] = 1; f[1] = st.pool.C; f[2] = st.pool.C; f[3] = 3; f[4] = 3; f[5] = st.pool.Hout; f[6] = 0; f[7] = 0; }
    else { f[0] = 0; f[1] = st.conv.Cin; f[2] = st.conv.Cout; f[3] = st.conv.kh; f[4] = st.conv.kw; f[5] = st.conv.Hout;
           f[6] = st.conv.block_n; f[7] = st.conv.block_k; }
  }
  cudaStreamSynchronize(s);
  for (int i = 0; i < ns; ++i) cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
  for (auto& x : ev) cudaEventDestroy(x);
  return ns;
}

// debug / validation access to intermediate activations (tests compare the tcgen05 path layer by layer)
extern "C" int trl_debug_facenet_num_layers(trl_ctx* c) { return c->facenet ? (int)c->facenet->layers.size() : 0; }

extern "C" int trl_debug_facenet_layer(trl_ctx* c, int idx, char* name, int name_len, int* dims /*[H, W, Cout, nseg]*/) {
  FaceNetEngine* e = c->facenet;
  if (!e || idx < 0 || idx >= (int)e->steps.size()) return TRL_E_INVALID;
  const LayerSpec& L = e->layers[idx];
  snprintf(name, name_len, "%s", L.name.c_str());
  const int h = e->sz[e->bufs[L.seg[0].buf].lvl];
  dims[0] = h; dims[1] = h; dims[2] = L.Cout; dims[3] = L.nseg;
  return TRL_OK;
}

// copies layer idx's output (all segments concatenated along channels) for the first n images to host as bf16 bits
extern "C" int trl_debug_facenet_output(trl_ctx* c, int idx, int n, uint16_t* h_out) {
  FaceNetEngine* e = c->facenet;
  if (!e || idx < 0 || idx >= (int)e->steps.size()) return TRL_E_INVALID;
  const LayerSpec& L = e->layers[idx];
  size_t col = 0;
  const int h = e->sz[e->bufs[L.seg[0].buf].lvl];
  const size_t pix = (size_t)n * h * h;
  for (int s = 0; s < L.nseg; ++s) {
    const SegSpec& sg = L.seg[s];
    const BufSpec& db = e->bufs[sg.buf];
    const int wdt = sg.n_end - sg.n_begin;
    TRL_CUDA(c, cudaMemcpy2D(h_out + col, (size_t)L.Cout * 2, e->d_act[sg.buf] + sg.coff, (size_t)db.C * 2, (size_t)wdt * 2, pix,
                             cudaMemcpyDeviceToHost));
    col += wdt;
  }
  return TRL_OK;
}
