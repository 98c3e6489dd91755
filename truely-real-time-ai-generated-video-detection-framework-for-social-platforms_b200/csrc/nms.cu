// K4 / K5 / K6: greedy NMS (warp-ballot bitmask, blocked by 32) fused with the box arithmetic that follows it
// at each point of the MTCNN cascade.  One CTA per NMS group (frame x scale, or frame).
//
// upstream semantics:
//   mode 0  torchvision.ops.nms : stable sort by score descending, area (x2-x1)(y2-y1), suppress IoU > thr
//   mode 1  nms_numpy(.., 'Min'): argsort ascending taken from the back (ties: later index first),
//           area (x2-x1+1)(y2-y1+1), overlap inter/min(area), keep o <= thr
// Groups of up to MAXN = 2048 candidates run entirely in shared memory (the product path: capacities <= 2048).  When the
// host raises the capacities beyond that (trl_set_capacity, the overflow-retry path of model.analyze_stream: upstream
// detect_face has no cap at all), groups with more than MAXN live candidates are handled by cascade_nms_big_kernel, the
// same algorithm with its arrays in a global-memory scratch region (up to BIGN = 16384 per group).
// Box arithmetic is written with explicit round-to-nearest intrinsics so that nvcc cannot contract a*b+c into an FMA:
// the reference evaluates every product and sum as a separate fp32 op, and pad() truncates the result to pixels.
#include "common.cuh"

namespace nms {

constexpr int MAXN = 2048;
constexpr int BIGN = 16384;   // largest group the global-memory variant accepts
constexpr int THREADS = 512;
constexpr int PER_T = MAXN / THREADS;

struct Smem {
  unsigned long long key[MAXN];
  float4 sbox[MAXN];      // sorted
  float sarea[MAXN];
  int sslot[MAXN];        // sorted position -> input slot
  int kept[MAXN];         // pick order -> sorted position
  unsigned char removed[MAXN];
  unsigned int keepmask;
  int nkeep;
  int nlive;
  int base;
};

template <int MODE>
__device__ __forceinline__ float box_area(const float4& b) {
  if (MODE == 0) return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
}

// true if `b` (lower score) is suppressed by `a`
template <int MODE>
__device__ __forceinline__ bool suppresses(const float4& a, float aa, const float4& b, float ab, float thr) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  if (MODE == 0) {
    const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
    return ovr > thr;
  } else {
    const float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f)), h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
    const float inter = __fmul_rn(w, h);
    const float o = __fdiv_rn(inter, fminf(aa, ab));
    return !(o <= thr);
  }
}

__device__ __forceinline__ unsigned long long make_key(float score, uint32_t tie) {
  // ascending sort of this key == score descending, then tie ascending.  Scores are probabilities (> 0).
  return ((unsigned long long)(~__float_as_uint(score)) << 32) | tie;
}
constexpr unsigned long long KEY_DEAD = ~0ull;

// Sort the n live entries (key != KEY_DEAD) and run greedy NMS.  Each thread passes in the boxes/keys of the
// slots it owns (slot = tid + k*THREADS).  On return sm.kept[0..nkeep) holds sorted positions in pick order,
// sm.sslot maps sorted position -> slot, sm.sbox the sorted boxes.  Returns nkeep.
template <int MODE>
__device__ int sort_and_nms(Smem& sm, int n, const float4 (&mybox)[PER_T], const unsigned long long (&mykey)[PER_T],
                            float thr) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int k = 0; k < PER_T; ++k) {
    const int i = tid + k * THREADS;
    if (i < n) sm.key[i] = mykey[k];
  }
  if (tid == 0) { sm.nkeep = 0; sm.nlive = 0; }
  __syncthreads();
  // rank sort (keys of live entries are unique)
  int rank[PER_T];
#pragma unroll
  for (int k = 0; k < PER_T; ++k) rank[k] = 0;
  for (int j = 0; j < n; ++j) {
    const unsigned long long kj = sm.key[j];
#pragma unroll
    for (int k = 0; k < PER_T; ++k) rank[k] += (kj < mykey[k]) ? 1 : 0;
  }
  int live = 0;
#pragma unroll
  for (int k = 0; k < PER_T; ++k) {
    const int i = tid + k * THREADS;
    if (i < n && mykey[k] != KEY_DEAD) {
      sm.sbox[rank[k]] = mybox[k];
      sm.sarea[rank[k]] = box_area<MODE>(mybox[k]);
      sm.sslot[rank[k]] = i;
      sm.removed[rank[k]] = 0;
      ++live;
    }
  }
  // number of live entries (sm.nlive was zeroed before the first barrier)
  if (live) atomicAdd(&sm.nlive, live);
  __syncthreads();
  const int nl = sm.nlive;

  const int lane = tid & 31, warp = tid >> 5;
  for (int blk = 0; blk * 32 < nl; ++blk) {
    if (warp == 0) {
      const int i = blk * 32 + lane;
      const bool valid = i < nl;
      float4 bx = valid ? sm.sbox[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float ar = valid ? sm.sarea[i] : 0.f;
      bool alive = valid && !sm.removed[i];
      unsigned int supby = 0;   // bit k: entry k of this block (higher score) suppresses me
      for (int k = 0; k < 32; ++k) {
        float4 o;
        o.x = __shfl_sync(0xffffffffu, bx.x, k);
        o.y = __shfl_sync(0xffffffffu, bx.y, k);
        o.z = __shfl_sync(0xffffffffu, bx.z, k);
        o.w = __shfl_sync(0xffffffffu, bx.w, k);
        const float oa = __shfl_sync(0xffffffffu, ar, k);
        if (valid && lane > k && suppresses<MODE>(o, oa, bx, ar, thr)) supby |= 1u << k;
      }
      for (int k = 0; k < 32; ++k) {
        const bool ak = __shfl_sync(0xffffffffu, (int)alive, k) != 0;
        if (ak && ((supby >> k) & 1u)) alive = false;
      }
      const unsigned int km = __ballot_sync(0xffffffffu, alive);
      const int pos = sm.nkeep + __popc(km & ((1u << lane) - 1u));
      if (alive) sm.kept[pos] = i;
      __syncwarp();
      if (lane == 0) { sm.keepmask = km; sm.nkeep += __popc(km); }
    }
    __syncthreads();
    const unsigned int km = sm.keepmask;
    if (km) {
      for (int j = (blk + 1) * 32 + tid; j < nl; j += THREADS) {
        if (sm.removed[j]) continue;
        const float4 bj = sm.sbox[j];
        const float aj = sm.sarea[j];
        unsigned int mk = km;
        while (mk) {
          const int k = __ffs(mk) - 1;
          mk &= mk - 1;
          const int i = blk * 32 + k;
          if (suppresses<MODE>(sm.sbox[i], sm.sarea[i], bj, aj, thr)) { sm.removed[j] = 1; break; }
        }
      }
    }
    __syncthreads();
  }
  return sm.nkeep;
}

// ---- box arithmetic (fp32, one rounding per op, exactly as the torch expressions evaluate)

__device__ __forceinline__ float4 rerec(float4 b) {
  const float h = __fsub_rn(b.w, b.y), w = __fsub_rn(b.z, b.x);
  const float l = fmaxf(w, h);
  float4 o;
  o.x = __fsub_rn(__fadd_rn(b.x, __fmul_rn(w, 0.5f)), __fmul_rn(l, 0.5f));
  o.y = __fsub_rn(__fadd_rn(b.y, __fmul_rn(h, 0.5f)), __fmul_rn(l, 0.5f));
  o.z = __fadd_rn(o.x, l);
  o.w = __fadd_rn(o.y, l);
  return o;
}

__device__ __forceinline__ float4 bbreg_plus1(float4 b, float r0, float r1, float r2, float r3) {
  const float w = __fadd_rn(__fsub_rn(b.z, b.x), 1.f), h = __fadd_rn(__fsub_rn(b.w, b.y), 1.f);
  float4 o;
  o.x = __fadd_rn(b.x, __fmul_rn(r0, w));
  o.y = __fadd_rn(b.y, __fmul_rn(r1, h));
  o.z = __fadd_rn(b.z, __fmul_rn(r2, w));
  o.w = __fadd_rn(b.w, __fmul_rn(r3, h));
  return o;
}

__device__ __forceinline__ void pad_box(float4 b, int W, int H, int* out4) {
  int x = (int)b.x, y = (int)b.y, ex = (int)b.z, ey = (int)b.w;   // trunc toward zero
  if (x < 1) x = 1;
  if (y < 1) y = 1;
  if (ex > W) ex = W;
  if (ey > H) ey = H;
  out4[0] = y; out4[1] = ey; out4[2] = x; out4[3] = ex;
}

__device__ __forceinline__ void flag_overflow(CapFlag* f, int stage, int frame, int count, int cap) {
  if (f) { f->overflow = 1; f->stage = stage; f->frame = frame; f->count = count; f->capacity = cap; }
}


// Epilogue of one cascade stage, shared by the shared-memory and the global-memory kernels: kept[0..nkeep) = sorted
// positions in pick order, sslot = sorted position -> input slot, sbox = sorted boxes, sscore = score by input slot.
template <int STAGE>
__device__ __forceinline__ void stage_epilogue(const StageParams& p, int b, int lvl, int seg, const Cand* in, int nkeep,
                                               const int* kept, const int* sslot, const float4* sbox, const float* sscore,
                                               int* s_base) {
  const int tid = threadIdx.x;
  if (STAGE == 1) {
    if (tid == 0) (*s_base) = atomicAdd(&p.cnt_out[b], nkeep);
    __syncthreads();
    const int base = (*s_base);
    if (base + nkeep > p.cap_out && tid == 0) flag_overflow(p.capflag, 2, b, base + nkeep, p.cap_out);
    for (int r = tid; r < nkeep; r += THREADS) {
      if (base + r >= p.cap_out) break;
      const int sp = kept[r];
      Cand c = in[sslot[sp]];
      c.key = ((uint32_t)lvl << 16) | (uint32_t)r;   // position in upstream's concatenated scale_picks list
      p.out[(size_t)b * p.cap_out + base + r] = c;
    }
  } else if (STAGE == 2 || STAGE == 3) {
    if (nkeep > p.cap_out && tid == 0) flag_overflow(p.capflag, STAGE == 2 ? 2 : 3, b, nkeep, p.cap_out);
    const int nk = min(nkeep, p.cap_out);
    for (int r = tid; r < nk; r += THREADS) {
      const int sp = kept[r];
      const int slot = sslot[sp];
      const float4 bx = sbox[sp];
      float4 q;
      if (STAGE == 2) {
        const Cand c = in[slot];
        const float regw = __fsub_rn(bx.z, bx.x), regh = __fsub_rn(bx.w, bx.y);
        q.x = __fadd_rn(bx.x, __fmul_rn(c.r0, regw));
        q.y = __fadd_rn(bx.y, __fmul_rn(c.r1, regh));
        q.z = __fadd_rn(bx.z, __fmul_rn(c.r2, regw));
        q.w = __fadd_rn(bx.w, __fmul_rn(c.r3, regh));
      } else {
        const float* rg = p.reg + ((size_t)seg * p.cap_in + slot) * 4;
        q = bbreg_plus1(bx, rg[0], rg[1], rg[2], rg[3]);
      }
      q = rerec(q);
      Cand o;
      o.x1 = q.x; o.y1 = q.y; o.x2 = q.z; o.y2 = q.w;
      o.score = sscore[slot];
      o.r0 = o.r1 = o.r2 = o.r3 = 0.f;
      o.key = (uint32_t)r;
      p.out[(size_t)b * p.cap_out + r] = o;
      pad_box(q, p.W, p.H, p.pad_out + ((size_t)b * p.cap_out + r) * 4);
    }
    if (tid == 0) p.cnt_out[b] = nk;
  } else {   // STAGE 4
    const int nk = min(nkeep, p.cap_out);
    if (nkeep > p.cap_out && tid == 0) flag_overflow(p.capflag, 4, b, nkeep, p.cap_out);
    // MTCNN.detect select_largest: order = argsort(area)[::-1]  (ties: later pick index first)
    for (int r = tid; r < nk; r += THREADS) {
      const float4 bx = sbox[kept[r]];
      const float ar = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
      int pos = 0;
      for (int q = 0; q < nk; ++q) {
        const float4 bq = sbox[kept[q]];
        const float aq = __fmul_rn(__fsub_rn(bq.z, bq.x), __fsub_rn(bq.w, bq.y));
        pos += (aq > ar || (aq == ar && q > r)) ? 1 : 0;
      }
      float* o = p.boxes_out + ((size_t)b * p.cap_out + pos) * 5;
      o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
      o[4] = sscore[sslot[kept[r]]];
    }
    if (tid == 0) p.cnt_out[b] = nk;
  }
}

// STAGE 1: per (frame, level) NMS 0.5 on P-Net candidates -> append to the frame list
// STAGE 2: per frame NMS 0.7 across levels -> stage-1 regression, rerec, pad  (R-Net inputs)
// STAGE 3: per frame: R-Net score > thr, NMS 0.7 -> bbreg, rerec, pad         (O-Net inputs)
// STAGE 4: per frame: O-Net score > thr, bbreg, NMS 0.7 'Min' -> sort largest-area first (final boxes)
template <int STAGE>
__global__ void __launch_bounds__(THREADS) cascade_nms_kernel(const StageParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  constexpr int MODE = (STAGE == 4) ? 1 : 0;
  const int tid = threadIdx.x;
  const int b = (STAGE == 1) ? blockIdx.y : blockIdx.x;
  const int lvl = (STAGE == 1) ? blockIdx.x : 0;
  const int seg = (STAGE == 1) ? b * p.n_levels + lvl : b;
  int n = p.cnt_in[seg];
  if (n > p.cap_in) n = p.cap_in;   // overflow already flagged by the producer
  if (n > MAXN) return;             // raised capacities only: cascade_nms_big_kernel owns this group
  const Cand* in = p.in + (size_t)seg * p.cap_in;

  float4 mybox[PER_T];
  unsigned long long mykey[PER_T];
  float myscore[PER_T];
#pragma unroll
  for (int k = 0; k < PER_T; ++k) {
    const int i = tid + k * THREADS;
    mykey[k] = KEY_DEAD;
    mybox[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    myscore[k] = 0.f;
    if (i < n) {
      const Cand c = in[i];
      float4 bx = make_float4(c.x1, c.y1, c.x2, c.y2);
      float score = c.score;
      bool pass = true;
      uint32_t tie = c.key;
      if (STAGE == 3 || STAGE == 4) {
        score = p.prob[(size_t)seg * p.cap_in + i];
        const int* pd = p.pad_in + ((size_t)seg * p.cap_in + i) * 4;
        const bool crop_ok = (pd[1] > pd[0] - 1) && (pd[3] > pd[2] - 1);
        pass = crop_ok && (score > p.thr_score);
        tie = (uint32_t)i;
        if (STAGE == 4) {
          const float* r = p.reg + ((size_t)seg * p.cap_in + i) * 4;
          bx = bbreg_plus1(bx, r[0], r[1], r[2], r[3]);
          tie = 0xffffffffu - (uint32_t)i;     // nms_numpy: among equal scores the later index is picked first
        }
      }
      mybox[k] = bx;
      myscore[k] = score;
      if (pass) mykey[k] = make_key(score, tie);
    }
  }
  const int nkeep = sort_and_nms<MODE>(sm, n, mybox, mykey, p.thr_nms);
  // scores by slot for the epilogue
  float* sscore = reinterpret_cast<float*>(sm.key);   // keys are dead now; reuse as score-by-slot
  __syncthreads();
#pragma unroll
  for (int k = 0; k < PER_T; ++k) {
    const int i = tid + k * THREADS;
    if (i < n) sscore[i] = myscore[k];
  }
  __syncthreads();

  stage_epilogue<STAGE>(p, b, lvl, seg, in, nkeep, sm.kept, sm.sslot, sm.sbox, sscore, &sm.base);
}

// ---- global-memory variant for groups of MAXN < n <= BIGN candidates (raised capacities, see the file header).
// Same rank sort + blocked greedy NMS, arrays in a per-group scratch region instead of shared memory.
struct BigScratch {       // one group's arrays, each `cap` entries
  unsigned long long* key;   // by input slot
  float4* ubox;              // by input slot
  float* uscore;             // by input slot
  float4* sbox;              // sorted
  float* sarea;
  int* sslot;
  int* kept;
  unsigned char* removed;
};
constexpr size_t BIG_BYTES_PER_ENTRY = 64;   // 8 + 16 + 4 + 16 + 4 + 4 + 4 + 1, rounded up
__device__ __forceinline__ BigScratch big_scratch(unsigned char* base, int group, int cap) {
  unsigned char* q = base + (size_t)group * cap * BIG_BYTES_PER_ENTRY;
  BigScratch g;
  g.ubox = reinterpret_cast<float4*>(q);                 q += (size_t)cap * 16;
  g.sbox = reinterpret_cast<float4*>(q);                 q += (size_t)cap * 16;
  g.key = reinterpret_cast<unsigned long long*>(q);      q += (size_t)cap * 8;
  g.uscore = reinterpret_cast<float*>(q);                q += (size_t)cap * 4;
  g.sarea = reinterpret_cast<float*>(q);                 q += (size_t)cap * 4;
  g.sslot = reinterpret_cast<int*>(q);                   q += (size_t)cap * 4;
  g.kept = reinterpret_cast<int*>(q);                    q += (size_t)cap * 4;
  g.removed = q;
  return g;
}

template <int STAGE>
__global__ void __launch_bounds__(THREADS) cascade_nms_big_kernel(const StageParams p, unsigned char* scratch) {
  constexpr int MODE = (STAGE == 4) ? 1 : 0;
  __shared__ unsigned int s_keepmask;
  __shared__ int s_nkeep, s_nlive, s_base;
  const int tid = threadIdx.x;
  const int b = (STAGE == 1) ? blockIdx.y : blockIdx.x;
  const int lvl = (STAGE == 1) ? blockIdx.x : 0;
  const int seg = (STAGE == 1) ? b * p.n_levels + lvl : b;
  int n = p.cnt_in[seg];
  if (n > p.cap_in) n = p.cap_in;
  if (n <= MAXN) return;            // the shared-memory kernel owns this group
  const Cand* in = p.in + (size_t)seg * p.cap_in;
  const BigScratch g = big_scratch(scratch, seg, p.cap_in);
  for (int i = tid; i < n; i += THREADS) {
    const Cand c = in[i];
    float4 bx = make_float4(c.x1, c.y1, c.x2, c.y2);
    float score = c.score;
    bool pass = true;
    uint32_t tie = c.key;
    if (STAGE == 3 || STAGE == 4) {
      score = p.prob[(size_t)seg * p.cap_in + i];
      const int* pd = p.pad_in + ((size_t)seg * p.cap_in + i) * 4;
      const bool crop_ok = (pd[1] > pd[0] - 1) && (pd[3] > pd[2] - 1);
      pass = crop_ok && (score > p.thr_score);
      tie = (uint32_t)i;
      if (STAGE == 4) {
        const float* r = p.reg + ((size_t)seg * p.cap_in + i) * 4;
        bx = bbreg_plus1(bx, r[0], r[1], r[2], r[3]);
        tie = 0xffffffffu - (uint32_t)i;
      }
    }
    g.ubox[i] = bx;
    g.uscore[i] = score;
    g.key[i] = pass ? make_key(score, tie) : KEY_DEAD;
  }
  if (tid == 0) { s_nkeep = 0; s_nlive = 0; }
  __syncthreads();
  int live = 0;
  for (int i = tid; i < n; i += THREADS) {
    const unsigned long long ki = g.key[i];
    if (ki == KEY_DEAD) continue;
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (g.key[j] < ki) ? 1 : 0;
    const float4 bx = g.ubox[i];
    g.sbox[rank] = bx;
    g.sarea[rank] = box_area<MODE>(bx);
    g.sslot[rank] = i;
    g.removed[rank] = 0;
    ++live;
  }
  if (live) atomicAdd(&s_nlive, live);
  __syncthreads();
  const int nl = s_nlive;
  const int lane = tid & 31, warp = tid >> 5;
  for (int blk = 0; blk * 32 < nl; ++blk) {
    if (warp == 0) {
      const int i = blk * 32 + lane;
      const bool valid = i < nl;
      float4 bx = valid ? g.sbox[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float ar = valid ? g.sarea[i] : 0.f;
      bool alive = valid && !g.removed[i];
      unsigned int supby = 0;
      for (int k = 0; k < 32; ++k) {
        float4 o;
        o.x = __shfl_sync(0xffffffffu, bx.x, k);
        o.y = __shfl_sync(0xffffffffu, bx.y, k);
        o.z = __shfl_sync(0xffffffffu, bx.z, k);
        o.w = __shfl_sync(0xffffffffu, bx.w, k);
        const float oa = __shfl_sync(0xffffffffu, ar, k);
        if (valid && lane > k && suppresses<MODE>(o, oa, bx, ar, p.thr_nms)) supby |= 1u << k;
      }
      for (int k = 0; k < 32; ++k) {
        const bool ak = __shfl_sync(0xffffffffu, (int)alive, k) != 0;
        if (ak && ((supby >> k) & 1u)) alive = false;
      }
      const unsigned int km = __ballot_sync(0xffffffffu, alive);
      const int pos = s_nkeep + __popc(km & ((1u << lane) - 1u));
      if (alive) g.kept[pos] = i;
      __syncwarp();
      if (lane == 0) { s_keepmask = km; s_nkeep += __popc(km); }
    }
    __syncthreads();
    const unsigned int km = s_keepmask;
    if (km) {
      for (int j = (blk + 1) * 32 + tid; j < nl; j += THREADS) {
        if (g.removed[j]) continue;
        const float4 bj = g.sbox[j];
        const float aj = g.sarea[j];
        unsigned int mk = km;
        while (mk) {
          const int k = __ffs(mk) - 1;
          mk &= mk - 1;
          const int i = blk * 32 + k;
          if (suppresses<MODE>(g.sbox[i], g.sarea[i], bj, aj, p.thr_nms)) { g.removed[j] = 1; break; }
        }
      }
    }
    __syncthreads();
  }
  stage_epilogue<STAGE>(p, b, lvl, seg, in, s_nkeep, g.kept, g.sslot, g.sbox, g.uscore, &s_base);
}

// stand-alone NMS over arrays (trl_nms)
template <int MODE>
__global__ void __launch_bounds__(THREADS) plain_nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                                            int n, float thr, int* __restrict__ keep, int* __restrict__ nkeep_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  float4 mybox[PER_T];
  unsigned long long mykey[PER_T];
#pragma unroll
  for (int k = 0; k < PER_T; ++k) {
    const int i = tid + k * THREADS;
    mykey[k] = KEY_DEAD;
    mybox[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
      mybox[k] = make_float4(boxes[i * 4 + 0], boxes[i * 4 + 1], boxes[i * 4 + 2], boxes[i * 4 + 3]);
      mykey[k] = make_key(scores[i], MODE == 0 ? (uint32_t)i : 0xffffffffu - (uint32_t)i);
    }
  }
  const int nk = sort_and_nms<MODE>(sm, n, mybox, mykey, thr);
  __syncthreads();
  for (int r = tid; r < nk; r += THREADS) keep[r] = sm.sslot[sm.kept[r]];
  if (tid == 0) *nkeep_out = nk;
}

}  // namespace nms

int nms_init(trl_ctx* c) {
  using namespace nms;
  const int bytes = (int)sizeof(Smem);
  TRL_CUDA(c, cudaFuncSetAttribute(cascade_nms_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  TRL_CUDA(c, cudaFuncSetAttribute(cascade_nms_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  TRL_CUDA(c, cudaFuncSetAttribute(cascade_nms_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  TRL_CUDA(c, cudaFuncSetAttribute(cascade_nms_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  TRL_CUDA(c, cudaFuncSetAttribute(plain_nms_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  TRL_CUDA(c, cudaFuncSetAttribute(plain_nms_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return TRL_OK;
}

int nms_max_n() { return nms::MAXN; }
int nms_big_max_n() { return nms::BIGN; }
size_t nms_big_scratch_bytes(int groups, int cap) { return cap > nms::MAXN ? (size_t)groups * cap * nms::BIG_BYTES_PER_ENTRY : 0; }

int launch_plain_nms(trl_ctx* c, const float* d_boxes, const float* d_scores, int n, float thr, int mode, int* d_keep,
                     int* d_nkeep, cudaStream_t s) {
  using namespace nms;
  if (n > MAXN) TRL_FAIL(c, TRL_E_CAPACITY, "trl_nms: n=%d exceeds %d", n, MAXN);
  if (mode == 0) plain_nms_kernel<0><<<1, THREADS, sizeof(Smem), s>>>(d_boxes, d_scores, n, thr, d_keep, d_nkeep);
  else plain_nms_kernel<1><<<1, THREADS, sizeof(Smem), s>>>(d_boxes, d_scores, n, thr, d_keep, d_nkeep);
  TRL_LAUNCH_CHECK(c);
  return TRL_OK;
}

int launch_cascade_stage(trl_ctx* c, int stage, const nms::StageParams& p, int B, cudaStream_t s) {
  using namespace nms;
  if (B <= 0) return TRL_OK;
  const size_t sh = sizeof(Smem);
  switch (stage) {
    case 1: cascade_nms_kernel<1><<<dim3(p.n_levels, B), THREADS, sh, s>>>(p); break;
    case 2: cascade_nms_kernel<2><<<B, THREADS, sh, s>>>(p); break;
    case 3: cascade_nms_kernel<3><<<B, THREADS, sh, s>>>(p); break;
    case 4: cascade_nms_kernel<4><<<B, THREADS, sh, s>>>(p); break;
    default: TRL_FAIL(c, TRL_E_INVALID, "bad cascade stage %d", stage);
  }
  TRL_LAUNCH_CHECK(c);
  if (p.cap_in > MAXN) {
    // raised capacities (overflow retry): groups beyond the shared-memory limit go through the global-memory variant
    if (!c->d_nms_big) TRL_FAIL(c, TRL_E_STATE, "cascade stage %d: no scratch for groups above %d candidates", stage, MAXN);
    switch (stage) {
      case 1: cascade_nms_big_kernel<1><<<dim3(p.n_levels, B), THREADS, 0, s>>>(p, c->d_nms_big); break;
      case 2: cascade_nms_big_kernel<2><<<B, THREADS, 0, s>>>(p, c->d_nms_big); break;
      case 3: cascade_nms_big_kernel<3><<<B, THREADS, 0, s>>>(p, c->d_nms_big); break;
      default: cascade_nms_big_kernel<4><<<B, THREADS, 0, s>>>(p, c->d_nms_big); break;
    }
    TRL_LAUNCH_CHECK(c);
  }
  return TRL_OK;
}
