// FaceNet (InceptionResnetV1) engine: shared op descriptors.  See facenet_plan.cu for the layer program,
// facenet_umma.cu for the tcgen05 implicit-GEMM kernel and facenet_simt.cu for the direct-conv kernels.
#pragma once
#include <cuda.h>
#include "common.cuh"

typedef __nv_bfloat16 bf16;

// Output columns [n_begin, n_end) of a GEMM go to channel slice dst_coff.. of an NHWC buffer with dst_ctot channels.
// (concat-free Inception blocks: branch convs write straight into their slice of the block's concat buffer)
struct OutSeg {
  int n_begin, n_end;
  bf16* dst;
  int dst_ctot, dst_coff;
};

enum EpiMode { EPI_RELU = 0, EPI_RESID_RELU = 1, EPI_RESID = 2 };

// One convolution (or several 1x1 convs sharing an input, fused along N) as an implicit GEMM:
//   D[m, n] = sum_{ky,kx,ci} A[pixel(m) shifted by (ky,kx), ci] * Wt[n, (ky,kx,ci)]
//   out = bf16( epi( D + bias[n] ) )           EPI_RELU:        relu(v)
//                                               EPI_RESID_RELU:  relu(x + scale * v)   (x = trunk, same [m, n])
//                                               EPI_RESID:       x + scale * v
struct ConvOp {
  char name[40];
  // input: NHWC bf16 buffer, channel slice [in_coff, in_coff + Cin) of in_ctot
  const bf16* in;
  int Hin, Win, Cin, in_ctot, in_coff;
  int Hout, Wout, Cout;
  int kh, kw, stride, pad_h, pad_w;
  const bf16* w;        // [Cout][kh*kw*Cin], K ordered (ky, kx, ci)
  const float* bias;    // [Cout]
  int epi;
  float scale;
  const bf16* resid;    // [N,Hout,Wout,Cout] (EPI_RESID*)
  int nseg;
  OutSeg seg[4];
  // ---- tcgen05 path
  int block_k;          // 64 (128B swizzle) or 32 (64B swizzle)
  int block_n;          // UMMA N per CTA tile
  int box_w, box_h, box_n;   // 128-row M tile = box_w x box_h x box_n output pixels
  int flat;             // 1x1 stride-1: rows are consecutive pixels (box_w = 128)
  CUtensorMap tmap_a;   // 4-D (C, W, H, N) over the input slice
  CUtensorMap tmap_w;   // 2-D (K, Cout)
  CUtensorMap tmap_r;   // 2-D (Cout, pixels) over the residual operand (EPI_RESID* only)
};

struct PoolOp {       // MaxPool2d(3, stride 2), NHWC
  const bf16* in; int Hin, Win, C, in_ctot, in_coff;
  bf16* out; int Hout, Wout, out_ctot, out_coff;
};

// d_n (may be NULL): device-side live batch size <= n (face-bearing crops after compaction); n sizes the launch
int launch_conv_simt(trl_ctx* c, const ConvOp& op, int n, cudaStream_t s, const int* d_n = nullptr);
int launch_conv_umma(trl_ctx* c, const ConvOp& op, int n, cudaStream_t s, const int* d_n = nullptr);
int umma_init(trl_ctx* c);
int umma_encode_maps(trl_ctx* c, ConvOp& op, int n_cap);
int launch_stem_conv(trl_ctx* c, const uint8_t* d_crops, int n, int S, const float* w, const float* bias, bf16* out,
                     int Ho, cudaStream_t s, const int* d_n = nullptr);
int launch_maxpool(trl_ctx* c, const PoolOp& op, int n, cudaStream_t s, const int* d_n = nullptr);
int launch_head(trl_ctx* c, const bf16* feat, int n, int hw, const float* w_t, const float* bias, float* emb, cudaStream_t s,
                const int* d_n = nullptr);
