// Model handle shared by the GPT-2 kernels (gpt2.cu: token-sequential; gpt2_dense.cu: tcgen05 dense).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpt {

constexpr int G_E = 32;
constexpr int G_FF = 128;
constexpr int G_MAX_L = 8;
constexpr int G_WARPS = 4;
constexpr int G_THREADS = G_WARPS * 32;

// Weight matrices [in][out] are repacked at create time into per-lane quads:
//   P[(in/4)][out/32][lane][in%4]  so that one LDG.128 gives a lane the 4 consecutive-input weights of its
//   output (out = group*32 + lane), consumed by two FFMA2 (fma.rn.f32x2) against an LDS.128 of the inputs.
struct LayerW {
  const float *ln1_w, *ln1_b, *attn_b, *proj_b, *ln2_w, *ln2_b, *fc_b, *fc2_b;
  const float4 *attn_wP, *proj_wP, *fc_wP, *fc2_wP;
  const float *attn_w, *proj_w, *fc_w, *fc2_w;   // original [in][out] layouts (dense fp32 kernel)
  const uint4* wimg;  // bf16 B-operand image for the tcgen05 dense forward (see gpt2_dense.cu), 40 KB
  const uint4* wfrag; // bf16 mma.sync B fragments for the KV-cached decode (see gpt2.cu, WF_*), 24 KB
};

struct Gpt2Dev {
  int L, dx, du, din, H, n_pos;
  const float *wpe, *embed_wT, *embed_b, *pred_wT, *pred_b, *lnf_w, *lnf_b;
  LayerW layer[G_MAX_L];
};

}  // namespace dpt

struct dpt_gpt2 {
  dpt::Gpt2Dev dev;
  float* blob;
};

namespace dpt {
// LayerW::wfrag, in uint4 units: for every Conv1D weight W[K][N] and every (k-step ks of 16 inputs, pair ntp of two
// 8-output tiles) one uint4 per lane = {b0, b1 of tile 2 ntp, b0, b1 of tile 2 ntp + 1} of mma.m16n8k16 (col-major B):
// lane (g = lane / 4, c = lane % 4): b0 = W[16 ks + 2c, +1][n], b1 = W[16 ks + 2c + 8, +9][n], n = 8 tile + g.
constexpr int WF_QKV = 0;      // K = 32, N = 96 : 2 x 6 groups of 32 lanes
constexpr int WF_PROJ = 384;   // K = 32, N = 32 : 2 x 2
constexpr int WF_FC = 512;     // K = 32, N = 128: 2 x 8
constexpr int WF_FC2 = 1024;   // K = 128, N = 32: 8 x 2
constexpr int WF_UINT4 = 1536; // 24576 B per layer
// byte layout of LayerW::wimg (K64 tiles, see umma.cuh): B[n][k] = W[k][n] for every Conv1D weight W[in][out]
constexpr int WIMG_QKV = 0;            // [96 rows x 64]   12288 B
constexpr int WIMG_PROJ = 12288;       // [32 rows x 64]    4096 B
constexpr int WIMG_FC = 16384;         // [128 rows x 64]  16384 B
constexpr int WIMG_FC2 = 32768;        // 2 x [32 rows x 64] 8192 B (k 0..63, 64..127)
constexpr int WIMG_BYTES = 40960;

struct DenseParams {
  Gpt2Dev m;
  const float *query, *cs, *ca, *cns, *cr;
  int B, T, Ts, test, share;   // share: sequences per context row
  float* out;
};
int gpt2_dense_launch(const DenseParams& p, cudaStream_t st);        // gpt2_dense.cu (tcgen05, bf16 operands)
int gpt2_dense_long_launch(const DenseParams& p, cudaStream_t st);   // gpt2_dense.cu (tcgen05, 129..512 tokens)
int gpt2_dense_fp32_launch(const DenseParams& p, cudaStream_t st);   // gpt2_dense_fp32.cu (CUDA cores, fp32)
void gpt2_pack_wimg(const float* attn_w, const float* proj_w, const float* fc_w, const float* fc2_w, unsigned char* img,
                    cudaStream_t st);
}  // namespace dpt
