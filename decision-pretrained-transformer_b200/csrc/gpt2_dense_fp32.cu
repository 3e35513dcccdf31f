// Dense Transformer.forward(x) in fp32 on the CUDA cores (precision = 0, sequences of <= 512 tokens).
//
// Same decomposition as the tensor-core kernel (gpt2_dense.cu) -- one CTA per sequence, thread t owns token
// row t, residual stream in fp32 registers -- but every contraction is fp32 FFMA2 (fma.rn.f32x2) so the
// result meets the 1e-5 logit bar against the reference.  The layer's weights (48 KB fp32) and the
// sequence's K and V (2 x 16 KB) live in shared memory; a thread's matvec reads weight rows as broadcast
// LDS.128 (all threads of a warp read the same address) and keeps 32 outputs in registers.  Attention is a
// per-row online softmax over the keys <= t (K/V rows broadcast from shared memory).  Compared with the
// token-sequential kernel there is no K/V traffic to HBM and all tokens of a sequence advance in parallel.
#include "common.cuh"
#include "gpt2_model.cuh"

namespace dpt {

// shared memory (floats); the CTA has THREADS = 128 / 256 / 512 threads = token rows
constexpr int DF_WQKV = 0;                     // [32][96]
constexpr int DF_WPROJ = DF_WQKV + 32 * 96;    // [32][32]
constexpr int DF_WFC = DF_WPROJ + 32 * 32;     // [32][128]
constexpr int DF_WFC2 = DF_WFC + 32 * 128;     // [128][32]
constexpr int DF_K = DF_WFC2 + 128 * 32;       // [THREADS][32], then V [THREADS][32]
constexpr int df_floats(int threads) { return DF_K + 2 * threads * 32; }   // 80 KB / 112 KB / 176 KB

__device__ __forceinline__ float2 ffma2_(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}

// acc[0..31] += sum_i xs[i] * W[i][col0 .. col0+31],  W in shared memory with row stride `ld` floats
template <int IN>
__device__ __forceinline__ void row_matvec32(const float* xs, const float* W, int ld, int col0, float2* acc) {
#pragma unroll 4
  for (int i = 0; i < IN; ++i) {
    const float2 xx = make_float2(xs[i], xs[i]);
    const float4* row = reinterpret_cast<const float4*>(W + i * ld + col0);
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const float4 w = row[o];
      acc[2 * o] = ffma2_(xx, make_float2(w.x, w.y), acc[2 * o]);
      acc[2 * o + 1] = ffma2_(xx, make_float2(w.z, w.w), acc[2 * o + 1]);
    }
  }
}

__device__ __forceinline__ void ln_row_f(const float* x, const float* w, const float* b, float* y) {
  float mean = 0.f;
#pragma unroll
  for (int c = 0; c < G_E; ++c) mean += x[c];
  mean *= (1.0f / G_E);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < G_E; ++c) var = fmaf(x[c] - mean, x[c] - mean, var);
  const float rs = 1.0f / sqrtf(var * (1.0f / G_E) + 1e-5f);
#pragma unroll
  for (int c = 0; c < G_E; ++c) y[c] = (x[c] - mean) * rs * __ldg(w + c) + __ldg(b + c);
}

__device__ __forceinline__ float gelu_new_f(float x) {
  return 0.5f * x * (1.0f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x)));
}

template <int DF_THREADS>
__global__ void __launch_bounds__(DF_THREADS) gpt2_dense_fp32_kernel(const DenseParams p) {
  constexpr int DF_V = DF_K + DF_THREADS * 32;
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const Gpt2Dev& m = p.m;
  const int S = p.T + 1;
  const bool valid = tid < S;
  const int dx = m.dx, du = m.du, din = m.din;

  float x[G_E];
  {
    float tok[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) tok[i] = 0.f;
    if (valid) {
      if (tid == 0) {
        for (int i = 0; i < dx; ++i) tok[i] = p.query[(size_t)b * dx + i];
      } else {
        const size_t row = (size_t)(b / p.share) * p.Ts + (tid - 1);
        for (int i = 0; i < dx; ++i) tok[i] = p.cs[row * dx + i];
        for (int i = 0; i < du; ++i) tok[dx + i] = p.ca[row * du + i];
        for (int i = 0; i < dx; ++i) tok[dx + du + i] = p.cns[row * dx + i];
        tok[2 * dx + du] = p.cr[row];
      }
    }
#pragma unroll
    for (int c = 0; c < G_E; ++c) x[c] = valid ? __ldg(m.embed_b + c) + __ldg(m.wpe + (size_t)tid * G_E + c) : 0.f;
    for (int i = 0; i < din; ++i) {
      const float tv = tok[i];
#pragma unroll
      for (int c = 0; c < G_E; ++c) x[c] = fmaf(tv, __ldg(m.embed_wT + i * G_E + c), x[c]);
    }
  }

  for (int l = 0; l < m.L; ++l) {
    const LayerW& w = m.layer[l];
    __syncthreads();   // previous layer's readers of the weights / K / V are done
    {
      auto cp = [&](const float* src, int off, int n) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(sm + off);
        for (int i = tid; i < n / 4; i += DF_THREADS) d4[i] = __ldg(s4 + i);
      };
      cp(w.attn_w, DF_WQKV, 32 * 96);
      cp(w.proj_w, DF_WPROJ, 32 * 32);
      cp(w.fc_w, DF_WFC, 32 * 128);
      cp(w.fc2_w, DF_WFC2, 128 * 32);
    }
    __syncthreads();
    float y[G_E];
    float2 acc[16];
    // ---- LN1, then k and v rows -> shared memory, q stays in registers ----
    ln_row_f(x, w.ln1_w, w.ln1_b, y);
#pragma unroll
    for (int part = 1; part <= 2; ++part) {      // 1: k, 2: v
#pragma unroll
      for (int o = 0; o < 16; ++o) acc[o] = make_float2(__ldg(w.attn_b + part * 32 + 2 * o), __ldg(w.attn_b + part * 32 + 2 * o + 1));
      row_matvec32<G_E>(y, sm + DF_WQKV, 96, part * 32, acc);
      float4* dst = reinterpret_cast<float4*>(sm + (part == 1 ? DF_K : DF_V) + tid * G_E);
#pragma unroll
      for (int o = 0; o < 8; ++o) dst[o] = make_float4(acc[2 * o].x, acc[2 * o].y, acc[2 * o + 1].x, acc[2 * o + 1].y);
    }
    float q[G_E];
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = make_float2(__ldg(w.attn_b + 2 * o), __ldg(w.attn_b + 2 * o + 1));
    row_matvec32<G_E>(y, sm + DF_WQKV, 96, 0, acc);
#pragma unroll
    for (int o = 0; o < 16; ++o) q[2 * o] = acc[o].x * 0.17677669529663687f, q[2 * o + 1] = acc[o].y * 0.17677669529663687f;
    __syncthreads();
    // ---- causal attention of row tid: online softmax over keys 0..tid ----
    {
      float mx = -INFINITY, lsum = 0.f;
#pragma unroll
      for (int o = 0; o < 16; ++o) acc[o] = make_float2(0.f, 0.f);
      const int jmax = __shfl_sync(0xffffffffu, tid, 31);   // last row of this warp: warp-uniform trip count
      for (int j = 0; j <= jmax; ++j) {
        const float4* kr = reinterpret_cast<const float4*>(sm + DF_K + j * G_E);
        float2 d0 = make_float2(0.f, 0.f), d1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const float4 kk = kr[o];
          d0 = ffma2_(make_float2(q[4 * o], q[4 * o + 1]), make_float2(kk.x, kk.y), d0);
          d1 = ffma2_(make_float2(q[4 * o + 2], q[4 * o + 3]), make_float2(kk.z, kk.w), d1);
        }
        const float s = (d0.x + d0.y) + (d1.x + d1.y);
        if (j <= tid) {
          if (s > mx) {   // rescale only when the running maximum moves
            const float sc = expf(mx - s);
            lsum *= sc;
#pragma unroll
            for (int o = 0; o < 16; ++o) acc[o].x *= sc, acc[o].y *= sc;
            mx = s;
          }
          const float pr = expf(s - mx);
          lsum += pr;
          const float2 pp = make_float2(pr, pr);
          const float4* vr = reinterpret_cast<const float4*>(sm + DF_V + j * G_E);
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const float4 vv = vr[o];
            acc[2 * o] = ffma2_(pp, make_float2(vv.x, vv.y), acc[2 * o]);
            acc[2 * o + 1] = ffma2_(pp, make_float2(vv.z, vv.w), acc[2 * o + 1]);
          }
        }
      }
      const float inv = 1.0f / lsum;
#pragma unroll
      for (int o = 0; o < 16; ++o) y[2 * o] = acc[o].x * inv, y[2 * o + 1] = acc[o].y * inv;
    }
    // ---- x += o Wproj + b ----
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = make_float2(__ldg(w.proj_b + 2 * o), __ldg(w.proj_b + 2 * o + 1));
    row_matvec32<G_E>(y, sm + DF_WPROJ, 32, 0, acc);
#pragma unroll
    for (int o = 0; o < 16; ++o) x[2 * o] += acc[o].x, x[2 * o + 1] += acc[o].y;
    // ---- MLP: 4 slices of 32 hidden units, each consumed by fc2 as soon as it is activated ----
    ln_row_f(x, w.ln2_w, w.ln2_b, y);
    float2 out2[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) out2[o] = make_float2(__ldg(w.fc2_b + 2 * o), __ldg(w.fc2_b + 2 * o + 1));
#pragma unroll 1
    for (int part = 0; part < 4; ++part) {
#pragma unroll
      for (int o = 0; o < 16; ++o) acc[o] = make_float2(__ldg(w.fc_b + part * 32 + 2 * o), __ldg(w.fc_b + part * 32 + 2 * o + 1));
      row_matvec32<G_E>(y, sm + DF_WFC, 128, part * 32, acc);
      float g[G_E];
#pragma unroll
      for (int o = 0; o < 16; ++o) g[2 * o] = gelu_new_f(acc[o].x), g[2 * o + 1] = gelu_new_f(acc[o].y);
      row_matvec32<G_E>(g, sm + DF_WFC2 + part * 32 * G_E, 32, 0, out2);
    }
#pragma unroll
    for (int o = 0; o < 16; ++o) x[2 * o] += out2[o].x, x[2 * o + 1] += out2[o].y;
  }

  if (p.test ? (tid == p.T) : (tid >= 1 && tid <= p.T)) {
    float y[G_E];
    ln_row_f(x, m.lnf_w, m.lnf_b, y);
    float* o = p.test ? p.out + (size_t)b * du : p.out + ((size_t)b * p.T + (tid - 1)) * du;
    for (int j = 0; j < du; ++j) {
      float lg = __ldg(m.pred_b + j);
#pragma unroll
      for (int c = 0; c < G_E; ++c) lg = fmaf(y[c], __ldg(m.pred_wT + c * du + j), lg);
      o[j] = lg;
    }
  }
}

template <int THREADS>
static cudaError_t launch_df(const DenseParams& p, cudaStream_t st) {
  const int smem = df_floats(THREADS) * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(gpt2_dense_fp32_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  gpt2_dense_fp32_kernel<THREADS><<<p.B, THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

int gpt2_dense_fp32_launch(const DenseParams& p, cudaStream_t st) {
  const int S = p.T + 1;
  cudaError_t e = S <= 128 ? launch_df<128>(p, st) : S <= 256 ? launch_df<256>(p, st) : launch_df<512>(p, st);
  if (e != cudaSuccess) {
    set_error("gpt2_dense_fp32 launch failed: %s", cudaGetErrorString(e));
    return DPT_ERR_CUDA;
  }
  return DPT_OK;
}

}  // namespace dpt
