// Shared definitions of the fused online-loop kernels (online_loop.cu: one warp does everything, the general
// kernel; online_loop_ws.cu: warp-specialised controller / helper warps, the fast kernel for d <= 10).
#pragma once
#include <type_traits>

#include "common.cuh"
#include "philox.cuh"

namespace dpt {

constexpr int OL_WARPS = 2;    // max warps per CTA (small CTAs: 11 x 2 warps fit one SM's shared memory at d <= 5)
constexpr int OL_THREADS = OL_WARPS * 32;
constexpr int OL_T = 32;       // steps buffered per flush
constexpr int OL_MAX_LD = 8;   // max lin_d

enum { K_OPT = 0, K_EMP = 1, K_UCB = 2, K_THOMPSON = 3, K_LINUCB = 4, K_LINUCB2 = 5 /* internal: lin_d == 2, state in registers */ };

struct OnlineParams {
  double p0, p1, p2, var;
  const float* means;
  const double* arms;
  int lin_d;
  Key key;
  uint64_t env_id0;
  int N, H, d;
  int rtype;         // DPT_REWARD_*
  uint32_t magic_d;  // ceil(2^32 / d): floor(x / d) == umulhi(x, magic_d) for the small x used here
  float *ctx_s, *ctx_a, *ctx_ns, *ctx_r, *cum_means;
  double* regret;    // [regret_reps][H][4] accumulators (replicated to spread same-address atomics)
  int regret_reps;   // power of two
  dpt_online_inject_t in;
  dpt_online_dump_t out;
  bool vec;  // float4 flush allowed
  double* creg_carry;  // fused fast kernel: [ceil(H/128)][N] cumulative regret before every 128th step (NULL: not wanted)
};

struct WarpTile {
  // env-major tiles written by lane = env and read back by lane = step; element (e, t) of the float tiles
  // lives in column (t + e) & 31, which keeps both access directions bank-conflict free without padding
  unsigned char acts[32][OL_T];
  float rew[32][OL_T];
  float creg[32][OL_T];       // cumulative regret of each env after each buffered step
};

template <int DMAX>
struct ArmState {
  double sum[DMAX];   // reward sum b
  double aux0[DMAX];  // EMP/UCB: mean; THOMPSON: posterior mean
  double aux1[DMAX];  // UCB: bonus; THOMPSON: posterior std
  int cnt[DMAX];
};

// 4 standard normals per Philox block (two Box-Muller pairs)
__device__ __forceinline__ void normals4(uint4 w, float z[4]) {
  box_muller(w.x, w.y, z[0], z[1]);
  box_muller(w.z, w.w, z[2], z[3]);
}

template <int LD>
__device__ __forceinline__ void inv_small(const double* S, double* Si, int ld) {
  // Gauss-Jordan with partial pivoting on [S | I] (S is SPD = I + A^T A, so it never fails)
  double a[OL_MAX_LD][2 * OL_MAX_LD];
  for (int i = 0; i < ld; ++i)
    for (int j = 0; j < ld; ++j) a[i][j] = S[i * ld + j], a[i][ld + j] = (i == j) ? 1.0 : 0.0;
  for (int c = 0; c < ld; ++c) {
    int piv = c;
    for (int r = c + 1; r < ld; ++r)
      if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    if (piv != c)
      for (int j = 0; j < 2 * ld; ++j) {
        const double t = a[c][j];
        a[c][j] = a[piv][j];
        a[piv][j] = t;
      }
    const double inv = 1.0 / a[c][c];
    for (int j = 0; j < 2 * ld; ++j) a[c][j] *= inv;
    for (int r = 0; r < ld; ++r)
      if (r != c) {
        const double f = a[r][c];
        for (int j = 0; j < 2 * ld; ++j) a[r][j] -= f * a[c][j];
      }
  }
  for (int i = 0; i < ld; ++i)
    for (int j = 0; j < ld; ++j) Si[i * ld + j] = a[i][ld + j];
}


// reward noise of the 4 steps 4k..4k+3 of one env (both online-loop kernels): ONE Philox block per step quad, index = k on
// STREAM_ENV_REWARD; gaussian: Box-Muller of (x, y) -> steps 4k, 4k+1 and of (z, w) -> 4k+2, 4k+3; bernoulli: u24 of each word
__device__ __forceinline__ void reward_noise4(Key key, uint64_t gid, uint32_t k, int rtype, float z[4]) {
  const uint4 w = philox_words(key, gid, k, STREAM_ENV_REWARD);
  if (rtype == DPT_REWARD_GAUSSIAN)
    normals4(w, z);
  else
    z[0] = u24(w.x), z[1] = u24(w.y), z[2] = u24(w.z), z[3] = u24(w.w);
}

// online_loop_ws.cu: returns cudaErrorNotSupported when the shape is outside what the warp-specialised kernel
// handles (the caller then launches the general kernel); tab = [2][H + 1] doubles of count-indexed terms
// regret_out: the caller's [H,4] sums (+=) or NULL; p.regret then is zeroed scratch of at least regret_reps * H * 3 doubles
cudaError_t launch_online_ws(int kind, const OnlineParams& p, double* tab, double* regret_out, cudaStream_t st, bool fused);
bool online_ws_supported(int kind, const OnlineParams& p);
void keep_pool_memory();

}  // namespace dpt
