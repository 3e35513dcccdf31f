// Bandit task draw and the single-step GPUBanditEnv transition.
//   dpt_bandit_sample_means : envs/bandit_env.py:10-18 + :29-34, envs/gpu_bandit_env.py:18-28
//   dpt_bandit_opt_action   : argmax / one-hot of caller-provided means (LinearBanditEnv :162-165)
//   dpt_gpu_bandit_step     : envs/gpu_bandit_env.py:53-63 (argmax -> gather -> + randn * var | bernoulli)
// One thread per env; rows are d floats (<= 128 B), so a warp touches a contiguous 32*d*4-byte run.
#include "common.cuh"
#include "philox.cuh"

namespace dpt {

__device__ __forceinline__ void write_opt(const float* m, int d, int env, int32_t* opt_idx, float* opt_a) {
  int best = 0;
  float bv = m[0];
  for (int j = 1; j < d; ++j)
    if (m[j] > bv) bv = m[j], best = j;  // first maximum wins, like np.argmax / torch.argmax
  if (opt_idx) opt_idx[env] = best;
  if (opt_a)
    for (int j = 0; j < d; ++j) opt_a[(size_t)env * d + j] = (j == best) ? 1.f : 0.f;
}

__global__ void sample_means_kernel(Key key, uint64_t env_id0, int N, int d, float* means, int32_t* opt_idx,
                                    float* opt_a) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= N) return;
  float m[32];
  for (int b = 0; 4 * b < d; ++b) {
    const uint4 w = philox_words(key, env_id0 + (uint64_t)env, (uint32_t)b, STREAM_TASK);
    for (int i = 0; i < 4; ++i)
      if (4 * b + i < d) m[4 * b + i] = u24(word_of(w, i));
  }
  for (int j = 0; j < d; ++j) means[(size_t)env * d + j] = m[j];
  write_opt(m, d, env, opt_idx, opt_a);
}

__global__ void opt_action_kernel(const float* means, int N, int d, int32_t* opt_idx, float* opt_a) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= N) return;
  float m[32];
  for (int j = 0; j < d; ++j) m[j] = means[(size_t)env * d + j];
  write_opt(m, d, env, opt_idx, opt_a);
}

__global__ void gpu_bandit_step_kernel(const float* __restrict__ means, const float* __restrict__ actions, float var,
                                       int type, Key key, uint64_t env_id0, uint32_t step, int N, int d,
                                       float* __restrict__ reward, const float* __restrict__ inject,
                                       float* __restrict__ dump) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= N) return;
  const float* u = actions + (size_t)env * d;
  int a = 0;
  float bv = u[0];
  for (int j = 1; j < d; ++j) {
    const float v = u[j];
    if (v > bv) bv = v, a = j;
  }
  const float mu = means[(size_t)env * d + a];
  float noise;
  if (inject) {
    noise = inject[env];
  } else {  // word (step & 1) * 2 .. of pair block step / 2
    const uint4 w = philox_words(key, env_id0 + (uint64_t)env, step >> 1, STREAM_ENV_REWARD);
    if (type == 0) {
      float z0, z1;
      box_muller(w.z, w.w, z0, z1);
      noise = (step & 1) ? z1 : z0;
    } else {
      noise = u24((step & 1) ? w.y : w.x);
    }
  }
  if (dump) dump[env] = noise;
  reward[env] = (type == 0) ? fmaf(var, noise, mu) : (noise < mu ? 1.f : 0.f);
}

}  // namespace dpt

using namespace dpt;

extern "C" int dpt_bandit_sample_means(uint64_t seed, uint64_t env_id0, int N, int d, float* means,
                                       int32_t* opt_a_index, float* opt_a, void* stream) {
  DPT_CHECK_ARG(N >= 0 && d >= 1 && d <= 32, "dpt_bandit_sample_means: N=%d d=%d (d must be in [1,32])", N, d);
  if (N == 0) return DPT_OK;
  DPT_CHECK_ARG(means, "dpt_bandit_sample_means: null means");
  sample_means_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(Key{(uint32_t)seed, (uint32_t)(seed >> 32)},
                                                                         env_id0, N, d, means, opt_a_index, opt_a);
  DPT_LAUNCH_CHECK();
  return DPT_OK;
}

extern "C" int dpt_bandit_opt_action(const float* means, int N, int d, int32_t* opt_a_index, float* opt_a,
                                     void* stream) {
  DPT_CHECK_ARG(N >= 0 && d >= 1 && d <= 32, "dpt_bandit_opt_action: N=%d d=%d (d must be in [1,32])", N, d);
  if (N == 0) return DPT_OK;
  DPT_CHECK_ARG(means, "dpt_bandit_opt_action: null means");
  opt_action_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(means, N, d, opt_a_index, opt_a);
  DPT_LAUNCH_CHECK();
  return DPT_OK;
}

extern "C" int dpt_gpu_bandit_step(const float* means, const float* actions, float var, int type, uint64_t seed,
                                   uint64_t env_id0, int64_t step, int N, int d, float* reward,
                                   const float* inject_noise, float* dump_noise, void* stream) {
  DPT_CHECK_ARG(N >= 0 && d >= 1, "dpt_gpu_bandit_step: N=%d d=%d", N, d);
  DPT_CHECK_ARG(type == 0 || type == 1, "dpt_gpu_bandit_step: unknown type %d (0 uniform, 1 bernoulli)", type);
  DPT_CHECK_ARG(step >= 0, "dpt_gpu_bandit_step: negative step");
  if (N == 0) return DPT_OK;
  DPT_CHECK_ARG(means && actions && reward, "dpt_gpu_bandit_step: null pointer");
  gpu_bandit_step_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      means, actions, var, type, Key{(uint32_t)seed, (uint32_t)(seed >> 32)}, env_id0, (uint32_t)step, N, d, reward,
      inject_noise, dump_noise);
  DPT_LAUNCH_CHECK();
  return DPT_OK;
}
