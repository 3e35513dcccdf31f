// Fused darkroom rollout: rollin_mdp (reference collect_data.py:83-111) + DarkroomEnv.transit
// (envs/darkroom_env.py:37-55), opt_action (:69-82), the permuted variant (:96-111) and the
// per-sample query state / optimal action of generate_mdp_histories_from_envs (:200-201).
//
// 'uniform' rollin resamples (state, action) every step, so all N*H steps are independent; 'expert'
// rollin from the reset state (0,0) has the closed form x_h = min(h, gx), y_h = min(max(h-gx,0), gy).
// Either way a step is a pure function of (env, h): the CTA's envs are one flat run of steps, a lane
// owns 4 consecutive steps (one Philox call: one word per step, split into (x, y, a) by a multiply
// chain -- exact integers, restated in oracle/philox.py).  Grid transitions are integer register
// arithmetic.  Output is 40 B per env-step in four fp32 streams; every stream is re-tiled across the
// warp with shuffles of bit-packed integers so each store instruction writes one contiguous,
// 16 B-aligned 512 B run (states/next_states 8 B per step, actions 20 B, rewards 4 B).
#include "common.cuh"
#include "philox.cuh"

namespace dpt {

constexpr int DK_THREADS = 256;
constexpr int DK_WARPS = DK_THREADS / 32;
constexpr int DK_MAX_ENVS = 64;

struct DarkroomParams {
  const int32_t* goals;
  const int32_t* perm_index;
  int dim, mode;
  Key key;
  uint64_t env_id0;
  int N, H, S, envs_per_cta;
  uint32_t magic_H;  // ceil(2^32 / H) when floor(t / H) == umulhi(t, magic_H) for all CTA-local t, else 0
  float *ctx_s, *ctx_a, *ctx_ns, *ctx_r, *query, *opt;
  dpt_darkroom_inject_t in;
  dpt_darkroom_dump_t out;
  bool has_in, has_out;
};

// perm_index -> the perm_index-th permutation of (0..4) in itertools.permutations order
// (envs/darkroom_env.py:96-98), packed 3 bits per entry; inverse packed the same way.
__device__ __forceinline__ void make_perm(int idx, uint32_t& perm, uint32_t& inv) {
  int avail[5] = {0, 1, 2, 3, 4};
  int fact = 24;
  perm = 0, inv = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int k = idx / fact;
    idx -= k * fact;
    if (i < 4) fact /= (4 - i);
    int v = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      if (j == k) v = avail[j];
      if (j >= k && j < 4) avail[j] = avail[j + 1];
    }
    perm |= (uint32_t)v << (3 * i);
    inv |= (uint32_t)i << (3 * v);
  }
}

struct EnvInfo {  // gx | gy<<8 in goal, perm / inverse perm 3 bits per entry
  uint32_t goal, perm, inv;
};

__device__ __forceinline__ EnvInfo load_env(const DarkroomParams& p, int env) {
  EnvInfo e;
  e.goal = (uint32_t)p.goals[2 * (size_t)env] | ((uint32_t)p.goals[2 * (size_t)env + 1] << 16);
  if (p.perm_index)
    make_perm(p.perm_index[env], e.perm, e.inv);
  else
    e.perm = e.inv = 0 | (1 << 3) | (2 << 6) | (3 << 9) | (4 << 12);
  return e;
}

// envs/darkroom_env.py:69-82: fix x first, then y, else stay (base action, before permutation)
__device__ __forceinline__ int base_opt_action(int sx, int sy, int gx, int gy) {
  return sx < gx ? 0 : sx > gx ? 1 : sy < gy ? 2 : sy > gy ? 3 : 4;
}

// One word -> (x, y, a): successive digits of w / 2^32 in the mixed radix (dim, dim, 5).
__device__ __forceinline__ void split_word(uint32_t w, uint32_t dim, int& sx, int& sy, int& a) {
  uint64_t t = (uint64_t)w * dim;
  sx = (int)(t >> 32);
  t = (uint64_t)(uint32_t)t * dim;
  sy = (int)(t >> 32);
  a = (int)(((uint64_t)(uint32_t)t * 5u) >> 32);
}

struct Step {
  int sx, sy, a, nx, ny, r;
};

// (state, action label) -> transit; a is the env-facing action label, perm maps it to the base move
__device__ __forceinline__ void transit(Step& s, const EnvInfo& e, int dim) {
  const int pa = (e.perm >> (3 * s.a)) & 7;  // :100-103
  const int gx = e.goal & 0xffff, gy = e.goal >> 16;
  int nx = s.sx + (pa == 0) - (pa == 1);       // :41-48
  int ny = s.sy + (pa == 2) - (pa == 3);
  s.nx = min(max(nx, 0), dim - 1);             // :49
  s.ny = min(max(ny, 0), dim - 1);
  s.r = (s.nx == gx) & (s.ny == gy);           // :51-54
}

__device__ __forceinline__ Step make_step(const DarkroomParams& p, const EnvInfo& e, int env, int h, uint32_t word) {
  Step s;
  if (p.mode == 0) {
    if (p.has_in) {
      const size_t row = (size_t)env * p.H + h;
      s.sx = p.in.states[2 * row], s.sy = p.in.states[2 * row + 1], s.a = p.in.actions[row];
    } else {
      split_word(word, (uint32_t)p.dim, s.sx, s.sy, s.a);
    }
  } else {  // expert from reset state (0,0): collect_data.py:89,95,104
    const int gx = e.goal & 0xffff, gy = e.goal >> 16;
    s.sx = min(h, gx);
    s.sy = min(max(h - gx, 0), gy);
    s.a = (e.inv >> (3 * base_opt_action(s.sx, s.sy, gx, gy))) & 7;  // :105-111
  }
  transit(s, e, p.dim);
  return s;
}

__device__ __forceinline__ void query_phase(const DarkroomParams& p, int env0, int ne, int tid, int nthreads) {
  for (int i = tid; i < ne * p.S; i += nthreads) {
    const int env = env0 + i / p.S, smp = i % p.S;
    const EnvInfo e = load_env(p, env);
    int qx, qy, unused;
    const size_t qrow = (size_t)env * p.S + smp;
    if (p.has_in) {
      qx = p.in.query[2 * qrow], qy = p.in.query[2 * qrow + 1];
    } else {
      const uint4 w = philox_words(p.key, p.env_id0 + (uint64_t)env, (uint32_t)smp, STREAM_DARKROOM_QUERY);
      split_word(w.x, (uint32_t)p.dim, qx, qy, unused);
    }
    if (p.has_out && p.out.query) p.out.query[2 * qrow] = qx, p.out.query[2 * qrow + 1] = qy;
    const int oa = (e.inv >> (3 * base_opt_action(qx, qy, e.goal & 0xffff, e.goal >> 16))) & 7;
    p.query[2 * qrow] = (float)qx;
    p.query[2 * qrow + 1] = (float)qy;
    for (int j = 0; j < 5; ++j) p.opt[5 * qrow + j] = (j == oa) ? 1.f : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// Fast path: H % 4 == 0, dim <= 256, 16 B-aligned outputs.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DK_THREADS) darkroom_rollin_fast(const DarkroomParams p) {
  __shared__ EnvInfo s_env[DK_MAX_ENVS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int env0 = blockIdx.x * p.envs_per_cta;
  const int ne = min(p.envs_per_cta, p.N - env0);
  const int H = p.H;
  if (tid < ne) s_env[tid] = load_env(p, env0 + tid);
  query_phase(p, env0, ne, tid, DK_THREADS);
  __syncthreads();

  const int total = ne * H;                     // steps of this CTA, a flat contiguous run
  const int chunks = (total + 127) >> 7;        // 128 steps per warp-chunk
  const size_t base = (size_t)env0 * H;
  for (int c = warp; c < chunks; c += DK_WARPS) {
    const int t0 = c * 128 + 4 * lane;          // first of this lane's 4 steps (CTA-local)
    uint32_t spk0 = 0, spk1 = 0, npk0 = 0, npk1 = 0, amask = 0, rmask = 0;
    if (t0 < total) {
      const int el = p.magic_H ? (int)__umulhi((uint32_t)t0, p.magic_H) : t0 / H;
      const int h0 = t0 - el * H;  // H % 4 == 0: all 4 steps are in env el
      const int env = env0 + el;
      const EnvInfo e = s_env[el];
      uint4 w = make_uint4(0, 0, 0, 0);
      if (p.mode == 0 && !p.has_in)
        w = philox_words(p.key, p.env_id0 + (uint64_t)env, (uint32_t)(h0 >> 2), STREAM_DARKROOM_STEP);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const Step s = make_step(p, e, env, h0 + i, word_of(w, i));
        const uint32_t sp = (uint32_t)s.sx | ((uint32_t)s.sy << 8);
        const uint32_t np = (uint32_t)s.nx | ((uint32_t)s.ny << 8);
        if (i < 2) {
          spk0 |= sp << (16 * i), npk0 |= np << (16 * i);
        } else {
          spk1 |= sp << (16 * (i - 2)), npk1 |= np << (16 * (i - 2));
        }
        amask |= 1u << (5 * i + s.a);
        rmask |= (uint32_t)s.r << i;
        if (p.has_out && p.mode == 0) {
          const size_t row = (size_t)env * H + h0 + i;
          if (p.out.states) p.out.states[2 * row] = s.sx, p.out.states[2 * row + 1] = s.sy;
          if (p.out.actions) p.out.actions[row] = s.a;
        }
      }
    }
    const size_t cbase = base + (size_t)c * 128;            // first step of the chunk (global flat)
    const int nsteps = min(128, total - c * 128);           // multiple of 4
    // rewards: float4 #lane = this lane's own 4 steps
    st_stream_if(4 * lane < nsteps, reinterpret_cast<float4*>(p.ctx_r + cbase) + lane,
                make_float4((float)(rmask & 1), (float)((rmask >> 1) & 1), (float)((rmask >> 2) & 1),
                            (float)((rmask >> 3) & 1)));
    // states / next_states: 2 floats per step -> float4 #f holds steps 2f, 2f+1 = lane f/2, half f%2
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int f = it * 32 + lane;
      const int src = f >> 1;
      const uint32_t s0 = __shfl_sync(0xffffffffu, spk0, src), s1 = __shfl_sync(0xffffffffu, spk1, src);
      const uint32_t n0 = __shfl_sync(0xffffffffu, npk0, src), n1 = __shfl_sync(0xffffffffu, npk1, src);
      const uint32_t sv = (f & 1) ? s1 : s0, nv = (f & 1) ? n1 : n0;
      st_stream_if(2 * f < nsteps, reinterpret_cast<float4*>(p.ctx_s + 2 * cbase) + f,
                   make_float4((float)(sv & 255), (float)((sv >> 8) & 255), (float)((sv >> 16) & 255),
                               (float)(sv >> 24)));
      st_stream_if(2 * f < nsteps, reinterpret_cast<float4*>(p.ctx_ns + 2 * cbase) + f,
                   make_float4((float)(nv & 255), (float)((nv >> 8) & 255), (float)((nv >> 16) & 255),
                               (float)(nv >> 24)));
    }
    // actions: 5 floats per step, 20 per lane -> float4 #f = lane f/5, bits 4*(f%5)..+3
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int f = it * 32 + lane;
      const int src = f / 5;
      const uint32_t bits = __shfl_sync(0xffffffffu, amask, src) >> (4 * (f - 5 * src));
      st_stream_if(4 * f < 5 * nsteps, reinterpret_cast<float4*>(p.ctx_a + 5 * cbase) + f,
                   make_float4((bits & 1u) ? 1.f : 0.f, (bits & 2u) ? 1.f : 0.f, (bits & 4u) ? 1.f : 0.f,
                               (bits & 8u) ? 1.f : 0.f));
    }
  }
}

// Generic path: one thread per step, any H / dim / alignment; same Philox counters.
__global__ void __launch_bounds__(DK_THREADS) darkroom_rollin_generic(const DarkroomParams p) {
  const int env0 = blockIdx.x * p.envs_per_cta;
  const int ne = min(p.envs_per_cta, p.N - env0);
  query_phase(p, env0, ne, threadIdx.x, DK_THREADS);
  const int total = ne * p.H;
  for (int t = threadIdx.x; t < total; t += DK_THREADS) {
    const int el = t / p.H, h = t - el * p.H;
    const int env = env0 + el;
    const EnvInfo e = load_env(p, env);
    uint32_t word = 0;
    if (p.mode == 0 && !p.has_in)
      word = word_of(philox_words(p.key, p.env_id0 + (uint64_t)env, (uint32_t)(h >> 2), STREAM_DARKROOM_STEP), h & 3);
    const Step s = make_step(p, e, env, h, word);
    const size_t row = (size_t)env * p.H + h;
    if (p.has_out && p.mode == 0) {
      if (p.out.states) p.out.states[2 * row] = s.sx, p.out.states[2 * row + 1] = s.sy;
      if (p.out.actions) p.out.actions[row] = s.a;
    }
    p.ctx_s[2 * row] = (float)s.sx, p.ctx_s[2 * row + 1] = (float)s.sy;
    p.ctx_ns[2 * row] = (float)s.nx, p.ctx_ns[2 * row + 1] = (float)s.ny;
    p.ctx_r[row] = (float)s.r;
    for (int j = 0; j < 5; ++j) p.ctx_a[5 * row + j] = (j == s.a) ? 1.f : 0.f;
  }
}

__global__ void darkroom_step_kernel(const int32_t* states, const float* actions, const int32_t* goals,
                                     const int32_t* perm_index, int dim, int N, int32_t* next_states,
                                     int32_t* rewards) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= N) return;
  DarkroomParams p{};
  p.goals = goals, p.perm_index = perm_index;
  const EnvInfo e = load_env(p, env);
  Step s;
  s.sx = states[2 * (size_t)env], s.sy = states[2 * (size_t)env + 1];
  const float* u = actions + 5 * (size_t)env;
  int a = 0;
  float bv = u[0];
  for (int j = 1; j < 5; ++j)
    if (u[j] > bv) bv = u[j], a = j;  // np.argmax: first maximum
  s.a = a;
  transit(s, e, dim);
  next_states[2 * (size_t)env] = s.nx, next_states[2 * (size_t)env + 1] = s.ny;
  rewards[env] = s.r;
}

__global__ void darkroom_opt_kernel(const int32_t* states, const int32_t* goals, const int32_t* perm_index, int N,
                                    float* actions) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= N) return;
  DarkroomParams p{};
  p.goals = goals, p.perm_index = perm_index;
  const EnvInfo e = load_env(p, env);
  const int b = base_opt_action(states[2 * (size_t)env], states[2 * (size_t)env + 1], e.goal & 0xffff, e.goal >> 16);
  const int oa = (e.inv >> (3 * b)) & 7;
  for (int j = 0; j < 5; ++j) actions[5 * (size_t)env + j] = (j == oa) ? 1.f : 0.f;
}

// One episode per env from a logits table: thread = env.
__global__ void darkroom_policy_rollout_kernel(const float* __restrict__ logits, const int32_t* goals,
                                               const int32_t* perm_index, int dim, int horizon, int sample, Key key,
                                               uint64_t env_id0, uint32_t episode, int N, float* st, float* ac, float* ns,
                                               float* rw, float* returns, const double* inject_u, double* dump_u) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= N) return;
  DarkroomParams p{};
  p.goals = goals, p.perm_index = perm_index;
  const EnvInfo e = load_env(p, env);
  Step s;
  s.sx = 0, s.sy = 0;                                     // reset (envs/darkroom_env.py:32-35)
  float ret = 0.f;
  for (int t = 0; t < horizon; ++t) {
    const float* lg = logits + ((size_t)env * dim * dim + (size_t)s.sx * dim + s.sy) * 5;
    float l[5];
    float lm = -INFINITY;
    for (int j = 0; j < 5; ++j) l[j] = lg[j], lm = fmaxf(lm, l[j]);
    int a = 0;
    if (sample) {   // scipy softmax (float64, temp = 1) + np.random.choice(p): ctrls/ctrl_darkroom.py:49-54
      double pe[5], tot = 0.0;
      for (int j = 0; j < 5; ++j) pe[j] = exp((double)l[j] - (double)lm), tot = __dadd_rn(tot, pe[j]);
      double acc = 0.0, cdf[5];
      for (int j = 0; j < 5; ++j) {
        const double pj = __ddiv_rn(pe[j], tot);
        acc = (j == 0) ? pj : __dadd_rn(acc, pj);
        cdf[j] = acc;
      }
      double u;
      if (inject_u) {
        u = inject_u[(size_t)t * N + env];
      } else {
        const uint4 w = philox_words(key, env_id0 + (uint64_t)env, episode * 65536u + (uint32_t)t, STREAM_CTRL);
        u = ((double)(w.x >> 5) * 67108864.0 + (double)(w.y >> 6)) * (1.0 / 9007199254740992.0);
      }
      if (dump_u) dump_u[(size_t)t * N + env] = u;
      for (int j = 0; j < 4; ++j) a += (__ddiv_rn(cdf[j], cdf[4]) <= u);
    } else {
      for (int j = 1; j < 5; ++j)
        if (l[j] > l[a]) a = j;                            // np.argmax: first maximum
    }
    s.a = a;
    transit(s, e, dim);
    const size_t row = (size_t)env * horizon + t;
    st[2 * row] = (float)s.sx, st[2 * row + 1] = (float)s.sy;
    ns[2 * row] = (float)s.nx, ns[2 * row + 1] = (float)s.ny;
    rw[row] = (float)s.r;
    for (int j = 0; j < 5; ++j) ac[5 * row + j] = (j == a) ? 1.f : 0.f;
    ret += (float)s.r;
    s.sx = s.nx, s.sy = s.ny;
  }
  if (returns) returns[env] = ret;
}

}  // namespace dpt

using namespace dpt;

extern "C" int dpt_darkroom_policy_rollout(const float* logits, const int32_t* goals, const int32_t* perm_index, int dim,
                                           int horizon, int sample, uint64_t seed, uint64_t env_id0, int64_t episode,
                                           int N, float* states, float* actions, float* next_states, float* rewards,
                                           float* returns, const double* inject_u, double* dump_u, void* stream) {
  DPT_CHECK_ARG(N >= 0 && horizon >= 0 && dim >= 1 && horizon < 65536 && episode >= 0,
                "dpt_darkroom_policy_rollout: N=%d horizon=%d dim=%d", N, horizon, dim);
  if (N == 0 || horizon == 0) return DPT_OK;
  DPT_CHECK_ARG(logits && goals && states && actions && next_states && rewards, "dpt_darkroom_policy_rollout: null pointer");
  darkroom_policy_rollout_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      logits, goals, perm_index, dim, horizon, sample, Key{(uint32_t)seed, (uint32_t)(seed >> 32)}, env_id0,
      (uint32_t)episode, N, states, actions, next_states, rewards, returns, inject_u, dump_u);
  DPT_LAUNCH_CHECK();
  return DPT_OK;
}

extern "C" int dpt_darkroom_rollin(const int32_t* goals, const int32_t* perm_index, int dim, int mode, uint64_t seed,
                                   uint64_t env_id0, int N, int H, int n_samples, float* ctx_states,
                                   float* ctx_actions, float* ctx_next_states, float* ctx_rewards,
                                   float* query_states, float* optimal_actions, const dpt_darkroom_inject_t* inject,
                                   const dpt_darkroom_dump_t* dump, void* stream) {
  DPT_CHECK_ARG(N >= 0 && H >= 0 && n_samples >= 0, "dpt_darkroom_rollin: negative size");
  DPT_CHECK_ARG(dim >= 1 && dim <= 65535, "dpt_darkroom_rollin: dim=%d outside [1,65535]", dim);
  DPT_CHECK_ARG(mode == 0 || mode == 1, "dpt_darkroom_rollin: unknown rollin mode %d (0 uniform, 1 expert)", mode);
  if (N == 0) return DPT_OK;
  DPT_CHECK_ARG(goals, "dpt_darkroom_rollin: null goals");
  DPT_CHECK_ARG(H == 0 || (ctx_states && ctx_actions && ctx_next_states && ctx_rewards),
                "dpt_darkroom_rollin: null context pointer");
  DPT_CHECK_ARG(n_samples == 0 || (query_states && optimal_actions), "dpt_darkroom_rollin: null query pointer");
  DarkroomParams p{};
  p.goals = goals, p.perm_index = perm_index, p.dim = dim, p.mode = mode;
  p.key = Key{(uint32_t)seed, (uint32_t)(seed >> 32)};
  p.env_id0 = env_id0;
  p.N = N, p.H = H, p.S = n_samples;
  p.ctx_s = ctx_states, p.ctx_a = ctx_actions, p.ctx_ns = ctx_next_states, p.ctx_r = ctx_rewards;
  p.query = query_states, p.opt = optimal_actions;
  if (inject) {
    DPT_CHECK_ARG(!dump, "dpt_darkroom_rollin: inject and dump are mutually exclusive");
    DPT_CHECK_ARG(mode == 1 || H == 0 || (inject->states && inject->actions),
                  "dpt_darkroom_rollin: inject needs states and actions");
    DPT_CHECK_ARG(n_samples == 0 || inject->query, "dpt_darkroom_rollin: inject needs query");
    p.in = *inject, p.has_in = true;
  }
  if (dump) p.out = *dump, p.has_out = true;
  // >= 3 waves with a nearly full last wave, at least ~2k steps per CTA
  const bool fast = H > 0 && (H % 4 == 0) && dim <= 256 && aligned16(ctx_states) && aligned16(ctx_actions) &&
                    aligned16(ctx_next_states) && aligned16(ctx_rewards);
  static int per_sm_fast = 0, per_sm_gen = 0;
  if (per_sm_fast == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_fast, darkroom_rollin_fast, DK_THREADS, 0) != cudaSuccess || per_sm_fast < 1) per_sm_fast = 4;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_gen, darkroom_rollin_generic, DK_THREADS, 0) != cudaSuccess || per_sm_gen < 1) per_sm_gen = 4;
  }
  int min_e = H > 0 ? (2048 + H - 1) / H : 1;
  if (min_e > DK_MAX_ENVS) min_e = DK_MAX_ENVS;
  const int e = pick_envs_per_cta(N, sm_count() * (fast ? per_sm_fast : per_sm_gen), min_e, DK_MAX_ENVS);
  p.envs_per_cta = e;
  p.magic_H = (H > 1 && (uint64_t)e * H * H < 0xffffffffull) ? (uint32_t)((0x100000000ull + (uint64_t)H - 1) / (uint64_t)H) : 0u;
  const int grid = (N + e - 1) / e;
  if (fast)
    darkroom_rollin_fast<<<grid, DK_THREADS, 0, (cudaStream_t)stream>>>(p);
  else
    darkroom_rollin_generic<<<grid, DK_THREADS, 0, (cudaStream_t)stream>>>(p);
  DPT_LAUNCH_CHECK();
  return DPT_OK;
}

extern "C" int dpt_darkroom_step(const int32_t* states, const float* actions, const int32_t* goals,
                                 const int32_t* perm_index, int dim, int N, int32_t* next_states, int32_t* rewards,
                                 void* stream) {
  DPT_CHECK_ARG(N >= 0 && dim >= 1, "dpt_darkroom_step: N=%d dim=%d", N, dim);
  if (N == 0) return DPT_OK;
  DPT_CHECK_ARG(states && actions && goals && next_states && rewards, "dpt_darkroom_step: null pointer");
  darkroom_step_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(states, actions, goals, perm_index, dim, N,
                                                                          next_states, rewards);
  DPT_LAUNCH_CHECK();
  return DPT_OK;
}

extern "C" int dpt_darkroom_opt_action(const int32_t* states, const int32_t* goals, const int32_t* perm_index, int N,
                                       float* actions, void* stream) {
  DPT_CHECK_ARG(N >= 0, "dpt_darkroom_opt_action: N=%d", N);
  if (N == 0) return DPT_OK;
  DPT_CHECK_ARG(states && goals && actions, "dpt_darkroom_opt_action: null pointer");
  darkroom_opt_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(states, goals, perm_index, N, actions);
  DPT_LAUNCH_CHECK();
  return DPT_OK;
}
