"""Philox4x32-10 (Salmon et al., SC'11) in numpy -- TEST INFRASTRUCTURE.

Bit-exact integer restatement of the counter-based generator the CUDA path
uses (csrc/philox.cuh).  The reference itself draws from numpy's global
MT19937 stream (SURVEY.md §5 "Seeds / RNG"); the new build replaces that with
Philox keyed by (seed) and counted by (index, env_lo, env_hi, stream) so results
are independent of launch geometry and of the number of GPUs.

Only the INTEGER layer is restated here (raw words, bounded integers, 24-bit
uniforms): those must match the device bit-for-bit.  Float transforms
(-log u, Box-Muller) are not restated: parity tests read the noise the device
actually used (the ``dump`` outputs of the C ABI) and feed it to the oracle.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

# stream tags (must match csrc/philox.cuh)
STREAM_TASK = 0          # task draw: bandit means / linear-bandit theta
STREAM_ROLLIN_SETUP = 1  # rollin_bandit per-env behaviour policy (cov, dirichlet, rand_index)
STREAM_ROLLIN_STEP = 2   # rollin_bandit per-step (categorical uniform, reward normal)
STREAM_DARKROOM_STEP = 3 # rollin_mdp 'uniform' (state, action)
STREAM_DARKROOM_QUERY = 4
STREAM_ENV_REWARD = 5    # env.step reward noise (GPUBanditEnv.step, online loop)
STREAM_CTRL = 6          # controller draws (Thompson normals, LinUCB first arm, transformer sampling)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32-valued array, key: (2,) ints -> (..., 4) uint32."""
    c = np.asarray(ctr).astype(np.uint64)
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def words(seed, env_ids, index, stream):
    """Raw words for counter (index, env_lo, env_hi, stream); broadcasting over env_ids/index."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    index = np.asarray(index, dtype=np.uint64)
    env_ids, index = np.broadcast_arrays(env_ids, index)
    ctr = np.stack([index & MASK, env_ids & MASK, env_ids >> np.uint64(32),
                    np.full(env_ids.shape, stream, dtype=np.uint64)], axis=-1)
    seed = int(seed)
    return philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))


def mulhi(w, n):
    """Bounded integer in [0, n): (w * n) >> 32 (bias < n / 2^32)."""
    return ((np.asarray(w).astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def u24(w):
    """Uniform in [0,1) with 24 bits: exactly representable in fp32 and fp64."""
    return (np.asarray(w) >> np.uint32(8)).astype(np.float64) * 2.0 ** -24


def bandit_means(seed, env_ids, dim):
    """means ~ U[0,1)^dim, fp32-exact: block b of STREAM_TASK gives arms 4b..4b+3."""
    nb = (dim + 3) // 4
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    w = words(seed, env_ids[:, None], np.arange(nb)[None, :], STREAM_TASK)  # [N, nb, 4]
    return u24(w.reshape(len(env_ids), nb * 4)[:, :dim])


def rollin_setup_ints(seed, env_ids, dim):
    """cov_idx in [0,11) and rand_idx in [0,dim) of rollin_bandit (block 0 of STREAM_ROLLIN_SETUP)."""
    w = words(seed, env_ids, 0, STREAM_ROLLIN_SETUP)
    return mulhi(w[..., 0], 11), mulhi(w[..., 1], dim)


def rollin_step_k(seed, env_ids, H):
    """31-bit integers k of the per-step categorical uniforms u = k * 2^-31: pair p = h // 2 of
    STREAM_ROLLIN_STEP, word 0 for even h, word 1 for odd h.  Returns [N,H] int64."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    npair = (H + 1) // 2
    w = words(seed, env_ids[:, None], np.arange(npair)[None, :], STREAM_ROLLIN_STEP)   # [N, npair, 4]
    k = (w[..., :2] >> np.uint32(1)).reshape(len(env_ids), 2 * npair)[:, :H]
    return k.astype(np.int64)


def split_word(w, dim):
    """One word -> (x, y, a): successive digits of w / 2^32 in the mixed radix (dim, dim, 5)."""
    w = np.asarray(w).astype(np.uint64)
    t = w * np.uint64(dim)
    x = t >> np.uint64(32)
    t = (t & MASK) * np.uint64(dim)
    y = t >> np.uint64(32)
    a = ((t & MASK) * np.uint64(5)) >> np.uint64(32)
    return x.astype(np.int64), y.astype(np.int64), a.astype(np.int64)


def darkroom_draws(seed, env_ids, H, dim):
    """(states [N,H,2], actions [N,H]) of the 'uniform' rollin: word h % 4 of block h // 4."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    nb = (H + 3) // 4
    w = words(seed, env_ids[:, None], np.arange(nb)[None, :], STREAM_DARKROOM_STEP).reshape(len(env_ids), nb * 4)[:, :H]
    x, y, a = split_word(w, dim)
    return np.stack([x, y], -1), a


def darkroom_query(seed, env_ids, n_samples, dim):
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    w = words(seed, env_ids[:, None], np.arange(n_samples)[None, :], STREAM_DARKROOM_QUERY)[..., 0]
    x, y, _ = split_word(w, dim)
    return np.stack([x, y], -1)
