"""Seed management for the Philox streams.

The reference seeds numpy's global generator once (collect_data.py:352) and every draw advances
that hidden state.  Here each randomised call gets a 64-bit Philox key derived from a base seed and
a call counter, and addresses its draws by (global env id, step, stream) -- so a call's result
depends only on (key, env ids), not on launch geometry or the number of GPUs.
"""
_state = {"seed": 0, "calls": 0}
_GOLDEN = 0x9E3779B97F4A7C15
_MASK = (1 << 64) - 1


def seed(s):
    """Counterpart of np.random.seed(s): restart the key sequence."""
    _state["seed"] = int(s) & _MASK
    _state["calls"] = 0


def _mix(x):  # splitmix64 finaliser
    x = (x + _GOLDEN) & _MASK
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _MASK
    return x ^ (x >> 31)


def key_for(base_seed, call):
    return _mix((int(base_seed) & _MASK) ^ _mix(int(call)))


def next_key():
    """Key for the next randomised call (advances the call counter)."""
    k = key_for(_state["seed"], _state["calls"])
    _state["calls"] += 1
    return k
