"""Import the UNMODIFIED reference (dev container only) behind import shims.

The reference needs four non-numerical modules that are absent here (gym,
IPython, skimage, matplotlib -- SURVEY.md §8c).  None carries hot-path
arithmetic, so they are replaced by empty stand-ins.  The reference's
top-level module names (``envs``, ``ctrls``, ``evals``, ``models``,
``collect_data``, ``utils`` ...) are moved out of ``sys.modules`` after import so
that they never collide with the drop-in modules of this repo.

``/root/reference`` does not exist on the GPU box.  ``stage()`` (called by ``__graft_entry__.build()``
in the dev container) copies the reference's ``*.py`` files, unmodified, to the git-ignored
``baseline/_ref/`` -- the place the bench contract reserves for the reference install -- which DOES
travel to the box; ``bench.py``'s CPU legs (``cpu_baseline`` and ``--impl reference``) then time the
reference's OWN functions from there.  The `-m gpu` tests and smoke() never call this module.
"""
import importlib
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(_REPO, "baseline", "_ref")
LIVE = "/root/reference"


def _pick_root():
    env = os.environ.get("DPT_REF")
    if env:
        return env
    return LIVE if os.path.isdir(os.path.join(LIVE, "envs")) else STAGED


REF_ROOT = _pick_root()
_TOP = ("envs", "ctrls", "evals", "models", "collect_data", "utils", "common_args", "dataset")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "envs"))


def stage(src=LIVE, dst=STAGED):
    """Copy the reference's Python sources, byte for byte, to ``baseline/_ref`` (git-ignored, not
    gpurun-ignored).  Returns the number of files copied (0 when ``src`` is absent)."""
    import shutil
    if not os.path.isdir(os.path.join(src, "envs")):
        return 0
    n = 0
    for base, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if not d.startswith(".")]
        for f in files:
            if f.endswith((".py", ".txt", ".md")):
                rel = os.path.relpath(os.path.join(base, f), src)
                out = os.path.join(dst, rel)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                shutil.copyfile(os.path.join(base, f), out)
                n += 1
    return n


def _shims():
    mods = {}
    gym = types.ModuleType("gym")

    class Env:  # gym.Env is only used as a base class (envs/base_env.py:8)
        pass

    class Box:  # only constructed, never read (envs/bandit_env.py:36-37)
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.low, self.high, self.shape = low, high, shape

    class Discrete:  # only ``.n`` is read (envs/darkroom_env.py:28,39)
        def __init__(self, n):
            self.n = n

    spaces = types.ModuleType("gym.spaces")
    spaces.Box, spaces.Discrete = Box, Discrete
    gym.Env, gym.spaces = Env, spaces
    mods["gym"], mods["gym.spaces"] = gym, spaces
    ipy = types.ModuleType("IPython")
    ipy.embed = lambda *a, **k: None
    mods["IPython"] = ipy
    sk, skt = types.ModuleType("skimage"), types.ModuleType("skimage.transform")
    skt.resize = None
    sk.transform = skt
    mods["skimage"], mods["skimage.transform"] = sk, skt
    class _Null:   # matplotlib is plotting only (evals/*.py): every attribute / call is a no-op
        def __getattr__(self, name):
            return self

        def __call__(self, *a, **k):
            return self

        def __iter__(self):
            return iter((self, self))
    mpl, plt = types.ModuleType("matplotlib"), _Null()
    mpl.pyplot = plt
    mods["matplotlib"], mods["matplotlib.pyplot"] = mpl, plt
    return mods


_cache = None


def load():
    """Returns a namespace with the reference modules as attributes:
    ``ref.collect_data``, ``ref.bandit_env``, ``ref.darkroom_env``,
    ``ref.ctrl_bandit``, ``ref.eval_bandit``, ``ref.eval_linear_bandit``, ``ref.net``."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    saved = {k: v for k, v in sys.modules.items()
             if k.split(".")[0] in _TOP or k.split(".")[0] in ("gym", "IPython", "skimage", "matplotlib")}
    for k in saved:
        del sys.modules[k]
    sys.modules.update(_shims())
    sys.path.insert(0, REF_ROOT)
    try:
        ns = types.SimpleNamespace()
        ns.collect_data = importlib.import_module("collect_data")
        ns.bandit_env = importlib.import_module("envs.bandit_env")
        ns.gpu_bandit_env = importlib.import_module("envs.gpu_bandit_env")
        ns.darkroom_env = importlib.import_module("envs.darkroom_env")
        ns.ctrl_bandit = importlib.import_module("ctrls.ctrl_bandit")
        ns.eval_bandit = importlib.import_module("evals.eval_bandit")
        ns.eval_linear_bandit = importlib.import_module("evals.eval_linear_bandit")
        ns.ctrl_darkroom = importlib.import_module("ctrls.ctrl_darkroom")
        ns.eval_darkroom = importlib.import_module("evals.eval_darkroom")
        ns.net = importlib.import_module("models.net")
        ns.dataset = importlib.import_module("dataset")
    finally:
        sys.path.remove(REF_ROOT)
        for k in list(sys.modules):
            top = k.split(".")[0]
            if top in _TOP or top in ("gym", "IPython", "skimage", "matplotlib"):
                del sys.modules[k]
        sys.modules.update(saved)
    _cache = ns
    return ns
