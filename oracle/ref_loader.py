"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference (ochsnerd/ip_mcmc) in-process.

The reference lives read-only under /root/reference and exists only in the build container
(never on the GPU box).  This module is used by ``oracle/make_golden.py`` (to generate the
fixtures committed under tests/golden/) and by the optional ``-m "not gpu"`` tests that
re-validate the restatement in ``oracle/*.py`` against the live reference when it is present.
Nothing in the product package (ip_mcmc_b200/) may import this file.

The reference's scripts do not import cleanly on a modern stack, so five shims are applied
*before* import; the reference sources are never edited or copied:

  1. ``matplotlib`` / ``matplotlib.pyplot`` stubs (imported at rusanov.py:3, lorenz.py:2,
     utilities.py:2; matplotlib is not installed here),
  2. ``np.float = float`` (alias removed from NumPy; used at rusanov.py:32,
     utilities.py:92,103),
  3. ``ip_mcmc.pCNProposer = ip_mcmc.ConstSteppCNProposer`` (stale name imported at
     lorenz.py:6-10, lorenz_mcmc.py:6-10, burgers_mcmc.py:4-8),
  4. a stub ``helpers`` module (the real one imports POT and runs a test at import,
     helpers.py:1,123),
  5. ``sys.path`` entries for report/scripts and report/scripts/burgers (the scripts
     hard-code /home/david/... at utilities.py:7,13 and lorenz_mcmc.py:13).
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("IPMCMC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ip_mcmc", "ip_mcmc"))


_loaded = None


def load():
    """Return a namespace with the reference modules: .ip_mcmc, .rusanov, .utilities,
    .lorenz, .lorenz_mcmc, .burgers_mcmc."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import numpy as np

    sys.dont_write_bytecode = True  # the reference tree is read-only
    # (1) matplotlib stubs
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")

        def _noop(*a, **k):
            return None

        plt.__getattr__ = lambda name: _noop
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    # (2) removed NumPy alias
    if not hasattr(np, "float"):
        np.float = float
    # (5) paths
    for p in (os.path.join(REFERENCE_ROOT, "ip_mcmc"),
              os.path.join(REFERENCE_ROOT, "report", "scripts"),
              os.path.join(REFERENCE_ROOT, "report", "scripts", "burgers")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ip_mcmc  # the reference package, unmodified

    # (3) stale name
    if not hasattr(ip_mcmc, "pCNProposer"):
        ip_mcmc.pCNProposer = ip_mcmc.ConstSteppCNProposer
    # (4) helpers stub
    if "helpers" not in sys.modules:
        helpers = types.ModuleType("helpers")
        for name in ("store_figure", "load_or_compute", "autocorrelation", "wasserstein_distance"):
            setattr(helpers, name, lambda *a, **k: None)
        sys.modules["helpers"] = helpers

    with contextlib.redirect_stdout(io.StringIO()):
        import rusanov
        import utilities
        import lorenz
        import lorenz_mcmc
        import burgers_mcmc

    ns = types.SimpleNamespace(ip_mcmc=ip_mcmc, rusanov=rusanov, utilities=utilities,
                               lorenz=lorenz, lorenz_mcmc=lorenz_mcmc, burgers_mcmc=burgers_mcmc)
    _loaded = ns
    return ns


@contextlib.contextmanager
def quiet():
    """The reference sampler prints one line per sample (sampler.py:24)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


class TapeRNG:
    """Noise-tape RNG: same injection seam as the reference's MockRNG (test_utilities.py:11-26)
    but wrapping a real generator and recording every draw, so that a chain run by the
    reference can be replayed step by step by the oracle and by the CUDA engine.

    ``multivariate_normal`` draws are recorded at the w level (after the SVD map,
    SURVEY.md section 9 item 6) and ``random`` draws at the U level.
    """

    def __init__(self, rng):
        self._rng = rng
        self.normals = []
        self.uniforms = []

    def multivariate_normal(self, mean, cov, *a, **k):
        w = self._rng.multivariate_normal(mean=mean, cov=cov, *a, **k)
        self.normals.append(w.copy())
        return w

    def random(self, *a, **k):
        u = self._rng.random(*a, **k)
        self.uniforms.append(u)
        return u

    def normal(self, *a, **k):
        w = self._rng.normal(*a, **k)
        self.normals.append(w)
        return w
