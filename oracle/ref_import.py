"""Import the UNMODIFIED reference (baryshnikova-lab/safepy) from /root/reference.  TEST INFRASTRUCTURE ONLY.

matplotlib and statsmodels are not installed in this image and cannot be (no network); the reference imports them
at module level (safepy/safe.py:16-30, safe_io.py:9, safe_colormaps.py:1-4) although the neighborhood/enrichment
path never calls them.  Throw-away stub modules are placed in sys.modules only when the real ones are missing.
Used by oracle/make_golden.py (this container only: /root/reference does not exist on the GPU box).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SAFEPY_REFERENCE", "/root/reference")


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def import_reference():
    """Returns the reference's `safepy.safe` module (class SAFE, run_permutations, ...)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "safepy")):
        raise ImportError("reference checkout not found at %s" % REFERENCE_ROOT)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        class _Anything:
            def __init__(self, *a, **k):
                pass

            def __call__(self, *a, **k):
                return _Anything()

            def __getattr__(self, name):
                return _Anything()

        def anything(name):
            if name.startswith("__"):       # inspect / importlib probe modules for __file__, __path__, ...
                raise AttributeError(name)
            return _Anything()

        mpl = _stub("matplotlib", use=lambda *a, **k: None, rcParams={})
        mpl.pyplot = _stub("matplotlib.pyplot", __getattr__=anything)
        mpl.colors = _stub("matplotlib.colors", Normalize=_Anything, LinearSegmentedColormap=_Anything,
                           __getattr__=anything)
        mpl.cm = _stub("matplotlib.cm", __getattr__=anything)
        mpl.patches = _stub("matplotlib.patches", __getattr__=anything)
        mpl.collections = _stub("matplotlib.collections", __getattr__=anything)
    try:
        import statsmodels.stats.multitest  # noqa: F401
    except ImportError:
        def fdrcorrection(*a, **k):
            raise NotImplementedError("statsmodels is not installed (multiple_testing=True unavailable)")

        sm = _stub("statsmodels")
        sm.stats = _stub("statsmodels.stats")
        sm.stats.multitest = _stub("statsmodels.stats.multitest", fdrcorrection=fdrcorrection)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from safepy import safe  # noqa: E402
    return safe
