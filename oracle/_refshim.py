"""Loader for the *live* reference package (test infrastructure only).

The reference (`/root/reference`, read-only, Python) imports `pyvista` and
`matplotlib` at module top (utilities/visualization_utils.py:21-25); neither is
installed.  This shim pre-seeds `sys.modules` with MagicMock stand-ins so the
numeric part of the reference imports unchanged.  It is used ONLY by
`tests/golden/make_golden.py` (fixture generation, in the build container), by
the not-gpu tests that pin the oracle against the live reference when
`/root/reference` happens to exist, and by `oracle/ref_driver.py` (the CPU-baseline
legs of bench.py).  Nothing in the product imports this file.

Where the reference is looked for: $VET_REFERENCE_SRC, then /root/reference/src (the
build container), then baseline/_ref (the offline install of `__graft_entry__.build()`:
git-ignored, but it travels to the GPU box with the snapshot).
"""
import os
import sys
from unittest import mock

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.environ.get("VET_REFERENCE_SRC"), "/root/reference/src", os.path.join(_ROOT, "baseline", "_ref")]
REFERENCE_SRC = next((c for c in _CANDIDATES if c and os.path.isdir(os.path.join(c, "viewport_entropy_toolkit"))),
                     "/root/reference/src")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "viewport_entropy_toolkit"))


def load_reference():
    """Returns the imported `viewport_entropy_toolkit` reference package."""
    if not reference_available():
        raise ImportError(f"reference sources not found under {REFERENCE_SRC}")
    if "viewport_entropy_toolkit" in sys.modules:
        return sys.modules["viewport_entropy_toolkit"]
    for name in ("pyvista", "matplotlib"):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    mpl = sys.modules["matplotlib"]
    for sub in ("pyplot", "animation", "figure", "axes"):
        full = f"matplotlib.{sub}"
        if full not in sys.modules:
            m = mock.MagicMock(name=full)
            sys.modules[full] = m
            setattr(mpl, sub, m)
    # PlotManager.__init__ unpacks `fig, ax = plt.subplots(...)`
    sys.modules["matplotlib.pyplot"].subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import viewport_entropy_toolkit  # noqa: E402

    return viewport_entropy_toolkit
