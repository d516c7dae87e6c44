"""Import the UNMODIFIED reference from /root/reference -- build-container only.

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so nothing that runs
there may import this module; it is used by oracle/make_golden.py (which freezes vectors
into tests/golden/) and by CPU tests that skip when the reference tree is absent.

Shims: vmvo/utils/trajectory.py:6 imports matplotlib (absent here),
and optimize_trajectory_v2.py:119-120 opens a GUI window per loop iteration.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VMVO_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "vmvo", "bicycle_model.py"))


def _install_shims():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    import cv2
    cv2.imshow = lambda *a, **k: None
    cv2.waitKey = lambda *a, **k: -1


def load():
    """Returns a namespace with the reference modules on the hot path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _install_shims()
    import vmvo.bicycle_model as bicycle_model
    import vmvo.constants as constants
    import vmvo.schema as schema
    import vmvo.utils.mpc as mpc
    ns = types.SimpleNamespace(bicycle_model=bicycle_model, constants=constants,
                               schema=schema, mpc=mpc)
    return ns


def load_v2():
    """The sliding-window driver module (needs the cv2/matplotlib shims)."""
    load()
    with contextlib.redirect_stdout(io.StringIO()):
        import vmvo.scripts.optimize_trajectory_v2 as v2
    # tqdm bar -> plain range, to keep test logs clean
    v2.tqdm = lambda it, *a, **k: it
    return v2


@contextlib.contextmanager
def quiet():
    """The reference prints two lines per window (optimize_trajectory_v2.py:65-66)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
