"""Headless loader for the upstream reference (TEST INFRASTRUCTURE ONLY).

Loads ``/root/reference/JacketAnalysisGUI_v2.py`` without Tk / matplotlib /
pip so that its analysis classes (GUI.py:115-803) can be executed to
(1) validate ``oracle/jacket_oracle.py`` and (2) generate the golden vectors
under ``tests/golden/`` (see ``tests/golden/make_golden.py``).

The reference only exists in the build container; on the GPU box this module
reports ``available() == False`` and nothing may depend on it at run time.
Nothing under the product package imports this file.
"""
from __future__ import annotations

import importlib.util
import io
import contextlib
import os
import subprocess
import sys
import types

REFERENCE_FILE = os.environ.get("JK_REFERENCE_FILE", "/root/reference/JacketAnalysisGUI_v2.py")

_cached = None


def available() -> bool:
    return os.path.isfile(REFERENCE_FILE)


def load(allow_raschii=False):
    """Return the reference module (cached). Raises FileNotFoundError if absent.  The golden vectors are pinned on the
    Airy fallback, so by default the module must have come up WITHOUT raschii (GUI.py:96-100); the raschii parity hook
    (oracle/raschii_hook.py) passes allow_raschii=True on machines that have the package."""
    global _cached
    if _cached is not None:
        assert allow_raschii or _cached.RASCHII_AVAILABLE is False, "golden vectors are pinned on the Airy fallback path"
        return _cached
    if not available():
        raise FileNotFoundError(REFERENCE_FILE)

    # (1) stub the GUI-only imports (GUI.py:84-90)
    stubs = ["matplotlib", "matplotlib.pyplot", "matplotlib.lines", "mpl_toolkits",
             "mpl_toolkits.mplot3d", "tkinter", "tkinter.ttk", "tkinter.messagebox",
             "tkinter.filedialog", "tkinter.simpledialog"]
    saved = {}
    for name in stubs:
        saved[name] = sys.modules.get(name)
        mod = types.ModuleType(name)
        mod.__path__ = []  # behave like a package for "from tkinter import ttk"
        sys.modules[name] = mod
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    for sub in ("ttk", "messagebox", "filedialog", "simpledialog"):
        setattr(sys.modules["tkinter"], sub, sys.modules["tkinter." + sub])
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    # (2) neutralise the import-time pip installer (GUI.py:23-37, run at :77)
    real_check_call = subprocess.check_call

    def _no_pip(cmd, *a, **k):
        raise subprocess.CalledProcessError(1, cmd)

    subprocess.check_call = _no_pip
    try:
        spec = importlib.util.spec_from_file_location("_jk_reference", REFERENCE_FILE)
        mod = importlib.util.module_from_spec(spec)
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
    finally:
        subprocess.check_call = real_check_call
        for name in stubs:
            if saved[name] is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = saved[name]
    assert allow_raschii or mod.RASCHII_AVAILABLE is False, "golden vectors are pinned on the Airy fallback path"
    _cached = mod
    return mod
