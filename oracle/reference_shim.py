"""Load the UNMODIFIED reference (/root/reference) behind import shims.

TEST INFRASTRUCTURE ONLY.  This module works only in the build container, where
``/root/reference`` is mounted read-only; it is used by
``tests/golden/generate_golden.py`` to produce the committed golden fixtures and
by a few ``-m "not gpu"`` tests that are skipped when the reference is absent.
Nothing in the product package, ``bench.py`` or the ``-m gpu`` tests imports it.

Shims:
  * ``matplotlib``/``matplotlib.pyplot``/``mpl_toolkits.mplot3d``/``openpyxl``
    are absent in this image -> stub modules (used by the reference for plots /
    xlsx only: scatter_.py:14, main_.py:19,24, poisson_.py:9,14).
  * ``np.product`` was removed in NumPy 2 -> alias to ``np.prod``
    (poisson_.py:31, main_.py:218,256).
No reference source is copied or edited; ``main_``/``poisson_`` (module-level
scripts) are exec'd from their source text with flags flipped by the harness.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("EMC_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "scatter_.py"))


def _install_stubs() -> None:
    import numpy as np

    if not hasattr(np, "product"):
        np.product = np.prod  # type: ignore[attr-defined]

    def _stub(name: str) -> types.ModuleType:
        from unittest import mock

        m = types.ModuleType(name)
        m.__dict__["__stub__"] = True
        # any plotting / spreadsheet call (plt.subplots(), ax.plot(), Workbook() ...) is a no-op:
        # main_.py:405-434 plots after the time loop and must not abort the run
        m.__getattr__ = lambda attr: mock.MagicMock(name="%s.%s" % (name, attr))   # PEP 562
        sys.modules[name] = m
        return m

    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits",
                 "mpl_toolkits.mplot3d", "openpyxl"):
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)
    m3d = sys.modules["mpl_toolkits.mplot3d"]
    if getattr(m3d, "__stub__", False) and not hasattr(m3d, "Axes3D"):
        m3d.Axes3D = object  # poisson_.py:14 `from mpl_toolkits.mplot3d import Axes3D`
    mpl = sys.modules["matplotlib"]
    if getattr(mpl, "__stub__", False):
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
    plt = sys.modules["matplotlib.pyplot"]
    if getattr(plt, "__stub__", False):
        from unittest import mock
        plt.subplots = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())   # main_.py:405 unpacks two


_cache: dict[str, types.ModuleType] = {}


def load_init_scatter():
    """Import the reference's ``init_`` and ``scatter_`` modules (they import cleanly)."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
    if "init_" in _cache:
        return _cache["init_"], _cache["scatter_"]
    _install_stubs()
    saved = {k: sys.modules.get(k) for k in ("init_", "scatter_")}
    sys.path.insert(0, REFERENCE_DIR)
    try:
        for k in ("init_", "scatter_"):
            sys.modules.pop(k, None)
        ref_init = importlib.import_module("init_")
        ref_scatter = importlib.import_module("scatter_")
    finally:
        sys.path.remove(REFERENCE_DIR)
        # keep the reference modules out of the global namespace so that the
        # product's own `init_`/`scatter_` mirrors are never shadowed
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache["init_"], _cache["scatter_"] = ref_init, ref_scatter
    return ref_init, ref_scatter


def exec_main(flags_on: bool = True, extra_globals: dict | None = None,
              run_loop: bool = False) -> dict:
    """exec main_.py's source text and return its namespace.

    With ``run_loop=False`` the three condition flags stay False except
    ``scat_cond`` (so tables exist) and the namespace gives access to the
    reference's own ``calc_energy``/``free_flight``/``carrier_drift``/... functions
    (main_.py:92-213).  With ``run_loop=True`` ``drift_cond`` is flipped too and
    ``sheet = {}`` is injected (main_.py:341 needs it when save_cond is False),
    which runs the whole 50-step loop (minutes).
    """
    import logging

    ref_init, ref_scatter = load_init_scatter()
    src = open(os.path.join(REFERENCE_DIR, "main_.py")).read()
    if flags_on:
        src = src.replace("scat_cond = False", "scat_cond = True")
    if run_loop:
        src = src.replace("drift_cond = False", "drift_cond = True")
        src = src.replace("scat_cond = False", "scat_cond = True")
    ns: dict = {"__name__": "ref_main_", "sheet": {}}
    if extra_globals:
        ns.update(extra_globals)
    saved = {k: sys.modules.get(k) for k in ("init_", "scatter_")}
    sys.modules["init_"], sys.modules["scatter_"] = ref_init, ref_scatter
    logging.disable(logging.CRITICAL)
    try:
        exec(compile(src, os.path.join(REFERENCE_DIR, "main_.py"), "exec"), ns)
    finally:
        logging.disable(logging.NOTSET)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ns


def exec_poisson(extra_globals: dict | None = None) -> dict:
    """exec poisson_.py (a script: init + NGP assignment + SOR solve + field)."""
    ref_init, _ = load_init_scatter()
    src = open(os.path.join(REFERENCE_DIR, "poisson_.py")).read()
    ns: dict = {"__name__": "ref_poisson_"}
    if extra_globals:
        ns.update(extra_globals)
    saved = sys.modules.get("init_")
    sys.modules["init_"] = ref_init
    try:
        exec(compile(src, os.path.join(REFERENCE_DIR, "poisson_.py"), "exec"), ns)
    finally:
        if saved is None:
            sys.modules.pop("init_", None)
        else:
            sys.modules["init_"] = saved
    return ns
