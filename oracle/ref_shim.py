"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference module
``/root/reference/HiCHap/matrixBuilding.py`` (Python 2.7 source) into this
Python 3 interpreter so its own functions can be executed as the parity oracle.

Nothing in ``hichap_master_b200/`` may import this file.  It only works in the
build container, where ``/root/reference`` exists; the GPU box never sees it.
It is used by ``oracle/make_golden.py`` (golden-vector generation) and by the
``not gpu`` tests that pin ``oracle/hichap_oracle.py`` against the reference.

What the shim does (nothing is copied into the repo; the source is read where
it lies and exec'd in memory):

* ``print x`` statements -> ``print(x)`` (all 37 are single-line statements,
  e.g. matrixBuilding.py:172, :571, :622).
* stub modules for the un-vendored third-party ``cooler`` package imported at
  matrixBuilding.py:10-14 (``cooler``, ``cooler.util.binnify``,
  ``cooler.reduce.CoolerMerger``, ``cooler.api.Cooler``,
  ``cooler.create_cooler``).  The numeric functions never call them.
* ``np.int`` / ``np.float`` aliases (removed in NumPy >= 1.24; used at
  matrixBuilding.py:246, :461, :512, :559, :565, :808, :814, :1115, :1123).

Streams handed to the reference must be *text* (``io.StringIO``): the code does
``line[1].lstrip('chr')`` (matrixBuilding.py:575) which fails on bytes.
"""
from __future__ import annotations

import os
import re
import sys
import types

REFERENCE_ROOT = os.environ.get("HICHAP_REFERENCE_ROOT", "/root/reference")
_REF_FILE = os.path.join(REFERENCE_ROOT, "HiCHap", "matrixBuilding.py")

_cached = None


def available() -> bool:
    return os.path.isfile(_REF_FILE)


def _install_cooler_stubs():
    if "cooler" in sys.modules and not getattr(sys.modules["cooler"], "_hc_stub", False):
        return  # a real cooler is importable: use it
    def _unavailable(*a, **k):
        raise RuntimeError("cooler is not installed; the shim only stubs its import")
    cooler = types.ModuleType("cooler")
    cooler._hc_stub = True
    cooler.create_cooler = _unavailable
    cooler.merge_coolers = _unavailable
    util = types.ModuleType("cooler.util")
    util.binnify = _unavailable
    reduce_ = types.ModuleType("cooler.reduce")
    reduce_.CoolerMerger = _unavailable
    api = types.ModuleType("cooler.api")
    api.Cooler = _unavailable
    cooler.util, cooler.reduce, cooler.api = util, reduce_, api
    sys.modules.update({"cooler": cooler, "cooler.util": util,
                        "cooler.reduce": reduce_, "cooler.api": api})


def load():
    """Return the reference ``matrixBuilding`` module object (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise FileNotFoundError(_REF_FILE)
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int          # noqa: NPY001 - what the 2019 reference expects
    if not hasattr(np, "float"):
        np.float = float      # noqa: NPY001
    try:
        import cooler  # noqa: F401
    except Exception:
        _install_cooler_stubs()
    with open(_REF_FILE, "r") as fh:
        src = fh.read()
    src = re.sub(r"^(\s*)print (.+)$", r"\1print(\2)", src, flags=re.M)
    mod = types.ModuleType("hichap_reference_matrixBuilding")
    mod.__file__ = _REF_FILE
    exec(compile(src, _REF_FILE, "exec"), mod.__dict__)
    _cached = mod
    return mod


_REF_SF = os.path.join(REFERENCE_ROOT, "HiCHap", "StructureFind.py")
_cached_sf = None


def load_structure():
    """The reference ``StructureFind`` module (HiCHap/StructureFind.py), unmodified except for the Python-2 surface:
    ``print x`` statements, ``xrange``, and stub modules for imports that are absent here (matplotlib, statsmodels,
    ghmm, cooler) -- none of them is touched by the methods the goldens use (Distance_Decay, Get_PCA's O/E loop,
    Get_DI, bias_handle)."""
    global _cached_sf
    if _cached_sf is not None:
        return _cached_sf
    if not os.path.isfile(_REF_SF):
        raise FileNotFoundError(_REF_SF)
    import numpy as np
    for name, typ in (("int", int), ("float", float), ("bool", bool)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    try:
        import cooler  # noqa: F401
    except Exception:
        _install_cooler_stubs()

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m._hc_stub = True
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    na = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stubbed import"))
    mpl = stub("matplotlib", use=lambda *a, **k: None)
    stub("matplotlib.backends")
    stub("matplotlib.backends.backend_pdf", PdfPages=na)
    stub("matplotlib.colors", LinearSegmentedColormap=na)
    mpl.pyplot = stub("matplotlib.pyplot")
    stub("statsmodels"); stub("statsmodels.sandbox"); stub("statsmodels.sandbox.stats")
    stub("statsmodels.sandbox.stats.multicomp", multipletests=na)
    stub("ghmm")
    with open(_REF_SF, "r") as fh:
        src = fh.read()
    src = re.sub(r"^(\s*)print (.+)$", r"\1print(\2)", src, flags=re.M)
    mod = types.ModuleType("hichap_reference_StructureFind")
    mod.__file__ = _REF_SF
    mod.__dict__["xrange"] = range
    exec(compile(src, _REF_SF, "exec"), mod.__dict__)
    _cached_sf = mod
    return mod
