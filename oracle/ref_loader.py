"""oracle/ref_loader.py — TEST INFRASTRUCTURE.

Imports the reference's OWN code without copying it:

* ``ref_core()``  -> the reference's compiled native module (``oracle/_ref/dp_core*.so``, built by
  ``make -C oracle ref`` from ``/root/reference/svecalign/vecalign/dp_core.pyx``).  The ``.so`` travels
  to the GPU box, so this works there too.
* ``ref_dp_utils()`` -> ``svecalign.vecalign.dp_utils`` imported straight from ``/root/reference``
  (build container only; returns None elsewhere).  The compiled core is pre-registered in
  ``sys.modules`` so pyximport never tries to write into the read-only reference tree.
"""
import glob
import importlib.machinery
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("SVX_REFERENCE_ROOT", "/root/reference")
_MOD = "svecalign.vecalign.dp_core"


def ref_core_path():
    hits = sorted(glob.glob(os.path.join(_HERE, "_ref", "dp_core*.so")))
    return hits[0] if hits else None


def ref_core():
    """The reference's compiled dp_core module, or None if oracle/_ref has not been built."""
    if _MOD in sys.modules:
        return sys.modules[_MOD]
    so = ref_core_path()
    if so is None:
        return None
    loader = importlib.machinery.ExtensionFileLoader(_MOD, so)
    spec = importlib.util.spec_from_file_location(_MOD, so, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    if have_reference():
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import svecalign.vecalign  # noqa: F401  (parent package must exist before registering)
        sys.modules[_MOD] = mod
    return mod


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "svecalign", "vecalign"))


def ref_dp_utils():
    """The reference's Python driver module, or None when /root/reference (or _ref) is absent."""
    if not have_reference() or ref_core() is None:
        return None
    import svecalign.vecalign.dp_utils as du
    return du


def ref_package(name):
    """Import another reference module by dotted name (build container only)."""
    if ref_dp_utils() is None:
        return None
    return importlib.import_module(name)
