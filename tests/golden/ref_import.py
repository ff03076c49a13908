"""Import the UNMODIFIED reference (QInfer, /root/reference) in this container.

Test tooling only: used by ``make_golden.py`` to generate the committed golden
vectors and by the optional ``-m refcheck`` tests. Nothing in the product, the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this module; the GPU box
has no /root/reference.

The reference is pure Python but targets Python 2/3 + NumPy 1.x, so importing
it under Python 3.12 / NumPy 2.3 / SciPy 1.18 needs the shims SURVEY.md §8c
lists.  The reference tree is read-only and lacks the generated
``qinfer/version.py`` (setup.py:18-33), so a scratch copy is made under /tmp.
No reference source enters this repository.
"""
import os
import shutil
import sys
import types

REFERENCE_SRC = "/root/reference/src/qinfer"
SCRATCH = "/tmp/qinfer_ref_scratch"


def reference_available():
    return os.path.isdir(REFERENCE_SRC)


def _write(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(text)


def import_reference():
    """Return the imported reference ``qinfer`` module (scratch copy + shims)."""
    if "qinfer" in sys.modules and getattr(sys.modules["qinfer"], "_b200_ref_shim", False):
        return sys.modules["qinfer"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_SRC)

    import numpy as np
    import scipy.integrate
    import scipy.ndimage

    dst = os.path.join(SCRATCH, "qinfer")
    if not os.path.isdir(dst):
        shutil.copytree(REFERENCE_SRC, dst)
        _write(os.path.join(dst, "version.py"), 'version = "1.0"\n')
    # `future` / `past` are not installed; the reference uses three names.
    _write(os.path.join(SCRATCH, "future", "__init__.py"), "")
    _write(os.path.join(SCRATCH, "future", "utils.py"),
           "def with_metaclass(meta, *bases):\n"
           "    class metaclass(meta):\n"
           "        def __new__(cls, name, this_bases, d):\n"
           "            return meta(name, bases, d)\n"
           "    return type.__new__(metaclass, 'temporary_class', (), {})\n"
           "def iteritems(d):\n"
           "    return iter(d.items())\n")
    _write(os.path.join(SCRATCH, "past", "__init__.py"), "")
    _write(os.path.join(SCRATCH, "past", "builtins.py"), "basestring = str\n")

    # NumPy 2 / SciPy 1.18 renames used by the reference.
    if not hasattr(scipy.integrate, "cumtrapz"):
        scipy.integrate.cumtrapz = scipy.integrate.cumulative_trapezoid
    if "scipy.ndimage.filters" not in sys.modules:
        m = types.ModuleType("scipy.ndimage.filters")
        m.gaussian_filter1d = scipy.ndimage.gaussian_filter1d
        sys.modules["scipy.ndimage.filters"] = m
        scipy.ndimage.filters = m
    for name, val in (("float", float), ("int", int), ("complex", complex), ("bool", bool)):
        if not hasattr(np, name):
            setattr(np, name, val)
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid

    if SCRATCH not in sys.path:
        sys.path.insert(0, SCRATCH)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import qinfer
    qinfer._b200_ref_shim = True
    return qinfer
