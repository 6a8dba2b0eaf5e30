"""CPU-side checks of the C ABI: the library builds, loads, and exports every symbol the header declares."""
import ctypes
import os


def test_library_exports_every_declared_symbol():
    from eavit_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _lib.declared_symbols()
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.eavit_version() >= 100


def test_product_never_imports_oracle():
    """The oracle is test infrastructure; the product package must not reference it."""
    from eavit_b200 import _lib
    pkg = os.path.dirname(_lib.LIB_PATH)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
