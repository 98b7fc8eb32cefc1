"""C-ABI hygiene that needs no GPU: the library builds, loads, exports every symbol include/*.h declares,
refuses to compute without a device, and the product never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "flowavenet_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fwn_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from tf_flowavenet_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from tf_flowavenet_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes table and header disagree"
    assert _lib.lib().fwn_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        return
    from tf_flowavenet_b200 import _lib
    cfg, h = _lib.FwnConfig(), ctypes.c_void_p()
    assert _lib.lib().fwn_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"no CPU fallback" in _lib.lib().fwn_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tf_flowavenet_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert "oracle" not in src.replace("flowavenet_oracle", "oracle") or fn == "NONE", "%s mentions the oracle" % fn
    code = "import sys; import tf_flowavenet_b200; assert not any(m.startswith('oracle') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], cwd=ROOT, check=True)
