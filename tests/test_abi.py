"""The C-ABI library loads and exports every symbol include/bd_b200.h declares; without a GPU the
product path fails loudly instead of falling back to anything."""
import os
import re
import subprocess

import pytest

from conftest import ROOT, have_gpu

HEADER = os.path.join(ROOT, "include", "bd_b200.h")
LIB = os.path.join(ROOT, "building_detection_b200", "libbd_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "run __graft_entry__.build() first"
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if l.strip()}
    decl = declared_symbols()
    assert len(decl) >= 30
    missing = [s for s in decl if s not in exported]
    assert not missing, missing


def test_ctypes_binding_covers_header():
    from building_detection_b200 import runtime
    runtime.lib()
    assert sorted(runtime._SIGS) == declared_symbols()


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import numpy as np
    from building_detection_b200 import runtime
    from building_detection_b200.predict_model.hrnet import HRNet
    with pytest.raises(runtime.NativeError):
        runtime.context(0)
    with pytest.raises(runtime.NativeError):
        HRNet().predict(np.zeros((1, 512, 512, 3), np.float32))
