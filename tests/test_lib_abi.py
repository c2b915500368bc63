"""The C-ABI shared library loads without a GPU and exports every symbol include/hp3d.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "hp3d.h")).read()
    declared = sorted(set(re.findall(r"\b(hp3d_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    from hierarchicalprobabilistic3dhuman_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared
    assert _lib.lib().hp3d_version() == 200


def test_argument_errors_do_not_need_a_gpu(built_lib):
    from hierarchicalprobabilistic3dhuman_b200 import _lib
    L = _lib.lib()
    rc = L.hp3d_rot6d_to_rotmat(None, 0, None, None)
    assert rc < 0 and b"bad argument" in L.hp3d_last_error()
    rc = L.hp3d_mf_sample(None, None, None, 1, 23, 8, 1.5, 0, 0, None, None, 8, None, None, None)
    assert rc < 0
