"""CPU: the C-ABI library loads and exports every symbol include/conp_b200.h
declares (no compute calls without a GPU), and fails loudly without a device."""
import os
import re

import pytest

from conp_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return abi.load_library()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "conp_b200.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|void|const char|void) \*?(conp_[a-z_A-Z0-9]+)\(", hdr, re.M)))
    assert sorted(abi.SYMBOLS) == declared
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.conp_abi_version() == 2


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(abi.ConpError) as e:
        abi.Context()
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lammps-user-conp2_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "conp_oracle" not in txt and "oracle/" not in txt, os.path.join(d, f)
