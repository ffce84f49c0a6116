"""CPU tests: the kernel bodies + C-ABI plumbing (host emulation build, tests/emu) against
the golden vectors of the real reference.  No GPU involved; the GPU parity tests proper are
in test_gpu_parity.py."""
import pytest

import cases
import parity
from emu_util import emu_lib


@pytest.fixture(scope="module")
def lib():
    return emu_lib()


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_case_against_golden(lib, name):
    parity.run_case(name, lib_path=lib)
