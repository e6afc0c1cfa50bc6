"""The XLA FFI shim (mccube_b200/csrc/xla_ffi_shim.cc) cannot run here -- no JAX, no XLA headers -- but it can be
COMPILED: against include/mccube_b200.h (every C-ABI call type-checked argument by argument) and against a
stand-in of xla/ffi/api/ffi.h (tests/ffi_stub) whose XLA_FFI_DEFINE_HANDLER_SYMBOL static_asserts that each
binding's parameter list is the one its handler takes."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SHIM = ROOT / "mccube_b200" / "csrc" / "xla_ffi_shim.cc"
CUDA_INC = Path("/usr/local/cuda/include")


def _compile(src: Path):
    cxx = shutil.which("g++")
    if cxx is None or not CUDA_INC.exists():
        pytest.skip("g++ or the CUDA headers are not available")
    return subprocess.run([cxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", f"-I{ROOT / 'tests' / 'ffi_stub'}",
                           f"-I{ROOT / 'include'}", f"-I{CUDA_INC}", str(src)], capture_output=True, text=True)


def test_shim_compiles_against_the_c_abi():
    r = _compile(SHIM)
    assert r.returncode == 0, r.stderr


def test_binding_check_catches_a_wrong_signature(tmp_path):
    """The stand-in is not vacuous: swapping one argument type in a binding must fail to compile."""
    bad = tmp_path / "shim_bad.cc"
    text = SHIM.read_text()
    needle = "MCCB_STREAM.Arg<F32>().Arg<U32>().Ret<F32>(), {ffi::Traits::kCmdBufferCompatible});"
    assert needle in text
    bad.write_text(text.replace(needle, needle.replace(".Arg<U32>()", ".Arg<F32>()")))
    r = _compile(bad)
    assert r.returncode != 0 and "binding and handler differ" in r.stderr


def test_every_handler_wraps_a_declared_entry_point():
    """Each handler calls one function of include/mccube_b200.h, and the ops the reference's step needs are all
    there (the list VERDICT r1 asked for)."""
    header = (ROOT / "include" / "mccube_b200.h").read_text()
    declared = set(re.findall(r"\b(mccb200_\w+)\s*\(", header))
    shim = SHIM.read_text()
    called = set(re.findall(r"Fail\((mccb200_\w+)\(", shim))
    assert called and called <= declared, called - declared
    handlers = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((mccb200_\w+)_ffi,", shim))
    need = {"mccb200_mc_indices", "mccb200_expand_select", "mccb200_box_child_bounds", "mccb200_box_prk_tables",
            "mccb200_box_prk_selection", "mccb200_box_prk_count", "mccb200_box_prk_rank_gather",
            "mccb200_gaussian_full_drift_tc", "mccb200_gaussian_full_drift_tc_prepare", "mccb200_moments_sum",
            "mccb200_moments_m2", "mccb200_stratified_keys", "mccb200_sort_pairs", "mccb200_gather_rows"}
    assert need <= handlers, need - handlers
    # the reference-side binding registers exactly these targets
    py = (ROOT / "integration" / "mccube" / "_b200.py").read_text()
    registered = set(re.findall(r'"(mccb200_\w+)_ffi"', py))
    assert handlers <= registered, handlers - registered
