"""The drop-in boundary: libdftbp_b200.so loads, exports every symbol include/dftbp_b200.h declares,
and refuses to run without a CUDA device (there is no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

from helpers import ROOT
from dftbplus_b200 import _lib, build as build_mod


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.SO_PATH):
        build_mod.build()
    return _lib.load()


def test_header_and_library_agree(lib):
    header = open(os.path.join(ROOT, "include", "dftbp_b200.h")).read()
    declared = set(re.findall(r"\b(dftbp_b200_[a-z0-9_]+)\s*\(", header))
    declared.discard("dftbp_b200_ctx")
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.dftbp_b200_version()


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "dftbp_b200.h"\nint main(void){return DFTBP_B200_OK;}\n')
    import subprocess
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is for CPU-only hosts")
    h = ctypes.c_void_p()
    rc = lib.dftbp_b200_create(ctypes.byref(h), 0, 0, 1)
    assert rc == 2                                   # DFTBP_B200_ENODEVICE
    assert b"no CPU fallback" in lib.dftbp_b200_last_error(h)
    lib.dftbp_b200_destroy(h)
    from dftbplus_b200.api import HotPath, DftbB200Error
    with pytest.raises(DftbB200Error):
        HotPath()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dftbplus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "dftb_oracle" not in text and "liboracle" not in text, f


def _build_c_driver(tmp_path):
    """tests/c_driver/chain.c: the hot path from plain C99 through include/dftbp_b200.h."""
    import subprocess
    exe = str(tmp_path / "chain")
    libdir = os.path.dirname(_lib.SO_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_driver", "chain.c"), "-o", exe,
                           "-L" + libdir, "-ldftbp_b200", "-lm", "-Wl,-rpath," + libdir])
    return exe


def test_c_driver_links_and_refuses_without_gpu(lib, tmp_path):
    import subprocess
    import torch
    exe = _build_c_driver(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the run itself is test_c_driver_chain")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 77, p.stdout + p.stderr      # DFTBP_B200_ENODEVICE -> 77
    assert "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_c_driver_chain(lib, tmp_path):
    """build_h0_s -> unpack -> hegv -> density_matrix -> pack -> mulliken -> derivative_shift from C."""
    import subprocess
    exe = _build_c_driver(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "chain.c: OK" in p.stdout
