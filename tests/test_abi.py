"""The C-ABI library loads without a GPU and exports every symbol include/uhsdr_b200.h declares.
No compute calls here: engine creation must fail loudly (no CPU fallback) when no device exists."""
import ctypes
import os
import re

import pytest

from uhsdr_b200.config import ChanCfg, default_cfg
from uhsdr_b200.engine import EXPORTS, Engine, UhsdrError, load_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "uhsdr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uhsdr_[a-z0-9_]+)\s*\(", src)))


@pytest.mark.parametrize("exact", [False, True])
def test_library_exports_every_declared_symbol(built, exact):
    lib = load_library(exact=exact)
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert set(EXPORTS) == set(names)
    assert lib.uhsdr_b200_abi_version() == 1
    assert lib.uhsdr_b200_backend().decode().startswith("cuda-sm100a")


def test_default_cfg_matches_python_mirror(built):
    lib = load_library()
    c = ChanCfg()
    assert lib.uhsdr_default_chan_cfg(ctypes.byref(c)) == 0
    assert bytes(c) == bytes(default_cfg())
    assert c.struct_size == ctypes.sizeof(ChanCfg)


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(UhsdrError) as ei:
        Engine(4)
    assert ei.value.code == -2      # UHSDR_ERR_NO_DEVICE


def test_bad_tables_rejected_before_any_device_work(built):
    lib = load_library()
    h = ctypes.c_void_p()
    rc = lib.uhsdr_engine_create(ctypes.byref(h), 0, 0, None, 0)
    assert rc == -1 and not h.value
    assert lib.uhsdr_strerror(-5).decode() == "configuration not implemented"


def test_library_does_not_link_the_oracle(built):
    """The product library must not reference the oracle (or any CPU chain)."""
    import subprocess
    for so in ("libuhsdr_b200.so", "libuhsdr_b200_exact.so"):
        out = subprocess.run(["nm", "-D", os.path.join(ROOT, "uhsdr_b200", "csrc", so)], capture_output=True, text=True).stdout
        assert "port_rx" not in out and "ref_rx" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "uhsdr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "uhsdr_port" not in txt, f
