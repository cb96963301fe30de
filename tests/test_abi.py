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
    assert lib.uhsdr_b200_abi_version() == 2
    assert lib.uhsdr_b200_backend().decode().startswith("cuda-sm100a")


def test_channel_range_in_c_matches_the_python_partition(built):
    from uhsdr_b200.partition import channel_range
    lib = load_library()
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            for rank in range(world):
                f, c = ctypes.c_int(), ctypes.c_int()
                assert lib.uhsdr_channel_range(rank, world, total, ctypes.byref(f), ctypes.byref(c)) == 0
                lo, hi = channel_range(rank, world, total)
                assert (f.value, c.value) == (lo, hi - lo)
    assert lib.uhsdr_channel_range(2, 2, 10, ctypes.byref(f), ctypes.byref(c)) == -1


def test_c_host_example_is_built_against_the_abi(built):
    """examples/host.c is compiled by a plain C compiler against include/uhsdr_b200.h and linked to the shipping library
    (the GPU test runs it); without a GPU it must fail loudly at engine creation, not fall back."""
    import subprocess
    exe = os.path.join(ROOT, "examples", "host")
    assert os.path.exists(exe)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, os.path.join(ROOT, "uhsdr_b200", "data", "uhsdr_tables.bin"), "8", "4", "0"], capture_output=True, text=True)
        assert r.returncode == 1 and "uhsdr_multi_create" in r.stderr


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


def test_corrupted_table_blobs_are_rejected_without_a_device(built):
    """Every section offset / count and every cross-index of the blob is bounds-checked at load time
    (uhsdr_tables_validate runs the same check as uhsdr_engine_create; no device needed)."""
    import random
    import struct
    from uhsdr_b200.tables import DEFAULT_BLOB
    lib = load_library()
    blob = open(DEFAULT_BLOB, "rb").read()

    def check(b):
        buf = ctypes.create_string_buffer(bytes(b), len(b))
        return lib.uhsdr_tables_validate(buf, len(b))

    assert check(blob) == 0
    hdr = struct.unpack("<14I", blob[:56])
    paths_off = hdr[6]
    cases = {"arrays_off": (16, len(blob) - 8), "paths_off": (24, len(blob) - 4), "extras_off": (52, len(blob)),
             "num_arrays": (12, 1 << 28), "path0.id": (paths_off, -3), "path0.fir_i_array": (paths_off + 16, 999),
             "path0.pre_lattice": (paths_off + 40, 4000)}
    for name, (off, val) in cases.items():
        b = bytearray(blob)
        struct.pack_into("<i" if val < 0 else "<I", b, off, val)
        assert check(b) == -4, name
        assert b"table blob" in lib.uhsdr_last_error(None)
    assert check(blob[:-4]) == -4 and check(blob[:40]) == -4
    rng = random.Random(7)
    for _ in range(500):          # random word corruptions in the index sections: rejected or harmless, never a crash
        b = bytearray(blob)
        struct.pack_into("<i", b, rng.randrange(0, 9900) // 4 * 4, rng.choice([-1, -2, 0x7fffffff, 1 << 20, rng.randrange(-5, 400)]))
        assert check(b) in (0, -4)


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
