"""The C-ABI shared library builds, loads and exports every symbol include/batch_drones.h declares.
No compute call is made (no GPU here): only argument validation paths that return before CUDA."""
import ctypes as C
import os
import re
import shutil

import pytest

from marl_gym_pybullet_drones_b200 import _native
from marl_gym_pybullet_drones_b200 import build as bd_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "batch_drones.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_native.LIB_PATH) or not bd_build.up_to_date():
        if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
            pytest.skip("nvcc not available and library not prebuilt")
        bd_build.build()
    return _native.load()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bd_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    assert set(syms) == set(_native.EXPORTS), (syms, _native.EXPORTS)


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} missing from {_native.LIB_PATH}"
    assert lib.bd_version() == 1


def test_config_struct_layout_matches_header(lib):
    # field list of the ctypes mirror == field list of the C struct, in order
    text = open(HEADER).read()
    body = re.search(r"typedef struct bd_config \{(.*?)\} bd_config;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(int32_t|uint64_t|double)\s+", "", decl)
        for part in decl.split(","):
            names.append(re.sub(r"\[.*\]", "", part).strip())
    assert names == [f[0] for f in _native.BdConfig._fields_]
    assert C.sizeof(_native.BdConfig) == 352


def test_create_rejects_bad_configs_with_messages(lib):
    h = C.c_void_p()
    cfg = _native.BdConfig()
    assert lib.bd_create(C.byref(cfg), C.byref(h)) == -1          # struct_size mismatch
    assert b"struct_size" in lib.bd_last_error()
    cfg.struct_size = C.sizeof(cfg)
    cfg.n_envs, cfg.n_drones = 4, 200
    assert lib.bd_create(C.byref(cfg), C.byref(h)) == -1 and b"n_drones" in lib.bd_last_error()
    cfg.n_drones = 2
    cfg.task = 1
    cfg.pyb_freq, cfg.ctrl_freq = 240, 50
    cfg.mass = cfg.kf = cfg.ixx = cfg.iyy = cfg.izz = 1.0
    assert lib.bd_create(C.byref(cfg), C.byref(h)) == -1
    assert b"pyb_freq is not divisible by env_freq" in lib.bd_last_error()   # BaseAviary.py:79-80 wording
    cfg.ctrl_freq = 30
    cfg.act_type = 7
    assert lib.bd_create(C.byref(cfg), C.byref(h)) == -1 and b"act_type" in lib.bd_last_error()
    cfg.act_type, cfg.drone_model = 3, 2          # VEL on the racer: BaseRLAviary.py:73-78 has no controller
    assert lib.bd_create(C.byref(cfg), C.byref(h)) == -1 and b"no controller is available" in lib.bd_last_error()
    cfg.drone_model = 0
    assert not h.value
    # null-handle calls fail cleanly instead of crashing
    assert lib.bd_step(None, None, None, None, None, None, None, None) == -1
    assert lib.bd_obs_dim(None) == -1 and lib.bd_launch_count(None) == 0
    lib.bd_destroy(None)


def test_sass_is_sm100a(lib):
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    import subprocess
    out = subprocess.run([cuobjdump, "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
