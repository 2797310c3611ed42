"""The C-ABI library loads and exports every symbol include/jbsim.h declares (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import pytest

from jitterbug_b200 import _native, build as jb_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    jb_build.build()
    return _native.load()


def _declared():
    text = open(os.path.join(ROOT, "include", "jbsim.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/jbsim.h but not exported"
        assert n in _native.SYMBOLS, f"{n} has no ctypes prototype in _native.SYMBOLS"
    assert sorted(_native.SYMBOLS) == names


def test_version_and_struct_sizes(lib):
    assert b"jbsim" in lib.jb_version()
    from jitterbug_b200.model import JbModelF
    assert C.sizeof(JbModelF) % 4 == 0
    assert C.sizeof(_native.JbSimConfig) == 40


def test_no_cpu_fallback(lib):
    """Without a CUDA device creation must fail loudly, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from jitterbug_b200.batched_env import BatchedJitterbugEnv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchedJitterbugEnv(num_envs=2)
    from jitterbug_b200.model import JbModelF, fill_struct, flatten_model
    from jitterbug_b200 import mjcf
    s = fill_struct(flatten_model(mjcf.compile_xml()))
    cfg = _native.JbSimConfig(2, 0, 50, 1000, 1, 15, 0, 0, 0)
    h = C.c_void_p()
    rc = lib.jb_sim_create(C.byref(h), C.byref(cfg), C.byref(s), 1, None)
    assert rc != 0 and b"no CUDA device" in lib.jb_last_error()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "jitterbug_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "jb_oracle" not in text, f
