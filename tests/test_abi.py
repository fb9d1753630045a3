"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol
include/dcc_b200.h declares; status strings; argument validation that needs no device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import dcc_b200.build as b
    b.build()
    from dcc_b200 import _lib
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "dcc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from dcc_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libdcc_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes binding missing for %s" % n


def test_abi_version_and_status_strings(lib):
    assert lib.dcc_abi_version() == 7
    assert lib.dcc_status_string(0) == b"ok"
    assert b"invalid" in lib.dcc_status_string(-1)
    assert lib.dcc_env_obs_dim(4, 20) == 110 and lib.dcc_env_obs_dim(8, 64) == 338
    assert lib.dcc_env_obs_dim(16, 256) == 1314


def test_cfg_default_and_validation(lib):
    from dcc_b200 import _lib
    cfg = _lib.EnvCfg()
    assert lib.dcc_env_cfg_default(C.byref(cfg)) == 0
    assert (cfg.n_agents, cfg.n_pois, cfg.n_envs, cfg.max_ep_len) == (4, 20, 16, 150)
    assert (cfg.r_cover, cfg.r_comm, cfg.comm_r_scale, cfg.comm_force_scale) == (0.2, 0.4, 0.95, 0.0)
    assert cfg.reference_compat == 1
    h = C.c_void_p()
    poi = (C.c_double * 40)()
    assert lib.dcc_env_create(None, poi, 0, C.byref(h)) == -1
    cfg.n_agents = 33
    assert lib.dcc_env_create(C.byref(cfg), poi, 0, C.byref(h)) == -1
    assert lib.dcc_env_destroy(None) == -1
    assert lib.dcc_env_step(None, None, None, None, None, None, None, None, None, None) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from dcc_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.DccError):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dynamic-coverage-control_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert "import oracle" not in s and "from oracle" not in s and "libdcc_oracle" not in s, f
