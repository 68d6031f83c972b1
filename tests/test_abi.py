"""CPU: the C-ABI library loads, exports every symbol include/icrl_b200.h declares, agrees with the ctypes mirror on
struct layouts, and validates arguments without touching a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from icrl_b200 import build, _lib
    build.build()
    return _lib.lib()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "icrl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(icrl_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from icrl_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/icrl_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for s in syms:
        assert re.search(rf"\bT {s}\b", out), f"{s} is not a defined text symbol of the shared library"


def test_struct_layouts_match_the_header(lib, tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirror."""
    from icrl_b200 import _lib
    probe = tmp_path / "probe.c"
    probe.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "icrl_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                     'sizeof(icrl_cn_desc), offsetof(icrl_cn_desc, params), sizeof(icrl_cn_train_cfg), offsetof(icrl_cn_train_cfg, lr),'
                     'sizeof(icrl_cn_train_metrics), sizeof(icrl_ppo_cfg), offsetof(icrl_ppo_cfg, lr), sizeof(icrl_ppo_data),'
                     'offsetof(icrl_cn_train_cfg, perm), sizeof(icrl_cn_dist), offsetof(icrl_cn_dist, episode_base),'
                     'sizeof(icrl_ppo_dist));return 0;}\n')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.CnDesc), _lib.CnDesc.params.offset, C.sizeof(_lib.CnTrainCfg), _lib.CnTrainCfg.lr.offset,
            C.sizeof(_lib.CnTrainMetrics), C.sizeof(_lib.PpoCfg), _lib.PpoCfg.lr.offset, C.sizeof(_lib.PpoData),
            _lib.CnTrainCfg.perm.offset, C.sizeof(_lib.CnDist), _lib.CnDist.episode_base.offset, C.sizeof(_lib.PpoDist)]
    assert got == want


def test_argument_validation_without_gpu(lib):
    from icrl_b200 import _lib
    assert lib.icrl_abi_version() == _lib.ABI_VERSION == 2
    d = _lib.CnDesc()
    assert lib.icrl_cn_param_count(C.byref(d)) == -1                     # empty descriptor is rejected
    assert b"obs_dim" in lib.icrl_last_error()
    d.obs_dim, d.acs_dim, d.n_select, d.n_hidden = 18, 6, 24, 1
    d.hidden[0] = 20
    for i in range(24):
        d.select[i] = i
    d.params = 1                                                           # never dereferenced by param_count
    assert lib.icrl_cn_param_count(C.byref(d)) == 24 * 20 + 20 + 20 + 1   # 521, SURVEY §8 table
    d.hidden[0] = 65
    assert lib.icrl_cn_param_count(C.byref(d)) == -1
    cfg = _lib.PpoCfg()
    cfg.obs_dim, cfg.act_dim, cfg.is_discrete = 113, 8, 0
    cfg.hidden[0] = cfg.hidden[1] = 64
    cfg.T = cfg.E = 1
    assert lib.icrl_ppo_param_count(C.byref(cfg)) == 35026               # AntWall policy, SURVEY §8 table
    cfg.obs_dim, cfg.act_dim = 18, 6
    assert lib.icrl_ppo_param_count(C.byref(cfg)) == 16654               # HalfCheetah
    cfg.obs_dim, cfg.act_dim, cfg.is_discrete = 1, 2, 1
    assert lib.icrl_ppo_param_count(C.byref(cfg)) == 13124               # LapGrid (no log_std)
    assert lib.icrl_dual_gae(None, None, None, None, None, None, None, None, 4, 4, 0.99, 0.95, 0.99, 0.95,
                             None, None, None, None, None) == -1           # NULL arrays -> ICRL_EINVAL, no launch
    assert lib.icrl_launch_count() == 0


def test_product_has_no_cpu_fallback_and_never_imports_the_oracle():
    """The package must not reference oracle/ anywhere, and must refuse to run without CUDA."""
    import torch as th
    pkg = os.path.join(ROOT, "icrl_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f
    if not th.cuda.is_available():
        from icrl_b200.device import resolve_device
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            resolve_device("cuda")
