"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every symbol that
include/tmpc.h declares, and the host-side mirror of the reference interface fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_problem


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "tmpc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tmpc_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built):
    from tunempc_b200.lib import ModelLib, EXPORTS
    syms = _declared_symbols()
    assert set(syms) == set(EXPORTS)
    for name in ("lq", "cstr", "unicycle", "evaporation", "chain", "dims9"):
        lib = ModelLib(name)
        for s in syms:
            assert hasattr(lib.lib, s), "%s missing in %s" % (s, lib.path)
        assert lib.model_name == name


def test_model_info_and_host_stage_eval(built):
    from tunempc_b200.lib import ModelLib
    from oracle import reference_port as rp
    lib = ModelLib("cstr")
    assert (lib.nx, lib.nu, lib.rk_steps, lib.dt) == (4, 2, 20, 1.0)
    z = np.array([2.1402, 1.0903, 114.191, 112.9066, 14.19, -1113.5])
    a = lib.stage_eval(z[:4], z[4:], 2)
    b = rp.StageLib("cstr").F(z[:4], z[4:], 2)
    for x, y in zip(a, b):
        assert np.max(np.abs(x - y)) <= 1e-12 * max(1.0, np.max(np.abs(y)))


def test_sm100a_code_in_library(built):
    import subprocess
    from tunempc_b200.lib import lib_path
    out = subprocess.run(["cuobjdump", "-lelf", lib_path("cstr")], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_options_and_no_cpu_fallback(built):
    import torch
    from tunempc_b200.pmpc import Pmpc, default_options
    pb = load_problem("cstr")
    assert default_options()["hessian_approximation"] == "exact" and default_options()["max_iter"] == 2000
    with pytest.raises(ValueError, match="Unknown option for Pmpc class instance"):
        Pmpc(pb, options={"no_such_option": 1})
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            Pmpc(pb)


def test_tables_follow_reference_layout():
    from tunempc_b200.problem import build_tables
    pb = load_problem("cstr")
    tab = build_tables(pb)
    assert tab.ref.shape == (1, pb.n_w) and tab.ref_du.shape == (1, pb.n_g)
    assert pb.n_w == 124 and pb.n_g == 168                       # SURVEY.md section 8.0 dimension table, cfg #2
    assert np.array_equal(tab.ref[0, :6], pb.wref[0]) and np.array_equal(tab.ref[0, -4:], pb.wref[0, :4])
    lbg, ubg = pb.bounds()
    assert np.isinf(ubg[pb.g_h(3)]).all() and (lbg[pb.g_h(3)] == 0).all() and (ubg[pb.g_dyn(3)] == 0).all()
    assert pb.h_x_idx == []                                      # CSTR: input bounds only (cstr_model.py:154-159)
