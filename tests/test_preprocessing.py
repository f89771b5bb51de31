"""The reference's own unit tests (test/test_processing.py, the only tests in the reference tree) restated for the sympy
restatement of tunempc/preprocessing.py: same cases, same known answers (h = [1,2,3], g = [-3]; h = [4,2,3], scost = 5000)."""
import collections

import numpy as np
import sympy as sp

from oracle import preprocessing_port as preprocessing
from oracle.preprocessing_port import SymFunction


def _xu():
    return sp.symbols("x0:1"), sp.symbols("u0:2")


def test_input_formatting_no_constraints():                             # test_processing.py:32-49
    x, u = _xu()
    sys = {"f": SymFunction("f", [x, u], [x[0]])}
    sys = preprocessing.input_formatting(sys)
    assert "h" not in sys and "g" not in sys and "vars" in sys
    assert len(sys["vars"]["x"]) == 1 and len(sys["vars"]["u"]) == 2


def test_input_formatting_lin_constraints():                            # test_processing.py:51-72
    x, u = _xu()
    sys = {"f": SymFunction("f", [x, u], [x[0]]), "h": SymFunction("h", [x, u], [x[0] + u[0], u[1]])}
    sys = preprocessing.input_formatting(sys)
    assert "h" in sys and sys["h"].size1_out(0) == 2 and "g" not in sys
    assert len(sys["vars"]["x"]) == 1 and len(sys["vars"]["u"]) == 2


def test_input_formatting_mixed_constraints():                          # test_processing.py:74-104
    x, u = _xu()
    sys = {"f": SymFunction("f", [x, u], [x[0]]), "h": SymFunction("h", [x, u], [x[0] + u[0], u[1], x[0] ** 2 * u[0]])}
    sys = preprocessing.input_formatting(sys)
    assert sys["h"].size1_out(0) == 3 and sys["g"].size1_out(0) == 1
    assert len(sys["vars"]["us"]) == 1
    h_eval = sys["h"](1.0, [0.0, 2.0], [3.0])
    g_eval = sys["g"](1.0, [0.0, 2.0], [3.0])
    assert h_eval.ravel().tolist() == [1.0, 2.0, 3.0]                    # :103
    assert g_eval.ravel().tolist() == [-3.0]                             # :104


def _sys_with_vars(h=None):
    x, u = _xu()
    sys = {"f": SymFunction("f", [x, u], [x[0]])}
    if h:
        sys["h"] = SymFunction("h", [x, u], [x[0] + u[0], u[1]])
    sys["vars"] = collections.OrderedDict()
    sys["vars"]["x"] = x
    sys["vars"]["u"] = u
    return sys


def test_add_mpc_slacks_no_constraints():                               # test_processing.py:106-123
    sys = preprocessing.add_mpc_slacks(_sys_with_vars(), None, None, slack_flag="active")
    assert "h" not in sys and "usc" not in sys["vars"]


def test_add_mpc_slacks_no_active_constraints():                        # test_processing.py:125-145
    sys = preprocessing.add_mpc_slacks(_sys_with_vars(h=True), None, [[], []], slack_flag="active")
    assert "h" in sys and "usc" not in sys["vars"]


def test_add_mpc_slacks_active_constraints():                           # test_processing.py:147-184
    lam_g = {"h": np.zeros((2, 2))}
    lam_g["h"][0, 0] = -5.0
    sys = preprocessing.add_mpc_slacks(_sys_with_vars(h=True), lam_g, [[0], []], slack_flag="active")
    assert "h" in sys and "usc" in sys["vars"] and len(sys["vars"]["usc"]) == 1 and "scost" in sys
    h_eval = sys["h"](1.0, [0.0, 2.0], [3.0])
    assert h_eval.ravel().tolist() == [4.0, 2.0, 3.0]                    # :183
    assert sys["scost"][0][0] == 5000.0                                  # :184
