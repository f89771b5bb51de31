"""B200-native batched solver for TuneMPC's MPC feedback solve (Pmpc.step + the SQP loop of sqp_method.py).

    from tunempc_b200 import Tuner, Pmpc, closed_loop_tools

`Tuner` / `Pmpc` / `closed_loop_tools` mirror `tunempc.Tuner`, `tunempc.pmpc.Pmpc` and `tunempc.closed_loop_tools`;
the arithmetic lives in libtmpc_<model>.so (CUDA, sm_100a) behind the C ABI of include/tmpc.h.  Imports are lazy so
that the host-side pieces (model cards, code generation, problem tables) can be used without torch or a GPU."""

__all__ = ["Tuner", "Pmpc", "MpcProblem", "closed_loop_tools", "configs"]


def __getattr__(name):
    if name == "Tuner":
        from .tuner import Tuner
        return Tuner
    if name == "Pmpc":
        from .pmpc import Pmpc
        return Pmpc
    if name == "MpcProblem":
        from .problem import MpcProblem
        return MpcProblem
    if name in ("closed_loop_tools", "configs", "tuning", "modelgen", "problem", "sharding", "lib", "pmpc", "tuner"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
