"""Batch drivers of the reference (`tunempc/closed_loop_tools.py`) with the Python loops turned into the batch axis.

Reference                                             here
  check_equivalence(controllers,cost,h,x0,dx,alpha)     same signature; the `for alph in alpha` loop (:43-68) is ONE batched
                                                        `ctrl.step(X_init)` per controller, X_init = x0 + alpha[:,None]*dx
  closed_loop_sim(controllers,cost,h,F,x0,N)            same signature; x0 may be (nx,) or (B,nx): B independent rollouts,
                                                        one batched step + plant step per time step (:84-102)
  initialize_log(controllers,x0)                        same

`cost`, `h` and `F` are CasADi Functions in the reference.  Here they default to the compiled model's own device
functions (stage cost `tmpc_stage_cost`, h = C z + c, the RK4 plant `tmpc_plant_step`), reached through the controller;
callables taking/returning torch CUDA tensors may be passed instead.  Everything stays on the GPU; the log holds
torch tensors with the batch as leading axis.  `reduce_rollout_stats` is the only collective of a multi-GPU run
(NCCL all_reduce of a fixed-size statistics vector).
"""
from __future__ import annotations

import numpy as np


def initialize_log(controllers, x0=None):                              # closed_loop_tools.py:106-121
    log = {"u": {}, "l": {}, "h": {}, "log": {}}
    if x0 is not None:
        log["x"] = {}
    for name in list(controllers.keys()):
        for log_key in log.keys():
            log[log_key][name] = []
        if x0 is not None:
            log["x"][name] = [x0]
    return log


def _as_cuda(a, ctrl):
    import torch
    if isinstance(a, torch.Tensor):
        return a.to(dtype=torch.float64, device="cuda:%d" % ctrl.device)
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device="cuda:%d" % ctrl.device)


def check_equivalence(controllers, cost, h, x0, dx, alpha, flag="tunempc"):
    """closed_loop_tools.py:30-70.  Returns ONE log dict (instead of a list with one dict per alpha): for every
    controller `u` (B,N,nu), `x` (B,N,nx) predicted trajectories, `l` (B,N) stage costs and `h` (B,N) first constraint
    row along them (the reference logs `h(...)[0][0]`, :65), B = len(alpha); plus `x_init` (B,nx) and `status`."""
    if flag != "tunempc":
        raise NotImplementedError("the acados export path (step_acados) is out of scope")
    import torch
    names = list(controllers.keys())
    c0 = controllers[names[0]]
    x0 = _as_cuda(np.asarray(x0, dtype=np.float64).reshape(1, -1), c0)
    dxv = _as_cuda(np.asarray(dx, dtype=np.float64).reshape(1, -1), c0)
    al = _as_cuda(np.asarray(alpha, dtype=np.float64).reshape(-1, 1), c0)
    X_init = (x0 + al * dxv).contiguous()                               # :47
    log = initialize_log(controllers, X_init)
    log["status"] = {}
    for name in names:
        ctrl = controllers[name]
        pb = ctrl.problem
        ctrl.reset()
        ctrl.step(X_init)                                               # :56
        w = ctrl.w_sol
        B, N, nz, nx = w.shape[0], pb.N, pb.nz, pb.nx
        Z = w[:, : N * nz].reshape(B, N, nz)
        Xp, Up = Z[:, :, :nx].contiguous(), Z[:, :, nx:nx + pb.nu].contiguous()   # the model inputs (slack variables follow them in a stage)
        log["u"][name] = Up                                             # :62-63
        log["x"][name] = Xp
        if cost is None or h is None:
            l_dev, h_dev = ctrl.stage_log(Xp.reshape(B * N, nx), Up.reshape(B * N, pb.nu))
        log["l"][name] = (l_dev if cost is None else cost(Xp.reshape(B * N, nx), Up.reshape(B * N, pb.nu))).reshape(B, N)  # :64
        if h is None:
            log["h"][name] = h_dev.reshape(B, N, -1)[:, :, 0] if pb.nh else torch.zeros((B, N), dtype=torch.float64, device=w.device)
        else:
            log["h"][name] = h(Xp.reshape(B * N, nx), Up.reshape(B * N, pb.nu)).reshape(B, N, -1)[:, :, 0]             # :65
        log["status"][name] = ctrl.status
        ctrl.reset()                                                    # :68
    log["x_init"] = X_init
    return log


def closed_loop_sim(controllers, cost, h, F, x0, N, flag="tunempc", disturbance=None):
    """closed_loop_tools.py:72-104 for B rollouts at once.  x0: (nx,) or (B,nx).  Log per controller: `x` list of N+1
    tensors (B,nx), `u` list of N tensors (B,nu), `l` list of N tensors (B,), `h` list of N tensors (B,nh), and
    `status` list of N int32 tensors (B,).  `disturbance(i, X) -> X` (optional) perturbs the state before step i
    (examples/unicycle/main.py:196-200)."""
    if flag != "tunempc":
        raise NotImplementedError("the acados export path (step_acados) is out of scope")
    names = list(controllers.keys())
    c0 = controllers[names[0]]
    X0 = _as_cuda(np.asarray(x0.detach().cpu().numpy() if hasattr(x0, "detach") else x0, dtype=np.float64).reshape(-1, c0.problem.nx), c0)
    log = initialize_log(controllers, X0)
    log["status"] = {name: [] for name in names}
    for name in names:
        controllers[name].reset()
    for i in range(N):                                                  # :84
        for name in names:
            ctrl = controllers[name]
            X = log["x"][name][-1]
            if disturbance is not None:
                X = disturbance(i, X)
                log["x"][name][-1] = X
            U = ctrl.step(X, outputs="u0")                              # :95 (the (B,n_w) / (B,n_g) solution tensors are not kept per step)
            log["u"][name].append(U)
            log["status"][name].append(ctrl.status)
            if cost is None or h is None:
                l_dev, h_dev = ctrl.stage_log(X, U)
            log["l"][name].append(l_dev if cost is None else cost(X, U))   # :98
            log["h"][name].append(h_dev if h is None else h(X, U))         # :99
            log["x"][name].append(ctrl.plant_step(X, U) if F is None else F(X, U))   # :102
    return log


def rollout_stats(log, name, l_ref=None):
    """fixed-size statistics vector of one controller's rollouts on this rank:
    [0] rollouts, [1] steps, [2] sum of stage costs, [3] sum of (l - l_ref) (transient cost, paper eq. (23)),
    [4] non-converged solves, [5] max constraint violation max(0, -h)"""
    import torch
    L = torch.stack(log["l"][name], dim=1)                              # (B, N)
    v = torch.zeros(6, dtype=torch.float64, device=L.device)
    v[0] = L.shape[0]
    v[1] = L.shape[1]
    v[2] = L.sum()
    if l_ref is not None:
        v[3] = (L - torch.as_tensor(l_ref, dtype=torch.float64, device=L.device).reshape(1, -1)[:, : L.shape[1]]).sum()
    v[4] = float(sum(int((s != 0).sum()) for s in log["status"][name]))
    hs = [t for t in log["h"][name] if t.numel()]
    if hs:
        v[5] = torch.clamp(-torch.stack(hs, dim=1), min=0.0).max()
    return v


def reduce_rollout_stats(v, dist=None):
    """sum over ranks of entries [0], [2..4]; max of [1] and [5].  `dist` = torch.distributed (NCCL on GPUs, gloo in
    the CPU tests); the only collective of a sharded closed-loop run."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return v
    s = v.clone()
    m = v.clone()
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    s[1] = m[1]
    s[5] = m[5]
    return s
