"""Model code generation: sympy ODE -> one C header usable from nvcc (device) and gcc (host).

The reference hands user dynamics to the solver as CasADi Functions wrapping an integrator
(`examples/cstr/cstr_model.py:76-99`, `examples/unicycle/main.py:51-67`) and differentiates them
symbolically (`tunempc/sqp_method.py:82-98`).  CasADi is not available here, so the front-end is sympy;
the emitted bundle is the same one acados generates from a CasADi model
(`external/acados/interfaces/acados_template/acados_template/generate_c_code_explicit_ode.py:66-96`):
ODE, its Jacobian and the second derivatives needed for the exact Lagrangian Hessian.

Emitted symbols (all `TMPC_HD static inline`, TMPC_HD = `__host__ __device__` under nvcc):
  tmpc_ode(x,u,f)                     f = xdot
  tmpc_ode_jac(x,u,f,J)               J row-major [NX][NZ], NZ = NX+NU
  tmpc_ode_d2(x,u,f,J,H)              H[TMPC_NHESS] structurally non-zero d2 f_a / dz_b dz_c, b<=c
  tmpc_ode_bilin(H,v,w,out)           out[a] = sum_bc H_abc v_b w_c   (straight-line over the non-zeros)
  tmpc_stage_cost(x,u)                economic stage cost l(x,u) of the model card (closed-loop log only)
plus the integrator spec (RK4 step count and step length) and op counts for the roofline.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Sequence

import sympy as sp
from sympy.printing.c import C99CodePrinter


class _Printer(C99CodePrinter):
    """C printer that keeps small integer powers as products (no pow() calls in device code)."""

    def _print_Pow(self, expr):
        b, e = expr.as_base_exp()
        if e.is_Integer and 2 <= int(e) <= 4:
            s = self._print(b)
            if not b.is_Atom:
                s = "(" + s + ")"
            return "(" + "*".join([s] * int(e)) + ")"
        if e.is_Integer and -4 <= int(e) <= -1:
            s = self._print(b)
            if not b.is_Atom:
                s = "(" + s + ")"
            return "(1.0/(" + "*".join([s] * int(-e)) + "))"
        return super()._print_Pow(expr)


_P = _Printer()


@dataclass
class OdeModel:
    name: str
    x: Sequence[sp.Symbol]
    u: Sequence[sp.Symbol]
    xdot: Sequence[sp.Expr]
    rk_steps: int = 1        # CasADi 'rk' number_of_finite_elements
    tf: float = 1.0          # integrator horizon (one MPC interval)
    discrete: bool = False   # True: xdot IS the map x+ = f(x,u) (no integrator)
    integrator: str = "rk4"  # 'rk4' (CasADi 'rk') or 'collocation' (CasADi 'collocation': Radau IIA, 3 nodes per element)
    hess_nz: List[tuple] = field(default_factory=list)
    cost: object = None      # economic stage cost l(x,u) (sympy), emitted as tmpc_stage_cost for the closed-loop log
    gnl: Sequence[sp.Expr] = ()   # nonlinear path constraints h_nl(x,u) >= 0: each gets a slack us_i and the equality row
                                  # h_nl,i(x,u) - us_i = 0 (tunempc/preprocessing.py:78-118); ns = len(gnl)
    nsc: int = 0             # soft-constraint slacks usc of the MPC (tunempc/preprocessing.py:120-155); stage variables
                             # are (x, u, us, usc), the dynamics and gnl depend on (x, u) only

    @property
    def nx(self):
        return len(self.x)

    @property
    def nu(self):
        return len(self.u)

    @property
    def ns(self):
        return len(self.gnl)


def _count_ops(exprs) -> int:
    return int(sum(sp.count_ops(e, visual=False) for e in exprs))


def _emit_block(lines, assigns, outputs):
    """CSE a list of (target, expr) and append C statements."""
    targets = [t for t, _ in outputs]
    exprs = [sp.nsimplify(e, rational=False) if False else e for _, e in outputs]
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("t_"), optimizations="basic")
    nops = 0
    for s, e in repl:
        lines.append("  const double %s = %s;" % (s, _P.doprint(e)))
        nops += sp.count_ops(e)
    for t, e in zip(targets, red):
        lines.append("  %s = %s;" % (t, _P.doprint(e)))
        nops += sp.count_ops(e)
    return int(nops)


def lin_groups(nz, d=3):
    """Cover all direction pairs (i<=j) of nz directions by groups of <= d directions (a pair covering design, found
    exhaustively) and assign every pair to exactly one group, balanced.  One CUDA thread integrates one group: the ODE,
    its Jacobian and second derivatives are evaluated once per group instead of once per pair."""
    import itertools
    d = min(d, nz)
    pairs = [(i, j) for i in range(nz) for j in range(i + 1, nz)]
    triples = list(itertools.combinations(range(nz), d))
    groups = None
    if nz <= 6:
        for ng in range(1, len(pairs) + 2):
            for gs in itertools.combinations(triples, ng):
                cov = set(p for g in gs for p in itertools.combinations(g, 2))
                if len(cov) == len(pairs) and set(x for g in gs for x in g) == set(range(nz)):
                    groups = list(gs)
                    break
            if groups:
                break
    else:
        # greedy set cover (the exhaustive search explodes beyond 6 directions)
        left = set(pairs)
        groups = []
        while left:
            g = max(triples, key=lambda t: len(left & set(itertools.combinations(t, 2))))
            groups.append(g)
            left -= set(itertools.combinations(g, 2))
        for x in range(nz):
            if not any(x in g for g in groups):
                groups.append(tuple(sorted((x, (x + 1) % nz, (x + 2) % nz))))
    allp = [(i, j) for i in range(nz) for j in range(i, nz)]
    load = [[] for _ in groups]
    opts = lambda p: [g for g, t in enumerate(groups) if p[0] in t and p[1] in t]
    for p in sorted(allp, key=lambda p: len(opts(p))):
        g = min(opts(p), key=lambda g: len(load[g]))
        load[g].append(p)
    return groups, load


def lin2_roles(nz, npart=3):
    """Role tables of the warp-specialised linearisation kernel (csrc/tmpc_lin2.cuh).  Every consumer warp runs the
    SAME code on a 'star' of direction pairs: a centre direction d and up to `npart` partners j (j == d is the diagonal
    pair), so that one operand column of every pair is shared.  The first nz roles own (integrate and publish) the
    first-order column of their centre; further roles only read it.  Returns [(centre, owns, [partners])]."""
    import random
    rng = random.Random(0)
    lower = -(-(nz * (nz + 1) // 2) // npart)
    best = None
    for attempt in range(400):
        pairs = set((i, j) for i in range(nz) for j in range(i + 1, nz))
        roles = [[d, 1, [d]] for d in range(nz)]
        order = list(range(nz))
        if attempt:
            rng.shuffle(order)
        for d in order:                          # owners pick partners among their free incident pairs
            inc = sorted(e for e in pairs if d in e)
            if attempt:
                rng.shuffle(inc)
            for e in inc[: npart - 1]:
                roles[d][2].append(e[0] if e[1] == d else e[1])
                pairs.discard(e)
        while pairs:                             # the rest is grouped into stars (most frequent endpoint first)
            cnt = {}
            for (i, j) in pairs:
                cnt[i] = cnt.get(i, 0) + 1
                cnt[j] = cnt.get(j, 0) + 1
            d = max(sorted(cnt), key=lambda k: cnt[k])
            mine = sorted(e for e in pairs if d in e)[:npart]
            roles.append([d, 0, [e[0] if e[1] == d else e[1] for e in mine]])
            for e in mine:
                pairs.discard(e)
        if best is None or len(roles) < len(best):
            best = roles
        if len(best) <= max(lower, nz):
            break
    roles = best
    # warp w of the CTA runs on scheduler w % 4; warp 0 is the producer (the heaviest role), consumer role r is warp r + 1.
    # Put the lightest roles (no first-order column to integrate) on the producer's scheduler: positions r with (r+1) % 4 == 0.
    light = [r for r in roles if not r[1]]
    heavy = [r for r in roles if r[1]]
    out = []
    for pos in range(len(roles)):
        if (pos + 1) % 4 == 0 and light:
            out.append(light.pop(0))
        elif heavy:
            out.append(heavy.pop(0))
        else:
            out.append(light.pop(0))
    return [(r[0], r[1], list(r[2])) for r in out]


def generate_header(model: OdeModel, out_path: str) -> dict:
    nx, nu = model.nx, model.nu
    nz = nx + nu
    z = list(model.x) + list(model.u)
    f = [sp.sympify(e) for e in model.xdot]
    J = [[sp.diff(f[a], z[b]) for b in range(nz)] for a in range(nx)]
    hess = []  # (a,b,c,expr)
    for a in range(nx):
        for b in range(nz):
            for c in range(b, nz):
                e = sp.diff(f[a], z[b], z[c])
                if sp.simplify(e) != 0:
                    hess.append((a, b, c, e))
    model.hess_nz = [(a, b, c) for a, b, c, _ in hess]

    sub = {}
    for i, s in enumerate(model.x):
        sub[s] = sp.Symbol("x[%d]" % i)
    for i, s in enumerate(model.u):
        sub[s] = sp.Symbol("u[%d]" % i)

    def S(e):
        return sp.sympify(e).xreplace(sub)

    L = []
    guard = "TMPC_MODEL_%s_H" % model.name.upper()
    L.append("/* generated by tunempc_b200/modelgen.py -- do not edit */")
    L.append("#ifndef %s\n#define %s" % (guard, guard))
    L.append("#include <math.h>")
    L.append("#ifndef TMPC_HD\n#ifdef __CUDACC__\n#define TMPC_HD __host__ __device__ __forceinline__\n#else\n#define TMPC_HD static inline\n#endif\n#endif")
    L.append('#define TMPC_MODEL_NAME "%s"' % model.name)
    # stage variables z = (x, u, us, usc): TMPC_NU / TMPC_NZ count every free variable of a stage block, TMPC_NUM / TMPC_NZM
    # only what the dynamics (and gnl) depend on -- equal when the problem has no slacks
    ns, nsc = model.ns, int(model.nsc)
    L.append("#define TMPC_NX %d\n#define TMPC_NUM %d\n#define TMPC_NZM %d\n#define TMPC_NS %d\n#define TMPC_NSC %d" % (nx, nu, nz, ns, nsc))
    L.append("#define TMPC_NU %d\n#define TMPC_NZ %d" % (nu + ns + nsc, nz + ns + nsc))
    L.append("#define TMPC_DISCRETE %d" % (1 if model.discrete else 0))
    L.append("#define TMPC_COLLOCATION %d" % (1 if (model.integrator == "collocation" and not model.discrete) else 0))
    L.append("#define TMPC_RK_STEPS %d" % model.rk_steps)
    L.append("#define TMPC_RK_DT %s" % repr(float(model.tf) / model.rk_steps))
    L.append("#define TMPC_NHESS %d" % len(hess))
    na = max(1, len(hess))
    L.append("#define TMPC_HESS_A {%s}" % ",".join(str(h[0]) for h in hess) if hess else "#define TMPC_HESS_A {0}")
    L.append("#define TMPC_HESS_B {%s}" % ",".join(str(h[1]) for h in hess) if hess else "#define TMPC_HESS_B {0}")
    L.append("#define TMPC_HESS_C {%s}" % ",".join(str(h[2]) for h in hess) if hess else "#define TMPC_HESS_C {0}")

    # linearisation task groups (see lin_groups)
    groups, load = lin_groups(nz)
    D = max(len(g) for g in groups)
    PP = max(len(l) for l in load)
    owner = {}
    for gi, g in enumerate(groups):
        for a in g:
            owner.setdefault(a, gi)
    gd, gown, gpa, gpb, gpi = [], [], [], [], []
    pidx = lambda i, j: i * nz - i * (i - 1) // 2 + (j - i)
    for gi, g in enumerate(groups):
        gd += list(g) + [-1] * (D - len(g))
        gown += [1 if owner[a] == gi else 0 for a in g] + [0] * (D - len(g))
        for (i, j) in load[gi]:
            gpa.append(g.index(i)); gpb.append(g.index(j)); gpi.append(pidx(i, j))
        pad = PP - len(load[gi])
        gpa += [0] * pad; gpb += [0] * pad; gpi += [-1] * pad
    L.append("#define TMPC_LIN_NG %d\n#define TMPC_LIN_D %d\n#define TMPC_LIN_PP %d" % (len(groups), D, PP))
    L.append("#define TMPC_LIN_GND {%s}" % ",".join(str(len(g)) for g in groups))
    L.append("#define TMPC_LIN_GNP {%s}" % ",".join(str(len(l)) for l in load))
    L.append("#define TMPC_LIN_GD {%s}" % ",".join(map(str, gd)))
    L.append("#define TMPC_LIN_GOWN {%s}" % ",".join(map(str, gown)))
    L.append("#define TMPC_LIN_GPA {%s}" % ",".join(map(str, gpa)))
    L.append("#define TMPC_LIN_GPB {%s}" % ",".join(map(str, gpb)))
    L.append("#define TMPC_LIN_GPI {%s}" % ",".join(map(str, gpi)))
    # warp-specialised kernel: consumer roles and the structural sparsity of J
    roles = lin2_roles(nz, int(os.environ.get("TMPC_L2_NPART", "3")))
    np2 = max(len(r[2]) for r in roles)
    part = []
    for _, _, js in roles:
        part += list(js) + [-1] * (np2 - len(js))
    L.append("#define TMPC_L2_NCW %d\n#define TMPC_L2_NP %d" % (len(roles), np2))
    L.append("#define TMPC_L2_CEN {%s}" % ",".join(str(r[0]) for r in roles))
    L.append("#define TMPC_L2_OWN {%s}" % ",".join(str(r[1]) for r in roles))
    L.append("#define TMPC_L2_PART {%s}" % ",".join(map(str, part)))
    L.append("#define TMPC_JNZ {%s}" % ",".join("0" if sp.simplify(J[a][b]) == 0 else "1"
                                                  for a in range(nx) for b in range(nz)))
    # ode
    L.append("TMPC_HD void tmpc_ode(const double* x, const double* u, double* f) {")
    L.append("  (void)x; (void)u;")
    c_f = _emit_block(L, None, [("f[%d]" % a, S(f[a])) for a in range(nx)])
    L.append("}")
    # ode + jac
    L.append("TMPC_HD void tmpc_ode_jac(const double* x, const double* u, double* f, double* J) {")
    L.append("  (void)x; (void)u;")
    outs = [("f[%d]" % a, S(f[a])) for a in range(nx)]
    outs += [("J[%d]" % (a * nz + b), S(J[a][b])) for a in range(nx) for b in range(nz)]
    c_J = _emit_block(L, None, outs)
    L.append("}")
    # ode + jac + second derivatives
    L.append("TMPC_HD void tmpc_ode_d2(const double* x, const double* u, double* f, double* J, double* H) {")
    L.append("  (void)x; (void)u; (void)H;")
    outs2 = list(outs) + [("H[%d]" % i, S(h[3])) for i, h in enumerate(hess)]
    c_H = _emit_block(L, None, outs2)
    L.append("}")
    # bilinear form over the non-zeros
    L.append("TMPC_HD void tmpc_ode_bilin(const double* H, const double* v, const double* w, double* out) {")
    L.append("  (void)H; (void)v; (void)w;")
    for a in range(nx):
        terms = []
        for i, (aa, b, c, _) in enumerate(hess):
            if aa != a:
                continue
            if b == c:
                terms.append("H[%d]*v[%d]*w[%d]" % (i, b, c))
            else:
                terms.append("H[%d]*(v[%d]*w[%d]+v[%d]*w[%d])" % (i, b, c, c, b))
        L.append("  out[%d] = %s;" % (a, " + ".join(terms) if terms else "0.0"))
    L.append("}")
    # the same bilinear form in two steps, for several w sharing one v:  G = d2f . v  (compact over the structural
    # non-zeros (a,c)),  out = G w
    gpos = {}
    gterms = {}
    for i, (a, b, c, _) in enumerate(hess):
        for (cc, bb) in ((c, b),) if b == c else ((c, b), (b, c)):
            gpos.setdefault((a, cc), len(gpos))
            gterms.setdefault((a, cc), []).append("H[%d]*v[%d]" % (i, bb))
    L.append("#define TMPC_NG %d" % max(1, len(gpos)))
    L.append("TMPC_HD void tmpc_ode_hv(const double* H, const double* v, double* G) {")
    L.append("  (void)H; (void)v; (void)G;")
    for key, gi in gpos.items():
        L.append("  G[%d] = %s;" % (gi, " + ".join(gterms[key])))
    L.append("}")
    L.append("TMPC_HD void tmpc_ode_gw(const double* G, const double* w, double* out) {")
    L.append("  (void)G; (void)w;")
    for a in range(nx):
        terms = ["G[%d]*w[%d]" % (gi, c) for (aa, c), gi in gpos.items() if aa == a]
        L.append("  out[%d] = %s;" % (a, " + ".join(terms) if terms else "0.0"))
    L.append("}")
    # economic stage cost l(x,u): logged along closed-loop runs (tunempc/closed_loop_tools.py:64,98)
    L.append("#define TMPC_HAS_COST %d" % (1 if model.cost is not None else 0))
    L.append("TMPC_HD double tmpc_stage_cost(const double* x, const double* u) {")
    L.append("  (void)x; (void)u;")
    L.append("  double l = 0.0;")
    if model.cost is not None:
        _emit_block(L, None, [("l", S(model.cost))])
    L.append("  return l;")
    L.append("}")
    # gradient and Hessian of l w.r.t. z = (x,u): the economic-MPC stage cost on the device (pmpc.py:97-107,299-301)
    L.append("TMPC_HD void tmpc_cost_grad(const double* x, const double* u, double* g) {")
    L.append("  (void)x; (void)u;")
    if model.cost is not None:
        cg = [sp.diff(model.cost, zz) for zz in z]
        _emit_block(L, None, [("g[%d]" % i, S(cg[i])) for i in range(nz)])
    else:
        L.append("  for (int i = 0; i < %d; ++i) g[i] = 0.0;" % nz)
    L.append("}")
    L.append("TMPC_HD void tmpc_cost_hess(const double* x, const double* u, double* H) {")
    L.append("  (void)x; (void)u;")
    if model.cost is not None:
        _emit_block(L, None, [("H[%d]" % (i * nz + j), S(sp.diff(model.cost, z[i], z[j]))) for i in range(nz) for j in range(nz)])
    else:
        L.append("  for (int i = 0; i < %d; ++i) H[i] = 0.0;" % (nz * nz))
    L.append("}")
    # slacked nonlinear path constraints (tunempc/preprocessing.py:78-118): value, Jacobian (row-major [NS][NZM]) and the
    # Hessian of lam'gnl w.r.t. (x,u) (row-major [NZM][NZM])
    if ns:
        gn = [sp.sympify(e) for e in model.gnl]
        L.append("TMPC_HD void tmpc_gnl(const double* x, const double* u, double* g) {")
        L.append("  (void)x; (void)u;")
        _emit_block(L, None, [("g[%d]" % i, S(gn[i])) for i in range(ns)])
        L.append("}")
        L.append("TMPC_HD void tmpc_gnl_jac(const double* x, const double* u, double* g, double* J) {")
        L.append("  (void)x; (void)u;")
        _emit_block(L, None, [("g[%d]" % i, S(gn[i])) for i in range(ns)] +
                    [("J[%d]" % (i * nz + b), S(sp.diff(gn[i], z[b]))) for i in range(ns) for b in range(nz)])
        L.append("}")
        lam_s = [sp.Symbol("lam[%d]" % i) for i in range(ns)]
        lg = sum(lam_s[i] * gn[i] for i in range(ns))
        L.append("TMPC_HD void tmpc_gnl_hess(const double* x, const double* u, const double* lam, double* H) {")
        L.append("  (void)x; (void)u; (void)lam;")
        _emit_block(L, None, [("H[%d]" % (i * nz + j), S(sp.diff(lg, z[i], z[j]))) for i in range(nz) for j in range(nz)])
        L.append("}")
    c_B = sum(3 if b == c else 5 for _, b, c, _ in hess)
    L.append("#define TMPC_OPS_F %d\n#define TMPC_OPS_J %d\n#define TMPC_OPS_H %d\n#define TMPC_OPS_BILIN %d" % (c_f, c_J, c_H, c_B))
    L.append("#endif")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    txt = "\n".join(L) + "\n"
    old = None
    if os.path.exists(out_path):
        with open(out_path) as fh:
            old = fh.read()
    if old != txt:
        with open(out_path, "w") as fh:
            fh.write(txt)
    return {"c_f": c_f, "c_J": c_J, "c_H": c_H, "c_bilin": c_B, "nhess": len(hess)}


def lambdify_ode(model: OdeModel):
    """numpy callables (f, J) straight from sympy -- used by tests to cross-check the generated C."""
    z = list(model.x) + list(model.u)
    f = sp.Matrix(model.xdot)
    Jm = f.jacobian(z)
    ff = sp.lambdify([z], f, "numpy")
    jj = sp.lambdify([z], Jm, "numpy")
    return ff, jj
