"""CPU oracle: numpy restatement of hmclab's single-chain HMC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``hmclab_b200`` imports this module; it is used
by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s CPU legs as the
checker / CPU baseline, never as the product path.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here
against outputs of the unmodified reference (imported from /root/reference with GUI
stubs, driven with replayed random draws) stored under ``tests/golden/*.npz`` by
``tests/golden/make_golden.py``.

The restatement follows the reference's arithmetic *order* on (d,1) column vectors,
one chain at a time, so that on the same machine it reproduces the reference to the
last bit wherever the reference itself is deterministic.  File:line citations are
relative to the hmclab repository (mounted at /root/reference while building).

Input format: the plain-dict tree of ``hmclab_b200._lowering.describe`` (arrays are
flat float64 vectors); the caller passes it in, this module imports nothing from the
product package.
"""
from __future__ import annotations

import numpy as np

# ------------------------------------------------------------------ distributions ----


def _col(v):
    return np.asarray(v, dtype=np.float64).reshape(-1, 1)


def bounds_penalty(lb, ub, m) -> float:
    """``misfit_bounds``: +inf if any coordinate is outside [lb, ub] (base.py:361-374)."""
    if (lb is not None and np.any(m < _col(lb))) or (
        ub is not None and np.any(m > _col(ub))
    ):
        return np.inf
    return 0.0


def _children_slices(node):
    off = 0
    for child in node["children"]:
        yield child, slice(off, off + child["dims"])
        off += child["dims"]


def misfit(node, m: np.ndarray) -> float:
    """chi(m) for a column vector m of shape (d,1)."""
    kind = node["kind"]
    own = bounds_penalty(node["lb"], node["ub"], m)
    if kind == "normal":
        # base.py:539-550 (diagonal branch)
        r = _col(node["means"]) - m
        return own + 0.5 * (r.T @ (_col(node["inv_cov"]) * r)).flatten()[0] + node["const"]
    if kind == "laplace":
        # base.py:689-700
        return (
            node["const"]
            + own
            + np.sum(np.abs(m - _col(node["means"])) * _col(node["inv_disp"])).flatten()[0]
        )
    if kind == "uniform":
        return own  # base.py:780-782
    if kind == "additive":
        # base.py:1036-1043
        total = 0.0
        for child in node["children"]:
            total += misfit(child, m)
        return total + own
    if kind == "composite":
        # base.py:867-879
        total = 0.0
        for child, sl in _children_slices(node):
            total += misfit(child, m[sl])
        return total + own
    if kind in ("linear_dense", "linear_csr"):
        wrapper = bounds_penalty(node.get("wrapper_lb"), node.get("wrapper_ub"), m)
        if node["premult"]:
            # LinearMatrix.py:185-191 / 390-396
            GtG = _matrix(node, "GtG")
            inner = own + 0.5 * (
                m.T @ (GtG @ m - 2 * _col(node["Gtd0"])) + node["dtd"]
            ).item()
        else:
            # LinearMatrix.py:192-202 / 406-415
            G = _matrix(node, "G")
            if node.get("cov_csc") is not None:
                # sparse G, sparse data covariance (LinearMatrix.py:470-480): LU solve per call
                Gs, solve, dcol = _sparse_cov_parts(node)
                res = Gs @ m - dcol
                inner = own + 0.5 * ((m.T @ Gs.T.tocsr() - dcol.T) @ solve(res)).item()
                return inner + wrapper
            if node.get("chol_upper") is not None:
                # dense data covariance, direct form (LinearMatrix.py:267-279)
                res = node["chol_upper"] @ (G @ m - _col(node["d"]))
            else:
                res = (G @ m - _col(node["d"])) / _col(node["sigma"])
            inner = own + (0.5 * np.linalg.norm(res) ** 2).item()
        return inner + wrapper  # LinearMatrix.py:114-116
    if kind == "srcloc2d":
        # SourceLocation.py:100-109
        x, z, T, v = _split_events_2d(node, m)
        dist = ((x - node["rx"][None, :]) ** 2.0 + (z - node["rz"][None, :]) ** 2.0) ** 0.5
        return own + 0.5 * np.nansum(((node["tobs"] - (T + dist / v)) / node["std"]) ** 2)
    if kind == "srcloc3d":
        # SourceLocation.py:482-493
        x, y, z, T, v = _split_events(node, m)
        dist = (
            (x - node["rx"][None, :]) ** 2.0
            + (y - node["ry"][None, :]) ** 2.0
            + (z - node["rz"][None, :]) ** 2.0
        ) ** 0.5
        return own + 0.5 * np.nansum(((node["tobs"] - (T + dist / v)) / node["std"]) ** 2)
    raise NotImplementedError(kind)


def gradient(node, m: np.ndarray) -> np.ndarray:
    """grad chi(m), shape (d,1)."""
    kind = node["kind"]
    if kind == "normal":
        # base.py:564-570
        return -_col(node["inv_cov"]) * (_col(node["means"]) - m) + bounds_penalty(
            node["lb"], node["ub"], m
        )
    if kind == "laplace":
        # base.py:703-710
        return bounds_penalty(node["lb"], node["ub"], m) + np.sign(
            m - _col(node["means"])
        ) * _col(node["inv_disp"])
    if kind == "uniform":
        return np.zeros((node["dims"], 1)) + bounds_penalty(node["lb"], node["ub"], m)
    if kind == "additive":
        # base.py:1045-1054
        g = np.zeros((node["dims"], 1))
        for child in node["children"]:
            g += gradient(child, m)
        return g + bounds_penalty(node["lb"], node["ub"], m)
    if kind == "composite":
        # base.py:881-898
        parts = [gradient(child, m[sl]) for child, sl in _children_slices(node)]
        return np.vstack(parts) + bounds_penalty(node["lb"], node["ub"], m)
    if kind in ("linear_dense", "linear_csr"):
        # no bounds term in the gradient (LinearMatrix.py:118-120, 204-208, 417-426)
        if node["premult"]:
            return _matrix(node, "GtG") @ m - _col(node["Gtd0"])
        if node.get("cov_csc") is not None:
            # LinearMatrix.py:482-485
            Gs, solve, dcol = _sparse_cov_parts(node)
            return Gs.T.tocsr() @ solve(Gs @ m - dcol)
        G, Gt = _matrix(node, "G"), _matrix(node, "Gt")
        return Gt @ ((G @ m - _col(node["d"])) / _col(node["var"]))
    if kind == "srcloc2d":
        # SourceLocation.py:111-158 (plain sums: no missing picks in the pinned cases)
        x, z, T, v = _split_events_2d(node, m)
        dx = x - node["rx"][None, :]
        dz = z - node["rz"][None, :]
        d = (dx**2.0 + dz**2.0) ** 0.5
        w = ((T + d / v) - node["tobs"]) / (node["std"] ** 2)
        g = np.zeros_like(m)
        stop = 3 * node["events"]
        g[0:stop:3, 0] = np.sum(w * (dx / (v * d)), axis=1)
        g[1:stop:3, 0] = np.sum(w * (dz / (v * d)), axis=1)
        g[2:stop:3, 0] = np.sum(w * np.ones_like(dx), axis=1)
        if node["infer_velocity"]:
            g[-1, 0] = np.sum(w * (-d / (v * v)))
        return g
    if kind == "srcloc3d":
        # SourceLocation.py:495-540
        x, y, z, T, v = _split_events(node, m)
        dx = x - node["rx"][None, :]
        dy = y - node["ry"][None, :]
        dz = z - node["rz"][None, :]
        d = (dx**2.0 + dy**2.0 + dz**2.0) ** 0.5
        t_calc = T + d / v
        w = (t_calc - node["tobs"]) / (node["std"] ** 2)
        g = np.zeros_like(m)
        E = node["events"]
        stop = 4 * E
        g[0:stop:4, 0] = np.nansum(w * (dx / (v * d)), axis=1)
        g[1:stop:4, 0] = np.nansum(w * (dy / (v * d)), axis=1)
        g[2:stop:4, 0] = np.nansum(w * (dz / (v * d)), axis=1)
        g[3:stop:4, 0] = np.nansum(w * np.ones_like(dx), axis=1)
        if node["infer_velocity"]:
            g[-1, 0] = np.nansum(w * (-d / (v * v)))
        return g
    raise NotImplementedError(kind)


def _split_events(node, m):
    E = node["events"]
    stop = 4 * E
    x, y, z, T = (m[i:stop:4] for i in range(4))  # each (E,1); SourceLocation.py:697-713
    v = m[-1] if node["infer_velocity"] else node["velocity"]
    return x, y, z, T, v


def _split_events_2d(node, m):
    stop = 3 * node["events"]
    x, z, T = (m[i:stop:3] for i in range(3))
    v = m[-1] if node["infer_velocity"] else node["velocity"]
    return x, z, T, v


def _cache_of(node):
    """Per-node cache of rebuilt matrices, stored IN the node (a global dict keyed by id(node) would hand
    a recycled id the matrices of a tree that no longer exists)."""
    return node.setdefault("_oracle_cache", {})


def _sparse_cov_parts(node):
    """(G as CSR, solve(rhs) of the factorised covariance, d) of the sparse-covariance LinearMatrix:
    ``scipy.sparse.linalg.factorized`` of the CSC covariance, right-hand side cast to the
    covariance's dtype first (LinearMatrix.py:462-464, 476, 484)."""
    cache, key = _cache_of(node), "sparse_cov"
    if key not in cache:
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla

        N, d = node["N"], node["dims"]
        indptr, indices, data = node["G_csr"]
        Gs = sp.csr_matrix((data, indices, indptr), shape=(N, d))
        cdata, cind, cptr = node["cov_csc"]
        cov = sp.csc_matrix((cdata, cind, cptr), shape=(N, N))
        lu = spla.factorized(cov)
        dtype = cov.dtype
        cache[key] = (Gs, lambda rhs: lu(rhs.astype(dtype)), np.asarray(node["d_stored"]).reshape(N, 1))
    return cache[key]


def _matrix(node, name):
    """Dense ndarray or scipy CSR/CSC rebuilt from the plain arrays of the tree."""
    cache, key = _cache_of(node), name
    if key in cache:
        return cache[key]
    if node["kind"] == "linear_dense":
        if name == "Gt":
            out = node["Gt"] if node.get("Gt") is not None else node["G"].T
        else:
            out = node[name]
    else:
        import scipy.sparse as sp

        d, N = node["dims"], node["N"]
        if name in ("G", "GtG"):
            out = sp.csr_matrix(
                (node["data"], node["indices"], node["indptr"]), shape=(N, d)
            )
        else:
            # the reference holds Gt as the CSC view of G (LinearMatrix.py:359)
            out = sp.csr_matrix(
                (node["t_data"], node["t_indices"], node["t_indptr"]), shape=(d, N)
            ).tocsc()
    cache[key] = out
    return out


def reflection_bounds(tree):
    """(lb, ub) that the top-level object's ``corrector`` reflects on: its own bounds
    (base.py:239-270, 1111-1142); a bound-less composite falls through to its direct
    children's bounds (base.py:946-978)."""
    n = tree["dims"]
    lb, ub = tree["lb"], tree["ub"]
    if "wrapper_lb" in tree:
        lb, ub = tree["wrapper_lb"], tree["wrapper_ub"]
    if tree["kind"] == "composite" and lb is None and ub is None:
        lo, hi = np.full(n, -np.inf), np.full(n, np.inf)
        for child, sl in _children_slices(tree):
            c_lb = child.get("wrapper_lb", child["lb"]) if "wrapper_lb" in child else child["lb"]
            c_ub = child.get("wrapper_ub", child["ub"]) if "wrapper_ub" in child else child["ub"]
            if c_lb is not None:
                lo[sl] = c_lb
            if c_ub is not None:
                hi[sl] = c_ub
        return lo, hi
    return lb, ub


def corrector(lb, ub, q: np.ndarray, p: np.ndarray) -> None:
    """One-shot mirror reflection, in place; the upper test sees the already
    lower-corrected coordinates (base.py:258-270)."""
    if lb is not None:
        lbc = _col(lb)
        low = q < lbc
        q[low] += 2 * (lbc[low] - q[low])
        p[low] *= -1.0
    if ub is not None:
        ubc = _col(ub)
        high = q > ubc
        q[high] += 2 * (ubc[high] - q[high])
        p[high] *= -1.0


# ------------------------------------------------------------------ mass matrices ----


def _cho_solve_lower(chol, b):
    """scipy.linalg.cho_solve((L, lower=True), b), as MassMatrices.Full calls it."""
    from scipy.linalg import cho_solve

    return cho_solve((chol, True), b)


def momentum_from_normal(mass, z: np.ndarray) -> np.ndarray:
    """MassMatrices.py:135-142 (Unit), :220-227 (Diagonal); z is the N(0,1) draw."""
    if mass["kind"] == "unit":
        return z
    if mass["kind"] == "full":
        return mass["cholesky"] @ z  # MassMatrices.py:311-317
    return np.sqrt(_col(mass["diagonal"])) * z


def kinetic_energy(mass, p: np.ndarray) -> float:
    if mass["kind"] == "unit":
        return 0.5 * (p.T @ p).item(0)  # MassMatrices.py:100-114
    if mass["kind"] == "full":
        return 0.5 * np.vdot(p, _cho_solve_lower(mass["cholesky"], p))  # :269-284
    return 0.5 * np.vdot(p, _col(mass["inverse_diagonal"]) * p)  # :185-199


def kinetic_gradient(mass, p: np.ndarray) -> np.ndarray:
    if mass["kind"] == "unit":
        return p  # MassMatrices.py:116-133
    if mass["kind"] == "full":
        return _cho_solve_lower(mass["cholesky"], p)  # :286-309
    return _col(mass["inverse_diagonal"]) * p  # :201-218


# --------------------------------------------------------------------- integrators ----

# Position ("a") / momentum ("b") coefficients in units of the step size.
_A1_3, _B1_3 = 0.11888010966548, 0.29619504261126  # Samplers.py:1666-1669
_A1_4, _A2_4, _B1_4 = 0.071353913450279725904, 0.268548791161230105820, 0.1916678


def stage_schedule(integrator: str, steps: int, eps: float):
    """Flat list of ("a", coeff) / ("b", coeff) sub-steps for one trajectory.

    lf : position-first leapfrog, Samplers.py:1524-1584
    3s : (a1,b1,a2,b2,a2,b1,a1) per step, Samplers.py:1663-1726
    4s : (a1,b1,a2,b2,a3,b2,a2,b1,a1) per step, Samplers.py:1586-1661
    Coefficients are formed exactly like the reference forms them (``0.5 * eps`` first;
    ``a1 *= eps``)."""
    if integrator == "lf":
        half = 0.5 * eps
        seq = [("a", half)]
        for _ in range(steps - 1):
            seq += [("b", eps), ("a", eps)]
        seq += [("b", eps), ("a", half)]
        return seq
    if integrator == "3s":
        a1 = _A1_3
        a2 = 1.0 / 2.0 - a1
        b1 = _B1_3
        b2 = 1.0 - 2.0 * b1
        a1, a2, b1, b2 = a1 * eps, a2 * eps, b1 * eps, b2 * eps
        one = [("a", a1), ("b", b1), ("a", a2), ("b", b2), ("a", a2), ("b", b1), ("a", a1)]
        return one * steps
    if integrator == "4s":
        a1, a2, b1 = _A1_4, _A2_4, _B1_4
        a3 = 1.0 - 2.0 * a1 - 2.0 * a2
        b2 = 1.0 / 2.0 - b1
        a1, a2, a3, b1, b2 = a1 * eps, a2 * eps, a3 * eps, b1 * eps, b2 * eps
        one = [("a", a1), ("b", b1), ("a", a2), ("b", b2), ("a", a3),
               ("b", b2), ("a", a2), ("b", b1), ("a", a1)]
        return one * steps
    raise ValueError(f"Unknown integrator used. Choices are: lf, 3s, 4s (got {integrator})")


def grads_per_step(integrator: str) -> int:
    return {"lf": 1, "3s": 3, "4s": 4}[integrator]


def propagate(tree, mass, integrator, steps, eps, q0, p0, trace=None):
    """One Hamiltonian trajectory; returns (q1, p1).  ``trace`` (a list) receives
    (q_at_gradient, gradient) copies for every gradient evaluation."""
    q, p = q0.copy(), p0.copy()
    lb, ub = reflection_bounds(tree)
    for what, coeff in stage_schedule(integrator, steps, eps):
        if what == "a":
            q += coeff * kinetic_gradient(mass, p)
            corrector(lb, ub, q, p)
        else:
            g = gradient(tree, q)
            if trace is not None:
                trace.append((q.copy(), g.copy()))
            p -= coeff * g
    return q, p


# ------------------------------------------------------------------- sampler loop ----


class ReplayDraws:
    """Random draws of one chain supplied up front, in the order the reference consumes
    them per proposal: normal(size=(d,1)), uniform(0.5,1.5) iff randomize, uniform(0,1)
    (Samplers.py:1463-1469, 1533-1536, 1486)."""

    def __init__(self, z, u_step, u_acc):
        self.z, self.u_step, self.u_acc = z, u_step, u_acc
        self.k = 0

    def normal(self, d):
        return np.array(self.z[self.k], dtype=np.float64).reshape(d, 1)

    def step_factor(self):
        return float(self.u_step[self.k])

    def accept_uniform(self):
        u = float(self.u_acc[self.k])
        self.k += 1
        return u


class GeneratorDraws:
    """Draws from a numpy Generator the way the reference's sampler does."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def normal(self, d):
        return self.rng.normal(size=(d, 1))

    def step_factor(self):
        return self.rng.uniform(0.5, 1.5)

    def accept_uniform(self):
        return self.rng.uniform(0, 1)


def run_chain(tree, mass, *, integrator="lf", steps=10, stepsize=0.1, randomize=True,
              q0=None, proposals=1, draws=None, thinning=1, record_trace=False,
              autotuning=False, target_acceptance_rate=0.65, learning_rate=0.75,
              proposal_offset=0):
    """Single Markov chain, ``proposals`` HMC proposals (Samplers.py:579-587, 675-678,
    1463-1492).  Returns a dict of per-proposal arrays; ``samples`` holds the stored
    (d+1)-rows [model, misfit] after every ``thinning``-th proposal."""
    d = tree["dims"]
    q = np.zeros((d, 1)) if q0 is None else np.array(q0, dtype=np.float64).reshape(d, 1)
    x = misfit(tree, q)
    out = {"accept": [], "H0": [], "H1": [], "samples": [], "q_prop": [], "p_prop": [],
           "trace_q": [], "trace_g": [], "stepsizes": []}
    for k in range(proposals):
        out["stepsizes"].append(stepsize)
        p0 = momentum_from_normal(mass, draws.normal(d))
        eps = draws.step_factor() * stepsize if randomize else stepsize
        trace = [] if record_trace else None
        q1, p1 = propagate(tree, mass, integrator, steps, eps, q, p0, trace)
        # Samplers.py:1471-1492
        x0 = misfit(tree, q)
        h0 = x0 + kinetic_energy(mass, p0)
        x1 = misfit(tree, q1)
        h1 = x1 + kinetic_energy(mass, p1)
        with np.errstate(all="ignore"):
            rate = np.exp(h0 - h1)
        if autotuning:
            # HMC.autotune, Samplers.py:1494-1522 (called before the Metropolis test)
            weight = (proposal_offset + k + 1) ** (-learning_rate)
            r = 0 if np.isnan(rate) else rate
            stepsize -= weight * (target_acceptance_rate - min(r, 1))
            if stepsize <= 0:
                stepsize = max(stepsize, 1e-18)
        accepted = bool(rate > draws.accept_uniform())
        if accepted:
            q, x = q1.copy(), x1
        else:
            x = x0
        out["accept"].append(accepted)
        out["H0"].append(h0)
        out["H1"].append(h1)
        out["q_prop"].append(q1[:, 0].copy())
        out["p_prop"].append(p1[:, 0].copy())
        if record_trace:
            out["trace_q"].append(np.stack([t[0][:, 0] for t in trace]))
            out["trace_g"].append(np.stack([t[1][:, 0] for t in trace]))
        if k % thinning == 0:
            out["samples"].append(np.concatenate([q[:, 0], [x]]))
    res = {k_: np.array(v) for k_, v in out.items() if len(v)}
    res["final_q"] = q[:, 0].copy()
    res["final_x"] = x
    res["final_stepsize"] = stepsize
    return res


def run_chains(tree, mass, *, q0, z, u_step, u_acc, **kw):
    """Batch of independent chains with injected draws.

    q0 [C,d]; z [K,C,d] standard-normal draws; u_step [K,C] in [0.5,1.5); u_acc [K,C].
    Returns arrays stacked proposal-major: accept [K,C], H0/H1 [K,C], q_prop [K,C,d],
    samples [K/thin, C, d+1], trace_q/trace_g [K, n_grad, C, d] when recorded."""
    C = q0.shape[0]
    per_chain = []
    for c in range(C):
        draws = ReplayDraws(z[:, c], u_step[:, c], u_acc[:, c])
        per_chain.append(
            run_chain(tree, mass, q0=q0[c], proposals=z.shape[0], draws=draws, **kw)
        )
    out = {}
    for key in per_chain[0]:
        stacked = np.stack([r[key] for r in per_chain])  # [C, ...]
        if key in ("final_q", "final_x", "final_stepsize"):
            out[key] = stacked
        elif key in ("trace_q", "trace_g"):
            out[key] = np.transpose(stacked, (1, 2, 0, 3))
        else:
            out[key] = np.moveaxis(stacked, 0, 1)
    return out


# ------------------------------------------------------------------------------ RWMH ----


def run_chain_rwmh(tree, *, stepsize=1.0, step_vector=None, q0=None, proposals=1, draws=None,
                   thinning=1, autotuning=False, target_acceptance_rate=0.65, learning_rate=0.75):
    """Random Walk Metropolis-Hastings chain (Samplers.py:1060-1086, autotune :1029-1058).
    ``stepsize`` scalar, ``step_vector`` (d,) per-coordinate factors or None (the reference's
    ``_stepsize_non_scalar_part``)."""
    d = tree["dims"]
    q = np.zeros((d, 1)) if q0 is None else np.array(q0, dtype=np.float64).reshape(d, 1)
    x = misfit(tree, q)
    nonscalar = 1.0 if step_vector is None else _col(step_vector)
    out = {"accept": [], "samples": [], "stepsizes": [], "x_prop": []}
    for k in range(proposals):
        out["stepsizes"].append(stepsize)
        qp = q + stepsize * nonscalar * draws.normal(d)
        xp = misfit(tree, qp)
        with np.errstate(all="ignore"):
            rate = np.exp(x - xp)
        if autotuning:
            weight = (k + 1) ** (-learning_rate)
            r = 0 if np.isnan(rate) else rate
            stepsize -= weight * (target_acceptance_rate - min(r, 1))
            if stepsize <= 0:
                stepsize = max(stepsize, 1e-18)
        accepted = bool(rate > draws.accept_uniform())
        if accepted:
            q, x = qp.copy(), xp
        out["accept"].append(accepted)
        out["x_prop"].append(xp)
        if k % thinning == 0:
            out["samples"].append(np.concatenate([q[:, 0], [x]]))
    res = {k_: np.array(v) for k_, v in out.items()}
    res["final_stepsize"] = stepsize
    return res
