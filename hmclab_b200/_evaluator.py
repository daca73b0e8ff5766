"""Device evaluation of the single-vector protocol methods (``misfit(m)``,
``gradient(m)``, ``corrector``, ``kinetic_energy`` ...) on small batches.

The distribution / mass-matrix objects of this package hold parameters only; these
helpers lower them once per batch size, keep the engine on the object and run the same
CUDA kernels the sampler uses.  There is no CPU implementation to fall back to.
"""
from __future__ import annotations

import numpy as np


class _Evaluator:
    def __init__(self, plan, mass):
        self.plan, self.mass = plan, mass
        self.engines = {}

    def _engine(self, chains: int):
        from hmclab_b200._engine import Engine

        eng = self.engines.get(chains)
        if eng is None:
            eng = Engine(self.plan, self.mass, chains)
            self.engines[chains] = eng
        return eng

    def _run(self, name, *arrays, outputs=1):
        import torch

        arrays = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays]
        chains = arrays[0].shape[0]
        eng = self._engine(chains)
        dev = [torch.as_tensor(a).to(eng.device) for a in arrays]
        res = getattr(eng, name)(*dev)
        if res is None:  # in-place
            return tuple(t.cpu().numpy() for t in dev)
        return res.cpu().numpy()

    def misfit_batch(self, q):
        return self._run("misfit", q)

    def gradient_batch(self, q):
        return self._run("gradient", q)

    def corrector_batch(self, q, p):
        return self._run("reflect_", q, p)

    def kinetic_energy_batch(self, p):
        return self._run("kinetic_energy", p)

    def kinetic_gradient_batch(self, p):
        return self._run("kinetic_gradient", p)

    def scale_momentum_batch(self, z):
        return self._run("scale_momentum", z)


def _fingerprint(dist, depth=0):
    """Everything `update_bounds`, `normalize`, `collapse_bounds` or `add_distribution` can change on a
    distribution or any of its children: those methods rebind the attributes, so the OBJECTS are the
    identity (big operators are not hashed).  The cache keeps these references, so a rebound array cannot
    be mistaken for its predecessor through a recycled ``id``."""
    own = [dist, getattr(dist, "lower_bounds", None), getattr(dist, "upper_bounds", None),
           repr(getattr(dist, "normalization_constant", None))]
    children = getattr(dist, "separate_distributions", None) or ()
    if depth <= 16:
        for c in children:
            own.extend(_fingerprint(c, depth + 1))
    return own


def _same(a, b) -> bool:
    return len(a) == len(b) and all(x is y or (isinstance(x, str) and x == y) for x, y in zip(a, b))


def evaluator_for(dist) -> _Evaluator:
    from hmclab_b200._lowering import describe, flatten

    key = _fingerprint(dist)
    cached = dist.__dict__.get("_hmcb_evaluator")
    if cached is None or not _same(cached[0], key):
        plan = flatten(describe(dist))
        cached = (key, _Evaluator(plan, {"kind": "unit", "dims": plan["dims"]}))
        dist.__dict__["_hmcb_evaluator"] = cached
    return cached[1]


def mass_evaluator_for(mass) -> _Evaluator:
    from hmclab_b200._lowering import describe_mass

    ev = mass.__dict__.get("_hmcb_evaluator")
    if ev is None:
        n = int(mass.dimensions)
        plan = {"dims": n, "terms": [], "checks": [], "likelihood": None,
                "reflect_lb": None, "reflect_ub": None}
        ev = _Evaluator(plan, describe_mass(mass))
        mass.__dict__["_hmcb_evaluator"] = ev
    return ev
