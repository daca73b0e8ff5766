"""ctypes binding of ``libhmcb.so`` (the C ABI in ``include/hmcb.h``).

PyTorch tensors are the batch container: every device buffer handed to the engine is a
``torch.float64`` CUDA tensor whose ``data_ptr()`` crosses the boundary, and work is
enqueued on torch's current stream.  There is no fallback: if the library has not been
built, or no B200 is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Dict, Optional

import numpy as np

from hmclab_b200 import _build

INTEGRATORS = {"lf": 0, "3s": 1, "4s": 2}
GRADS_PER_STEP = {"lf": 1, "3s": 3, "4s": 4}
PATH_NAMES = {0: "fused_priors", 1: "fused_srcloc", 2: "staged", 3: "fused_dense"}

_c_double_p = C.POINTER(C.c_double)
_c_int32_p = C.POINTER(C.c_int32)


class HmcbError(RuntimeError):
    pass


class _Block(C.Structure):
    _fields_ = [
        ("proposals", C.c_int64), ("thinning", C.c_int64), ("proposal_offset", C.c_int64),
        ("chain_offset", C.c_int64), ("seed", C.c_uint64), ("stepsize", C.c_double),
        ("randomize_stepsize", C.c_int32), ("reserved", C.c_int32),
        ("q", C.c_void_p), ("x", C.c_void_p),
        ("z_in", C.c_void_p), ("u_step_in", C.c_void_p), ("u_accept_in", C.c_void_p),
        ("out_samples", C.c_void_p), ("out_accept", C.c_void_p), ("out_h0", C.c_void_p),
        ("out_h1", C.c_void_p), ("accepted_total", C.c_void_p),
        ("out_q_prop", C.c_void_p), ("out_p_prop", C.c_void_p),
        ("trace_q", C.c_void_p), ("trace_g", C.c_void_p),
        ("stepsize_chain", C.c_void_p), ("autotune", C.c_int32), ("reserved2", C.c_int32),
        ("target_acceptance_rate", C.c_double), ("learning_rate", C.c_double),
        ("out_stepsize", C.c_void_p),
    ]


_LIB = None

# name -> (restype, argtypes); every symbol declared in include/hmcb.h
SIGNATURES = {
    "hmcb_abi_version": (C.c_int, []),
    "hmcb_debug_spmm_tables": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, _c_int32_p, _c_int32_p, _c_double_p,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                         _c_double_p, _c_double_p, C.POINTER(C.c_int64)]),
    "hmcb_debug_spmm_block_tables": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, _c_int32_p, _c_int32_p, _c_double_p,
                                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                               _c_double_p, _c_double_p, C.POINTER(C.c_int64)]),
    "hmcb_last_error": (C.c_char_p, []),
    "hmcb_create": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]),
    "hmcb_destroy": (C.c_int, [C.c_void_p]),
    "hmcb_set_integrator": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "hmcb_set_exact_arithmetic": (C.c_int, [C.c_void_p, C.c_int]),
    "hmcb_set_mass_unit": (C.c_int, [C.c_void_p]),
    "hmcb_set_mass_diagonal": (C.c_int, [C.c_void_p, _c_double_p, _c_double_p]),
    "hmcb_set_mass_full": (C.c_int, [C.c_void_p, _c_double_p, _c_double_p]),
    "hmcb_debug_i8_gemm": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmcb_debug_crt_product": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64, _c_double_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "hmcb_debug_i8_gather_gemm": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                            C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "hmcb_debug_oz_slice_rows": (C.c_double, [_c_double_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "hmcb_clear_target": (C.c_int, [C.c_void_p]),
    "hmcb_add_prior": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, _c_double_p,
                                 _c_double_p, C.c_double]),
    "hmcb_add_bound_check": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _c_double_p,
                                       _c_double_p, C.c_int]),
    "hmcb_set_reflection": (C.c_int, [C.c_void_p, _c_double_p, _c_double_p]),
    "hmcb_set_likelihood_dense_premult": (C.c_int, [C.c_void_p, _c_double_p, _c_double_p,
                                                    C.c_double]),
    "hmcb_set_likelihood_dense_direct": (C.c_int, [C.c_void_p, C.c_int64, _c_double_p,
                                                   _c_double_p, _c_double_p, _c_double_p,
                                                   _c_double_p]),
    "hmcb_set_likelihood_dense_direct_cov": (C.c_int, [C.c_void_p, C.c_int64, _c_double_p, _c_double_p, _c_double_p,
                                                       _c_double_p, _c_double_p]),
    "hmcb_set_likelihood_csr_direct": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _c_int32_p,
                                                 _c_int32_p, _c_double_p, _c_int32_p, _c_int32_p,
                                                 _c_double_p, _c_double_p, _c_double_p,
                                                 _c_double_p]),
    "hmcb_set_likelihood_csr_premult": (C.c_int, [C.c_void_p, C.c_int64, _c_int32_p, _c_int32_p,
                                                  _c_double_p, _c_double_p, C.c_double]),
    "hmcb_set_likelihood_srcloc3d": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _c_double_p,
                                               _c_double_p, _c_double_p, _c_double_p,
                                               _c_double_p, C.c_int, C.c_double]),
    "hmcb_set_likelihood_srcloc2d": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _c_double_p,
                                               _c_double_p, _c_double_p, _c_double_p, C.c_int,
                                               C.c_double]),
    "hmcb_finalize": (C.c_int, [C.c_void_p]),
    "hmcb_path": (C.c_int, [C.c_void_p]),
    "hmcb_dense_products_on_tcgen05": (C.c_int, [C.c_void_p]),
    "hmcb_grads_per_proposal": (C.c_int64, [C.c_void_p]),
    "hmcb_launch_count": (C.c_int64, [C.c_void_p]),
    "hmcb_kernel_timing_begin": (C.c_int, [C.c_void_p]),
    "hmcb_kernel_timing_end": (C.c_int, [C.c_void_p, _c_double_p, C.POINTER(C.c_int64)]),
    "hmcb_debug_fp64_peak": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _c_double_p, _c_double_p]),
    "hmcb_misfit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmcb_gradient": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmcb_reflect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmcb_scale_momentum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmcb_kinetic_energy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmcb_kinetic_gradient": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hmcb_run_block": (C.c_int, [C.c_void_p, C.POINTER(_Block), C.c_void_p]),
    "hmcb_run_block_rwmh": (C.c_int, [C.c_void_p, C.POINTER(_Block), C.c_void_p, C.c_void_p]),
    "hmcb_sample_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                   C.c_double, C.c_int, C.c_uint64, C.c_int64, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
}


def load_library(path: Optional[str] = None):
    """dlopen libhmcb.so and declare every prototype.  Raises if it has not been built."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or os.environ.get("HMCB_LIBRARY") or _build.LIB_PATH
    if not os.path.exists(path):
        raise HmcbError(
            f"The CUDA engine library `{path}` is missing. Build it with "
            "`python -m hmclab_b200._build` (needs nvcc); hmclab_b200 has no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = restype, argtypes
    if lib.hmcb_abi_version() != 2:
        raise HmcbError("libhmcb.so ABI version mismatch; rebuild the library")
    _LIB = lib
    return lib


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_c_double_p)


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_c_int32_p)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class Engine:
    """One engine = one target distribution x one mass matrix x ``chains`` chains on one GPU."""

    def __init__(self, plan: Dict[str, Any], mass: Dict[str, Any], chains: int, *,
                 integrator: str = "lf", amount_of_steps: int = 10, device: Optional[int] = None,
                 exact: Optional[bool] = None):
        import torch

        if not torch.cuda.is_available():
            raise HmcbError("hmclab_b200 needs a CUDA device (B200); there is no CPU fallback.")
        self.lib = load_library()
        self.torch = torch
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.chains, self.dims = int(chains), int(plan["dims"])
        if integrator not in INTEGRATORS:
            raise ValueError("Unknown integrator used. Choices are: lf, 3s, 4s")
        if mass["dims"] != self.dims:
            raise ValueError("Mass matrix dimensions do not match the distribution.")
        handle = C.c_void_p()
        self._handle = None
        self._ok(self.lib.hmcb_create(self.device_index, self.chains, self.dims, C.byref(handle)))
        self._handle = handle
        self.integrator, self.amount_of_steps = integrator, int(amount_of_steps)
        self._ok(self.lib.hmcb_set_integrator(handle, INTEGRATORS[integrator], int(amount_of_steps)))
        if mass["kind"] == "unit":
            self._ok(self.lib.hmcb_set_mass_unit(handle))
        elif mass["kind"] == "full":
            self._ok(self.lib.hmcb_set_mass_full(handle, _dp(_f64(mass["cholesky"])), _dp(_f64(mass["inverse"]))))
        else:
            self._ok(self.lib.hmcb_set_mass_diagonal(
                handle, _dp(_f64(mass["diagonal"])), _dp(_f64(mass["inverse_diagonal"]))))
        if exact is not None:
            self._ok(self.lib.hmcb_set_exact_arithmetic(handle, int(bool(exact))))
        self._lower(plan)
        self._ok(self.lib.hmcb_finalize(handle))
        self.path = PATH_NAMES[self.lib.hmcb_path(handle)]

    # ------------------------------------------------------------------ plumbing -----
    def _ok(self, status: int):
        if status != 0:
            raise HmcbError(self.lib.hmcb_last_error().decode())

    def close(self):
        if getattr(self, "_handle", None) is not None:
            self.lib.hmcb_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _lower(self, plan):
        h, lib = self._handle, self.lib
        kinds = {"normal": 0, "laplace": 1}
        for term in plan["terms"]:
            self._ok(lib.hmcb_add_prior(h, kinds[term["kind"]], term["offset"], term["len"],
                                        _dp(_f64(term["a"])), _dp(_f64(term["b"])),
                                        float(term["const"])))
        for chk in plan["checks"]:
            lb = None if chk["lb"] is None else _f64(chk["lb"])
            ub = None if chk["ub"] is None else _f64(chk["ub"])
            self._ok(lib.hmcb_add_bound_check(h, chk["offset"], chk["len"], _dp(lb), _dp(ub),
                                              int(bool(chk["in_gradient"]))))
        rlb = None if plan["reflect_lb"] is None else _f64(plan["reflect_lb"])
        rub = None if plan["reflect_ub"] is None else _f64(plan["reflect_ub"])
        self._ok(lib.hmcb_set_reflection(h, _dp(rlb), _dp(rub)))
        lik = plan["likelihood"]
        if lik is None:
            return
        kind = lik["kind"]
        if kind == "linear_dense" and lik["premult"]:
            self._ok(lib.hmcb_set_likelihood_dense_premult(
                h, _dp(_f64(lik["GtG"])), _dp(_f64(lik["Gtd0"])), float(lik["dtd"])))
        elif kind == "linear_dense" and lik.get("misfit_G") is not None:
            self._ok(lib.hmcb_set_likelihood_dense_direct_cov(
                h, int(lik["N"]), _dp(_f64(lik["G"])), _dp(_f64(lik["Gt"])), _dp(_f64(lik["d"])),
                _dp(_f64(lik["misfit_G"])), _dp(_f64(lik["misfit_d"]))))
        elif kind == "linear_dense":
            Gt = None if lik.get("Gt") is None else _f64(lik["Gt"])
            self._ok(lib.hmcb_set_likelihood_dense_direct(
                h, int(lik["N"]), _dp(_f64(lik["G"])), _dp(Gt), _dp(_f64(lik["d"])),
                _dp(_f64(lik["var"])), _dp(_f64(lik["sigma"]))))
        elif kind == "linear_csr" and lik["premult"]:
            self._ok(lib.hmcb_set_likelihood_csr_premult(
                h, int(lik["data"].size), _ip(lik["indptr"]), _ip(lik["indices"]),
                _dp(_f64(lik["data"])), _dp(_f64(lik["Gtd0"])), float(lik["dtd"])))
        elif kind == "linear_csr":
            self._ok(lib.hmcb_set_likelihood_csr_direct(
                h, int(lik["N"]), int(lik["data"].size), _ip(lik["indptr"]), _ip(lik["indices"]),
                _dp(_f64(lik["data"])), _ip(lik["t_indptr"]), _ip(lik["t_indices"]),
                _dp(_f64(lik["t_data"])), _dp(_f64(lik["d"])), _dp(_f64(lik["var"])),
                _dp(_f64(lik["sigma"]))))
        elif kind == "srcloc3d":
            v = float(lik["velocity"]) if not lik["infer_velocity"] else 0.0
            self._ok(lib.hmcb_set_likelihood_srcloc3d(
                h, int(lik["events"]), int(lik["stations"]), _dp(_f64(lik["rx"])),
                _dp(_f64(lik["ry"])), _dp(_f64(lik["rz"])), _dp(_f64(lik["tobs"])),
                _dp(_f64(lik["std"])), int(bool(lik["infer_velocity"])), v))
        elif kind == "srcloc2d":
            v = float(lik["velocity"]) if not lik["infer_velocity"] else 0.0
            self._ok(lib.hmcb_set_likelihood_srcloc2d(
                h, int(lik["events"]), int(lik["stations"]), _dp(_f64(lik["rx"])),
                _dp(_f64(lik["rz"])), _dp(_f64(lik["tobs"])), _dp(_f64(lik["std"])),
                int(bool(lik["infer_velocity"])), v))
        else:
            raise NotImplementedError(kind)

    # ---------------------------------------------------------------- properties -----
    @property
    def grads_per_proposal(self) -> int:
        return int(self.lib.hmcb_grads_per_proposal(self._handle))

    @property
    def tcgen05_slice_pairs(self) -> int:
        """0, or the int8 slice products per gradient evaluation when the dense products run on tcgen05."""
        return int(self.lib.hmcb_dense_products_on_tcgen05(self._handle))

    @property
    def launch_count(self) -> int:
        return int(self.lib.hmcb_launch_count(self._handle))

    def kernel_timing_begin(self):
        self._ok(self.lib.hmcb_kernel_timing_begin(self._handle))

    def kernel_timing_end(self):
        """[(total_ms, passes)] for the gradient passes and the misfit passes since _begin."""
        ms, n = (C.c_double * 2)(), (C.c_int64 * 2)()
        self._ok(self.lib.hmcb_kernel_timing_end(self._handle, ms, n))
        return [(float(ms[i]), int(n[i])) for i in range(2)]

    def set_exact_arithmetic(self, on: bool):
        """True: the priors-only fused kernel rounds multiply and add separately, as numpy does."""
        self._ok(self.lib.hmcb_set_exact_arithmetic(self._handle, int(bool(on))))

    def set_integrator(self, integrator: str, amount_of_steps: int):
        if integrator not in INTEGRATORS:
            raise ValueError("Unknown integrator used. Choices are: lf, 3s, 4s")
        self._ok(self.lib.hmcb_set_integrator(self._handle, INTEGRATORS[integrator],
                                              int(amount_of_steps)))
        self.integrator, self.amount_of_steps = integrator, int(amount_of_steps)

    # --------------------------------------------------------------- evaluation ------
    def _check_batch(self, t, name):
        torch = self.torch
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64
                and t.is_contiguous() and tuple(t.shape) == (self.chains, self.dims)
                and t.device.index == self.device_index):
            raise ValueError(f"{name}: expected a contiguous float64 CUDA tensor of shape "
                             f"{(self.chains, self.dims)} on cuda:{self.device_index}")

    def _new(self, *shape, dtype=None):
        return self.torch.empty(*shape, dtype=dtype or self.torch.float64, device=self.device)

    def misfit(self, q):
        self._check_batch(q, "q")
        x = self._new(self.chains)
        self._ok(self.lib.hmcb_misfit(self._handle, q.data_ptr(), x.data_ptr(), self._stream()))
        return x

    def gradient(self, q):
        self._check_batch(q, "q")
        g = self._new(self.chains, self.dims)
        self._ok(self.lib.hmcb_gradient(self._handle, q.data_ptr(), g.data_ptr(), self._stream()))
        return g

    def reflect_(self, q, p):
        self._check_batch(q, "q")
        self._check_batch(p, "p")
        self._ok(self.lib.hmcb_reflect(self._handle, q.data_ptr(), p.data_ptr(), self._stream()))

    def scale_momentum(self, z):
        self._check_batch(z, "z")
        p = self._new(self.chains, self.dims)
        self._ok(self.lib.hmcb_scale_momentum(self._handle, z.data_ptr(), p.data_ptr(), self._stream()))
        return p

    def kinetic_energy(self, p):
        self._check_batch(p, "p")
        k = self._new(self.chains)
        self._ok(self.lib.hmcb_kinetic_energy(self._handle, p.data_ptr(), k.data_ptr(), self._stream()))
        return k

    def kinetic_gradient(self, p):
        self._check_batch(p, "p")
        g = self._new(self.chains, self.dims)
        self._ok(self.lib.hmcb_kinetic_gradient(self._handle, p.data_ptr(), g.data_ptr(), self._stream()))
        return g

    # ------------------------------------------------------------------ sampling -----
    def stored_rows(self, proposals: int, thinning: int, proposal_offset: int = 0) -> int:
        first = -(-proposal_offset // thinning)
        last = -(-(proposal_offset + proposals) // thinning)
        return last - first

    def run_block(self, q, x, proposals: int, *, stepsize: float, randomize_stepsize: bool = True,
                  thinning: int = 1, proposal_offset: int = 0, chain_offset: int = 0, seed: int = 0,
                  z=None, u_step=None, u_accept=None, out_samples=None, out_accept=None,
                  out_h0=None, out_h1=None, accepted_total=None, out_q_prop=None, out_p_prop=None,
                  trace_q=None, trace_g=None, stepsize_chain=None, autotune=False,
                  target_acceptance_rate=0.65, learning_rate=0.75, out_stepsize=None,
                  rwmh=False, step_vector=None):
        """Advance every chain by ``proposals`` proposals in place (q [C,d], x [C])."""
        torch = self.torch
        self._check_batch(q, "q")
        B, Cn, d = int(proposals), self.chains, self.dims
        G = self.grads_per_proposal

        def ptr(t, shape, dtype, name):
            if t is None:
                return None
            if not (t.is_cuda and t.dtype == dtype and t.is_contiguous() and tuple(t.shape) == shape):
                raise ValueError(f"{name}: expected contiguous {dtype} CUDA tensor of shape {shape}, "
                                 f"got {t.dtype} {tuple(t.shape)}")
            return t.data_ptr()

        f64 = torch.float64
        rows = self.stored_rows(B, thinning, proposal_offset)
        blk = _Block(
            proposals=B, thinning=int(thinning), proposal_offset=int(proposal_offset),
            chain_offset=int(chain_offset), seed=int(seed) & (2**64 - 1), stepsize=float(stepsize),
            randomize_stepsize=int(bool(randomize_stepsize)), reserved=0,
            q=q.data_ptr(), x=ptr(x, (Cn,), f64, "x"),
            z_in=ptr(z, (B, Cn, d), f64, "z"), u_step_in=ptr(u_step, (B, Cn), f64, "u_step"),
            u_accept_in=ptr(u_accept, (B, Cn), f64, "u_accept"),
            out_samples=ptr(out_samples, (rows, Cn, d + 1), f64, "out_samples"),
            out_accept=ptr(out_accept, (B, Cn), torch.uint8, "out_accept"),
            out_h0=ptr(out_h0, (B, Cn), f64, "out_h0"), out_h1=ptr(out_h1, (B, Cn), f64, "out_h1"),
            accepted_total=ptr(accepted_total, (Cn,), torch.int32, "accepted_total"),
            out_q_prop=ptr(out_q_prop, (B, Cn, d), f64, "out_q_prop"),
            out_p_prop=ptr(out_p_prop, (B, Cn, d), f64, "out_p_prop"),
            trace_q=ptr(trace_q, (B, G, Cn, d), f64, "trace_q"),
            trace_g=ptr(trace_g, (B, G, Cn, d), f64, "trace_g"),
            stepsize_chain=ptr(stepsize_chain, (Cn,), f64, "stepsize_chain"),
            autotune=int(bool(autotune)), reserved2=0,
            target_acceptance_rate=float(target_acceptance_rate), learning_rate=float(learning_rate),
            out_stepsize=ptr(out_stepsize, (B, Cn), f64, "out_stepsize"),
        )
        if rwmh:
            sv = None
            if step_vector is not None:
                sv = ptr(step_vector, (d,), f64, "step_vector")
            self._ok(self.lib.hmcb_run_block_rwmh(self._handle, C.byref(blk), sv, self._stream()))
        else:
            self._ok(self.lib.hmcb_run_block(self._handle, C.byref(blk), self._stream()))

    def run_block_rwmh(self, q, x, proposals: int, *, stepsize: float, step_vector=None, **kw):
        """Random Walk Metropolis-Hastings proposals (hmcb_run_block_rwmh); same arguments as
        run_block, plus ``step_vector`` [dims] per-coordinate step factors."""
        self.run_block(q, x, proposals, stepsize=stepsize, randomize_stepsize=False, rwmh=True,
                       step_vector=step_vector, **kw)

    def sample_host(self, q0, proposals: int, *, stepsize: float, randomize_stepsize: bool = True,
                    thinning: int = 1, block_proposals: int = 0, seed: int = 0, chain_offset: int = 0,
                    samples=None, accepted=None, final_q=None, final_x=None):
        """Whole run with HOST (ideally pinned) torch/numpy buffers; see hmcb_sample_host."""

        def hptr(a):
            if a is None:
                return None
            if hasattr(a, "data_ptr"):
                assert not a.is_cuda and a.is_contiguous()
                return a.data_ptr()
            assert a.flags.c_contiguous
            return a.ctypes.data

        self._ok(self.lib.hmcb_sample_host(
            self._handle, hptr(q0), int(proposals), int(thinning), int(block_proposals),
            float(stepsize), int(bool(randomize_stepsize)), int(seed) & (2**64 - 1),
            int(chain_offset), hptr(samples), hptr(accepted), hptr(final_q), hptr(final_x)))
