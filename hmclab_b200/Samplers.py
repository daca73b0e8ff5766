"""Batched Hamiltonian Monte Carlo behind the reference's sampler API.

``HMC().sample(...)`` keeps the signature, argument meaning, validation and error
behaviour of ``hmclab.Samplers.HMC.sample`` (hmclab/Samplers.py:1180-1311, 321-479,
1313-1417) and adds ``chains=`` (and a few batching knobs).  What the reference does in a
per-proposal Python loop (``_sample_loop`` :579-587, ``_propose`` :1463, integrators
:1524-1726, ``_evaluate_acceptance`` :1471) runs on the GPU in blocks of proposals for all
chains at once through the C ABI (``include/hmcb.h``); between blocks the host streams the
stored samples to the reference's ``Samples`` file format and checks ``max_time`` and
Ctrl-C, so an interrupted run leaves a valid file exactly like the reference
(:684-704).

``autotuning=True`` adapts one step size per chain on the device with the reference's
update rule (Samplers.py:1494-1522).  ``diagnostic_mode=True`` prints block-level time shares
at the end of the run (the reference's per-function timers have no counterpart once the calls
are fused into kernels); ``get_diagnostics()`` returns the same numbers.
"""
from __future__ import annotations

import time as _time
from datetime import datetime as _datetime
from typing import Optional

import numpy as _numpy

from hmclab_b200 import parallel as _parallel
from hmclab_b200.Distributions.base import _AbstractDistribution
from hmclab_b200.MassMatrices import Unit as _Unit
from hmclab_b200.MassMatrices import _AbstractMassMatrix
from hmclab_b200.Samples import Samples as _Samples


def _is_distribution(obj) -> bool:
    """Objects of this package, or genuine hmclab distributions (read by attribute name)."""
    if isinstance(obj, _AbstractDistribution):
        return True
    return any(k.__name__ == "_AbstractDistribution" for k in type(obj).__mro__)


def _is_mass_matrix(obj) -> bool:
    if isinstance(obj, _AbstractMassMatrix):
        return True
    return any(k.__name__ == "_AbstractMassMatrix" for k in type(obj).__mro__)


class HMC:
    """Hamiltonian Monte Carlo over a batch of independent chains (hmclab/Samplers.py:1105)."""

    name = "Hamiltonian Monte Carlo"
    available_integrators = ["lf", "3s", "4s"]
    integrators_full_names = {"lf": "leapfrog integrator", "3s": "three stage integrator",
                              "4s": "four stage integrator"}

    def __init__(self, seed=None):
        # Samplers.py:316-319: one Generator per sampler.  It seeds the on-device counter RNG
        # (Philox keyed by seed, global chain id, proposal) and, with ``host_rng=True``,
        # supplies the draws itself in the reference's order.
        self.seed = seed
        self.rng = _numpy.random.default_rng(seed)
        self.samples = None
        self.accepted_proposals = 0
        self.current_proposal = 0
        self.amount_of_writes = 0
        self.engine = None
        self._timings = {}
        self._between_blocks = None

    # ------------------------------------------------------------------------- API ----
    def sample(
        self,
        samples_filename: str,
        distribution,
        stepsize: float = 0.1,
        randomize_stepsize: bool = True,
        amount_of_steps: int = 10,
        mass_matrix=None,
        integrator: str = "lf",
        initial_model: _numpy.ndarray = None,
        proposals: int = 100,
        online_thinning: int = 1,
        diagnostic_mode: bool = False,
        overwrite_existing_file: bool = False,
        max_time: float = None,
        autotuning: bool = False,
        target_acceptance_rate: float = 0.65,
        learning_rate: float = 0.75,
        queue=None,
        disable_progressbar=False,
        *,
        chains: Optional[int] = None,
        block_proposals: Optional[int] = None,
        host_rng: bool = False,
        device: Optional[int] = None,
        distributed: bool = False,
        **kwargs,
    ):
        """Sample ``chains`` independent Markov chains of ``proposals`` proposals each.

        Reference arguments: see hmclab/Samplers.py:1201-1287.  Additional arguments:

        chains: number of chains (default: the number of rows of ``initial_model`` if it is
            2-D ``[chains, dimensions]``, else 1).  All chains share the tuning settings.
        block_proposals: proposals per device launch sequence / host hand-over.
        host_rng: draw momenta, step-size factors and acceptance uniforms from
            ``self.rng`` on the host in the reference's order (normal(d,1), uniform(0.5,1.5),
            uniform(0,1) per proposal; chain c uses ``default_rng(seed + c)`` for c > 0).
            With ``chains=1`` this reproduces a reference run with the same seed to
            rounding.  Default: counter-based RNG on the device.
        distributed: under an initialised ``torch.distributed`` group, shard ``chains``
            over the ranks (one process per GPU); every rank writes ``<name>.rank<r><ext>``.
        """
        self.samples = None
        try:
            self._init_sampler(
                samples_filename=samples_filename, distribution=distribution,
                initial_model=initial_model, proposals=proposals,
                online_thinning=online_thinning,
                overwrite_existing_file=overwrite_existing_file, max_time=max_time,
                disable_progressbar=disable_progressbar, diagnostic_mode=diagnostic_mode,
                chains=chains, block_proposals=block_proposals, host_rng=host_rng,
                device=device, distributed=distributed,
                stepsize=stepsize, randomize_stepsize=randomize_stepsize,
                amount_of_steps=amount_of_steps, mass_matrix=mass_matrix, integrator=integrator,
                autotuning=autotuning, target_acceptance_rate=target_acceptance_rate,
                learning_rate=learning_rate, **kwargs)
        except Exception:
            if self.samples is not None:
                self.samples.close()
            raise
        self._sample_loop()
        if queue is not None:
            queue.put({"0": self._summary()})
        return self

    # ---------------------------------------------------------------------- set-up ----
    def _init_sampler(self, samples_filename, distribution, initial_model, proposals,
                      online_thinning, overwrite_existing_file, max_time, disable_progressbar,
                      diagnostic_mode, chains, block_proposals, host_rng, device, distributed,
                      **kwargs):
        assert type(samples_filename) == str, (
            f"First argument should be a string containing the path of the file to which to "
            f"write samples. It was an object of type {type(samples_filename)}.")
        assert _is_distribution(distribution), (
            "The passed target distribution should be a derived class of _AbstractDistribution.")
        self.distribution = distribution
        assert type(distribution.dimensions) == int and distribution.dimensions > 0, (
            "The passed target distribution should have an integer dimension larger than zero.")
        self.dimensions = d = distribution.dimensions
        assert type(proposals) == int and proposals > 0, (
            "The amount of proposal requested (`proposals`) should be an integer number larger "
            "than zero.")
        self.proposals = proposals
        assert type(online_thinning) == int and online_thinning > 0, (
            "The amount of online thinning (`online_thinning`) should be an integer number "
            "larger than zero.")
        self.online_thinning = online_thinning
        assert proposals % online_thinning == 0, (
            "The amount of proposals (`proposals`) needs to be a multiple of the online "
            "thinning (`online_thinning`) number, to prevent sample wastage.")
        self.proposals_after_thinning = proposals // online_thinning
        # Samplers.py:449-460, 1386-1417 time individual Python calls; here the calls are fused
        # into kernels, so diagnostic_mode reports the block-level shares instead (at close)
        self.diagnostic_mode = bool(diagnostic_mode)

        # initial models: (d,), (d,1) -> one chain (or broadcast to `chains`); [C, d] -> C chains
        if initial_model is None:
            n_chains = 1 if chains is None else int(chains)
            q0 = _numpy.zeros((n_chains, d))
        else:
            arr = _numpy.array(initial_model, dtype=_numpy.float64)
            if arr.ndim == 2 and arr.shape[1] == d and arr.shape != (d, 1):
                q0 = arr
            else:
                assert arr.size == d, (
                    f"The initial model (`initial_model`) dimension is incompatible with the "
                    f"target distribution. Supplied model shape: {arr.shape}."
                    f"Required shape: {(d, 1)}")
                q0 = arr.reshape(1, d)
            n_chains = q0.shape[0] if chains is None else int(chains)
            if q0.shape[0] == 1 and n_chains > 1:
                q0 = _numpy.repeat(q0, n_chains, axis=0)
            assert q0.shape[0] == n_chains, (
                f"`initial_model` holds {q0.shape[0]} chains but chains={n_chains}.")
        assert n_chains > 0, "`chains` should be an integer number larger than zero."
        self.total_chains = n_chains

        # chain sharding (one process per GPU)
        self.rank, self.world_size = _parallel.world() if distributed else (0, 1)
        lo, hi = _parallel.shard_range(n_chains, self.world_size, self.rank)
        self.chain_offset, self.chains = lo, hi - lo
        assert self.chains > 0, "more ranks than chains"
        q0 = _numpy.ascontiguousarray(q0[lo:hi])
        if self.world_size > 1:
            import os

            stem, ext = os.path.splitext(samples_filename)
            samples_filename = f"{stem}.rank{self.rank}{ext}"
        self.samples_filename = samples_filename
        self.samples = _Samples(samples_filename, mode="w", overwrite=overwrite_existing_file)

        if max_time is not None:
            max_time = float(max_time)
            assert max_time > 0.0, (
                "The maximal runtime (`max_time`) should be a float larger than zero.")
        self.max_time = max_time
        self.disable_progressbar = disable_progressbar
        self.host_rng = bool(host_rng)
        # private knobs of ParallelSampleSMP: exact block boundaries for the exchange schedule
        self._first_block = kwargs.pop("_first_block", None)
        exact_blocks = kwargs.pop("_exact_blocks", False)

        self._init_sampler_specific(**kwargs)

        # engine ---------------------------------------------------------------------------
        import torch

        from hmclab_b200._engine import Engine
        from hmclab_b200._lowering import describe, describe_mass, flatten

        self.engine = Engine(flatten(describe(distribution)), describe_mass(self.mass_matrix),
                             self.chains, integrator=self.integrator,
                             amount_of_steps=self.amount_of_steps, device=device)
        self._torch = torch
        self._q = torch.as_tensor(q0, dtype=torch.float64).to(self.engine.device).contiguous()
        self._x = self.engine.misfit(self._q)
        x0 = self._x.cpu().numpy()
        assert not _numpy.isnan(x0).any(), "Initial position in model space gives NaN probability"
        assert not _numpy.isinf(x0).any(), (
            "Initial position in model space gives inf/-inf probability")
        self.current_model = q0.T.copy()
        self.current_x = float(x0[0]) if self.chains == 1 else x0
        self.accepted_proposals = 0
        self.accepted_proposals_per_chain = _numpy.zeros(self.chains, dtype=_numpy.int64)
        self.amount_of_writes = 0
        self.current_proposal = 0

        # block size: whole stored rows, at most ~256 MiB of samples per block
        row_bytes = self.chains * (d + 1) * 8
        rows = max(1, min(self.proposals_after_thinning, (256 << 20) // row_bytes))
        if block_proposals is not None:
            assert type(block_proposals) == int and block_proposals > 0
            rows = max(1, min(self.proposals_after_thinning, block_proposals // online_thinning))
        self.block_proposals = rows * online_thinning
        if exact_blocks:
            self.block_proposals = int(block_proposals)

        self.samples.allocate(self.chains, self.proposals_after_thinning, d)
        self._write_tuning_settings()
        self.samples.write_attribute("proposals", self.proposals)
        self.samples.write_attribute("acceptance_rate", 0)
        self.samples.write_attribute("online_thinning", self.online_thinning)
        self.samples.write_attribute("start_time", _datetime.now().strftime("%d-%b-%Y (%H:%M:%S.%f)"))
        self.samples.write_attribute("sampler", self.name)
        self.samples.write_attribute("chain_offset", self.chain_offset)

    def _init_sampler_specific(self, **kwargs):
        for key in ("stepsize", "randomize_stepsize", "amount_of_steps", "mass_matrix",
                    "integrator", "autotuning", "target_acceptance_rate", "learning_rate"):
            setattr(self, key, kwargs.pop(key))
        if len(kwargs) != 0:
            raise TypeError(f"Unidentified argument(s) not applicable to sampler: {kwargs}")
        if self.autotuning:
            # Samplers.py:1334-1341; every chain adapts its own step size on the device
            assert self.learning_rate > 0.5 and self.learning_rate <= 1.0, (
                f"The learning rate should be larger than 0.5 and smaller than or equal to 1.0, "
                f"otherwise the Markov chain does not converge. Chosen: {self.learning_rate}")
        self.stepsize = float(self.stepsize)
        assert self.stepsize > 0.0, "Stepsize should be a float larger than zero."
        assert type(self.amount_of_steps) == int, (
            "The amount of steps (amount_of_steps) the HMC integrator should make should be an "
            "integer.")
        assert self.amount_of_steps > 0, (
            "The amount of steps (amount_of_steps) the HMC integrator should make should be "
            "larger than zero.")
        if self.mass_matrix is None:
            self.mass_matrix = _Unit(self.dimensions)
        assert _is_mass_matrix(self.mass_matrix), (
            "The passed mass matrix (mass_matrix) should be a class derived from "
            "_AbstractMassMatrix.")
        self.mass_matrix.rng = self.rng
        assert self.mass_matrix.dimensions == self.dimensions, (
            f"The passed mass matrix (mass_matrix) should have dimensions equal to the target "
            f"distribution. Passed: {self.mass_matrix.dimensions}, required: {self.dimensions}.")
        self.integrator = str(self.integrator)
        if self.integrator not in self.available_integrators:
            raise ValueError(f"Unknown integrator used. Choices are: {self.available_integrators}")

    def _write_tuning_settings(self):
        self.samples.write_attribute("stepsize", self.stepsize)
        self.samples.write_attribute("amount_of_steps", self.amount_of_steps)
        self.samples.write_attribute("mass_matrix", self.mass_matrix.name)
        self.samples.write_attribute("integrator", self.integrators_full_names[self.integrator])

    # ------------------------------------------------------------------- main loop ----
    def _host_draws(self, k0, B):
        """Draws for proposals [k0, k0+B) of every local chain, consumed in the reference's
        per-proposal order from one Generator per chain."""
        C, d = self.chains, self.dimensions
        if not hasattr(self, "_chain_rngs"):
            rngs = []
            for c in range(C):
                g = self.chain_offset + c
                if g == 0:
                    rngs.append(self.rng)
                else:
                    base = self.seed if isinstance(self.seed, (int, _numpy.integer)) else None
                    rngs.append(_numpy.random.default_rng(None if base is None else base + g))
            self._chain_rngs = rngs
        z = _numpy.empty((B, C, d))
        us = _numpy.ones((B, C))
        ua = _numpy.empty((B, C))
        for c, rng in enumerate(self._chain_rngs):
            for k in range(B):
                z[k, c] = rng.normal(size=(d, 1))[:, 0]
                if self.randomize_stepsize:
                    us[k, c] = rng.uniform(0.5, 1.5)
                ua[k, c] = rng.uniform(0, 1)
        return z, us, ua

    def _sample_loop(self):
        torch, eng = self._torch, self.engine
        C, d, thin = self.chains, self.dimensions, self.online_thinning
        dev = eng.device
        rows_max = self.block_proposals // thin + 1
        dbuf = [torch.empty(rows_max, C, d + 1, dtype=torch.float64, device=dev) for _ in range(2)]
        hbuf = [torch.empty(rows_max, C, d + 1, dtype=torch.float64).pin_memory() for _ in range(2)]
        accepted = torch.zeros(C, dtype=torch.int32, device=dev)
        copy_stream = torch.cuda.Stream(device=dev)
        tune = {}
        history = None
        self._history = None
        if self.autotuning:
            self._stepsize_chain = torch.full((C,), self.stepsize, dtype=torch.float64, device=dev)
            tune = dict(stepsize_chain=self._stepsize_chain, autotune=True,
                        target_acceptance_rate=self.target_acceptance_rate,
                        learning_rate=self.learning_rate)
            # per-proposal histories like the reference's (Samplers.py:1340-1341), kept on the
            # host; skipped when they would be larger than ~32 MB
            if self.proposals * C <= 4_000_000:
                history = {"stepsizes": [], "h0": [], "h1": []}
                self._history = history
        copied = [None, None]      # event: D2H of slot finished
        pending = None             # (slot, rows) waiting to be written to disk
        # the device RNG key: one draw from the sampler's Generator (reproducible per seed)
        device_seed = int(self.rng.integers(0, 2**63 - 1)) if not self.host_rng else 0

        self.start_time = _datetime.now()
        t_start = _time.time()
        done, nblock = 0, 0
        self._timings = {"device_blocks_s": 0.0, "host_write_s": 0.0, "blocks": 0}

        def drain(item):
            slot, rows = item
            copied[slot].synchronize()
            t0 = _time.time()
            if rows:
                self.samples.write_block(hbuf[slot][:rows].numpy())
            self.amount_of_writes += rows
            self._timings["host_write_s"] += _time.time() - t0

        try:
            while done < self.proposals:
                B = min(self.block_proposals, self.proposals - done)
                if nblock == 0 and self._first_block:
                    B = min(int(self._first_block), B)
                if self.max_time is not None:
                    # the reference checks max_time after every proposal (Samplers.py:700-704);
                    # here it is checked between device blocks, so blocks are sized from the
                    # measured rate to overshoot by at most ~5 % of max_time
                    if done == 0:
                        B = min(B, 4 * thin)
                    else:
                        per_proposal = max((_time.time() - t_start) / done, 1e-9)
                        left = max(self.max_time - (_time.time() - t_start), 0.0)
                        fit = int(min(0.05 * self.max_time, left) / per_proposal)
                        B = max(1, min(B, fit))
                rows = eng.stored_rows(B, thin, done)
                slot = nblock & 1
                draws = {}
                if self.host_rng:
                    z, us, ua = self._host_draws(done, B)
                    draws = dict(z=torch.as_tensor(z).to(dev), u_step=torch.as_tensor(us).to(dev),
                                 u_accept=torch.as_tensor(ua).to(dev))
                if copied[slot] is not None:
                    torch.cuda.current_stream(dev).wait_event(copied[slot])
                t0 = _time.time()
                self._run_block(self._q, self._x, B, stepsize=self.stepsize,
                              randomize_stepsize=self.randomize_stepsize, thinning=thin,
                              proposal_offset=done, chain_offset=self.chain_offset, seed=device_seed,
                              out_samples=dbuf[slot][:rows] if rows else None, accepted_total=accepted,
                              **draws,
                              **tune, **self._history_buffers(history, B))
                if history is not None:
                    for key in history:
                        history[key].append(self._hist[key][:B].cpu().numpy())
                if self._between_blocks is not None:   # e.g. replica exchange: permute chain states
                    self._between_blocks(done + B, dbuf[slot], rows)
                produced = torch.cuda.Event()
                produced.record(torch.cuda.current_stream(dev))
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(produced)
                    if rows:
                        hbuf[slot][:rows].copy_(dbuf[slot][:rows], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                copied[slot] = ev
                # while the GPU works on this block, put the previous one on disk
                if pending is not None:
                    drain(pending)
                pending = (slot, rows)
                produced.synchronize()
                self._timings["device_blocks_s"] += _time.time() - t0
                self._timings["blocks"] += 1
                done += B
                nblock += 1
                self.current_proposal = done - 1
                if self.max_time is not None and _time.time() - t_start > self.max_time:
                    raise TimeoutError
        except KeyboardInterrupt:
            pass  # Samplers.py:688-699: keep what was sampled, close the file cleanly
        except TimeoutError:
            pass
        finally:
            if pending is not None:
                drain(pending)
            torch.cuda.synchronize(dev)
            self.end_time = _datetime.now()
            self._close_sampler(accepted)

    def _run_block(self, q, x, B, **kw):
        self.engine.run_block(q, x, B, **kw)

    def _history_buffers(self, history, B):
        if history is None:
            return {}
        torch, dev = self._torch, self.engine.device
        if (not hasattr(self, "_hist") or self._hist["h0"].shape[0] < B
                or self._hist["h0"].shape[1] != self.chains or self._hist["h0"].device != dev):
            self._hist = {k: torch.empty(self.block_proposals, self.chains, dtype=torch.float64, device=dev)
                          for k in ("stepsizes", "h0", "h1")}
        return dict(out_stepsize=self._hist["stepsizes"][:B], out_h0=self._hist["h0"][:B],
                    out_h1=self._hist["h1"][:B])

    def _close_sampler(self, accepted):
        per_chain = accepted.cpu().numpy().astype(_numpy.int64)
        if self.autotuning:
            final = self._stepsize_chain.cpu().numpy()
            self.stepsize = float(final[0]) if self.chains == 1 else final
            self.samples.write_attribute("final_stepsizes", final)
            hist = getattr(self, "_history", None)
            if hist is not None and len(hist["stepsizes"]):
                self.stepsizes = _numpy.concatenate(hist["stepsizes"])          # [proposals, chains]
                with _numpy.errstate(all="ignore"):
                    self.acceptance_rates = _numpy.exp(_numpy.concatenate(hist["h0"])
                                                       - _numpy.concatenate(hist["h1"]))
                self.samples.write_attribute("acceptance_rates", self.acceptance_rates)
                self.samples.write_attribute("stepsizes", self.stepsizes)
        x_local = self._x
        if self.world_size > 1:
            acc_all, x_all = _parallel.gather_diagnostics(accepted, self._x, self.total_chains)
            self.accepted_proposals_all_chains = acc_all.cpu().numpy().astype(_numpy.int64)
            self.current_x_all_chains = x_all.cpu().numpy()
        self.accepted_proposals_per_chain = per_chain
        self.accepted_proposals = int(per_chain.sum())
        q = self._q.cpu().numpy()
        x = x_local.cpu().numpy()
        self.current_model = q.T.copy()
        self.current_x = float(x[0]) if self.chains == 1 else x
        n_done = self.current_proposal + 1
        self.samples.write_attribute("acceptance_rate",
                                     self.accepted_proposals / max(1, n_done * self.chains))
        self.samples.write_attribute("end_time", self.end_time.strftime("%d-%b-%Y (%H:%M:%S.%f)"))
        self.samples.write_attribute("runtime", str(self.end_time - self.start_time))
        self.samples.write_attribute("runtime_seconds",
                                     (self.end_time - self.start_time).total_seconds())
        self.samples.close()
        if getattr(self, "diagnostic_mode", False):
            total = max((self.end_time - self.start_time).total_seconds(), 1e-12)
            t = self._timings
            print("Detailed statistics:")
            print(f"Total runtime: {total:.2f} seconds")
            print("{:<34} {:<20}".format("Component", "percentage of time"))
            for label, key in (("device blocks (kernels, waited)", "device_blocks_s"),
                               ("host sample writes", "host_write_s")):
                print("{:<34} {:<20.2f}".format(label, 100 * t.get(key, 0.0) / total))
            print("{:<34} {:<20}".format("kernel launches", self.engine.launch_count))
            print("{:<34} {:<20}".format("engine path", self.engine.path))

    # ------------------------------------------------------------------ reporting ----
    def load_results(self, burn_in: int = 0) -> _numpy.ndarray:
        """Samples.py reader on the file just written (Samplers.py:766-774)."""
        with _Samples(self.samples_filename, burn_in=burn_in) as s:
            return _numpy.array(s.numpy)

    def get_diagnostics(self):
        return dict(self._timings, launches=self.engine.launch_count if self.engine else 0,
                    path=self.engine.path if self.engine else None)

    def _summary(self):
        return {"proposals": self.proposals, "chains": self.chains,
                "acceptance_rate": self.accepted_proposals / max(1, (self.current_proposal + 1) * self.chains),
                "path": self.engine.path if self.engine else None}


class RWMH(HMC):
    """Random Walk Metropolis-Hastings over a batch of chains (hmclab/Samplers.py:777-1102):
    ``proposed = current + stepsize * normal``, accepted iff ``exp(x - x') > u``.  ``stepsize`` is
    a positive float or a ``(dimensions, 1)`` array of per-coordinate steps; autotuning adapts a
    scalar step per chain (with an array step the array becomes the fixed per-coordinate factor
    and the adapted scalar starts at 1.0, as in the reference :935-944)."""

    name = "Random Walk Metropolis Hastings"

    def sample(self, samples_filename: str, distribution, stepsize=1.0, initial_model=None,
               proposals: int = 100, online_thinning: int = 1, diagnostic_mode: bool = False,
               overwrite_existing_file: bool = False, max_time: float = None,
               autotuning: bool = False, target_acceptance_rate: float = 0.65,
               learning_rate: float = 0.75, queue=None, disable_progressbar=False, *,
               chains: Optional[int] = None, block_proposals: Optional[int] = None,
               host_rng: bool = False, device: Optional[int] = None, distributed: bool = False,
               **kwargs):
        self.samples = None
        try:
            self._init_sampler(
                samples_filename=samples_filename, distribution=distribution,
                initial_model=initial_model, proposals=proposals, online_thinning=online_thinning,
                overwrite_existing_file=overwrite_existing_file, max_time=max_time,
                disable_progressbar=disable_progressbar, diagnostic_mode=diagnostic_mode,
                chains=chains, block_proposals=block_proposals, host_rng=host_rng, device=device,
                distributed=distributed, stepsize=stepsize, autotuning=autotuning,
                target_acceptance_rate=target_acceptance_rate, learning_rate=learning_rate, **kwargs)
        except Exception:
            if self.samples is not None:
                self.samples.close()
            raise
        self._sample_loop()
        if queue is not None:
            queue.put({"0": self._summary()})
        return self

    def _init_sampler_specific(self, **kwargs):
        for key in ("stepsize", "autotuning", "target_acceptance_rate", "learning_rate"):
            setattr(self, key, kwargs.pop(key))
        self._step_vector = None
        if self.autotuning:
            assert self.learning_rate > 0.5 and self.learning_rate <= 1.0, (
                f"The learning rate should be larger than 0.5 and smaller than or equal to 1.0, "
                f"otherwise the Markov chain does not converge. Chosen: {self.learning_rate}")
            if type(self.stepsize) == _numpy.ndarray:
                self._step_vector = self.stepsize
                self.stepsize = 1.0
            assert type(self.stepsize) == float, (
                "Autotuning RWMH is only implemented for scalar stepsizes. If you need it for "
                "non-scalar steps, write us an email.")
        if len(kwargs) != 0:
            raise TypeError(f"Unidentified argument(s) not applicable to sampler: {kwargs}")
        try:
            self.stepsize = float(self.stepsize)
            assert self.stepsize > 0.0, (
                "RW-MH step length should be a positive float or a numpy.ndarray. The passed "
                "argument is a float equal to or smaller than zero.")
        except TypeError:
            vec = _numpy.asarray(self.stepsize)
            assert vec.shape == (self.dimensions, 1), (
                "RW-MH step length should be a numpy.ndarray of shape (dimensions, 1) or a "
                "positive float. The passed argument is an ndarray of the wrong shape.")
            self._step_vector = vec
            self.stepsize = 1.0     # proposed = current + stepsize_array * 1.0 * normal (:1064-1069)
        if self._step_vector is not None:
            assert _numpy.asarray(self._step_vector).shape == (self.dimensions, 1)
        # the engine object also carries the HMC settings; they are not used by RWMH
        self.mass_matrix = _Unit(self.dimensions)
        self.integrator, self.amount_of_steps, self.randomize_stepsize = "lf", 1, False

    def _write_tuning_settings(self):
        self.samples.write_attribute(
            "stepsize", self.stepsize if self._step_vector is None else "ndarray")

    def _host_draws(self, k0, B):
        """normal(d,1) then uniform(0,1) per proposal (Samplers.py:1064-1081)."""
        C, d = self.chains, self.dimensions
        if not hasattr(self, "_chain_rngs"):
            base = self.seed if isinstance(self.seed, (int, _numpy.integer)) else None
            self._chain_rngs = [
                self.rng if self.chain_offset + c == 0
                else _numpy.random.default_rng(None if base is None else base + self.chain_offset + c)
                for c in range(C)]
        z, ua = _numpy.empty((B, C, d)), _numpy.empty((B, C))
        for c, rng in enumerate(self._chain_rngs):
            for k in range(B):
                z[k, c] = rng.normal(size=(d, 1))[:, 0]
                ua[k, c] = rng.uniform(0, 1)
        return z, _numpy.ones((B, C)), ua

    def _run_block(self, q, x, B, **kw):
        kw.pop("randomize_stepsize", None)
        kw.pop("u_step", None)
        if self._step_vector is not None and not hasattr(self, "_step_vector_dev"):
            self._step_vector_dev = self._torch.as_tensor(
                _numpy.ascontiguousarray(self._step_vector, dtype=_numpy.float64).reshape(-1)
            ).to(self.engine.device)
        self.engine.run_block_rwmh(q, x, B, step_vector=getattr(self, "_step_vector_dev", None), **kw)


class ParallelSampleSMP:
    """Front end with the reference's multi-chain API (hmclab/Samplers.py:1807-1998).

    The reference starts one OS process per chain; here the chains of one call become ONE
    batch on the GPU and every chain still gets its own samples file.  All chains must share
    the posterior object and the tuning ``kwargs`` (that is what a batch is).

    ``exchange=True``: the reference swaps the models of scheduled chain pairs when
    ``exp((x_i(m_i) - x_i(m_j)) + (x_j(m_j) - x_j(m_i))) > u`` (Samplers.py:589-669).  With one
    shared posterior the exponent is exactly zero, so every scheduled swap is accepted: the
    exchange is a permutation of chain states after every ``exchange_interval``-th proposal,
    following a schedule drawn like the reference's (``rng.choice`` without replacement per
    exchange round, :1872-1880).

    Chains with DIFFERENT posteriors (tempering ladders, competing models): the chains are grouped
    by posterior object, every group is one batch on its own engine, and an exchange round
    evaluates each scheduled chain's posterior at its partner's model on the device
    (``hmcb_misfit``), applies the reference's acceptance test per pair and swaps the models
    (``_sample_grouped``).  Under ``torchrun`` (an initialised ``torch.distributed`` group, one process
    per GPU) the chains are sharded over the ranks and the exchange round is the one collective of the
    path (``parallel.exchange_round``: NCCL all-gather of models and partner misfits); every rank
    writes the files of its own chains, and the result does not depend on the number of GPUs.
    """

    def __init__(self, seed=None):
        self.rng = _numpy.random.default_rng(seed)

    def sample(self, samplers, filenames, posteriors, overwrite_existing_files=False,
               proposals: int = 100, exchange: bool = True, exchange_interval: int = 1,
               initial_model=None, kwargs=None):
        import os
        import tempfile

        assert overwrite_existing_files, (
            "You have to manually enable overwriting samples. This is for safety. The existing "
            "file dialog doesn't work in the parallel case. Set `overwrite_existing_files=True`.")
        n = len(samplers)
        assert len(filenames) == n, (
            f"The number of supplied initial models ({len(filenames)}) is not equal to the amount "
            f"of chains ({n}). Supply {n} models.")
        assert len(posteriors) == n, (
            f"The number of supplied initial models ({len(posteriors)}) is not equal to the amount "
            f"of chains ({n}). Supply {n} posteriors.")
        if type(initial_model) == list:
            assert len(initial_model) == n, (
                f"The number of supplied initial models ({len(initial_model)}) is not equal to the "
                f"amount of chains ({n}). Supply either 1 or {n} models.")
        if type(kwargs) == list:
            assert len(kwargs) == n, (
                f"The number of supplied kwargs dictionaries ({len(kwargs)}) is not equal to the "
                f"amount of chains ({n}). Supply either 1 or {n} kwargs.")
            if any(k != kwargs[0] for k in kwargs[1:]):
                raise NotImplementedError("A batch of chains shares one set of tuning kwargs.")
            kwargs = kwargs[0]
        kwargs = dict(kwargs or {})
        if not all(isinstance(smp, HMC) for smp in samplers):
            raise NotImplementedError("Only HMC samplers run on the batched engine.")
        if any(p is not posteriors[0] for p in posteriors[1:]) or kwargs.pop("_grouped", False):
            return self._sample_grouped(samplers, filenames, posteriors, proposals, exchange,
                                        exchange_interval, initial_model, kwargs)
        if not all(isinstance(smp, HMC) for smp in samplers):
            raise NotImplementedError("Only HMC samplers run on the batched engine.")
        kwargs.pop("overwrite_existing_file", None)
        d = int(posteriors[0].dimensions)
        if initial_model is None:
            q0 = _numpy.zeros((n, d))
        elif type(initial_model) == list:
            q0 = _numpy.stack([_numpy.asarray(m, dtype=_numpy.float64).reshape(d) for m in initial_model])
        else:
            q0 = _numpy.repeat(_numpy.asarray(initial_model, dtype=_numpy.float64).reshape(1, d), n, axis=0)

        driver = samplers[0]
        self.samplers = list(samplers)
        self.exchange_schedule = None
        if exchange:
            assert type(exchange_interval) == int and exchange_interval > 0
            pairs = n // 2
            rounds = proposals // exchange_interval
            self.exchange_schedule = (
                _numpy.vstack([self.rng.choice(n, pairs * 2, replace=False) for _ in range(rounds)])
                if pairs and rounds else _numpy.zeros((0, 0), dtype=int))
            kwargs["block_proposals"] = exchange_interval
            kwargs["_exact_blocks"] = True

            thinning = int(kwargs.get("online_thinning", 1))

            def swap(done, block_rows, rows, schedule=self.exchange_schedule):
                # proposal k = done - 1 just finished; the reference exchanges when k % interval == 0
                # and stores the sample of proposal k after the exchange (Samplers.py:589-678)
                k = done - 1
                if schedule.size == 0 or k % exchange_interval != 0:
                    return
                row = k // exchange_interval
                if row >= schedule.shape[0]:
                    return
                perm = _numpy.arange(n)
                for a, b in schedule[row].reshape(-1, 2):
                    perm[a], perm[b] = b, a
                idx = driver._torch.as_tensor(perm, device=driver.engine.device)
                driver._q.copy_(driver._q[idx])
                driver._x.copy_(driver._x[idx])
                if rows and k % thinning == 0:
                    # the stored row of proposal k shows the exchanged state; unlike the reference,
                    # whose misfit column keeps the value of the model that left the chain, model
                    # and misfit stay a consistent pair here
                    block_rows[rows - 1] = block_rows[rows - 1][idx]

            driver._between_blocks = swap
            if exchange_interval != 1:
                # blocks must end right after proposals k with k % interval == 0: k = 0, I, 2I, ...
                # -> first block of one proposal, then blocks of `interval`; done by the sampler
                # when block boundaries are requested explicitly
                kwargs["_first_block"] = 1
        try:
            with tempfile.TemporaryDirectory() as tmp:
                combined = os.path.join(tmp, "batch.npy")
                driver.sample(combined, posteriors[0], proposals=proposals, initial_model=q0,
                              chains=n, overwrite_existing_file=True, **kwargs)
                self._split(driver, combined, filenames)
        finally:
            driver._between_blocks = None
        return self

    def _sample_grouped(self, samplers, filenames, posteriors, proposals, exchange, exchange_interval,
                        initial_model, kwargs):
        """Chains whose posteriors differ: one engine per distinct posterior object, replica exchange
        between the groups exactly as the reference defines it (Samplers.py:589-669):

            after proposal k with k % exchange_interval == 0, for every scheduled pair (a, b):
            improvement_i = chi_i(m_i) - chi_i(m_partner);  swap iff exp(improvement_a +
            improvement_b) > u, u drawn by the pair's master (the odd position of the schedule row).

        The row stored for proposal k shows the exchanged model next to the misfit the chain held
        before the exchange -- what the reference writes, since it refreshes ``current_x`` only at the
        next proposal; the chain itself continues with chi_i(m_partner).  Settings in ``kwargs`` are
        shared by all chains; ``host_rng=True`` draws every chain's momenta and uniforms from its own
        sampler's Generator in the reference's order (and then reproduces a reference run)."""
        import torch

        from hmclab_b200._engine import Engine
        from hmclab_b200._lowering import describe, describe_mass, flatten

        n = len(samplers)
        known = {"stepsize", "amount_of_steps", "integrator", "randomize_stepsize", "online_thinning",
                 "mass_matrix", "disable_progressbar", "host_rng", "device", "diagnostic_mode"}
        unknown = set(kwargs) - known
        if unknown:
            raise TypeError(f"ParallelSampleSMP with several posteriors does not take {sorted(unknown)}")
        stepsize = float(kwargs.get("stepsize", 0.1))
        steps = int(kwargs.get("amount_of_steps", 10))
        integrator = kwargs.get("integrator", "lf")
        randomize = bool(kwargs.get("randomize_stepsize", True))
        thin = int(kwargs.get("online_thinning", 1))
        host_rng = bool(kwargs.get("host_rng", False))
        assert stepsize > 0 and steps > 0 and thin > 0 and proposals % thin == 0
        d = int(posteriors[0].dimensions)
        assert all(int(p.dimensions) == d for p in posteriors), "all posteriors need the same dimensions"
        mass = kwargs.get("mass_matrix")
        if mass is None:
            from hmclab_b200 import MassMatrices as _M

            mass = _M.Unit(d)
        if initial_model is None:
            q0 = _numpy.zeros((n, d))
        elif type(initial_model) == list:
            q0 = _numpy.stack([_numpy.asarray(m, dtype=_numpy.float64).reshape(d) for m in initial_model])
        else:
            q0 = _numpy.repeat(_numpy.asarray(initial_model, dtype=_numpy.float64).reshape(1, d), n, axis=0)
        dev_index = kwargs.get("device")
        dev = torch.device("cuda", torch.cuda.current_device() if dev_index is None else int(dev_index))

        # groups of chains sharing a posterior object: one engine each.  The chains are laid out group
        # after group ("positions"); a rank owns a contiguous range of positions (all of them without a
        # process group), so a group's local part is a contiguous slice of it and the device random
        # streams, keyed by position, do not depend on the number of GPUs.
        all_groups = []
        for i, post in enumerate(posteriors):
            for g in all_groups:
                if g["posterior"] is post:
                    g["chains"].append(i)
                    break
            else:
                all_groups.append({"posterior": post, "chains": [i]})
        order = [i for g in all_groups for i in g["chains"]]          # position -> chain
        position = _numpy.empty(n, dtype=_numpy.int64)                # chain -> position
        position[order] = _numpy.arange(n)
        rank, world = _parallel.world()
        lo, hi = _parallel.shard_range(n, world, rank)
        if world > 1:
            import hashlib
            import torch.distributed as dist

            mine = hashlib.sha256(repr(self.rng.bit_generator.state).encode()).hexdigest()
            states = [None] * world
            dist.all_gather_object(states, mine)
            if any(st != mine for st in states):
                raise ValueError("ParallelSampleSMP(seed=...) must be seeded identically on every rank: the "
                                 "exchange schedule and the device random streams derive from it")
        groups, start = [], 0
        for g in all_groups:
            a, b = max(lo, start), min(hi, start + len(g["chains"]))
            if a < b:
                groups.append({"posterior": g["posterior"], "chains": g["chains"][a - start: b - start],
                               "offset": a})
            start += len(g["chains"])
        for g in groups:
            ids = g["chains"]
            g["engine"] = Engine(flatten(describe(g["posterior"])), describe_mass(mass), len(ids),
                                 integrator=integrator, amount_of_steps=steps, device=dev.index)
            g["q"] = torch.as_tensor(q0[ids], dtype=torch.float64).to(dev).contiguous()
            g["x"] = g["engine"].misfit(g["q"])
            g["accepted"] = torch.zeros(len(ids), dtype=torch.int32, device=dev)
            assert bool(torch.isfinite(g["x"]).all()), "The initial model has a non-finite misfit."
        local_chains = [i for g in groups for i in g["chains"]]       # = order[lo:hi]

        self.samplers = list(samplers)
        self.exchange_schedule = None
        if exchange:
            assert type(exchange_interval) == int and exchange_interval > 0
            pairs, rounds = n // 2, proposals // exchange_interval
            self.exchange_schedule = (
                _numpy.vstack([self.rng.choice(n, pairs * 2, replace=False) for _ in range(rounds)])
                if pairs and rounds else _numpy.zeros((0, 0), dtype=int))
        device_seed = int(self.rng.integers(0, 2**63 - 1))
        rows_host = {i: [] for i in local_chains}       # stored rows per chain
        self.exchanges_accepted = 0

        def misfit_at(models):       # chain i's own posterior at the model in row i, on the device
            out, r = [], 0
            for g in groups:
                C = len(g["chains"])
                out.append(g["engine"].misfit(models[r: r + C].contiguous()))
                r += C
            return torch.cat(out) if out else models.new_zeros(0)

        done = 0
        while done < proposals:
            if exchange:   # blocks end right after the proposals k with k % interval == 0
                nxt = done if done % exchange_interval == 0 else (done // exchange_interval + 1) * exchange_interval
                B = min(nxt - done + 1, proposals - done)
            else:
                B = min(256, proposals - done)
            bufs = []
            for g in groups:
                eng, C = g["engine"], len(g["chains"])
                rows = eng.stored_rows(B, thin, done)
                buf = torch.empty(rows, C, d + 1, dtype=torch.float64, device=dev) if rows else None
                draws = {}
                if host_rng:
                    z, us, ua = _numpy.empty((B, C, d)), _numpy.ones((B, C)), _numpy.empty((B, C))
                    for l, i in enumerate(g["chains"]):
                        rng = samplers[i].rng
                        for k in range(B):
                            z[k, l] = rng.normal(size=(d, 1))[:, 0]
                            if randomize:
                                us[k, l] = rng.uniform(0.5, 1.5)
                            ua[k, l] = rng.uniform(0, 1)
                    draws = dict(z=torch.as_tensor(z).to(dev), u_step=torch.as_tensor(us).to(dev),
                                 u_accept=torch.as_tensor(ua).to(dev))
                eng.run_block(g["q"], g["x"], B, stepsize=stepsize, randomize_stepsize=randomize, thinning=thin,
                              proposal_offset=done, chain_offset=g["offset"], seed=device_seed, out_samples=buf,
                              accepted_total=g["accepted"], **draws)
                bufs.append(buf)
            k = done + B - 1
            if exchange and k % exchange_interval == 0 and k // exchange_interval < self.exchange_schedule.shape[0]:
                row = self.exchange_schedule[k // exchange_interval].reshape(-1, 2)
                # a: even entry of the schedule row, b: odd entry = the pair's master, which draws the
                # uniform -- from its sampler's generator with host_rng (the owner rank), else from the
                # front end's generator (identical on every rank, drawn for all pairs up front)
                shared_u = None if host_rng else self.rng.uniform(0, 1, size=len(row))
                if host_rng:
                    draw = lambda p, a, b: samplers[order[b]].rng.uniform(0, 1)
                else:
                    draw = lambda p, a, b: shared_u[p]
                q_loc = torch.cat([g["q"] for g in groups]) if groups else torch.zeros(0, d, dtype=torch.float64, device=dev)
                x_loc = torch.cat([g["x"] for g in groups]) if groups else torch.zeros(0, dtype=torch.float64, device=dev)
                accepted, q_before = _parallel.exchange_round(position[row], n, q_loc, x_loc, misfit_at, draw)
                self.exchanges_accepted += int(accepted.sum())
                r = 0
                for g, buf in zip(groups, bufs):
                    C = len(g["chains"])
                    g["q"].copy_(q_loc[r: r + C])
                    g["x"].copy_(x_loc[r: r + C])
                    r += C
                if k % thin == 0:      # the stored row: exchanged model, misfit from before the exchange
                    where = {i: (gi, l) for gi, g in enumerate(groups) for l, i in enumerate(g["chains"])}
                    for (a, b), ok in zip(row, accepted):
                        for mine, other in ((a, b), (b, a)):
                            if ok and int(mine) in where:
                                gi, l = where[int(mine)]
                                bufs[gi][-1, l, :d] = q_before[position[other]]
            for g, buf in zip(groups, bufs):
                if buf is not None:
                    host = buf.cpu().numpy()
                    for l, i in enumerate(g["chains"]):
                        rows_host[i].append(host[:, l, :])
            done += B
        torch.cuda.synchronize(dev)
        for g in groups:
            acc = g["accepted"].cpu().numpy()
            qf, xf = g["q"].cpu().numpy(), g["x"].cpu().numpy()
            for l, i in enumerate(g["chains"]):
                smp = samplers[i]
                smp.accepted_proposals = int(acc[l])
                smp.current_proposal = proposals - 1
                smp.current_model = qf[l][:, None].copy()
                smp.current_x = float(xf[l])
            g["engine"].close()
        for i in local_chains:          # every rank writes the files of the chains it owns
            name = filenames[i]
            rows = _numpy.concatenate(rows_host[i]) if rows_host[i] else _numpy.zeros((0, d + 1))
            out = _Samples(name, mode="w", overwrite=True)
            out.allocate(1, rows.shape[0], d)
            if rows.shape[0]:
                out.write_block(_numpy.ascontiguousarray(rows[:, None, :]))
            out.write_attribute("proposals", proposals)
            out.write_attribute("online_thinning", thin)
            out.write_attribute("sampler", "Hamiltonian Monte Carlo")
            out.write_attribute("stepsize", stepsize)
            out.write_attribute("amount_of_steps", steps)
            out.write_attribute("integrator", integrator)
            out.write_attribute("acceptance_rate", samplers[i].accepted_proposals / max(1, proposals))
            out.write_attribute("chain_index", i)
            out.close()
        return self

    @staticmethod
    def _split(driver, combined, filenames):
        """One reference-format samples file per chain."""
        with _Samples(combined) as src:
            attrs = dict(src._attributes)
            per = int(attrs["samples_per_chain"])
            done = driver.current_proposal + 1
            for c, name in enumerate(filenames):
                out = _Samples(name, mode="w", overwrite=True)
                rows = _numpy.ascontiguousarray(src.chain(c).T)           # [per, d+1]
                out.allocate(1, per, rows.shape[1] - 1)
                out.write_block(rows[:, None, :])
                for key, value in attrs.items():
                    if key not in ("write_index", "last_written_sample", "chains", "samples_per_chain",
                                   "acceptance_rate", "final_stepsizes", "stepsizes", "acceptance_rates"):
                        out.write_attribute(key, value)
                out.write_attribute("acceptance_rate",
                                    float(driver.accepted_proposals_per_chain[c]) / max(1, done))
                out.write_attribute("chain_index", c)
                out.close()
