"""How far the Ozaki-sliced tcgen05 path sits inside the 1e-10 parity bar at BASELINE's config-3 size:
the oracle follows three chains (first, middle, last) of the full 8192-chain batch through one 4-stage
proposal (20 gradient evaluations); max relative error of the proposed position / momentum / Hamiltonian
for the native DMMA path and for 5, 6 (default) and 7 kept orders (one process per setting: the engine
reads the environment at set-up).  Test infrastructure: imports oracle/."""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
WORKER = r'''
import json, os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests")); sys.path.insert(0, os.path.join(%r, "tests", "golden"))
import numpy as np, torch
from hmclab_b200 import workloads
from hmclab_b200._engine import Engine
from hmclab_b200._lowering import describe, describe_mass, flatten
from oracle import hmc_oracle as oracle
w = workloads.dense_large()
C, d, K = w.chains, w.dims, 1
rng = np.random.default_rng(0)
z = rng.normal(size=(K, C, d)); us = rng.uniform(0.5, 1.5, size=(K, C)); ua = rng.uniform(size=(K, C))
tree, mtree = describe(w.posterior), describe_mass(w.mass_matrix)
eng = Engine(flatten(tree), mtree, C, integrator=w.integrator, amount_of_steps=w.amount_of_steps)
q = torch.as_tensor(w.initial_models).cuda().contiguous(); x = eng.misfit(q)
out = dict(out_h0=torch.zeros(K, C, dtype=torch.float64, device="cuda"), out_h1=torch.zeros(K, C, dtype=torch.float64, device="cuda"),
           out_q_prop=torch.zeros(K, C, d, dtype=torch.float64, device="cuda"), out_p_prop=torch.zeros(K, C, d, dtype=torch.float64, device="cuda"),
           out_accept=torch.zeros(K, C, dtype=torch.uint8, device="cuda"))
eng.run_block(q, x, K, stepsize=w.stepsize, randomize_stepsize=True, z=torch.as_tensor(z).cuda(),
              u_step=torch.as_tensor(us).cuda(), u_accept=torch.as_tensor(ua).cuda(), **out)
got = {k: v.cpu().numpy() for k, v in out.items()}
sel = np.linspace(0, C - 1, 3).astype(int)
with np.errstate(all="ignore"):
    ref = oracle.run_chains(tree, mtree, q0=w.initial_models[sel], z=z[:, sel], u_step=us[:, sel], u_acc=ua[:, sel],
                            integrator=w.integrator, steps=w.amount_of_steps, stepsize=w.stepsize, randomize=True)
rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
print(json.dumps({"slice_products": eng.tcgen05_slice_pairs,
                  "q_prop": rel(got["out_q_prop"][:, sel], ref["q_prop"]), "p_prop": rel(got["out_p_prop"][:, sel], ref["p_prop"]),
                  "H1": rel(got["out_h1"][:, sel], ref["H1"]),
                  "decisions_equal": bool(np.array_equal(got["out_accept"][:, sel].astype(bool), ref["accept"]))}))
''' % (ROOT, ROOT, ROOT)
res = {}
for name, env in {"DMMA (HMCB_OZAKI=0)": {"HMCB_OZAKI": "0"},
                  "orders 0..4 (5 digits)": {"HMCB_OZAKI_ORDERS": "5"},
                  "orders 0..5 (6 digits, default)": {},
                  "orders 0..6 (7 digits)": {"HMCB_OZAKI_ORDERS": "7"},
                  "G q 6 digits, G^T r 13 modular products (HMCB_OZAKI_CRT=1)": {"HMCB_OZAKI_CRT": "1"}}.items():
    p = subprocess.run([sys.executable, "-c", WORKER], env={**os.environ, **env}, capture_output=True, text=True)
    res[name] = json.loads(p.stdout.strip().splitlines()[-1]) if p.returncode == 0 else {"error": p.stderr[-400:]}
print(json.dumps(res, indent=1))
