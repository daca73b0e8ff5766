"""The reference-side binding shown in INTEGRATION.md is executable: the code block is taken
from the document verbatim, pointed at the in-tree library, and must reproduce the package's
own engine on a BayesRule([Normal, LinearMatrix]) posterior."""
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_stub_runs_and_matches_the_engine():
    import torch

    from hmclab_b200 import Distributions as D
    from hmclab_b200 import _build
    from hmclab_b200._engine import Engine
    from hmclab_b200._lowering import describe, flatten

    with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
        text = f.read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "InterfaceHMCB" in b)
    stub = stub.replace('C.cdll.LoadLibrary("libhmcb.so")', f'C.cdll.LoadLibrary("{_build.LIB_PATH}")')
    ns = {}
    exec(compile(stub, "INTEGRATION.md", "exec"), ns)

    rng = np.random.default_rng(0)
    G, d = rng.normal(size=(60, 24)), rng.normal(size=(60, 1))
    post = D.BayesRule([D.Normal(np.zeros((24, 1)), 1.0), D.LinearMatrix(G, d, 2.0)])
    q0 = rng.normal(size=(7, 24)) * 0.1
    got = ns["run"](post, q0, 0.02, 5, 6)
    assert got.shape == (6, 7, 25) and np.all(np.isfinite(got))

    eng = Engine(flatten(describe(post)), {"kind": "unit", "dims": 24}, 7, integrator="lf",
                 amount_of_steps=5)
    q = torch.as_tensor(q0).cuda().contiguous()
    x = eng.misfit(q)
    ref = torch.zeros(6, 7, 25, dtype=torch.float64, device="cuda")
    eng.run_block(q, x, 6, stepsize=0.02, seed=1, out_samples=ref)
    assert np.array_equal(got, ref.cpu().numpy())
