"""Argument validation of HMC.sample mirrors the reference's (hmclab/Samplers.py:347-426,
1328-1384; tests/test_failed_sampling.py, tests/test_sampling.py:358-387).  Everything here
fails before the device is touched, so it runs without a GPU."""
import numpy as np
import pytest

from hmclab_b200 import Distributions as D
from hmclab_b200 import MassMatrices as M
from hmclab_b200.Samplers import HMC


@pytest.fixture
def target():
    return D.Normal(np.zeros((5, 1)), 1.0)


def _call(tmp_path, target, **kw):
    args = dict(samples_filename=str(tmp_path / "s.npy"), distribution=target, proposals=10,
                disable_progressbar=True)
    args.update(kw)
    return HMC(seed=1).sample(**args)


def test_filename_must_be_string(tmp_path, target):
    with pytest.raises(AssertionError, match="string"):
        _call(tmp_path, target, samples_filename=3)


def test_distribution_type(tmp_path):
    with pytest.raises(AssertionError, match="_AbstractDistribution"):
        _call(tmp_path, object())


@pytest.mark.parametrize("kw", [dict(proposals=0), dict(proposals=2.5), dict(online_thinning=0),
                                dict(proposals=10, online_thinning=3)])
def test_proposals_and_thinning(tmp_path, target, kw):
    with pytest.raises(AssertionError):
        _call(tmp_path, target, **kw)


def test_initial_model_shape(tmp_path, target):
    with pytest.raises(AssertionError, match="incompatible"):
        _call(tmp_path, target, initial_model=np.zeros((4, 1)))
    with pytest.raises(AssertionError, match="chains"):
        _call(tmp_path, target, initial_model=np.zeros((3, 5)), chains=4)


def test_sampler_specific_arguments(tmp_path, target):
    with pytest.raises(AssertionError, match="Stepsize"):
        _call(tmp_path, target, stepsize=-1.0)
    with pytest.raises(AssertionError, match="amount_of_steps"):
        _call(tmp_path, target, amount_of_steps=2.0)
    with pytest.raises(AssertionError, match="larger than zero"):
        _call(tmp_path, target, amount_of_steps=0)
    with pytest.raises(AssertionError, match="dimensions equal"):
        _call(tmp_path, target, mass_matrix=M.Unit(6))
    with pytest.raises(AssertionError, match="_AbstractMassMatrix"):
        _call(tmp_path, target, mass_matrix=np.eye(5))
    with pytest.raises(ValueError, match="Unknown integrator"):
        _call(tmp_path, target, integrator="rk4")
    with pytest.raises(TypeError, match="Unidentified argument"):
        _call(tmp_path, target, temperature=3.0)
    with pytest.raises(AssertionError, match="max_time"):
        _call(tmp_path, target, max_time=-1.0)
    with pytest.raises(AssertionError, match="learning rate"):
        _call(tmp_path, target, autotuning=True, learning_rate=0.4)


def test_existing_file_is_not_overwritten(tmp_path, target):
    path = tmp_path / "s.npy"
    path.write_bytes(b"x")
    with pytest.raises(FileExistsError):
        _call(tmp_path, target)


def test_failed_init_leaves_no_half_open_file(tmp_path, target):
    with pytest.raises(ValueError):
        _call(tmp_path, target, integrator="nope")
    assert not (tmp_path / "s.npy").exists() and not (tmp_path / "s.npy.pkl").exists()


def test_unsupported_distribution_is_named(tmp_path):
    class Himmelblau(D._AbstractDistribution):
        dimensions = 2

    import torch

    if torch.cuda.is_available():
        with pytest.raises(NotImplementedError, match="Himmelblau"):
            _call(tmp_path, Himmelblau())
    else:
        from hmclab_b200._lowering import describe

        with pytest.raises(NotImplementedError, match="Himmelblau"):
            describe(Himmelblau())


def test_without_gpu_the_sampler_raises_instead_of_falling_back(tmp_path, target):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hmclab_b200._engine import HmcbError

    with pytest.raises(HmcbError, match="no CPU fallback"):
        _call(tmp_path, target)
