"""GPU parity of the visibility-prior generator: vipnerf_visibility_prior through the
VisibilityWeightsComputer mirror against golden outputs of the UNMODIFIED reference class, and against the numpy
oracle on a larger frame pair."""
import numpy
import pytest

from oracle import visibility_prior_oracle as P
from tests.helpers import load_npz_raw
from tests.test_prior_oracle import case

pytestmark = pytest.mark.gpu

# fp64 with the reference's operation order; numpy's stacked small matmuls may fuse multiply-adds
TOL = 1e-9


@pytest.mark.parametrize('ci', [0, 1, 2])
def test_weights_match_reference_golden(ci, built_library):
    from vipnerf_b200.VisibilityPriorFused02 import VisibilityWeightsComputer
    args, dmin, dmax, planes, temp, ref = case(load_npz_raw('visibility_prior.npz'), ci)
    comp = VisibilityWeightsComputer({'num_depth_planes': planes, 'temperature': temp})
    got = comp.compute_weights(*args, dmin, dmax)
    assert got.dtype == numpy.float64 and got.shape == ref.shape
    assert numpy.abs(got - ref).max() <= TOL * max(1.0, numpy.abs(ref).max()), numpy.abs(got - ref).max()
    wd, mask = comp.compute_weights_device(*args, dmin, dmax)
    # the mask rule of start_generation; pixels within TOL of the 0.5 threshold may legitimately differ
    decided = numpy.abs(ref - 0.5) > TOL
    assert numpy.array_equal(mask.cpu().numpy()[decided], (ref > 0.5)[decided])


def test_larger_pair_against_oracle(built_library):
    """A 96 x 128 pair with 64 planes (the reference's plane count) against the numpy oracle."""
    from vipnerf_b200.VisibilityPriorFused02 import VisibilityWeightsComputer
    g = load_npz_raw('visibility_prior.npz')
    rng = numpy.random.default_rng(3)
    h, w = 96, 128
    base = rng.uniform(0, 255, size=(h // 8, w // 8, 3))
    f1 = numpy.round(numpy.kron(base, numpy.ones((8, 8, 1)))).astype('uint8')
    f2 = numpy.roll(f1, 3, axis=1)
    k = numpy.array([[0.9 * w, 0, w / 2], [0, 0.9 * w, h / 2], [0, 0, 1.0]])
    e1, e2 = g['c0.extrinsic1'], g['c0.extrinsic2']
    ref = P.compute_weights(f1, f2, e1, e2, k, k, 1.0, 30.0, 64, 10)
    got = VisibilityWeightsComputer({'num_depth_planes': 64, 'temperature': 10}).compute_weights(f1, f2, e1, e2, k, None, 1.0, 30.0)
    assert numpy.abs(got - ref).max() <= TOL, numpy.abs(got - ref).max()
