"""The visibility-prior oracle (oracle/visibility_prior_oracle.py) against golden outputs of the UNMODIFIED reference
VisibilityWeightsComputer.compute_weights (tests/golden/visibility_prior.npz, oracle/make_golden_prior.py)."""
import numpy
import pytest

from oracle import visibility_prior_oracle as P
from tests.helpers import load_npz_raw


def case(g, ci):
    dmin, dmax, planes, temp = g[f'c{ci}.params']
    args = [g[f'c{ci}.{k}'] for k in ('frame1', 'frame2', 'extrinsic1', 'extrinsic2', 'intrinsic1', 'intrinsic2')]
    return args, float(dmin), float(dmax), int(planes), float(temp), g[f'c{ci}.weights']


@pytest.mark.parametrize('ci', [0, 1, 2])
def test_weights_match_reference_bit_for_bit(ci):
    args, dmin, dmax, planes, temp, ref = case(load_npz_raw('visibility_prior.npz'), ci)
    got = P.compute_weights(*args, dmin, dmax, planes, temp)
    assert got.dtype == ref.dtype == numpy.float64 and got.shape == ref.shape
    numpy.testing.assert_array_equal(got, ref)


def test_fixture_has_both_mask_values():
    g = load_npz_raw('visibility_prior.npz')
    m = P.visibility_mask(g['c0.weights'])
    assert 0.05 < m.mean() < 0.95
    assert P.visibility_mask(g['c1.weights']).mean() > 0.5 and not P.visibility_mask(g['c2.weights']).any()
