"""Stage-wise, teacher-forced parity of the non-matmul CUDA stages against the oracle: every
stage is fed the oracle's own inputs so that errors cannot compound.  All calls go through the C ABI."""
import pytest
import torch

from oracle import vipnerf_oracle as O
from tests.helpers import load_npz, rel_err, to_cuda

pytestmark = pytest.mark.gpu

# fp32 elementwise stages evaluated op-for-op like the reference: differences come only from expf/sincosf
# (<= 2 ulp) and from summation order in the scans/reductions.
STAGE_TOL = 1e-5


@pytest.mark.parametrize('scene', ['fern', 'dtu', 're10k'])
@pytest.mark.parametrize('lindisp', [False, True])
def test_coarse_z_bit_exact(scene, lindisp, built_library):
    from vipnerf_b200 import renderpath
    ndc = O.SCENES[scene]['ndc']
    batch = O.make_rays(scene, 333, seed=4)
    near, far = (batch['near_ndc'], batch['far_ndc']) if ndc else (batch['near'], batch['far'])
    if lindisp and ndc:
        near = near + 0.25   # 1/near must be finite
        batch['near_ndc'] = near
    ref = O.coarse_z_vals(near, far, 64, lindisp=lindisp)
    got = renderpath.coarse_z_vals(to_cuda(batch), ndc=ndc, n_coarse=64, lindisp=lindisp).cpu()
    assert torch.equal(got, ref)


def test_coarse_z_stratified_jitter(built_library):
    from vipnerf_b200 import renderpath
    batch = O.make_rays('dtu', 257, seed=6)
    t_rand = torch.rand(257, 64, generator=torch.Generator().manual_seed(5))
    ref = O.coarse_z_vals(batch['near'], batch['far'], 64, t_rand=t_rand)
    b = to_cuda(batch)
    b['t_rand'] = t_rand.cuda()
    got = renderpath.coarse_z_vals(b, ndc=False, n_coarse=64).cpu()
    assert rel_err(got, ref)[0] <= 1e-6


def _oracle_pass(scene, n_rays, seed, n_sec_views=0, which='coarse'):
    ndc = O.SCENES[scene]['ndc']
    sd = O.synth_state_dict(0)
    batch = O.make_rays(scene, n_rays, seed=seed, n_sec_views=n_sec_views)
    with torch.no_grad():
        out = O.render(sd, batch, ndc=ndc, retraw=True, sec_views_vis=n_sec_views > 0)
    return ndc, batch, out


@pytest.mark.parametrize('scene', ['fern', 'dtu'])
@pytest.mark.parametrize('which', ['coarse', 'fine'])
def test_composite_teacher_forced(scene, which, built_library):
    from vipnerf_b200 import renderpath
    ndc, batch, ref = _oracle_pass(scene, 301, 8, n_sec_views=2)
    got = renderpath.volume_rendering(
        to_cuda(batch), ref[f'z_vals_{which}'].cuda(), ref[f'raw_sigma_{which}'].cuda(), ref[f'raw_rgb_{which}'].cuda(),
        ref[f'raw_visibility2_{which}'].cuda(), ndc=ndc)
    keys = ['rgb', 'acc', 'depth', 'depth_var', 'alpha', 'visibility', 'weights', 'visibility2']
    keys += ['depth_ndc', 'depth_var_ndc'] if ndc else []
    for k in keys:
        assert got[k].shape == ref[f'{k}_{which}'].shape, k
        err = rel_err(got[k], ref[f'{k}_{which}'])[0]
        assert err <= STAGE_TOL, (k, err)


def test_composite_white_background(built_library):
    from vipnerf_b200 import renderpath
    batch = O.make_rays('dtu', 64, seed=9)
    g = torch.Generator().manual_seed(3)
    z = O.coarse_z_vals(batch['near'], batch['far'], 64)
    sigma = torch.rand(64, 64, generator=g) * 2
    rgb = torch.rand(64, 64, 3, generator=g)
    ref = O.composite(sigma, rgb, z, batch['rays_d'], False, white_bkgd=True)
    got = renderpath.volume_rendering(to_cuda(batch), z.cuda(), sigma.cuda(), rgb.cuda(), ndc=False, white_bkgd=True)
    assert rel_err(got['rgb'], ref['rgb'])[0] <= STAGE_TOL


def _check_fine_z(z, ref):
    """Inverse-cdf samples that land in an almost empty bin are divided by denom ~ 1e-5 (VipNeRF01.py:257-259):
    last-ulp differences of the cdf (summation order) are amplified to ~1e-4 of the depth range there - the
    u = 1 sample of most rays is such a sample.  So: all but <= 4 of a ray's 192 depths within 1e-5, none off by
    more than 1e-3 (a whole bin would be 1.6e-2), median at rounding level."""
    d = (z - ref).abs() / ref.abs().max()
    assert ((d > STAGE_TOL).sum(dim=1) <= 4).all(), (d > STAGE_TOL).sum(dim=1).max().item()
    assert d.max().item() <= 1e-3, d.max().item()
    assert d.median().item() <= 1e-6


@pytest.mark.parametrize('scene', ['fern', 'dtu'])
def test_fine_z_teacher_forced(scene, built_library):
    """get_z_vals_fine from the oracle's coarse sigma/rgb: sorted, same multiset of coarse depths, and equal to
    the oracle's 192 depths.  The inverse-cdf has genuine discontinuities (denom < 1e-5 clamp, SURVEY 8a note 5),
    so the max is taken over all but the worst 0.1 % of the samples and the worst sample is bounded separately."""
    from vipnerf_b200 import renderpath
    ndc, batch, ref = _oracle_pass(scene, 512, 10)
    got = renderpath.volume_rendering(to_cuda(batch), ref['z_vals_coarse'].cuda(), ref['raw_sigma_coarse'].cuda(),
                                      ref['raw_rgb_coarse'].cuda(), ndc=ndc, n_fine=128)
    z = got['z_vals_fine'].cpu()
    assert z.shape == (512, 192)
    assert (z[:, 1:] >= z[:, :-1]).all()
    _check_fine_z(z, ref['z_vals_fine'])


def test_sample_pdf_edge_cases(built_library):
    """Degenerate pdfs of the reference fixture (all-zero weights, one-hot, plateaus, 1e-9 weights) through the
    CUDA re-sampler.  The kernel re-samples from composited weights, so the fixture's weight patterns are turned
    into densities; both sides then composite the same sigma and invert the resulting cdf."""
    from vipnerf_b200 import renderpath
    w = load_npz('stage_sample_pdf.npz')['weights']            # [48,62]
    batch = O.make_rays('dtu', 48, seed=12)
    z = O.coarse_z_vals(batch['near'], batch['far'], 64)
    sigma = torch.cat([torch.zeros(48, 1), w, torch.zeros(48, 1)], -1) * 3.0
    rgb = torch.zeros(48, 64, 3)
    comp = O.composite(sigma, rgb, z, batch['rays_d'], False)
    ref = O.fine_z_vals(z, comp['weights'], 128)
    got = renderpath.volume_rendering(to_cuda(batch), z.cuda(), sigma.cuda(), rgb.cuda(), ndc=False, n_fine=128)
    _check_fine_z(got['z_vals_fine'].cpu(), ref)


def test_fine_z_random_u(built_library):
    """Training-mode re-sampling (random u, unsorted): bitonic sort + merge."""
    from vipnerf_b200 import renderpath
    ndc, batch, ref = _oracle_pass('dtu', 130, 13)
    u = torch.rand(130, 128, generator=torch.Generator().manual_seed(2))
    want = O.fine_z_vals(ref['z_vals_coarse'], ref['weights_coarse'], 128, u=u)
    got = renderpath.volume_rendering(to_cuda(batch), ref['z_vals_coarse'].cuda(), ref['raw_sigma_coarse'].cuda(),
                                      ref['raw_rgb_coarse'].cuda(), ndc=ndc, n_fine=128, u_rand=u.cuda())
    z = got['z_vals_fine'].cpu()
    assert (z[:, 1:] >= z[:, :-1]).all()
    _check_fine_z(z, want)


def test_empty_and_ragged_batches(built_library):
    from vipnerf_b200 import renderpath
    for n in (0, 1, 3, 5):
        batch = O.make_rays('dtu', max(n, 1), seed=14)
        batch = {k: v[:n] for k, v in batch.items()}
        z = renderpath.coarse_z_vals(to_cuda(batch), ndc=False)
        assert z.shape == (n, 64)
        if n:
            assert torch.equal(z.cpu(), O.coarse_z_vals(batch['near'], batch['far'], 64))
