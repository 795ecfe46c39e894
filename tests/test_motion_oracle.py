"""Motion oracle: pinned pieces against the reference's pure-torch files (when mounted) and
property tests for the restatements of absent third-party extensions (lietorch, lietorch_extras,
pytorch3d) — those stay "parity unpinned" (oracle/motion_oracle.py header)."""
import math
import warnings

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation

from oracle import motion_oracle as M
from oracle import ref_loader


def g(seed):
    return torch.Generator().manual_seed(seed)


# ---------------------------------------------------------------- SE3 properties (unpinned math)
def test_se3_exp_log_roundtrip_and_small_angle():
    xi = torch.randn(200, 6, generator=g(0)) * torch.tensor([2., 2., 2., .6, .6, .6])   # |phi| < pi: principal branch
    xi[:20, 3:] *= 1e-4                                   # small-angle branch
    T = M.se3_exp(xi.double())
    back = M.se3_log(T)
    torch.testing.assert_close(back, xi.double(), rtol=1e-8, atol=1e-8)
    assert torch.allclose(T[:, 3:].norm(dim=-1), torch.ones(200, dtype=torch.float64))


def test_se3_act_matches_scipy_and_group_law():
    xi = torch.randn(50, 6, generator=g(1)).double()
    T = M.se3_exp(xi)
    X = torch.randn(50, 3, generator=g(2)).double()
    R = Rotation.from_quat(T[:, 3:].numpy())              # scipy quats are (x, y, z, w) too
    ref = torch.from_numpy(R.apply(X.numpy())) + T[:, :3]
    torch.testing.assert_close(M.se3_act(T, X), ref, rtol=1e-10, atol=1e-10)
    A, B = T[:25], T[25:]
    torch.testing.assert_close(M.se3_act(M.se3_mul(A, B), X[:25]), M.se3_act(A, M.se3_act(B, X[:25])),
                               rtol=1e-10, atol=1e-10)
    # exp of a pure rotation about z by 90 degrees
    T90 = M.se3_exp(torch.tensor([[0., 0., 0., 0., 0., math.pi / 2]], dtype=torch.float64))
    torch.testing.assert_close(M.se3_act(T90, torch.tensor([[1., 0., 0.]], dtype=torch.float64)),
                               torch.tensor([[0., 1., 0.]], dtype=torch.float64), atol=1e-12, rtol=0)


def test_gn_step_recovers_rigid_motion():
    """All pixels move by one rigid transform; with uniform embeddings (affinity 0.5 everywhere) a few
    damped Gauss-Newton steps must drive every per-pixel transform to it."""
    n, h, w = 1, 6, 8
    intr = torch.tensor([[40., 40., 4., 3.]])
    depth = 2.0 + torch.rand(n, h, w, generator=g(3))
    T_true = M.se3_exp(torch.tensor([[0.05, -0.03, 0.02, 0.01, -0.02, 0.015]])).view(1, 1, 1, 7).expand(n, h, w, 7)
    target = M.project(M.se3_act(T_true, M.inv_project(depth, intr)), intr).permute(0, 3, 1, 2)
    Ts = M.se3_identity(n, h, w)
    ae = torch.zeros(n, 32, h, w)
    wgt = torch.ones(n, 3, h, w)
    for _ in range(8):
        Ts = M.gn_step(Ts, ae, target, wgt, depth, intr, radius=32, ep=1e-3)
    err = (M.se3_log(Ts) - M.se3_log(T_true.contiguous())).abs().max().item()
    assert err < 1e-3, err


def test_corr_lookup_integer_coords_reads_volume():
    f1 = torch.randn(1, 16, 6, 8, generator=g(4))
    f2 = torch.randn(1, 16, 6, 8, generator=g(5))
    pyr = M.all_pairs_correlation(f1, f2, 2)
    yy, xx = torch.meshgrid(torch.arange(6.), torch.arange(8.), indexing="ij")
    coords = torch.stack([xx, yy])[None]
    out = M.corr_lookup(pyr, coords, radius=1)            # [1, 2*9, 6, 8]
    # level 0, window centre (i=1, j=1) at integer coords = corr(p, p)
    centre = out[0, 4]
    ref = torch.stack([pyr[0][0, y, x, y, x] for y in range(6) for x in range(8)]).view(6, 8)
    torch.testing.assert_close(centre, ref)
    # i indexes x: (i=2, j=1) is the neighbour to the right
    right = out[0, 2 * 3 + 1]
    assert torch.allclose(right[2, 3], pyr[0][0, 2, 3, 2, 4])


def test_splat_identity_places_points_on_four_pixels():
    n, c, h, w = 1, 2, 5, 6
    intr = torch.tensor([[30., 30., 3., 2.5]])
    depth = torch.full((n, h, w), 3.0)
    feat = torch.randn(n, c, h, w, generator=g(6))
    out, z = M.splat_warp(M.se3_identity(n, h, w), depth, feat, intr, radius=2.0)
    assert out.shape == feat.shape and (z[:, :, 1:, 1:] == 3.0).all()


# ---------------------------------------------------------------- pinned against the reference
ref = pytest.mark.skipif(not ref_loader.available(), reason="reference not mounted")


@ref
def test_projective_and_sampler_ops_vs_reference():
    ns = ref_loader.load()
    assert ns.projective_ops is not None, getattr(ns, "projective_ops_error", None)
    intr = torch.tensor([[50., 55., 7.5, 5.0], [48., 50., 8.0, 5.5]])
    depth = 1.0 + torch.rand(2, 11, 16, generator=g(7)) * 5
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        X_ref = ns.projective_ops.inv_project(depth, intr)
    X = M.inv_project(depth, intr)
    assert torch.equal(X, X_ref)
    assert torch.equal(M.project(X + 0.1, intr), ns.projective_ops.project(X + 0.1, intr))
    coords = torch.rand(2, 11, 16, 2, generator=g(8)) * torch.tensor([17., 12.]) - 1
    d_ref, _ = ns.sampler_ops.depth_sampler(1.0 / depth, coords)
    assert torch.equal(M.depth_sampler(1.0 / depth, coords), d_ref)


@ref
def test_cvx_upsample_and_correlation_vs_reference():
    ns = ref_loader.load()
    assert ns.se3_field is not None, getattr(ns, "se3_field_error", None)
    data = torch.randn(2, 5, 7, 6, generator=g(9))
    mask = torch.randn(2, 576, 5, 7, generator=g(10))
    torch.testing.assert_close(M.cvx_upsample(data, mask), ns.se3_field.cvx_upsample(data, mask), rtol=1e-6, atol=1e-6)
    f1 = torch.randn(2, 32, 8, 16, generator=g(11))
    f2 = torch.randn(2, 32, 8, 16, generator=g(12))
    blk = ns.corr.CorrBlock(f1, f2, num_levels=3, radius=3)
    mine = M.all_pairs_correlation(f1, f2, 3)
    for a, b in zip(mine, blk.corr_pyramid):
        assert torch.equal(a, b)


def test_se3_exp_matches_matrix_exponential():
    """Independent statement of lietorch's SE3.exp (absent here): exp of the 4x4 twist matrix [[phi^, tau], [0, 0]]
    (scipy.linalg.expm) gives the rotation R = exp(phi^) and the translation t = V(phi) tau of the oracle."""
    from scipy.linalg import expm
    xi = (torch.randn(40, 6, generator=g(7)) * torch.tensor([1.5, 1.5, 1.5, .9, .9, .9])).double()
    xi[:5, 3:] *= 1e-5                                     # small-angle series branch
    T = M.se3_exp(xi)
    for i in range(xi.shape[0]):
        tau, phi = xi[i, :3].numpy(), xi[i, 3:].numpy()
        tw = np.zeros((4, 4))
        tw[:3, :3] = np.array([[0, -phi[2], phi[1]], [phi[2], 0, -phi[0]], [-phi[1], phi[0], 0]])
        tw[:3, 3] = tau
        E = expm(tw)
        R = Rotation.from_quat(T[i, 3:].numpy()).as_matrix()
        np.testing.assert_allclose(R, E[:3, :3], atol=1e-9)
        np.testing.assert_allclose(T[i, :3].numpy(), E[:3, 3], atol=1e-9)
