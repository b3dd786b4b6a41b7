"""GPU parity on a REAL LiDAR scene: tests/golden/realscan_16384.npz, 16 384 points of the reference's sample KITTI scan
(mayavi/kitti_sample_scan.txt, made by tests/golden/make_realscan.py).  Every comparison here is between two
evaluations of the SAME inputs (kernel vs reference kernel / CPU oracle), so all index results must match bit for bit.
(The file name sorts last on purpose: these checks come after the synthetic-data suite.)"""
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "realscan_16384.npz")


def scenes():
    """[2, 16384, 4]: the scan and its mirror image across the x = 0 plane (a second, distinct scene)."""
    p = np.load(FIXTURE)["points"].astype(np.float32)
    q = p.copy()
    q[:, 0] = -q[:, 0]
    return np.ascontiguousarray(np.stack([p, q], 0))


def test_realscan_fps_layer1_bit_exact(pkg, oracle_ops, cuda):
    pts = scenes()
    xyz = torch.from_numpy(np.ascontiguousarray(pts[..., :3])).to(cuda)
    got = pkg.farthest_point_sample(4096, xyz)
    np.testing.assert_array_equal(got.cpu().numpy(), oracle_ops.farthest_point_sample(4096, pts[..., :3]))
    assert (got[:, 0] == 0).all()


def test_realscan_fps_vs_reference_kernel(pkg, ref_ops, cuda):
    xyz = torch.from_numpy(np.ascontiguousarray(scenes()[..., :3])).to(cuda)
    assert torch.equal(pkg.farthest_point_sample(4096, xyz), ref_ops.farthest_point_sample(4096, xyz))


def test_realscan_ball_query_layer1_bit_exact(pkg, oracle_ops, cuda):
    pts = scenes()[:1]
    xyz = torch.from_numpy(np.ascontiguousarray(pts[..., :3])).to(cuda)
    fidx = pkg.farthest_point_sample(4096, xyz)
    q = pkg.gather_point(xyz, fidx)
    idxs, cnts = pkg.query_ball_point_multi([0.0, 0.2, 0.4], [0.2, 0.4, 0.8], [32, 32, 64], xyz, q, True)
    qn = q.cpu().numpy()
    for i, (lo, hi, k) in enumerate(((0.0, 0.2, 32), (0.2, 0.4, 32), (0.4, 0.8, 64))):
        eidx, ecnt = oracle_ops.query_ball_point_dilated(lo, hi, k, pts[..., :3], qn)
        np.testing.assert_array_equal(cnts[i].cpu().numpy(), ecnt)
        np.testing.assert_array_equal(idxs[i].cpu().numpy(), eidx * (ecnt > 0)[..., None])
        assert int(cnts[i].min()) >= 1                               # queries are input points: d == 0 self hit


def test_realscan_layer1_features_and_ffps_routes(pkg, oracle_ops, cuda):
    """Layer 1 of the 3DSSD table on the real scene against the oracle (indices exact, features within the budget), then
    the layer-2 F-FPS on the GPU-computed features: the matrix-free kernel and the two-op matrix route must pick the
    same 512 points."""
    from oracle import layers as olayers
    pts = scenes()
    arch = [pkg.config.ARCH_3DSSD[0]]
    params = pkg.params.init_params(pkg.config.ARCH_3DSSD, 1, seed=0)
    net = pkg.SABackbone(arch, params, in_channels=1, device=cuda)
    xyz_l, feat_l, fps_l, dbg = net.forward(torch.from_numpy(pts).to(cuda), return_debug=True)
    oxyz, ofeat, ofps, odbg = olayers.backbone_forward(arch, pts, params, return_debug=True)
    np.testing.assert_array_equal(fps_l[1].cpu().numpy(), ofps[1])
    for a, b in zip(dbg[0]["idx"], odbg[0]["idx"]):
        np.testing.assert_array_equal(a.cpu().numpy(), b)
    got, exp = feat_l[1].cpu().numpy(), ofeat[1]
    assert np.abs(got.astype(np.float64) - exp).max() <= 1e-3 * np.abs(exp).max()
    x1, f1 = xyz_l[1].contiguous(), feat_l[1].contiguous()
    direct = pkg.farthest_point_sample_features(512, x1, f1)
    matrix = pkg.farthest_point_sample_with_distance(512, pkg.calc_square_dist(torch.cat([x1, f1], -1).contiguous()))
    assert torch.equal(direct, matrix)
