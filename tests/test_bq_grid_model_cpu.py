"""CPU tests of the algorithm behind the culled ball query (3dssd_b200/csrc/ball_query_grid.cu).

oracle/bq_grid_model.c models the kernel's data path (fp32 grid arithmetic, counting sort into cell records, 3x3 cell
ranges, per-shell index bitmaps read back in index order, the dense-neighbourhood / non-finite-query fallback, unit
lists).  Here its neighbour lists, counts and unit lists are compared with the restatement of the reference
(oracle_query_ball_point / _dilated: /root/reference/lib/utils/tf_ops/grouping/tf_grouping_g.cu:215-255, :308-357) on the
layer shapes of the 3DSSD configuration and on adversarial clouds; tests/test_ops_gpu.py pins the CUDA code itself."""
import importlib

import numpy as np
import pytest

from oracle import ops as oracle_ops

synth = importlib.import_module("3dssd_b200.synth")


def _check(xyz1, xyz2, lows, highs, ks, dilated, expect_dense=None):
    idx, cnt, units, stats = oracle_ops.query_ball_point_grid_model(lows, highs, ks, xyz1, xyz2, dilated)
    b, m = xyz2.shape[:2]
    for s, (lo, hi, k) in enumerate(zip(lows, highs, ks)):
        with np.errstate(invalid="ignore"):
            ei, ec = (oracle_ops.query_ball_point_dilated(lo, hi, k, xyz1, xyz2) if dilated
                      else oracle_ops.query_ball_point(hi, k, xyz1, xyz2))
        np.testing.assert_array_equal(cnt[s], ec, err_msg="pts_cnt of shell %d" % s)
        mask = (ec > 0)[..., None]
        np.testing.assert_array_equal(idx[s] * mask, ei * mask, err_msg="neighbour lists of shell %d" % s)   # rows with cnt == 0
        assert (idx[s][~mask[..., 0]] == 0).all()                       # are undefined in the reference; the product zeroes them
        # unit list: ceil(cnt / 8) units per non-empty group, each exactly once, in any order
        u = units[s]
        nu = int(u[0])
        want = sorted((g << 4) | j for g, c in enumerate(ec.reshape(-1)) for j in range((int(c) + 7) // 8))
        assert nu == len(want) and sorted(u[1:1 + nu].tolist()) == want
    if expect_dense is not None:
        assert (stats[0] > 0) == expect_dense, stats
    return stats


def test_grid_model_layer1_shape_dilated_three_shells():
    """3DSSD layer 1 (3dssd.yaml:47-49): 16384 candidates, D-FPS subset as queries, shells 0.2 / 0.4 / 0.8 x 32 / 32 / 64."""
    pts = synth.kitti_like(2, 16384, seed=5)[..., :3].copy()
    q = oracle_ops.gather_point(pts, oracle_ops.farthest_point_sample(1024, pts))
    stats = _check(pts, q, [0.0, 0.2, 0.4], [0.2, 0.4, 0.8], [32, 32, 64], True, expect_dense=False)
    assert stats[1] < 0.02 * 2 * 1024 * 16384                           # the point of culling: < 2 % of the exhaustive tests


def test_grid_model_vote_centres_plain_query():
    """Layer 4 style (3dssd.yaml:64-66): plain query, the queries are NOT candidates (vote centres), some balls are empty."""
    rng = np.random.default_rng(3)
    pts = synth.kitti_like(2, 4096, seed=8)[..., :3].copy()
    q = (pts[:, :256] + rng.normal(0, 1.5, (2, 256, 3))).astype(np.float32)
    q[:, :8] += 500.0                                                   # far outside the cloud: empty balls, clamped cells
    _check(pts, q, [0.0, 0.0], [4.8, 6.4], [16, 32], False)


@pytest.mark.parametrize("dilated", [False, True])
def test_grid_model_dense_cube_takes_the_index_order_path(dilated):
    """U(0,1)^3 as in the reference's own op test (tf_grouping_op_test.py:11-14): every ball is full after a few hundred
    candidates, the 3x3 neighbourhood holds more than n/8 points -> index-order scan with early exit."""
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 1, (2, 4096, 3)).astype(np.float32)
    q = rng.uniform(0, 1, (2, 300, 3)).astype(np.float32)
    _check(pts, q, [0.0, 0.3], [0.3, 0.6], [16, 64], dilated, expect_dense=True)


@pytest.mark.parametrize("kind", ["line", "point", "duplicates", "boundary", "huge-extent", "tiny-radius"])
def test_grid_model_adversarial_clouds(kind):
    rng = np.random.default_rng(11)
    n, m = 3000, 200
    lows, highs, ks = [0.0, 0.4], [0.4, 0.9], [8, 32]
    if kind == "line":
        pts = np.zeros((1, n, 3), np.float32); pts[..., 2] = np.sort(rng.uniform(0, 300, (1, n)))
    elif kind == "point":
        pts = np.full((1, n, 3), -3.25, np.float32)
    elif kind == "duplicates":
        base = rng.uniform(-20, 20, (1, 150, 3)).astype(np.float32)
        pts = base[:, rng.integers(0, 150, n)]
    elif kind == "boundary":                                            # coordinates on exact multiples of the cell size
        pts = (rng.integers(0, 60, (1, n, 3)) * np.float32(0.9 * 1.01)).astype(np.float32)
    elif kind == "huge-extent":                                         # the cell count hits its cap, cells grow by 1.25x steps
        pts = (rng.uniform(-1, 1, (1, n, 3)) * np.array([3.0e4, 2.0e4, 5.0])).astype(np.float32)
        pts[0, 1000:2000] = pts[0, :1000] + rng.normal(0, 0.3, (1000, 3)).astype(np.float32)
    else:
        pts = rng.uniform(0, 1, (1, n, 3)).astype(np.float32)
        lows, highs = [0.0, 1e-6], [1e-6, 1e-3]
    q = (pts[:, rng.choice(n, m, replace=False)] + (rng.normal(0, 0.2, (1, m, 3)) * (rng.random((1, m, 1)) < 0.5)).astype(np.float32)).astype(np.float32)
    for dilated in (True, False):
        _check(pts, q, lows, highs, ks, dilated)


@pytest.mark.parametrize("where", ["candidate", "query"])
def test_grid_model_non_finite_coordinates_keep_the_reference_behaviour(where):
    """max(sqrt(NaN), 1e-20) < r is TRUE in the plain query (a NaN distance hits) and the dilated predicates are false for
    NaN (tf_grouping_g.cu:237-252, :337-353): a scene holding a non-finite point collapses to one cell, a non-finite query
    scans in index order -- same lists as the reference either way."""
    rng = np.random.default_rng(2)
    pts = rng.uniform(-5, 5, (1, 2500, 3)).astype(np.float32)
    q = pts[:, :64].copy()
    if where == "candidate":
        pts[0, 700, 1] = np.nan
        pts[0, 1900, 0] = np.inf
    else:
        q[0, 5, 2] = np.nan
        q[0, 9, 0] = -np.inf
    for dilated in (True, False):
        _check(pts, q, [0.0, 0.5], [0.5, 1.0], [16, 16], dilated)
