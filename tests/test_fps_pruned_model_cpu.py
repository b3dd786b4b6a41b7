"""CPU tests of the exactness claim behind the pruned D-FPS kernel (3dssd_b200/csrc/fps_bucket.cu).

oracle/fps_pruned_model.c models that kernel lane by lane (Morton buckets, boxes, the one-sided skip rule, the
(value, key-with-place) arg-max, the per-warp cache, the resume state).  Here the model is compared index for index
with the restatement of the reference (oracle_farthest_point_sample: /root/reference/lib/utils/tf_ops/sampling/
tf_sampling_g.cu:124-178) on every kind of scene the GPU parity test uses and on adversarial ones -- so the algorithm is
pinned without a GPU; tests/test_ops_gpu.py::test_fps_bucket_kernel_bit_exact pins the CUDA code itself."""
import importlib

import numpy as np
import pytest

from oracle import ops as oracle_ops

synth = importlib.import_module("3dssd_b200.synth")


def _scenes(kind, b, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "kitti":
        pts = synth.kitti_like(b, n, seed=seed)[..., :3].copy()
        if n >= 100:
            pts[:, n // 2: n // 2 + 40] = pts[:, 3:43]                   # exact duplicates (the loader pads with them)
    elif kind == "uniform":
        pts = rng.uniform(-40, 40, (b, n, 3)).astype(np.float32)
    elif kind == "lattice":                                              # many exactly equal distances: the tie-break decides
        pts = rng.integers(0, 12, (b, n, 3)).astype(np.float32) * 0.5
    elif kind == "line":                                                 # zero extent on two axes
        pts = np.zeros((b, n, 3), np.float32); pts[..., 1] = rng.uniform(0, 100, (b, n))
    elif kind == "point":                                                # one location only
        pts = np.full((b, n, 3), 2.5, np.float32)
    elif kind == "far":                                                  # large offsets: coarse fp32 grid, heavy cancellation
        pts = (rng.uniform(-1, 1, (b, n, 3)) + np.array([9000.0, -7000.0, 300.0])).astype(np.float32)
    elif kind == "tiny":                                                 # squared distances far below the 1e-30 guard
        pts = (rng.integers(0, 50, (b, n, 3)) * 1e-17).astype(np.float32)
    elif kind == "denormal":                                             # squared distances are subnormal fp32 numbers
        pts = (rng.integers(0, 40, (b, n, 3)) * 1e-21).astype(np.float32)
    elif kind == "clusters":                                             # tight clumps far apart: boxes straddle gaps
        c = rng.uniform(-60, 60, (b, 24, 3))
        pts = (c[:, rng.integers(0, 24, n)] + rng.normal(0, 0.02, (b, n, 3))).astype(np.float32)
    elif kind == "plane_ties":                                           # integer grid in a plane: exact ties everywhere
        g = np.stack(np.meshgrid(np.arange(128), np.arange(128), indexing="ij"), -1).reshape(-1, 2)
        pts = np.zeros((b, n, 3), np.float32)
        for s in range(b):
            pts[s, :, :2] = g[rng.permutation(len(g))[:n]]
    else:
        raise ValueError(kind)
    return pts


CASES = [("kitti", 16384, 4096), ("kitti", 12001, 3000), ("uniform", 9000, 700), ("lattice", 16384, 2500), ("line", 8200, 300),
         ("point", 10000, 260), ("kitti", 4096, 1024), ("uniform", 777, 200), ("lattice", 64, 64), ("far", 16384, 1500),
         ("tiny", 9000, 400), ("denormal", 8500, 300), ("clusters", 16384, 2048), ("plane_ties", 16384, 1024), ("uniform", 33, 33), ("uniform", 1, 1)]


@pytest.mark.parametrize("contract", [True, False])
@pytest.mark.parametrize("kind,n,m", CASES)
def test_pruned_fps_model_equals_reference_restatement(kind, n, m, contract):
    """Skipping buckets by the box bound never changes an index, whichever way the compiler rounds the bound."""
    pts = _scenes(kind, 2, n, seed=n + m)
    exp = oracle_ops.farthest_point_sample(m, pts)
    got, stats = oracle_ops.fps_pruned_model(m, pts, contract=contract)
    np.testing.assert_array_equal(got, exp)
    assert stats[1] == 2 * (m - 1)


def test_pruned_fps_model_prunes_on_lidar_like_scenes():
    """The point of the kernel: on a KITTI-like scene a round updates a handful of the 512 buckets."""
    pts = synth.kitti_like(1, 16384, seed=1000)[..., :3].copy()
    _, (updates, rounds, worst) = oracle_ops.fps_pruned_model(4096, pts)
    assert worst == 512                                                  # round 1 starts from 1e38 everywhere: nothing to skip
    assert updates / rounds < 12.0


@pytest.mark.parametrize("cuts", [(0, 1400, 2500, 3300, 4096), (0, 1, 2, 4096), (0, 4095, 4096)])
def test_pruned_fps_model_resumable_rounds(cuts):
    """Rounds in separate launches: distances in ORIGINAL order + the bucket permutation in temp [b, 2n] are all the state."""
    pts = _scenes("kitti", 2, 16384, seed=77)
    exp = oracle_ops.farthest_point_sample(4096, pts)
    out = np.full((2, 4096), -7, np.int32)
    temp = np.zeros((2, 2 * 16384), np.float32)
    for j0, j1 in zip(cuts[:-1], cuts[1:]):
        oracle_ops.fps_pruned_model(4096, pts, rounds=(j0, j1), temp=temp, out=out, idx_offset=1000)
        assert (out[:, j1:] == -7).all()
        np.testing.assert_array_equal(out[:, :j1], exp[:, :j1] + 1000)
    perm = temp[:, 16384:].view(np.uint32)                              # the permutation is one: every point exactly once
    assert all(np.array_equal(np.sort(perm[s]), np.arange(16384)) for s in range(2))


def test_pruned_fps_key_orders_like_the_reference_and_carries_the_place():
    """fb_key(o, p) = (o mod 1024, o div 1024, p): unique in its first two fields, so the place never decides a tie, and
    ordered exactly like fps_key (thread id of the reference's strided scan first, then the scan order within a thread:
    tf_sampling_g.cu:142-171)."""
    o = np.arange(16384, dtype=np.uint32)
    rng = np.random.default_rng(0)
    p = rng.permutation(16384).astype(np.uint32)
    key = ((o & 1023) << 21) | ((o >> 10) << 17) | p
    assert key.max() < 0x7FFFFFFF                                        # below KEY_INVALID
    ref_order = np.lexsort((o >> 10, o & 1023))                          # (k mod 1024 asc, k div 1024 asc)
    np.testing.assert_array_equal(np.argsort(key, kind="stable"), ref_order)
    np.testing.assert_array_equal((((key >> 17) & 15) << 10) | (key >> 21), o)
    np.testing.assert_array_equal(key & 0x3FFF, p)


def test_pruned_fps_model_randomised_sweep():
    """120 seeded random scenes of mixed kinds and sizes (n 1..3000, any m <= n + 5: sampling more points than there are
    distinct locations ends in all-zero distances, where only the tie-break decides), both roundings of the bound."""
    rng = np.random.default_rng(2024)
    kinds = ["kitti", "uniform", "lattice", "line", "point", "far", "tiny", "clusters"]
    for t in range(120):
        kind = kinds[t % len(kinds)]
        n = int(rng.integers(1, 3000))
        m = int(rng.integers(1, min(n + 5, 600) + 1))
        pts = _scenes(kind, 1, n, seed=10_000 + t)
        exp = oracle_ops.farthest_point_sample(m, pts)
        got, _ = oracle_ops.fps_pruned_model(m, pts, contract=bool(t & 1))
        np.testing.assert_array_equal(got, exp, err_msg="kind %s n %d m %d (case %d)" % (kind, n, m, t))
