"""Deterministic synthetic inputs (there is no dataset in this environment).

kitti_like(): the "synthetic KITTI" distribution of SURVEY.md section 8d / BASELINE.md 2c -- rect camera frame,
range of /root/reference/configs/kitti/3dssd/3dssd.yaml:3, 16384 points per scene like the reference loader
(lib/dataset/dataloader/kitti_dataloader.py:137-147), INCLUDING its duplicate padding: the loader re-samples
with replacement when a scan has fewer points, so exact duplicates (FPS / ball-query ties) are part of the
input distribution.
"""
import numpy as np


def kitti_like(batch, npoints=16384, seed=1000, dup_frac=0.02):
    out = np.empty((batch, npoints, 4), np.float32)
    for s in range(batch):
        rng = np.random.default_rng(seed + s)
        x = rng.uniform(-40.0, 40.0, npoints)
        z = 70.0 * rng.uniform(0.0, 1.0, npoints) ** 2            # denser near the sensor
        ground = rng.uniform(0.0, 1.0, npoints) < 0.7
        y = np.where(ground, 1.65 + 0.05 * rng.standard_normal(npoints), rng.uniform(-1.0, 1.6, npoints))
        inten = rng.uniform(0.0, 1.0, npoints)
        pts = np.stack([x, y, z, inten], axis=1).astype(np.float32)
        ndup = int(round(dup_frac * npoints))
        if ndup:
            src = rng.integers(0, npoints - ndup, ndup)
            pts[npoints - ndup:] = pts[src]                       # duplicate padding at the tail
        out[s] = pts
    return out


def uniform_cube(batch, npoints, channels=0, seed=0):
    """U(0,1)^3 coordinates (+ N(0,1) features), the distribution of the reference's own op test
    (lib/utils/tf_ops/grouping/tf_grouping_op_test.py:11-14)."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(0.0, 1.0, (batch, npoints, 3)).astype(np.float32)
    if channels == 0:
        return xyz
    feat = rng.standard_normal((batch, npoints, channels)).astype(np.float32)
    return np.concatenate([xyz, feat], axis=-1)


def lattice(batch, npoints, step=0.25, seed=0):
    """Points on a coarse lattice: many exactly equal distances and exact duplicates -- adversarial for
    arg-max tie-breaking (FPS) and for the d == 0 / d == r boundaries of the ball queries."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 12, (batch, npoints, 3)).astype(np.float32) * np.float32(step)
    return g
