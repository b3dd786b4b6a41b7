"""Writes gpurun_in/fps_check_cases.bin for tools/quick_check/fps_check: scenes + the CPU oracle's D-FPS indices
(oracle/ is used here as the checker only).  Layout: int32 ncases, then (b, n, m) per case, then per case the points
[b,n,3] float32 followed by the expected indices [b,m] int32."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ops
synth = importlib.import_module("3dssd_b200.synth")
rng = np.random.default_rng(7)
cases = []
pts = synth.kitti_like(8, 16384, seed=1000)[..., :3].copy()                         # the bench's layer-1 shape
cases.append((pts, 4096))
lat = (rng.integers(0, 12, (2, 16384, 3)).astype(np.float32) * 0.5)                  # ties everywhere
cases.append((lat, 2500))
k2 = synth.kitti_like(2, 12001, seed=15001)[..., :3].copy(); k2[:, 6000:6040] = k2[:, 3:43]
cases.append((k2, 3000))
cases.append((rng.uniform(-40, 40, (2, 9000, 3)).astype(np.float32), 700))
os.makedirs(os.path.join(ROOT, "gpurun_in"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_in", "fps_check_cases.bin"), "wb") as f:
    np.array([len(cases)] + [v for p, m in cases for v in (p.shape[0], p.shape[1], m)], np.int32).tofile(f)
    for p, m in cases:
        np.ascontiguousarray(p, np.float32).tofile(f)
        ops.farthest_point_sample(m, p).astype(np.int32).tofile(f)
print("wrote", len(cases), "cases")
