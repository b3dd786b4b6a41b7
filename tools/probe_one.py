"""One operator of the step on its bench-shaped input, a few launches (for ncu --set full captures).
usage: python tools/probe_one.py fps_l1 | bq_l1 | ffps_l2 | expand_l4 | nms"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dssd_b200")
synth = importlib.import_module("3dssd_b200.synth")


def main(which, reps=5):
    dev = torch.device("cuda:0")
    pts = torch.from_numpy(synth.kitti_like(8, 16384, seed=1000)).to(dev)
    xyz = pts[..., :3].contiguous()
    if which == "fps_l1":
        fn = lambda: pkg.farthest_point_sample(4096, xyz)
    elif which == "bq_l1":
        q = pkg.gather_point(xyz, pkg.farthest_point_sample(4096, xyz))
        fn = lambda: pkg.query_ball_point_multi([0.0, 0.2, 0.4], [0.2, 0.4, 0.8], [32, 32, 64], xyz, q, True)
    elif which == "ffps_l2":
        x2 = xyz[:, :4096].contiguous()
        f2 = torch.relu(torch.randn((8, 4096, 64), device=dev))
        fn = lambda: pkg.tf_ops.farthest_point_sample_features(512, x2, f2)
    elif which == "expand_l4":
        x = torch.rand((8, 512, 3), device=dev) * 40
        z = torch.randn((8, 512, 512), device=dev)
        idx = torch.randint(0, 512, (8, 256, 32), device=dev, dtype=torch.int32)
        wx = torch.randn((3, 256), device=dev)
        fn = lambda: pkg.tf_ops.hoist_expand_split(x, z, 256, wx, x[:, :256].contiguous(), idx)
    elif which == "nms":
        rng = np.random.default_rng(0)
        b = np.concatenate([rng.uniform(-30, 30, (8, 256, 3)), rng.uniform(1, 4, (8, 256, 3)), rng.uniform(-3, 3, (8, 256, 1))], -1).astype(np.float32)
        boxes, sc = torch.from_numpy(b).to(dev), torch.rand((8, 256), device=dev)
        fn = lambda: pkg.bev_nms(boxes, sc, 0.1, 100)
    else:
        raise SystemExit(__doc__)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    print("ok", which)


if __name__ == "__main__":
    main(sys.argv[1])
