"""Layer-1 D-FPS (16384 -> 4096, 8 scenes): the single-CTA bucket kernel (csrc/fps_bucket.cu) next to the cluster kernel,
CUDA events, and the identity of their outputs.  Usage (GPU box): python tools/fps_l1_compare.py [out.json]"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dssd_b200")
synth = importlib.import_module("3dssd_b200.synth")


def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main(out_path):
    dev = torch.device("cuda:0")
    res = {"gpu": torch.cuda.get_device_name(0), "shape": "B=8, N=16384 -> M=4096 (and M sweep)"}
    xyz = torch.from_numpy(synth.kitti_like(8, 16384, seed=1000)[..., :3].copy()).to(dev)
    a = pkg.farthest_point_sample(4096, xyz, bucket_kernel=True)
    b = pkg.farthest_point_sample(4096, xyz, bucket_kernel=False)
    res["identical"] = bool(torch.equal(a, b))
    for m in (256, 1024, 4096):
        res["bucket_m%d_ms" % m] = timeit(lambda: pkg.farthest_point_sample(m, xyz, bucket_kernel=True))
    res["bucket_ns_per_round"] = (res["bucket_m4096_ms"] - res["bucket_m256_ms"]) * 1e6 / (4096 - 256)
    res["bucket_setup_ms_est"] = res["bucket_m256_ms"] - 255 * res["bucket_ns_per_round"] * 1e-6
    for cl in (0, 4):
        res["cluster%d_m4096_ms" % cl] = timeit(lambda: pkg.farthest_point_sample(4096, xyz, bucket_kernel=False, cluster=cl))
    uni = torch.from_numpy(np.random.default_rng(0).uniform(-40, 40, (8, 16384, 3)).astype(np.float32)).to(dev)
    res["bucket_uniform_cube_m4096_ms"] = timeit(lambda: pkg.farthest_point_sample(4096, uni, bucket_kernel=True))
    res["cluster_uniform_cube_m4096_ms"] = timeit(lambda: pkg.farthest_point_sample(4096, uni, bucket_kernel=False))
    print(json.dumps(res, indent=1))
    if out_path:
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        with open(out_path, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "fps_l1_compare.json"))
