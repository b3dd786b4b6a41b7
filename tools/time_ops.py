"""Op-by-op device timings of libssd3d.so next to the reference's own kernels (oracle/_ref), CUDA events.
Usage (GPU box): python tools/time_ops.py [out.json]"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dssd_b200")
synth = pkg.synth if hasattr(pkg, "synth") else importlib.import_module("3dssd_b200.synth")
from oracle import ref_ops  # noqa: E402


def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main(out_path):
    dev = torch.device("cuda:0")
    res = {"gpu": torch.cuda.get_device_name(0)}
    B = 8
    pts4 = torch.from_numpy(synth.kitti_like(B, 16384, seed=1000)).to(dev)
    xyz = pts4[..., :3].contiguous()
    have_ref = ref_ops.available()

    # ---- D-FPS layer 1: 16384 -> 4096   (variant 0 = direct/scene-resident when it fits, 1 = xyz-in-packet)
    for variant in (0, 1):
        for cl in (0, 4, 8, 16):
            try:
                res["fps_L1_16384_4096_v%d_cl%d_ms" % (variant, cl)] = timeit(
                    lambda: pkg.farthest_point_sample(4096, xyz, cluster=cl, packet_kernel=bool(variant)))
            except Exception as e:  # noqa: BLE001
                res["fps_L1_16384_4096_v%d_cl%d_ms" % (variant, cl)] = "ERR " + str(e)
    if have_ref:
        res["ref_fps_L1_ms"] = timeit(lambda: ref_ops.farthest_point_sample(4096, xyz, sync=False), 1, 3)
    fidx = pkg.farthest_point_sample(4096, xyz)
    new_xyz = pkg.gather_point(xyz, fidx)

    # ---- D-FPS layer 2/3 shapes
    x2 = new_xyz
    for cl in (0, 1, 2, 4, 8):
        res["fps_L2_4096_512_cl%d_ms" % cl] = timeit(lambda: pkg.farthest_point_sample(512, x2, cluster=cl))
    if have_ref:
        res["ref_fps_L2_ms"] = timeit(lambda: ref_ops.farthest_point_sample(512, x2, sync=False), 1, 3)
    x3 = x2[:, :512].contiguous()
    for cl in (0, 1, 2):
        res["fps_L3_512_256_cl%d_ms" % cl] = timeit(lambda: pkg.farthest_point_sample(256, x3, cluster=cl))

    # ---- F-FPS layer 2: N=4096, 3+64 features -> 512: matrix route vs fused
    f2 = torch.randn((B, 4096, 67), device=dev)
    res["sqdist_L2_4096x67_ms"] = timeit(lambda: pkg.calc_square_dist(f2))
    d2 = pkg.calc_square_dist(f2)
    for cl in (0, 2, 4, 8, 16):
        res["fpsdist_L2_cl%d_ms" % cl] = timeit(lambda: pkg.farthest_point_sample_with_distance(512, d2, cluster=cl))
    for cl in (0, 8, 16):
        try:
            res["ffps_fused_L2_cl%d_ms" % cl] = timeit(lambda: pkg.farthest_point_sample(512, f2, cluster=cl))
        except Exception as e:  # noqa: BLE001
            res["ffps_fused_L2_cl%d_ms" % cl] = "ERR " + str(e)
    res["ffps_direct_L2_ms"] = timeit(lambda: pkg.farthest_point_sample_features(512, f2[..., :3].contiguous(), f2[..., 3:].contiguous()))
    x2c, p2c = f2[..., :3].contiguous(), f2[..., 3:].contiguous()
    res["ffps_direct_L2_ms"] = timeit(lambda: pkg.farthest_point_sample_features(512, x2c, p2c))
    if have_ref:
        res["ref_fpsdist_L2_ms"] = timeit(lambda: ref_ops.farthest_point_sample_with_distance(512, d2, sync=False), 1, 3)
        res["ref_fps_generic_L2_ms"] = timeit(lambda: ref_ops.farthest_point_sample(512, f2, sync=False), 1, 3)
    del d2
    f3 = torch.randn((B, 512, 131), device=dev)
    res["sqdist_L3_512x131_ms"] = timeit(lambda: pkg.calc_square_dist(f3))
    d3 = pkg.calc_square_dist(f3)
    res["fpsdist_L3_ms"] = timeit(lambda: pkg.farthest_point_sample_with_distance(256, d3))
    res["ffps_fused_L3_ms"] = timeit(lambda: pkg.farthest_point_sample(256, f3))
    x3c, p3c = f3[..., :3].contiguous(), f3[..., 3:].contiguous()
    res["ffps_direct_L3_ms"] = timeit(lambda: pkg.farthest_point_sample_features(256, x3c, p3c))

    # ---- ball query layer 1: 3 dilated shells
    lows, highs, ks = [0.0, 0.2, 0.4], [0.2, 0.4, 0.8], [32, 32, 64]
    res["bq_L1_multi3_ms"] = timeit(lambda: pkg.query_ball_point_multi(lows, highs, ks, xyz, new_xyz, True))
    res["bq_L1_single_0.4_0.8_ms"] = timeit(lambda: pkg.query_ball_point_dilated(0.4, 0.8, 64, xyz, new_xyz))
    if have_ref:
        def ref_bq():
            for lo, hi, k in zip(lows, highs, ks):
                ref_ops.query_ball_point_dilated(lo, hi, k, xyz, new_xyz, sync=False)
        res["ref_bq_L1_3shells_ms"] = timeit(ref_bq, 1, 3)

    # ---- group + MLP layer 1 scale 3 (K=64, Cin=4 -> 32,32,64) and layer 2 scale 3 shape
    idxs, cnts = pkg.query_ball_point_multi(lows, highs, ks, xyz, new_xyz, True)
    feat1 = pts4[..., 3:].contiguous()
    res["group_concat_L1s3_ms"] = timeit(lambda: pkg.group_concat(xyz, feat1, new_xyz, idxs[2]))
    g = pkg.group_concat(xyz, feat1, new_xyz, idxs[2])
    one = lambda c: (torch.ones(c, device=dev), torch.zeros(c, device=dev))
    w1 = torch.randn((4, 32), device=dev); w2 = torch.randn((32, 32), device=dev); w3 = torch.randn((32, 64), device=dev)
    res["linear_L1s3_4x32_ms"] = timeit(lambda: pkg.linear_bn_relu(g, w1, *one(32)))
    h = pkg.linear_bn_relu(g, w1, *one(32))
    res["linear_L1s3_32x32_ms"] = timeit(lambda: pkg.linear_bn_relu(h, w2, *one(32)))
    res["linear_L1s3_32x64_pool_ms"] = timeit(lambda: pkg.linear_bn_relu(h, w3, *one(64), pool=64, rowmask=cnts[2]))
    # layer-4-like GEMM: rows = 8*256*32 = 65536, 512 -> 1024
    xa = torch.randn((B, 256, 32, 512), device=dev); wa = torch.randn((512, 1024), device=dev) * 0.05
    res["linear_L4_65536x512x1024_pool_ms"] = timeit(lambda: pkg.linear_bn_relu(xa, wa, *one(1024), pool=32, rowmask=None), 2, 5)
    torch.backends.cuda.matmul.allow_tf32 = False
    res["torch_fp32_matmul_65536x512x1024_ms"] = timeit(lambda: torch.matmul(xa.view(-1, 512), wa), 2, 5)

    # ---- three_nn
    res["three_nn_16384_4096_ms"] = timeit(lambda: pkg.three_nn(xyz, new_xyz))
    if have_ref:
        res["ref_three_nn_ms"] = timeit(lambda: ref_ops.three_nn(xyz, new_xyz, sync=False), 1, 3)

    print(json.dumps(res, indent=1))
    if out_path:
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        with open(out_path, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "time_ops.json"))
