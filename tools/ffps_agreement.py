"""F-FPS index agreement between the product's pinned arithmetic and the reference's route.

The reference feeds farthest_point_sample_with_distance from  |a|^2 + |b|^2 - 2 a.b^T  computed by a TF-1.4 cuBLAS SGEMM
(/root/reference/lib/utils/model_util.py:144-160): its summation order is unspecified, so bit parity of the matrix -- and
therefore of the sampled indices once a near-tie flips -- is undefinable.  This tool measures how often that happens:
the same layer-2 / layer-3 inputs (xyz + features produced by the product's own layers on synthetic KITTI scenes) go
through (a) the product's matrix-free F-FPS (pinned sequential fma chains) and (b) a torch fp32 matmul (cuBLAS,
allow_tf32=False) restatement of calc_square_dist + the SAME sampling kernel; reported per layer over >= 32 scenes:
scenes with identical index lists, position-wise agreement, set overlap (a flip reorders the suffix, the chosen SET
moves much less), and the first position where a scene diverges.
usage: python tools/ffps_agreement.py [scenes=32] [out.json]"""
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


def torch_sqdist(a):
    a_sq = (a * a).sum(-1, keepdim=True)
    return (a_sq + a_sq.transpose(1, 2) - 2.0 * torch.matmul(a, a.transpose(1, 2))).contiguous()


def main():
    nscenes = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "r02_ffps_agreement.json")
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    net = pkg.SABackbone(device=dev)
    stats = {"layer2 (N=4096, 67 ch, 512 samples)": [], "layer3 (N=512, 131 ch, 256 samples)": []}
    for b0 in range(0, nscenes, 8):
        pts = torch.from_numpy(synth.kitti_like(8, 16384, seed=5000 + b0)).to(dev)
        xyz_l, feat_l, fps_l = net.forward(pts)
        cases = [("layer2 (N=4096, 67 ch, 512 samples)", xyz_l[1], feat_l[1], 512),
                 ("layer3 (N=512, 131 ch, 256 samples)", xyz_l[2][:, :512].contiguous(), feat_l[2][:, :512].contiguous(), 256)]
        for name, xyz, feat, m in cases:
            both = torch.cat([xyz, feat], -1).contiguous()
            ours = pkg.farthest_point_sample_with_distance(m, pkg.calc_square_dist(both))      # pinned arithmetic (== matrix-free kernel)
            if pkg.tf_ops.ffps_supported(xyz.shape[1], both.shape[2]) and both.shape[2] <= 68:
                assert torch.equal(ours, pkg.tf_ops.farthest_point_sample_features(m, xyz, feat))
            ref = pkg.farthest_point_sample_with_distance(m, torch_sqdist(both))               # cuBLAS SGEMM route
            o, r = ours.cpu().numpy(), ref.cpu().numpy()
            for s in range(o.shape[0]):
                same = o[s] == r[s]
                first = int(np.argmin(same)) if not same.all() else -1
                stats[name].append({"identical": bool(same.all()), "positionwise": float(same.mean()),
                                    "set_overlap": len(set(o[s]) & set(r[s])) / float(m), "first_divergence": first})
    res = {"scenes": nscenes, "note": __doc__.split("usage")[0].strip()}
    for name, rows in stats.items():
        res[name] = {"scenes_identical": int(sum(x["identical"] for x in rows)), "scenes": len(rows),
                     "mean_positionwise_agreement": float(np.mean([x["positionwise"] for x in rows])),
                     "mean_set_overlap": float(np.mean([x["set_overlap"] for x in rows])),
                     "first_divergence_positions": sorted(x["first_divergence"] for x in rows if x["first_divergence"] >= 0)}
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
