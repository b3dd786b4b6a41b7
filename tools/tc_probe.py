"""Times linear_tc on representative layer shapes (CUDA events) -- python tools/tc_probe.py [shape-index]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dssd_b200")
P = pkg.params

SHAPES = [  # rows, cin, cout, pool, mode
    (131072, 131, 128, 1, "split"),      # L3 first layer
    (131072, 128, 128, 1, "split"),
    (131072, 256, 256, 32, "pool"),
    (65536, 512, 1024, 32, "pool"),      # L4 biggest
    (65536, 259, 256, 1, "split"),
    (1048576, 4, 16, 1, "split"),        # L1
    (1048576, 16, 32, 32, "pool"),
    (524288, 67, 64, 1, "split"),        # L2
    (131072, 128, 128, 1, "f32"),
    (131072, 128, 128, 1, "hoist"),      # 9: L3 scale 1 second conv fed by the hoisted first conv (n1 = cin)
    (131072, 128, 256, 1, "hoist"),      # 10: L3 scale 3
    (65536, 256, 512, 1, "hoist"),       # 11: L4 scale 2
]


def main():
    dev = torch.device("cuda:0")
    sel = [int(a) for a in sys.argv[1:] if "=" not in a] or range(len(SHAPES))
    kv = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
    rng = np.random.default_rng(0)
    for i in sel:
        rows, cin, cout, pool, mode = SHAPES[i]
        prm = {}
        P._conv_init(rng, prm, "s", cin, cout, True)
        f = P.fold(prm, "s", True, dev)
        shape = (rows // pool, pool, cin) if pool > 1 else (rows, cin)
        kw = dict(pool=pool) if pool > 1 else {}
        if mode == "split":
            kw.update(want_f32=False, want_split=True)
        if mode == "hoist":
            B, ns, npts = 8, 32, 1024
            m = rows // (B * ns)
            xyz = torch.rand((B, npts, 3), device=dev) * 40
            new_xyz = xyz[:, :m].contiguous()
            idx = torch.randint(0, npts, (B, m, ns), device=dev, dtype=torch.int32)
            z = torch.randn((B, npts, cin), device=dev)
            wx = torch.randn((3, cin), device=dev)

            def run():
                return pkg.linear_tc_hoisted(xyz, z, 0, wx, new_xyz, idx, f)
        else:
            x = torch.randn(shape, device=dev)
            hi, lo = pkg.split_rows(x)

            def run():
                return pkg.linear_tc(hi, lo, f, **kw)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        fl = 2.0 * rows * cin * cout
        print("shape %d rows=%d cin=%d cout=%d pool=%d %s: %.1f us  %.1f TFLOP/s (fp32-equivalent)  kp=%d"
              % (i, rows, cin, cout, pool, mode, ms * 1e3, fl / ms / 1e9, f.kp), flush=True)


if __name__ == "__main__":
    main()
