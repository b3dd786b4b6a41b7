"""CPU-side tests (-m "not gpu") of the host logic: the C-ABI library loads and exports every symbol that
include/ssd3d.h declares, shape/config plumbing, BN folding, batch sharding, and the world_size-2 gather of
detection blocks over gloo."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "ssd3d.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ssd3d_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 17
    import importlib
    b = importlib.import_module("3dssd_b200.build")
    b.build()
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libssd3d.so does not export %s" % name
    assert declared == set(pkg.EXPORTS), declared ^ set(pkg.EXPORTS)
    lib.ssd3d_version.restype = ctypes.c_int
    assert lib.ssd3d_version() == 1


def test_no_cpu_fallback(pkg):
    import torch
    with pytest.raises(ValueError, match="CUDA"):
        pkg.farthest_point_sample(4, torch.zeros((1, 8, 3)))
    with pytest.raises(ValueError):
        pkg.query_ball_point(0.0, 4, torch.zeros((1, 8, 3)), torch.zeros((1, 2, 3)))
    with pytest.raises(NotImplementedError):
        pkg.pointnet_sa_module_msg(None, None, [], [], [], True, None, True, [], [], [], None, False, "s", False, params={})


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "3dssd_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("CPU oracle", "").replace("the oracle", "").lower() or f in ("sqdist.cu",), f


def test_arch_channels_and_params(pkg):
    cfg, P = pkg.config, pkg.params
    assert cfg.layer_channels(cfg.ARCH_3DSSD, 1) == [1, 64, 128, 256, 256, 128, 512]
    prm = P.init_params(cfg.ARCH_3DSSD, 1, seed=0)
    assert prm["layer1/conv0_0/weights"].shape == (4, 16)          # C + 3 = 4 input channels
    assert prm["layer2/conv2_1/weights"].shape == (64, 96)
    assert prm["layer4/conv1_2/weights"].shape == (512, 1024)
    assert prm["layer4/ensemble/weights"].shape == (1536, 512)
    assert prm["vote/vote_offsets/weights"].shape == (128, 3) and "vote/vote_offsets/bn/gamma" not in prm
    # BN folding: y = (xW + b - mean) * gamma / sqrt(var + 1e-3) + beta
    f = P.fold(prm, "layer1/conv0_0", True, "cpu")
    x = np.random.default_rng(0).standard_normal((5, 4)).astype(np.float32)
    ref = ((x @ prm["layer1/conv0_0/weights"] + prm["layer1/conv0_0/biases"]) - prm["layer1/conv0_0/bn/moving_mean"]) \
        * prm["layer1/conv0_0/bn/gamma"] / np.sqrt(prm["layer1/conv0_0/bn/moving_variance"] + 1e-3) + prm["layer1/conv0_0/bn/beta"]
    got = (x @ f.w.numpy()) * f.scale.numpy() + f.shift.numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)


def test_shard_bounds(pkg):
    d = pkg.dist
    for total in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 4, 8):
            spans = [d.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_oracle_backbone_runs_on_cpu(pkg, oracle_ops):
    """BASELINE config 1 (plumbing, no GPU): one SA layer N=4096 -> 1024, r=0.4, K=32, C=64, MLP [64,64,128],
    B=2, through the CPU restatement only -- shapes and invariants."""
    from oracle import layers as olayers
    synth = __import__("importlib").import_module("3dssd_b200.synth")
    rng = np.random.default_rng(0)
    pts = np.concatenate([synth.uniform_cube(2, 4096, seed=1), rng.standard_normal((2, 4096, 64)).astype(np.float32)], -1)
    arch = pkg.config.ARCH_SINGLE_SA
    prm = pkg.params.init_params(arch, 64, seed=1)
    xyz_l, feat_l, fps_l, dbg = olayers.backbone_forward(arch, pts, prm, return_debug=True)
    assert xyz_l[1].shape == (2, 1024, 3) and feat_l[1].shape == (2, 1024, 128) and fps_l[1].shape == (2, 1024)
    assert (fps_l[1][:, 0] == 0).all() and len(set(fps_l[1][0].tolist())) == 1024      # distinct samples
    assert (dbg[0]["cnt"][0] >= 1).all() and (feat_l[1] >= 0).all()


_WORKER = r'''
import importlib, os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
d = importlib.import_module("3dssd_b200.dist")
rank, world, _ = d.init_from_env("gloo")
total = 5                                        # uneven: rank 0 gets 3 scenes, rank 1 gets 2
full = torch.arange(total * 100 * 9, dtype=torch.float32).view(total, 100, 9)
cnt = torch.arange(total, dtype=torch.int32) + 10
lo, hi = d.shard_bounds(total, rank, world)
blk, c = d.gather_detections(full[lo:hi].clone(), cnt[lo:hi].clone(), total_scenes=total)
assert torch.equal(blk, full), "gathered blocks differ"
assert torch.equal(c, cnt), "gathered counts differ"
assert torch.equal(d.shard_batch(full, rank, world), full[lo:hi])
dist.barrier()
print("rank", rank, "ok")
'''


def test_gather_detections_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o)
        assert "ok" in o


def test_first_conv_hoisting_algebra_matches_literal_conv():
    """relu((concat(f_j, x_j - c_i) . W1 + b) folded with BN) == relu(z[j] + (x_j - c_i) . Wx') with the per-point table
    z = (f . Wf) * s + t  -- the identity behind ssd3d_linear_tc_hoisted / ssd3d_sa_mlp_fused_hoisted, checked on the CPU
    against the oracle's literal first conv with KITTI-sized coordinates."""
    import importlib
    from oracle import ops as oops
    P = importlib.import_module("3dssd_b200.params")
    rng = np.random.default_rng(3)
    b, n, c, m, k, n1 = 2, 300, 29, 20, 16, 48
    xyz = rng.uniform(0, 70, (b, n, 3)).astype(np.float32)
    feats = np.maximum(rng.standard_normal((b, n, c)), 0).astype(np.float32)
    new_xyz = np.array(xyz[:, :m], copy=True)
    idx = rng.integers(0, n, (b, m, k)).astype(np.int32)
    prm = {}
    P._conv_init(rng, prm, "s/conv0_0", c + 3, n1, True)
    prm["s/conv0_0/biases"] = rng.standard_normal(n1).astype(np.float32)
    g = np.concatenate([oops.group_point(feats, idx), oops.group_point(xyz, idx) - new_xyz[:, :, None]], -1)
    bnp = tuple(prm["s/conv0_0/bn/" + kk] for kk in ("gamma", "beta", "moving_mean", "moving_variance"))
    exp = oops.linear_bn_relu(g, prm["s/conv0_0/weights"], prm["s/conv0_0/biases"], bnp, True)
    zconv, wxs, n1s = P.prepare(prm, "cpu").hoisted(["s/conv0_0"], True, c)
    assert n1s == [n1] and tuple(wxs[0].shape) == (3, n1)
    z = (feats.reshape(-1, c).astype(np.float64) @ zconv.w.numpy().astype(np.float64)) * zconv.scale.numpy() + zconv.shift.numpy()
    z = z.reshape(b, n, n1).astype(np.float32)
    d = oops.group_point(xyz, idx) - new_xyz[:, :, None]                                  # exact fp32 subtraction first
    got = np.maximum(oops.group_point(z, idx) + d @ wxs[0].numpy(), 0)
    assert np.abs(got - exp).max() <= 1e-4 * np.abs(exp).max()


def test_c_abi_argument_validation_needs_no_gpu(pkg):
    """The attribute / shape checks of the reference's OpKernel::Compute (OP_REQUIRES -> InvalidArgument,
    tf_sampling.cpp:136-142, tf_grouping.cpp:275-288) live in the C ABI in front of any CUDA call: they answer with
    SSD3D_ERR_INVALID_ARGUMENT (-1) or SSD3D_ERR_UNSUPPORTED (-2) and a message, also on a machine without a GPU."""
    import ctypes
    L = pkg.lib()
    null = ctypes.c_void_p(None)
    one = ctypes.c_void_p(16)                                  # a non-null, never dereferenced address

    def err():
        return L.ssd3d_last_error().decode()

    assert L.ssd3d_farthest_point_sample(1, 0, 3, 4, one, null, one, null) == -1 and "bad shape" in err()
    assert L.ssd3d_farthest_point_sample(1, 8, 3, 4, null, null, one, null) == -1 and "null" in err()
    assert L.ssd3d_farthest_point_sample_features(1, 100, 3, 0, 4, null, null, one, null) == -1
    assert L.ssd3d_farthest_point_sample_features(1, 8192, 3, 64, 16, one, one, one, null) == -2 and "not covered" in err()
    assert L.ssd3d_ffps_supported(4096, 67) == 1 and L.ssd3d_ffps_supported(8192, 67) == 0 and L.ssd3d_ffps_supported(512, 131) == 1
    r = (ctypes.c_float * 1)(-1.0)
    k = (ctypes.c_int * 1)(16)
    ptrs = (ctypes.c_void_p * 1)(16)
    assert L.ssd3d_query_ball_point_multi(1, 8, 4, 1, 0, ctypes.cast(r, ctypes.c_void_p), ctypes.cast(r, ctypes.c_void_p),
                                          ctypes.cast(k, ctypes.c_void_p), one, one, ctypes.cast(ptrs, ctypes.c_void_p),
                                          ctypes.cast(ptrs, ctypes.c_void_p), null) == -1 and "positive radius" in err()
    nout = (ctypes.c_int * 3)(128, 128, 256)
    assert L.ssd3d_sa_mlp_fused(1, 64, 128, 8, 32, one, one, one, one, null, 3, ctypes.cast(nout, ctypes.c_void_p), one, one,
                                one, 256, null, null, 0, null) == -2 and "does not fit" in err()
    assert L.ssd3d_sa_mlp_fused(1, 64, 1, 8, 7, one, one, one, one, null, 3, ctypes.cast(nout, ctypes.c_void_p), one, one,
                                one, 256, null, null, 0, null) == -1 and "nsample" in err()
    assert L.ssd3d_linear_tc(128, 20, 16, one, one, one, one, one, one, 1, 1, null, one, 16, null, null, 0, null) == -1 and "multiple of 16" in err()
    assert L.ssd3d_version() > 0


def test_header_is_plain_c():
    """include/ssd3d.h is the drop-in boundary: it must compile as C99 and as C++ with nothing but the standard
    headers (no torch / CUDA types in any signature)."""
    import shutil
    import subprocess
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "ssd3d.h")
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", hdr])
    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)          # declarations only, comments stripped
    assert "torch" not in code.lower() and "cudaStream_t" not in code and "#include <cuda" not in code
