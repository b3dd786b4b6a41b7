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
    assert lib.ssd3d_version() == 2


def test_no_cpu_fallback(pkg):
    import torch
    with pytest.raises(ValueError, match="CUDA"):
        pkg.farthest_point_sample(4, torch.zeros((1, 8, 3)))
    with pytest.raises(ValueError):
        pkg.query_ball_point(0.0, 4, torch.zeros((1, 8, 3)), torch.zeros((1, 2, 3)))
    with pytest.raises(NotImplementedError):     # use_attention (query_ball_point_withidx) is the one unbuilt branch
        pkg.pointnet_sa_module_msg(None, None, [], [], [], False, None, True, [], [], [], None, True, "s", False, params={})
    with pytest.raises(ValueError, match="CUDA"):   # training mode exists (row f3) but, like everything, only on the GPU
        pkg.tf_ops.bn_train(torch.zeros((4, 3)), torch.ones(3), torch.zeros(3))


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
# the in-step form: the producer writes into views of the send buffer, one collective, result read through views
g = d.DetectionGather(total, "cpu")
ob, oc = g.out()
assert ob.shape == (hi - lo, 100, 9) and oc.shape == (hi - lo,) and g.b_max == 3
ob.copy_(full[lo:hi]); oc.copy_(cnt[lo:hi])
g.gather()
rb, rc = g.result()
assert torch.equal(rb, full) and torch.equal(rc, cnt) and g.raw.numel() == world * g.slice_bytes
try:
    d.gather_detections(full[:1], cnt[:1], total_scenes=total)
    raise SystemExit("a wrong local shard size must be rejected")
except ValueError:
    pass
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
    assert L.ssd3d_sa_mlp_fused(1, 64, 128, 8, 32, one, one, one, one, null, null, 3, ctypes.cast(nout, ctypes.c_void_p), one, one, 1,
                                one, 256, null, null, 0, null) == -2 and "does not fit" in err()
    assert L.ssd3d_sa_mlp_fused(1, 64, 1, 8, 7, one, one, one, one, null, null, 3, ctypes.cast(nout, ctypes.c_void_p), one, one, 0,
                                one, 256, null, null, 0, null) == -1 and "nsample" in err()
    assert L.ssd3d_sa_mlp_fused(1, 64, 1, 8, 32, one, one, one, one, null, one, 3, ctypes.cast(nout, ctypes.c_void_p), one, one, 0,
                                null, 256, one, one, 256, null) == -1 and "unit list" in err()       # units need the fp32 output only
    assert L.ssd3d_linear_tc(128, 20, 16, one, one, one, one, one, one, 1, 1, null, one, 16, null, null, 0, null) == -1 and "multiple of 16" in err()
    assert L.ssd3d_version() > 0
    # the explicit-placement entry points (round 2): strides, round ranges and the cluster request are validated up front
    ll = ctypes.c_longlong
    fps_ex = L.ssd3d_farthest_point_sample_ex
    assert fps_ex(1, 64, 3, 8, one, ll(64 * 3), null, one, 4, 0, 0, 8, 0, 0, null) == -1 and "strides" in err()      # ldo < m
    assert fps_ex(1, 64, 3, 8, one, ll(10), null, one, 8, 0, 0, 8, 0, 0, null) == -1 and "strides" in err()          # scene stride < n*c
    assert fps_ex(1, 64, 3, 8, one, ll(192), null, one, 8, 0, 5, 3, 0, 0, null) == -1 and "rounds" in err()
    assert fps_ex(1, 64, 3, 8, one, ll(192), null, one, 8, 0, 0, 8, 3, 0, null) == -1 and "cluster" in err()
    assert fps_ex(1, 4096, 3, 8, one, ll(4096 * 3), null, one, 8, 0, 0, 4, 0, 0, null) == -1 and "temp" in err()   # partial range, no state buffer
    assert fps_ex(1, 64, 3, 8, one, ll(192), one, one, 8, 0, 0, 4, 0, 0, null) == -1 and "resident-scene" in err()  # too small for a cluster
    assert fps_ex(1, 64, 7, 8, one, ll(64 * 7), one, one, 8, 0, 0, 4, 0, 0, null) == -1 and "c == 3" in err()
    assert fps_ex(1, 64, 3, 8, one, ll(192), one, one, 8, 0, 4, 4, 0, 0, null) == 0                                   # empty range: nothing to do
    assert L.ssd3d_fps_supports_rounds(16384, 3) == 1 and L.ssd3d_fps_supports_rounds(65536, 3) == 0 and L.ssd3d_fps_supports_rounds(4096, 4) == 0
    assert L.ssd3d_iota_idx(2, 8, 0, one, 4, null) == -1 and "ldo" in err()
    assert L.ssd3d_decode_dist_anchor_free(4, 12, one, one, 20, one, 1, one, one, null) == -1 and "ld_reg" in err()
    assert L.ssd3d_concat_rows(1, 9, one, one, 3, one, null) == -1


def test_round2_entry_points_validate_without_a_gpu(pkg):
    """Unit-list forms, the pruned-FPS selectors and the peer exchange answer bad arguments before any CUDA call."""
    import ctypes
    L = pkg.lib()
    null, one = ctypes.c_void_p(None), ctypes.c_void_p(16)
    err = lambda: L.ssd3d_last_error().decode()
    # linear_tc on unit lists: pooling needs ReLU + the fp32 output only; the list itself is required
    assert L.ssd3d_linear_tc_units(1024, 64, 32, one, one, one, one, one, one, 1, null, 0, one, 32, null, null, 0, null) == -1 and "null unit list" in err()
    assert L.ssd3d_linear_tc_units(1024, 64, 32, one, one, one, one, one, one, 0, one, 1, one, 32, null, null, 0, null) == -1 and "post-ReLU" in err()
    assert L.ssd3d_linear_tc_units(1021, 64, 32, one, one, one, one, one, one, 1, one, 0, one, 32, null, null, 0, null) == -1 and "multiple of 8" in err()
    assert L.ssd3d_linear_tc_hoisted_units(1, 64, 32, 8, 12, one, one, 32, one, one, one, one, 32, one, one, one, one, 1, 0, one, 32,
                                           null, null, 0, null) == -1 and "nsample" in err()
    assert L.ssd3d_hoist_expand_split_units(1, 64, 32, 8, 16, one, one, 32, one, one, one, null, one, one, 32, null) == -1
    # temp size of a resumable D-FPS: 2n floats per scene where the pruned kernel is taken, n otherwise
    assert L.ssd3d_fps_temp_elems(16384, 3, 4096, 0) == 2 * 16384          # default route of layer 1
    assert L.ssd3d_fps_temp_elems(16384, 3, 4096, 2) == 16384              # bit 1: pruned kernel forbidden
    assert L.ssd3d_fps_temp_elems(4096, 3, 1024, 0) == 4096 and L.ssd3d_fps_temp_elems(4096, 3, 1024, 4) == 2 * 4096
    assert L.ssd3d_fps_temp_elems(16384, 67, 4096, 4) == 16384             # xyz only
    assert L.ssd3d_fps_temp_elems(20000, 3, 4096, 4) == 20000              # beyond its capacity
    assert pkg.tf_ops.fps_temp_elems(16384, 3, 4096, bucket_kernel=False) == 16384
    # peer exchange
    assert L.ssd3d_peer_allgather(one, 4096, one, 0, 0, 0, 8192, 16384, 16400, one, one, null) == -1 and "world" in err()
    assert L.ssd3d_peer_allgather(one, 4100, one, 2, 0, 0, 8200, 16400, 16416, one, one, null) == -1 and "16-byte" in err()
    assert L.ssd3d_peer_allgather(one, 4096, null, 2, 0, 0, 8192, 16384, 16400, one, one, null) == -1 and "null" in err()


def test_peer_layout_and_padded_slices(pkg):
    """Byte layout of a pipeline's share of the symmetric allocation: parities and flag areas do not overlap, everything the
    kernel copies with 16-byte accesses is 16-byte aligned, for any world size and scenes-per-rank count."""
    D = pkg.dist
    for world in (1, 2, 3, 8):
        for total in (world, 8 * world, 8 * world + 3, 61):
            g = D.DetectionGather.__new__(D.DetectionGather)            # layout arithmetic only (no device)
            g.world, g.rank, g.total, g.max_output = world, 0, total, 100
            g.b_max = (total + world - 1) // world
            g.block_bytes = g.b_max * 100 * 9 * 4
            sb = (g.block_bytes + g.b_max * 4 + 15) // 16 * 16
            recv, flags, tot = D.peer_layout(world, sb)
            assert sb % 16 == 0 and sb >= g.block_bytes + 4 * g.b_max
            assert recv == (0, world * sb) and flags[0] == 2 * world * sb and flags[1] - flags[0] >= 4 * world
            assert tot >= flags[1] + 4 * world and all(v % 16 == 0 for v in recv + flags)
            assert D.PeerArena.share_bytes(world, sb) % 256 == 0 and D.PeerArena.share_bytes(world, sb) >= tot


def test_library_keeps_no_state(pkg):
    """include/ssd3d.h promises a library that 'keeps no state between calls': no tuning setters in the ABI (round 1
    had process-global ssd3d_tune_set_* knobs; every such choice is now an argument of an *_ex entry point) and no
    developer hooks in the shipped build."""
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", pkg.LIB_PATH]).decode()
    names = [ln.split()[-1] for ln in out.splitlines() if " T " in ln]
    assert not [n for n in names if "tune" in n or "_dev_" in n or "_set_" in n], names
    hdr = open(os.path.join(ROOT, "include", "ssd3d.h")).read()
    assert "tune" not in hdr and "keeps no state" in hdr
    bss = [ln for ln in subprocess.check_output(["nm", "--defined-only", pkg.LIB_PATH]).decode().splitlines()
           if (" b " in ln or " d " in ln or " B " in ln or " D " in ln) and "g_" in ln and "ssd3d" in ln]
    assert all("g_err" in ln for ln in bss), bss          # the only global: the thread-local last-error string


def test_oracle_c_nms_twin_matches_python_restatement(oracle_ops):
    from oracle import head as ohead
    rng = np.random.default_rng(2)
    for n, spread in ((256, 12.0), (64, 2.0), (5, 1.0)):
        ctr = rng.uniform(-spread, spread, (2, n, 3)).astype(np.float32)
        boxes = np.concatenate([ctr, rng.uniform(0.5, 4.5, (2, n, 3)).astype(np.float32),
                                rng.uniform(-np.pi, np.pi, (2, n, 1)).astype(np.float32)], -1)
        sc = rng.uniform(0, 1, (2, n)).astype(np.float32)
        sc[:, 1:3] = sc[:, :1]
        a, b = ohead.bev_nms(boxes, sc, 0.1, 50), oracle_ops.bev_nms(boxes, sc, 0.1, 50)
        np.testing.assert_array_equal(a[0], b[0]); np.testing.assert_array_equal(a[1], b[1])


def test_fps_part_bounds(pkg):
    pb = pkg.layers_util._part_bounds
    assert pb(4096, 4) == [(0, 1024), (1024, 2048), (2048, 3072), (3072, 4096)]
    assert pb(4096, [0.34, 0.28, 0.22, 0.16]) == [(0, 1408), (1408, 2560), (2560, 3456), (3456, 4096)]
    assert pb(512, 8)[-1][1] == 512 and all(a < b for a, b in pb(512, 8))
    assert pb(100, 3) == [(0, 100)]                       # smaller than one 128-row tile: a single part


def _bq_grid_cells(xyz, r_max):
    """float32 restatement of the cell arithmetic of csrc/ball_query_grid.cu (bq_grid_build_kernel / bqg_cell_coord)."""
    f = np.float32
    lo, hi = xyz.min(0), xyz.max(0)
    ext = (hi - lo).astype(np.float32)
    order = np.argsort(-ext, kind="stable")
    a0, a1 = int(order[0]), int(order[1])
    c = f(r_max) * f(1.01)
    while True:
        fa, fb = np.floor(ext[a0] / c) + f(1), np.floor(ext[a1] / c) + f(1)
        if fa * fb <= 8192:
            break
        c = f(c * f(1.25))
    inv_c = f(1) / c

    def cell(v, mn, nc):
        u = ((v.astype(np.float32) - f(mn)) * inv_c).astype(np.float32)
        return np.clip(np.floor(u).astype(np.int64), 0, int(nc) - 1)
    return (a0, a1), (lo[a0], lo[a1]), (int(fa), int(fb)), cell, float(c)


@pytest.mark.parametrize("name", ["kitti", "cube", "line", "huge-extent", "boundary"])
def test_ball_query_grid_culling_rule_never_drops_a_hit(pkg, name):
    """The culled ball query (csrc/ball_query_grid.cu) visits the 3x3 cell neighbourhood of a query.  Property behind its
    bit-exactness: for EVERY (query, candidate) pair the reference would accept for the largest radius, the cells differ
    by at most one on both grid axes -- checked here on the float32 cell arithmetic, incl. points placed exactly on cell
    boundaries and extents that push the cell count to its cap."""
    import importlib
    synth = importlib.import_module("3dssd_b200.synth")
    rng = np.random.default_rng(3)
    r_max = 0.8
    if name == "kitti":
        xyz = synth.kitti_like(1, 16384, seed=9)[0, :, :3]
    elif name == "cube":
        xyz = rng.uniform(0, 1, (4096, 3)).astype(np.float32); r_max = 0.2
    elif name == "line":
        xyz = np.stack([np.linspace(0, 5000, 8192, dtype=np.float32), np.zeros(8192, np.float32), np.zeros(8192, np.float32)], -1)
    elif name == "huge-extent":
        xyz = (rng.uniform(-1, 1, (4096, 3)) * np.array([3.0e4, 2.0e4, 5.0])).astype(np.float32); r_max = 0.3
    else:
        base = synth.kitti_like(1, 2048, seed=2)[0, :, :3]
        (a0, a1), (m0, m1), _, _, c = _bq_grid_cells(base, r_max)
        xyz = base.copy()
        xyz[::2, a0] = np.float32(m0) + np.round((xyz[::2, a0] - m0) / c).astype(np.float32) * np.float32(c)   # on cell boundaries
    (a0, a1), (m0, m1), (na, nb), cell, c = _bq_grid_cells(xyz, r_max)
    assert na * nb <= 8192 and c >= r_max * 1.0099
    q = xyz[rng.choice(len(xyz), 256, replace=False)] + rng.normal(0, r_max / 3, (256, 3)).astype(np.float32)
    d = np.sqrt(((q[:, None, :].astype(np.float64) - xyz[None].astype(np.float64)) ** 2).sum(-1))
    qi, ki = np.nonzero(d <= r_max * (1 + 1e-6))                  # everything the fp32 test could accept, with slack
    assert len(qi) > 0
    da = np.abs(cell(q[:, a0], m0, na)[qi] - cell(xyz[:, a0], m0, na)[ki])
    db = np.abs(cell(q[:, a1], m1, nb)[qi] - cell(xyz[:, a1], m1, nb)[ki])
    assert da.max() <= 1 and db.max() <= 1


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


def test_ctypes_signatures_match_the_header(pkg):
    """Every argtypes list of 3dssd_b200/_lib.py is the prototype of include/ssd3d.h, argument by argument (a count or a
    width mismatch -- long vs int, size_t vs int -- is silent in ctypes and corrupts the call): the header is parsed here
    and each parameter's C type mapped to its ctypes class."""
    lib_mod = __import__("importlib").import_module("3dssd_b200._lib")
    hdr = open(os.path.join(ROOT, "include", "ssd3d.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", "", hdr)
    protos = re.findall(r"\b([a-z_ ]+?[ \*])\s*(ssd3d_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)
    assert len(protos) >= 50

    def ctype(decl):
        decl = decl.strip()
        if decl in ("void", ""):
            return None
        if "*" in decl or decl.startswith("ssd3d_stream_t"):
            return ctypes.c_void_p
        base = re.sub(r"\b(const|unsigned)\b", "", decl)
        base = " ".join(base.split()[:-1]) if len(base.split()) > 1 else base.strip()     # drop the parameter name
        return {"int": ctypes.c_int, "long": ctypes.c_long, "long long": ctypes.c_longlong, "size_t": ctypes.c_size_t,
                "float": ctypes.c_float}[base.strip()]

    seen = set()
    for ret, name, args in protos:
        if name == "ssd3d_last_error":
            continue
        want = [t for t in (ctype(a) for a in args.split(",")) if t is not None]
        got = lib_mod._SIGNATURES[name]
        assert len(got) == len(want), "%s: %d argtypes, the header declares %d parameters" % (name, len(got), len(want))
        for i, (g, w) in enumerate(zip(got, want)):
            assert ctypes.sizeof(g) == ctypes.sizeof(w) and (g is ctypes.c_float) == (w is ctypes.c_float) and \
                (g is ctypes.c_void_p) == (w is ctypes.c_void_p), "%s: parameter %d is %s in the header, %s in _lib.py" % (name, i, w, g)
        seen.add(name)
    assert seen == set(lib_mod._SIGNATURES), seen ^ set(lib_mod._SIGNATURES)
