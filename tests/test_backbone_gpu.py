"""GPU parity tests (-m gpu) of the layer / backbone orchestration: pointnet_sa_module_msg and the whole 3DSSD
SA backbone against the CPU oracle (indices bit-exact, features <= 1e-3 relative fp32, BASELINE.json)."""
import copy
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

synth = importlib.import_module("3dssd_b200.synth")
TOL = 1e-3   # BASELINE.json north star: grouped-MLP features within 1e-3 relative fp32


def rel_err(got, exp, rtol=1e-3, atol_frac=2e-5):
    """max|got-exp| / max|exp| (the global norm the tolerances are quoted in), after ALSO asserting the elementwise
    bound |got - exp| <= rtol*|exp| + atol_frac*max|exp|: a small-magnitude feature may not hide behind a large one."""
    g, e = got.astype(np.float64), exp.astype(np.float64)
    scale = max(1e-12, np.abs(e).max())
    excess = np.abs(g - e) - (rtol * np.abs(e) + atol_frac * scale)
    if excess.size and excess.max() > 0:
        k = np.unravel_index(np.argmax(excess), excess.shape)
        raise AssertionError("elementwise bound exceeded at %s: got %r expected %r (max|exp| %g, %d of %d elements over)"
                             % (k, g[k], e[k], scale, int((excess > 0).sum()), excess.size))
    return float(np.abs(g - e).max() / scale)


def scaled_arch(pkg, div):
    """ARCH_3DSSD with every point count divided by `div` (same layer structure, radii, MLP widths)."""
    arch = copy.deepcopy(pkg.config.ARCH_3DSSD)
    for spec in arch:
        spec[6] = [r if r == -1 else max(1, r // div) for r in spec[6]]
        spec[8] = [p if p in (-1, 0) else max(1, p // div) for p in spec[8]]
    return arch


def compact_scene(batch, n, seed):
    """KITTI-like cloud squeezed so that a 1/8-size scene keeps the point density the radii were tuned for."""
    pts = synth.kitti_like(batch, n, seed=seed)
    pts[..., 0] *= 0.35
    pts[..., 2] *= 0.35
    return pts


def check_backbone(pkg, net_out, oracle_out, nlayers):
    xyz_l, feat_l, fps_l, dbg = net_out
    oxyz, ofeat, ofps, odbg = oracle_out
    for li in range(1, nlayers + 1):
        if ofps[li] is not None:
            np.testing.assert_array_equal(fps_l[li].cpu().numpy(), ofps[li], err_msg="fps_idx of layer %d" % li)
        d, od = dbg[li - 1], odbg[li - 1]
        for s, (a, b) in enumerate(zip(d.get("idx", []), od.get("idx", []))):
            np.testing.assert_array_equal(a.cpu().numpy(), b, err_msg="ball-query idx layer %d scale %d" % (li, s))
        for s, (a, b) in enumerate(zip(d.get("cnt", []), od.get("cnt", []))):
            np.testing.assert_array_equal(a.cpu().numpy(), b, err_msg="pts_cnt layer %d scale %d" % (li, s))
        if oxyz[li] is not None:
            e = rel_err(xyz_l[li].cpu().numpy(), oxyz[li])
            assert e < TOL, "xyz of layer %d: rel err %g" % (li, e)
        e = rel_err(feat_l[li].cpu().numpy(), ofeat[li])
        assert e < TOL, "features of layer %d: rel err %g" % (li, e)


def check_backbone_layerwise(pkg, arch, params, net_out, ffps_mode="matrix"):
    """Teacher-forced parity: every layer of the oracle is fed the GPU's OWN outputs of the layers before it, so its
    indices (FPS -- incl. F-FPS, which reads features --, ball query) must match bit for bit BY CONSTRUCTION of the
    arithmetic, not because no near-tie happened to flip under the ~1e-5 feature differences of the layers before."""
    from oracle import layers as olayers
    xyz_l, feat_l, fps_l, dbg = net_out
    npy = lambda t: None if t is None else t.cpu().numpy()
    for li, spec in enumerate(arch, start=1):
        (xyz_i, feat_i, radius, nsample, mlps, bn, rng, method, npoint, former, attn, ltype, scope, dilated,
         vote_idx, agg) = spec
        xin, fin = npy(xyz_l[xyz_i[0]]), npy(feat_l[feat_i[0]])
        if ltype == "SA_Layer":
            former_idx = npy(fps_l[former]) if former != -1 else None
            vote_ctr = npy(xyz_l[vote_idx]) if vote_idx != -1 else None
            ex, ef, ei, ed = olayers.pointnet_sa_module_msg(xin, fin, radius, nsample, mlps, False, None, bn, rng, method,
                                                            npoint, former_idx, attn, scope, dilated, vote_ctr, agg,
                                                            params=params, ffps_mode=ffps_mode, return_debug=True)
            np.testing.assert_array_equal(npy(fps_l[li]), ei, err_msg="fps_idx of layer %d" % li)
            for sc, (a, b) in enumerate(zip(dbg[li - 1].get("idx", []), ed.get("idx", []))):
                np.testing.assert_array_equal(npy(a), b, err_msg="ball-query idx layer %d scale %d" % (li, sc))
            for sc, (a, b) in enumerate(zip(dbg[li - 1].get("cnt", []), ed.get("cnt", []))):
                np.testing.assert_array_equal(npy(a), b, err_msg="pts_cnt layer %d scale %d" % (li, sc))
            np.testing.assert_array_equal(npy(xyz_l[li]), ex, err_msg="new_xyz of layer %d" % li)
        elif ltype == "Vote_Layer":
            ex, ef, _ = olayers.vote_layer(xin, fin, mlps, False, None, bn, scope, params=params)
            assert rel_err(npy(xyz_l[li]), ex) < TOL, "vote xyz of layer %d" % li
        else:
            continue
        e = rel_err(npy(feat_l[li]), ef)
        assert e < TOL, "features of layer %d: rel err %g" % (li, e)


@pytest.mark.parametrize("ffps_mode,mlp_mode,fuse", [("matrix", "tc", True), ("fused", "tc", True), ("matrix", "tc", False),
                                                     ("matrix", "fp32", False)])
def test_backbone_small_vs_oracle(pkg, oracle_ops, cuda, ffps_mode, mlp_mode, fuse):
    from oracle import layers as olayers
    arch = scaled_arch(pkg, 8)                                    # 2048 -> 512 -> 128 -> 64 -> 32 centres
    params = pkg.params.init_params(arch, 1, seed=4, random_bias=True)
    pts = compact_scene(2, 2048, seed=40)
    net = pkg.SABackbone(arch, params, in_channels=1, device=cuda, ffps_mode=ffps_mode, mlp_mode=mlp_mode,
                         fuse_scale=fuse)
    out = net.forward(torch.from_numpy(pts).to(cuda), return_debug=True)
    exp = olayers.backbone_forward(arch, pts, params, ffps_mode=ffps_mode, return_debug=True)
    check_backbone(pkg, out, exp, len(arch))
    assert out[1][-1].shape == (2, 32, 512)


@pytest.mark.parametrize("gather,hoist", [(False, 0), (True, 0), (True, 1), (True, 2)])
def test_backbone_mlp_route_variants_vs_oracle(pkg, oracle_ops, cuda, gather, hoist):
    """Every route of the grouped MLP (materialised operand, operand built in the kernel, first conv hoisted for the
    layer-by-layer scales only / for all scales) reproduces the oracle's literal conv stack."""
    from oracle import layers as olayers
    arch = scaled_arch(pkg, 8)
    params = pkg.params.init_params(arch, 1, seed=6, random_bias=True)
    pts = compact_scene(2, 2048, seed=41)
    net = pkg.SABackbone(arch, params, in_channels=1, device=cuda, gather_in_kernel=gather, hoist_first=hoist)
    out = net.forward(torch.from_numpy(pts).to(cuda), return_debug=True)
    exp = olayers.backbone_forward(arch, pts, params, return_debug=True)
    check_backbone(pkg, out, exp, len(arch))


def test_single_sa_layer_plain_query_vs_oracle(pkg, oracle_ops, cuda):
    """BASELINE configs[0]: one SA layer N=4096 -> 1024, plain ball query r=0.4, K=32, C=64, MLP [64,64,128]."""
    from oracle import layers as olayers
    arch = pkg.config.ARCH_SINGLE_SA
    rng = np.random.default_rng(0)
    pts = np.concatenate([synth.uniform_cube(2, 4096, seed=1), rng.standard_normal((2, 4096, 64)).astype(np.float32)], -1)
    params = pkg.params.init_params(arch, 64, seed=1, random_bias=True)
    net = pkg.SABackbone(arch, params, in_channels=64, device=cuda)
    out = net.forward(torch.from_numpy(pts).to(cuda), return_debug=True)
    exp = olayers.backbone_forward(arch, pts, params, return_debug=True)
    check_backbone(pkg, out, exp, 1)


def test_sa_module_with_empty_balls_and_vote_ctr(pkg, oracle_ops, cuda):
    """CG-layer shape: queries are free-floating centres (vote_ctr), plain query, some balls empty -> masked rows."""
    from oracle import layers as olayers
    rng = np.random.default_rng(5)
    xyz = rng.uniform(0, 4, (2, 256, 3)).astype(np.float32)
    feats = rng.standard_normal((2, 256, 32)).astype(np.float32)
    ctr = rng.uniform(-1, 5, (2, 64, 3)).astype(np.float32)      # some far outside -> empty balls
    arch = [[[0], [0], [0.6, 1.2], [16, 32], [[32, 32, 64], [32, 64, 64]], True, [-1], ['D-FPS'], [64],
             -1, False, 'SA_Layer', 'cg', False, -1, 48]]
    params = pkg.params.init_params(arch, 32, seed=2, random_bias=True)
    args = (arch[0][2], arch[0][3], arch[0][4], False, None, True, [-1], ['D-FPS'], [64], None, False, 'cg', False)
    got = pkg.pointnet_sa_module_msg(torch.from_numpy(xyz).to(cuda), torch.from_numpy(feats).to(cuda), *args,
                                     vote_ctr=torch.from_numpy(ctr).to(cuda), aggregation_channel=48, params=params,
                                     return_debug=True)
    exp = olayers.pointnet_sa_module_msg(xyz, feats, *args, vote_ctr=ctr, aggregation_channel=48, params=params,
                                         return_debug=True)
    assert (exp[3]["cnt"][0] == 0).any(), "the case must contain empty balls"
    np.testing.assert_array_equal(got[2].cpu().numpy(), exp[2])
    for a, b in zip(got[3]["idx"], exp[3]["idx"]):
        np.testing.assert_array_equal(a.cpu().numpy(), b)
    np.testing.assert_array_equal(got[0].cpu().numpy(), exp[0])
    assert rel_err(got[1].cpu().numpy(), exp[1]) < TOL


@pytest.mark.parametrize("decay", [None, 0.5])
def test_sa_module_training_mode_batchnorm_vs_oracle(pkg, oracle_ops, cuda, decay):
    """is_training=True: batch-statistics BatchNorm in every conv of the SA module, moving statistics
    updated in place (tf_util.py:424-444), against the float64 restatement; then inference with the committed
    statistics equals the oracle's inference on the updated parameters."""
    from oracle import layers as olayers
    rng = np.random.default_rng(11)
    pts = compact_scene(2, 1024, seed=61)
    xyz, feats = pts[..., :3].copy(), rng.standard_normal((2, 1024, 16)).astype(np.float32)
    arch = [[[0], [0], [0.4, 0.8], [16, 32], [[16, 16, 32], [16, 32, 48]], True, [-1], ['D-FPS'], [128],
             -1, False, 'SA_Layer', 'trn', True, -1, 64]]
    params = dict(pkg.params.init_params(arch, 16, seed=8, random_bias=True))
    args = (arch[0][2], arch[0][3], arch[0][4], True, decay, True, [-1], ['D-FPS'], [128], None, False, 'trn', True)
    pp = pkg.params.prepare(params, cuda)
    got = pkg.pointnet_sa_module_msg(torch.from_numpy(xyz).to(cuda), torch.from_numpy(feats).to(cuda), *args,
                                     aggregation_channel=64, params=pp, return_debug=True)
    upd = {}
    exp = olayers.pointnet_sa_module_msg(xyz, feats, *args, aggregation_channel=64, params=params, return_debug=True,
                                         bn_updates=upd)
    np.testing.assert_array_equal(got[2].cpu().numpy(), exp[2])
    for a, b in zip(got[3]["idx"], exp[3]["idx"]):
        np.testing.assert_array_equal(a.cpu().numpy(), b)
    assert rel_err(got[1].cpu().numpy(), exp[1]) < TOL
    assert set(upd) == {"trn/conv%d_%d" % (i, j) for i in range(2) for j in range(3)} | {"trn/ensemble"}
    for scope, u in upd.items():
        st = pp.bn_state(scope)
        assert rel_err(st["moving_mean"].cpu().numpy(), u["moving_mean"], atol_frac=1e-4) < 1e-4, scope
        assert rel_err(st["moving_variance"].cpu().numpy(), u["moving_variance"], atol_frac=1e-4) < 1e-4, scope
    # commit -> the parameter dict carries the new statistics and inference folds THEM
    before = params["trn/conv0_0/bn/moving_mean"].copy()
    pp.commit_bn()
    assert not np.array_equal(params["trn/conv0_0/bn/moving_mean"], before)
    args_inf = args[:3] + (False,) + args[4:]
    got_i = pkg.pointnet_sa_module_msg(torch.from_numpy(xyz).to(cuda), torch.from_numpy(feats).to(cuda), *args_inf,
                                       aggregation_channel=64, params=pp)
    exp_i = olayers.pointnet_sa_module_msg(xyz, feats, *args_inf, aggregation_channel=64, params=params)
    assert rel_err(got_i[1].cpu().numpy(), exp_i[1]) < TOL


def test_fp_and_global_sa_modules_vs_oracle(pkg, oracle_ops, cuda):
    from oracle import layers as olayers
    rng = np.random.default_rng(6)
    xyz1 = rng.uniform(0, 1, (2, 300, 3)).astype(np.float32); xyz2 = rng.uniform(0, 1, (2, 70, 3)).astype(np.float32)
    p1 = rng.standard_normal((2, 300, 8)).astype(np.float32); p2 = rng.standard_normal((2, 70, 24)).astype(np.float32)
    arch = [[[0, 1], [0, 1], -1, -1, [32, 16], True, [-1], [-1], [-1], -1, -1, 'FP_Layer', 'fp', False, -1, -1]]
    rngp = np.random.default_rng(1)
    params = {}
    P = pkg.params
    P._conv_init(rngp, params, "fp/conv_0", 32, 32, True); P._conv_init(rngp, params, "fp/conv_1", 32, 16, True)
    t = lambda a: torch.from_numpy(a).to(cuda)
    got = pkg.pointnet_fp_module(t(xyz1), t(xyz2), t(p1), t(p2), [32, 16], False, None, "fp", True, params=params)
    exp = olayers.pointnet_fp_module(xyz1, xyz2, p1, p2, [32, 16], False, None, "fp", True, params=params)
    assert rel_err(got.cpu().numpy(), exp) < TOL
    params = {}
    P._conv_init(rngp, params, "g/conv0", 11, 32, True); P._conv_init(rngp, params, "g/conv1", 32, 64, True)
    got = pkg.pointnet_sa_module(t(xyz1), t(p1), [32, 64], False, None, True, "g", params=params)
    exp = olayers.pointnet_sa_module(xyz1, p1, [32, 64], False, None, True, "g", params=params)
    assert got.shape == (2, 64) and rel_err(got.cpu().numpy(), exp) < TOL


def test_backbone_full_size_one_scene_vs_oracle(pkg, oracle_ops, cuda):
    """BASELINE configs[1] at full size (16384 points, real layer table), one scene so the CPU oracle finishes
    in about a minute."""
    from oracle import layers as olayers
    arch = pkg.config.ARCH_3DSSD
    params = pkg.params.init_params(arch, 1, seed=0)
    pts = synth.kitti_like(1, 16384, seed=1000)
    net = pkg.SABackbone(arch, params, in_channels=1, device=cuda)
    out = net.forward(torch.from_numpy(pts).to(cuda), return_debug=True)
    check_backbone_layerwise(pkg, arch, params, out)          # robust: teacher-forced, layer by layer
    # latency mode (sampling consumed in parts, resumable layer-1 FPS) is the same function: identical outputs
    net_l = pkg.SABackbone(arch, params, in_channels=1, device=cuda, latency_mode=True)
    out_l = net_l.forward(torch.from_numpy(pts).to(cuda), return_debug=True)
    for li in range(1, len(arch) + 1):
        if out[2][li] is not None:
            assert torch.equal(out_l[2][li], out[2][li]), "latency-mode fps_idx of layer %d" % li
        if out[0][li] is not None:
            assert torch.equal(out_l[0][li], out[0][li]), "latency-mode xyz of layer %d" % li
        assert torch.equal(out_l[1][li], out[1][li]), "latency-mode features of layer %d" % li
        for a, b in zip(out_l[3][li - 1].get("idx", []), out[3][li - 1].get("idx", [])):
            assert torch.equal(a, b)
    exp = olayers.backbone_forward(arch, pts, params, return_debug=True)   # end to end (indices may legitimately fork
    check_backbone(pkg, out, exp, len(arch))                               # at a near-tie; this seed does not)


@pytest.mark.parametrize("parts", [2, [0.5, 0.25, 0.25]])
def test_backbone_latency_mode_small_equals_default(pkg, cuda, parts):
    """Sampling consumed in parts (side streams, resumable D-FPS) is bit-identical to the plain schedule, eager and
    captured."""
    arch = scaled_arch(pkg, 4)                                    # 4096 -> 1024 -> 256 -> 128
    params = pkg.params.init_params(arch, 1, seed=9, random_bias=True)
    pts = torch.from_numpy(compact_scene(3, 4096, seed=44)).to(cuda)
    a = pkg.SABackbone(arch, params, in_channels=1, device=cuda).forward(pts, return_debug=True)
    net = pkg.SABackbone(arch, params, in_channels=1, device=cuda, latency_mode=True, fps_parts=parts)
    b = net.forward(pts, return_debug=True)
    replay = net.capture(pts)
    c, _ = replay()
    torch.cuda.synchronize()
    for li in range(1, len(arch) + 1):
        for o in (b, c):
            assert torch.equal(o[1][li], a[1][li]), "features of layer %d" % li
            if a[2][li] is not None:
                assert torch.equal(o[2][li], a[2][li]), "fps_idx of layer %d" % li
        for x, y in zip(b[3][li - 1].get("idx", []), a[3][li - 1].get("idx", [])):
            assert torch.equal(x, y)


def test_backbone_unit_lists_do_not_change_a_bit(pkg, cuda):
    """layers_util.COMPACT_GROUPS (grouped MLPs convolve only the 8-row units that hold distinct neighbours) is an
    execution detail: every layer's features are bit-identical with it on and off -- full-size layer table, real
    neighbour statistics of the synthetic KITTI-like clouds, and a dense cloud where most groups are full."""
    lu = pkg.layers_util
    arch = pkg.config.ARCH_3DSSD
    params = pkg.params.init_params(arch, 1, seed=3, random_bias=True)
    dense_cloud = compact_scene(2, 16384, seed=5)
    for pts_np in (synth.kitti_like(2, 16384, seed=1234), dense_cloud):
        pts = torch.from_numpy(pts_np).to(cuda)
        net = pkg.SABackbone(arch, params, in_channels=1, device=cuda)
        assert lu.COMPACT_GROUPS
        a = net.forward(pts)
        try:
            lu.COMPACT_GROUPS = False
            b = net.forward(pts)
        finally:
            lu.COMPACT_GROUPS = True
        for li in range(1, len(arch) + 1):
            assert torch.equal(a[1][li], b[1][li]), "features of layer %d" % li


def test_backbone_full_size_properties_and_graph_replay(pkg, cuda):
    """Full configs[1] batch: shapes, invariants the domain offers, and CUDA-graph replay == eager."""
    net = pkg.SABackbone(device=cuda)
    pts = torch.from_numpy(synth.kitti_like(8, 16384, seed=1000)).to(cuda)
    xyz_l, feat_l, fps_l, dbg = net.forward(pts, return_debug=True)
    assert [tuple(f.shape) for f in feat_l[1:]] == [(8, 4096, 64), (8, 1024, 128), (8, 512, 256), (8, 256, 256),
                                                    (8, 256, 128), (8, 256, 512)]
    f1 = fps_l[1]
    assert (f1[:, 0] == 0).all()
    # D-FPS samples of distinct locations are distinct while unsampled distinct points remain
    assert all(len(set(row.tolist())) > 4000 for row in f1.cpu())
    for d in dbg[:3]:                                             # dilated layers: self hit -> cnt >= 1
        for c in d["cnt"]:
            assert int(c.min()) >= 1
    for f in feat_l[1:]:
        assert torch.isfinite(f).all()
    blk, cnt = net.detection_block(xyz_l, feat_l)
    assert blk.shape == (8, 100, 9) and cnt.shape == (8,)
    replay = net.capture(pts)
    out2, (blk2, cnt2) = replay()
    torch.cuda.synchronize()
    assert torch.equal(out2[1][-1], feat_l[-1]) and torch.equal(blk2, blk)
    pts2 = torch.from_numpy(synth.kitti_like(8, 16384, seed=2000)).to(cuda)
    out3, _ = replay(pts2)
    torch.cuda.synchronize()
    eager = net.forward(pts2)
    assert torch.equal(out3[1][-1], eager[1][-1]) and torch.equal(out3[2][1], eager[2][1])
