"""One small launch of every kernel family of libssd3d.so, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py [names...]
Sizes are the unit-test sizes (racecheck slows kernels ~100x); results are checked against the oracle where cheap so a
silent wrong answer under the tool is also caught."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dssd_b200")
P = pkg.params
synth = importlib.import_module("3dssd_b200.synth")
from oracle import ops as oops  # noqa: E402

dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rng = np.random.default_rng(0)


def fps3_direct():                       # cluster of 4 and 8, resident-scene kernel, incl. a resumed launch
    pts = synth.kitti_like(2, 4096, seed=3)[..., :3].copy()
    exp = oops.farthest_point_sample(96, pts)
    for cl in (4, 8):
        got = pkg.farthest_point_sample(96, T(pts), cluster=cl)
        assert np.array_equal(got.cpu().numpy(), exp)
    buf = torch.zeros((2, 96), dtype=torch.int32, device=dev)
    temp = torch.empty((2, 4096), device=dev)
    for j0, j1 in ((0, 40), (40, 96)):
        pkg.farthest_point_sample(96, T(pts), out=(buf, 0), rounds=(j0, j1), temp=temp)
    assert np.array_equal(buf.cpu().numpy(), exp)


def fps3_packet():
    pts = synth.kitti_like(2, 4096, seed=4)[..., :3].copy()
    got = pkg.farthest_point_sample(64, T(pts), cluster=4, packet_kernel=True)
    assert np.array_equal(got.cpu().numpy(), oops.farthest_point_sample(64, pts))


def ffps_cluster():
    f = rng.standard_normal((2, 2048, 67)).astype(np.float32)
    xyz, feat = T(f[..., :3]), T(f[..., 3:])
    a = pkg.tf_ops.farthest_point_sample_features(48, xyz, feat)
    b = pkg.farthest_point_sample_with_distance(48, pkg.calc_square_dist(T(f)))
    assert torch.equal(a, b)


def fpsdist_and_generic():
    f = rng.standard_normal((2, 1024, 19)).astype(np.float32)
    got = pkg.farthest_point_sample(40, T(f))
    assert np.array_equal(got.cpu().numpy(), oops.farthest_point_sample(40, f))


def ball_query():
    pts = synth.kitti_like(2, 2048, seed=5)[..., :3].copy()
    q = pts[:, :128].copy()
    idxs, cnts = pkg.query_ball_point_multi([0.0, 0.4], [0.4, 0.8], [16, 32], T(pts), T(q), True)
    e = oops.query_ball_point_dilated(0.4, 0.8, 32, pts, q)
    assert np.array_equal(idxs[1].cpu().numpy(), e[0]) and np.array_equal(cnts[1].cpu().numpy(), e[1])
    pkg.query_ball_point(0.5, 16, T(pts), T(q))


def _stack(c, mlp):
    prm, scopes, cin = {}, [], c + 3
    for j, co in enumerate(mlp):
        P._conv_init(rng, prm, "s/conv0_%d" % j, cin, co, True)
        scopes.append("s/conv0_%d" % j)
        cin = co
    return prm, scopes


def sa_fused():
    for c, mlp, k in ((1, [16, 16, 32], 32), (64, [64, 64, 128], 32)):
        pts = synth.kitti_like(2, 1024, seed=6)[..., :3].copy()
        xyz, feat = T(pts), T(rng.standard_normal((2, 1024, c)).astype(np.float32))
        new_xyz = xyz[:, :64].contiguous()
        idx, cnt = pkg.query_ball_point(2.0, k, xyz, new_xyz)
        prm, scopes = _stack(c, mlp)
        pp = P.prepare(prm, dev)
        stack = pp.fused_stack(scopes, True, c + 3, limit=0)
        y = pkg.tf_ops.sa_mlp_fused(xyz, feat, new_xyz, idx, cnt, stack)
        zconv, wxs, n1s = pp.hoisted([scopes[0]], True, c)
        hst = pp.fused_stack(scopes[1:], True, n1s[0], limit=0)
        hi, lo = pkg.tf_ops.split_rows(feat)
        z, _ = pkg.tf_ops.linear_tc(hi, lo, zconv, relu=False, want_f32=True, want_split=False)
        y2 = pkg.tf_ops.sa_mlp_fused_hoisted(xyz, z, 0, wxs[0], new_xyz, idx, cnt, hst)
        assert float((y - y2).abs().max()) <= 1e-3 * float(y.abs().max())


def linear_tc():
    x = T(rng.standard_normal((4, 32, 256)).astype(np.float32))
    prm = {}
    P._conv_init(rng, prm, "a", 256, 512, True)
    f = P.fold(prm, "a", True, dev)
    hi, lo = pkg.tf_ops.split_rows(x)
    y, _ = pkg.tf_ops.linear_tc(hi, lo, f, pool=32)
    ref = torch.relu((x.reshape(-1, 256) @ f.w) * f.scale + f.shift).view(4, 32, 512).max(1).values
    assert float((y - ref).abs().max()) <= 1e-3 * float(ref.abs().max())
    # gather / hoisted producer modes
    pts = synth.kitti_like(2, 512, seed=7)[..., :3].copy()
    xyz, feat = T(pts), T(rng.standard_normal((2, 512, 128)).astype(np.float32))
    new_xyz = xyz[:, :32].contiguous()
    idx, cnt = pkg.query_ball_point(3.0, 32, xyz, new_xyz)
    prm, scopes = _stack(128, [128, 256])
    pp = P.prepare(prm, dev)
    pkg.tf_ops.linear_tc_gather(xyz, feat, new_xyz, idx, pp.conv(scopes[0], True))
    zconv, wxs, n1s = pp.hoisted([scopes[0]], True, 128)
    hi, lo = pkg.tf_ops.split_rows(feat)
    z, _ = pkg.tf_ops.linear_tc(hi, lo, zconv, relu=False, want_f32=True, want_split=False)
    pkg.tf_ops.linear_tc_hoisted(xyz, z, 0, wxs[0], new_xyz, idx, pp.conv(scopes[1], True), pool=32, rowmask=cnt, want_split=False,
                                 want_f32=True)


def small_ops():
    pts = T(synth.kitti_like(2, 512, seed=8))
    xyz, feat = pkg.tf_ops.split_points(pts)
    pkg.three_nn(xyz, xyz[:, :64].contiguous())
    boxes = T(np.concatenate([rng.uniform(-5, 5, (2, 200, 3)), rng.uniform(1, 3, (2, 200, 3)), rng.uniform(-3, 3, (2, 200, 1))], -1).astype(np.float32))
    pkg.bev_nms(boxes, T(rng.uniform(0, 1, (2, 200)).astype(np.float32)), 0.1, 100)
    pkg.tf_ops.decode_dist_anchor_free(xyz[:, :64].contiguous(), T(rng.standard_normal((2, 64, 30)).astype(np.float32)),
                                       T(rng.standard_normal((2, 64, 1)).astype(np.float32)))
    pkg.tf_ops.concat_rows([xyz[:, :100].contiguous(), xyz[:, 100:300].contiguous()])


def fps_bucket():                        # single-CTA D-FPS with spatial pruning, incl. a resumed launch
    pts = synth.kitti_like(1, 2048, seed=9)[..., :3].copy()
    exp = oops.farthest_point_sample(80, pts)
    got = pkg.farthest_point_sample(80, T(pts), bucket_kernel=True)
    assert np.array_equal(got.cpu().numpy(), exp)
    buf = torch.zeros((1, 80), dtype=torch.int32, device=dev)
    temp = torch.empty((1, pkg.tf_ops.fps_temp_elems(2048, 3, 80, bucket_kernel=True)), device=dev)
    for j0, j1 in ((0, 30), (30, 80)):
        pkg.farthest_point_sample(80, T(pts), out=(buf, 0), rounds=(j0, j1), temp=temp, bucket_kernel=True)
    assert np.array_equal(buf.cpu().numpy(), exp)


def unit_lists():                        # culled ball query emitting unit lists -> fused and layer-by-layer kernels on them
    pts = synth.kitti_like(2, 2048, seed=10)[..., :3].copy()
    xyz, feat = T(pts), T(np.maximum(rng.standard_normal((2, 2048, 64)), 0).astype(np.float32))
    new_xyz = xyz[:, :96].contiguous()
    for grid in (True, False):
        (idx,), (cnt,), (units,) = pkg.query_ball_point_multi([0.0], [1.0], [32], xyz, new_xyz, False, grid=grid, return_units=True)
        prm, scopes = _stack(64, [64, 64, 128])
        pp = P.prepare(prm, dev)
        zconv, wxs, n1s = pp.hoisted([scopes[0]], True, 64)
        hi, lo = pkg.tf_ops.split_rows(feat)
        z, _ = pkg.tf_ops.linear_tc(hi, lo, zconv, relu=False, want_f32=True, want_split=False)
        hst = pp.fused_stack(scopes[1:], True, n1s[0], limit=0)
        dense = torch.zeros((2, 96, 128), device=dev); comp = torch.zeros_like(dense); lay = torch.zeros_like(dense)
        pkg.tf_ops.sa_mlp_fused_hoisted(xyz, z, 0, wxs[0], new_xyz, idx, cnt, hst, out_f32=(dense, 0))
        pkg.tf_ops.sa_mlp_fused_hoisted(xyz, z, 0, wxs[0], new_xyz, idx, cnt, hst, out_f32=(comp, 0), units=units)
        _, (h2, l2) = pkg.tf_ops.linear_tc_hoisted(xyz, z, 0, wxs[0], new_xyz, idx, pp.conv(scopes[1], True), units=units)
        pkg.tf_ops.linear_tc(h2, l2, pp.conv(scopes[2], True), out_f32=(lay, 0), units=units, unit_pool=True)
        eh, el = pkg.tf_ops.hoist_expand_split(xyz, z, 0, wxs[0], new_xyz, idx, units=units)
        assert torch.equal(dense, comp) and torch.equal(dense, lay)
        nu = int(units[0].item()) * 8
        assert eh.shape[-1] == 64 and nu > 0


def peer_gather():                       # exchange kernel; two simulated ranks on two streams need CONCURRENT kernels, which
    # the sanitizer tools do not give (they serialise launches): under a tool run the one-rank form (SSD3D_PROBE_SERIAL=1)
    world, sb = (1 if os.environ.get("SSD3D_PROBE_SERIAL") else 2), 4096
    recv_off, flag_off, total = pkg.dist.peer_layout(world, sb)
    sym = [torch.zeros((total,), dtype=torch.uint8, device=dev) for _ in range(world)]
    peers = torch.tensor([t.data_ptr() for t in sym], dtype=torch.int64, device=dev)
    send = [torch.randint(0, 255, (sb,), dtype=torch.uint8, device=dev) for _ in range(world)]
    out = [torch.zeros((world * sb,), dtype=torch.uint8, device=dev) for _ in range(world)]
    state = [torch.zeros((4,), dtype=torch.int32, device=dev) for _ in range(world)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    torch.cuda.synchronize()
    for _ in range(3):
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                pkg.tf_ops.peer_allgather(send[r], peers, world, r, sb, recv_off, flag_off, state[r], out[r])
        torch.cuda.synchronize()
    assert all(torch.equal(o, torch.cat(send)) for o in out) and all(int(s_[2].item()) == 0 for s_ in state)


ALL = [fps3_direct, fps3_packet, ffps_cluster, fpsdist_and_generic, ball_query, sa_fused, linear_tc, small_ops, fps_bucket,
       unit_lists, peer_gather]
if __name__ == "__main__":
    sel = sys.argv[1:]
    for fn in ALL:
        if sel and fn.__name__ not in sel:
            continue
        fn()
        torch.cuda.synchronize()
        print("ok", fn.__name__, flush=True)
