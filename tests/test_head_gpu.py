"""GPU tests (-m gpu) of the detection head / decode / BEV-NMS rows against the numpy oracle."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
synth = importlib.import_module("3dssd_b200.synth")


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


def random_boxes(rng, b, n, spread=20.0):
    ctr = rng.uniform(-spread, spread, (b, n, 3)).astype(np.float32)
    lhw = rng.uniform(0.5, 4.5, (b, n, 3)).astype(np.float32)
    ry = rng.uniform(-np.pi, np.pi, (b, n, 1)).astype(np.float32)
    return np.concatenate([ctr, lhw, ry], -1)


@pytest.mark.parametrize("n,spread,max_out", [(256, 12.0, 100), (256, 3.0, 100), (100, 40.0, 20), (7, 1.0, 100), (512, 15.0, 50)])
def test_bev_nms_vs_oracle(pkg, cuda, n, spread, max_out):
    from oracle import head as ohead
    rng = np.random.default_rng(n)
    boxes = random_boxes(rng, 3, n, spread)
    scores = rng.uniform(0, 1, (3, n)).astype(np.float32)
    scores[:, 5:9] = scores[:, 4:5]                               # ties: lower index first
    eb, ec = ohead.bev_nms(boxes, scores, 0.1, max_out)
    gb, gc = pkg.bev_nms(torch.from_numpy(boxes).to(cuda), torch.from_numpy(scores).to(cuda), 0.1, max_out)
    np.testing.assert_array_equal(gc.cpu().numpy(), ec)
    np.testing.assert_array_equal(gb.cpu().numpy(), eb)
    # size-independent properties: scores descending, kept boxes mutually below the IoU threshold, zero padding
    blk, cnt = gb.cpu().numpy(), gc.cpu().numpy()
    for s in range(3):
        k = cnt[s]
        assert (np.diff(blk[s, :k, 7]) <= 0).all() and (blk[s, k:] == 0).all() and 1 <= k <= max_out


def test_head_decode_nms_vs_oracle(pkg, oracle_ops, cuda):
    from oracle import head as ohead
    rng = np.random.default_rng(3)
    xyz = rng.uniform(-10, 10, (2, 256, 3)).astype(np.float32)
    feat = np.maximum(rng.standard_normal((2, 256, 512)), 0).astype(np.float32)
    prm = pkg.params.init_head_params(512, seed=5)
    head = pkg.DetectionHead(params=prm, device=cuda)
    blk, cnt, raw = head.forward([None] * 6 + [torch.from_numpy(xyz).to(cuda)], [None] * 6 + [torch.from_numpy(feat).to(cuda)],
                                 return_raw=True)
    eboxes, escore, eraw = ohead.head_forward(xyz, feat, prm)
    assert rel_err(raw["feat"].cpu().numpy(), eraw["feat"]) < 1e-4
    assert rel_err(raw["cls"].cpu().numpy(), eraw["cls"]) < 1e-4
    # decoding is discontinuous in the angle bin (argmax): compare where the oracle's two best bins are well separated
    top2 = np.sort(eraw["reg"][..., 6:18], axis=-1)[..., -2:]
    stable = (top2[..., 1] - top2[..., 0]) > 1e-3
    gboxes = raw["boxes"][:, :, 0].cpu().numpy()
    assert stable.mean() > 0.95
    assert np.abs(gboxes[stable] - eboxes[stable]).max() < 2e-3
    assert rel_err(raw["score"][..., 0].cpu().numpy(), escore) < 1e-4
    # NMS on identical inputs (the GPU-decoded boxes) must agree exactly
    eb, ec = ohead.bev_nms(gboxes, raw["score"][..., 0].cpu().numpy(), 0.1, 100)
    np.testing.assert_array_equal(cnt.cpu().numpy(), ec)
    np.testing.assert_array_equal(blk.cpu().numpy(), eb)


def test_backbone_with_head_and_graph(pkg, cuda):
    cfg, P = pkg.config, pkg.params
    prm = dict(P.init_params(cfg.ARCH_3DSSD, 1, seed=0))
    prm.update(P.init_head_params(512, seed=1))
    head = pkg.DetectionHead(params=prm, device=cuda)
    net = pkg.SABackbone(params=prm, device=cuda, head=head)
    pts = torch.from_numpy(synth.kitti_like(4, 16384, seed=1000)).to(cuda)
    out = net.forward(pts)
    blk, cnt = net.detections(out[0], out[1])
    assert blk.shape == (4, 100, 9) and cnt.shape == (4,) and int(cnt.min()) >= 1 and torch.isfinite(blk).all()
    replay = net.capture(pts)
    _, (blk2, cnt2) = replay()
    torch.cuda.synchronize()
    assert torch.equal(blk2, blk) and torch.equal(cnt2, cnt)
