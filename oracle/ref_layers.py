"""The reference's SA backbone as the reference itself would execute it on a GPU: its OWN CUDA kernels
(oracle/_ref/libref_ops.so, compiled unmodified from /root/reference) for sampling / grouping, and -- because
TensorFlow 1.4 cannot be installed here -- PyTorch fp32 ops standing in one-for-one for the TF stock ops of the
MLP (tf.nn.conv2d, bias_add, unfused batch_norm, relu, reduce_max, mask multiply; lib/utils/tf_util.py:127-201,
:424-444, lib/utils/layers_util.py:160-185) and of calc_square_dist (lib/utils/model_util.py:144-160).

TEST / BENCH INFRASTRUCTURE ONLY: this is the `--impl reference` arm of bench.py and the "reference single-GPU
SA-layer latency" of BASELINE.md 2a.  None of the product's kernels are on this path.
"""
import torch

from . import ref_ops

BN_EPS = 1e-3


def _t(params, key, dev, cache):
    k = (key, str(dev))
    if k not in cache:
        cache[k] = torch.from_numpy(params[key]).to(dev)
    return cache[k]


def _conv(params, scope, x, bn, relu, cache):
    dev = x.device
    w = _t(params, scope + "/weights", dev, cache)
    y = torch.matmul(x, w)                                   # tf.nn.conv2d 1x1
    y = y + _t(params, scope + "/biases", dev, cache)        # tf.nn.bias_add
    if bn:                                                   # tf.contrib.layers.batch_norm(fused=False), inference
        g = _t(params, scope + "/bn/gamma", dev, cache)
        be = _t(params, scope + "/bn/beta", dev, cache)
        mu = _t(params, scope + "/bn/moving_mean", dev, cache)
        var = _t(params, scope + "/bn/moving_variance", dev, cache)
        inv = g * torch.rsqrt(var + BN_EPS)
        y = y * inv + (be - mu * inv)
    if relu:
        y = torch.relu(y)
    return y


def calc_square_dist(a):
    a_sq = (a * a).sum(-1, keepdim=True)
    return a_sq + a_sq.transpose(1, 2) - 2.0 * torch.matmul(a, a.transpose(1, 2))


def sa_module_msg(xyz, points, radius_list, nsample_list, mlp_list, bn, fps_sample_range_list, fps_method_list,
                  npoint_list, scope, dilated_group, vote_ctr, aggregation_channel, params, cache, aggregation=True):
    bs, n, _ = xyz.shape
    cur, last = [], 0
    for rng, method, npoint in zip(fps_sample_range_list, fps_method_list, npoint_list):
        end = n if rng == -1 else last + rng
        tmp_xyz = xyz[:, last:end].contiguous()
        tmp_points = points[:, last:end]
        if npoint == 0:
            last += rng
            continue
        if vote_ctr is not None:
            npoint = vote_ctr.shape[1]
            fps_idx = torch.arange(npoint, dtype=torch.int32, device=xyz.device).unsqueeze(0).repeat(bs, 1)
        elif method == "FS":
            f = torch.cat([tmp_xyz, tmp_points], -1)
            i1 = ref_ops.farthest_point_sample_with_distance(npoint, calc_square_dist(f).contiguous(), sync=False)
            i2 = ref_ops.farthest_point_sample(npoint, tmp_xyz, sync=False)
            fps_idx = torch.cat([i1, i2], -1)
        elif npoint == tmp_xyz.shape[1]:
            fps_idx = torch.arange(npoint, dtype=torch.int32, device=xyz.device).unsqueeze(0).repeat(bs, 1)
        elif method == "F-FPS":
            f = torch.cat([tmp_xyz, tmp_points], -1)
            fps_idx = ref_ops.farthest_point_sample_with_distance(npoint, calc_square_dist(f).contiguous(), sync=False)
        else:
            fps_idx = ref_ops.farthest_point_sample(npoint, tmp_xyz, sync=False)
        cur.append(fps_idx + last)
        last += rng
    fps_idx = torch.cat(cur, -1).contiguous()
    new_xyz = ref_ops.gather_point((vote_ctr if vote_ctr is not None else xyz).contiguous(), fps_idx, sync=False)
    outs = []
    for i, (radius, nsample) in enumerate(zip(radius_list, nsample_list)):
        if dilated_group:
            min_r = 0.0 if i == 0 else radius_list[i - 1]
            idx, cnt = ref_ops.query_ball_point_dilated(min_r, radius, nsample, xyz, new_xyz, sync=False)
        else:
            idx, cnt = ref_ops.query_ball_point(radius, nsample, xyz, new_xyz, sync=False)
        mask = (cnt > 0).to(torch.int32)
        idx = (idx * mask.unsqueeze(-1)).contiguous()
        g_xyz = ref_ops.group_point(xyz, idx, sync=False) - new_xyz.unsqueeze(2)
        g = torch.cat([ref_ops.group_point(points.contiguous(), idx, sync=False), g_xyz], -1)
        for j in range(len(mlp_list[i])):
            g = _conv(params, "%s/conv%d_%d" % (scope, i, j), g, bn, True, cache)
        outs.append(g.max(dim=2).values * mask.unsqueeze(-1).to(torch.float32))
    if outs:
        new_points = torch.cat(outs, -1)
        if aggregation and aggregation_channel is not None and aggregation_channel != -1:
            new_points = _conv(params, scope + "/ensemble", new_points, bn, True, cache)
    else:
        new_points = ref_ops.gather_point(points.contiguous(), fps_idx, sync=False)
    return new_xyz, new_points, fps_idx


def backbone_forward(arch, points_in, params, cache=None):
    """All launches go to the legacy default stream (the reference kernels have no stream argument), so the
    caller must run this on torch's default stream and synchronise around it."""
    cache = {} if cache is None else cache
    xyz_list = [points_in[..., :3].contiguous()]
    feat_list = [points_in[..., 3:].contiguous()]
    for spec in arch:
        (xyz_i, feat_i, radius, nsample, mlps, bn, rng, method, npoint, former, attn, ltype, scope, dilated,
         vote_idx, agg) = spec
        vote_ctr = xyz_list[vote_idx] if vote_idx != -1 else None
        if ltype == "SA_Layer":
            nx, nf, _ = sa_module_msg(xyz_list[xyz_i[0]], feat_list[feat_i[0]], radius, nsample, mlps, bn, rng, method,
                                      npoint, scope, dilated, vote_ctr, agg, params, cache)
            xyz_list.append(nx); feat_list.append(nf)
        elif ltype == "Vote_Layer":
            pts = feat_list[feat_i[0]]
            for i in range(len(mlps)):
                pts = _conv(params, "%s/vote_layer_%d" % (scope, i), pts, bn, True, cache)
            off = _conv(params, scope + "/vote_offsets", pts, False, False, cache)
            lo = torch.tensor([-3.0, -2.0, -3.0], device=off.device).view(1, 1, 3)
            xyz_list.append(xyz_list[xyz_i[0]] + torch.minimum(torch.maximum(off, lo), -lo)); feat_list.append(pts)
        else:
            raise ValueError("reference arm covers the 3DSSD backbone layer types only, got %r" % (ltype,))
    return xyz_list, feat_list


def head_forward(xyz, feat, params, cache=None, mlp=(128,), bn=True, scope="", angle_bins=12, max_output=100,
                 nms_threshold=0.1):
    """Detection head + decode + post-processing as the reference runs them (lib/modeling/head_builder.py:81-114,
    lib/utils/head_util.py:26-59, lib/utils/anchor_decoder.py:6-14, :86-112, single_stage_detector.py:210-211,
    lib/builder/postprocessor.py:52-120): conv1d stacks and the decode as eager fp32 ops one-for-one on the GPU, then
    -- like tf.image.non_max_suppression in the reference -- the greedy BEV NMS on the CPU (oracle C twin), which
    puts a device->host sync in the step.  Returns (block [B,100,9], count [B]) numpy."""
    import math

    from . import ops
    cache = {} if cache is None else cache
    pre = "" if scope == "" else scope + "/"
    y = feat
    for i in range(len(mlp)):
        y = _conv(params, "%sconv1d_%d" % (pre, i), y, bn, True, cache)
    cls = _conv(params, pre + "pred_cls", _conv(params, pre + "pred_cls_base", y, bn, True, cache), False, False, cache)
    reg = _conv(params, pre + "pred_reg", _conv(params, pre + "pred_reg_base", y, bn, True, cache), False, False, cache)
    off, acls, ares = reg[..., :6], reg[..., 6:6 + angle_bins], reg[..., 6 + angle_bins:]
    bins = torch.argmax(acls, dim=-1)
    res = torch.gather(ares, -1, bins.unsqueeze(-1)).squeeze(-1)
    angle = ((bins.to(torch.float32) + res + 0.0) * (2 * math.pi / angle_bins)).unsqueeze(-1)
    translate, half = off[..., :3], off[..., 3:6]
    ctr = xyz + translate
    pad = torch.zeros_like(half)
    pad[..., 1] = half[..., 1]
    ctr = ctr + pad
    lhw = torch.clamp_min(half * 2.0, 0.1)
    boxes = torch.cat([ctr, lhw, angle], dim=-1)
    score = torch.sigmoid(cls)[..., 0]
    return ops.bev_nms(boxes.cpu().numpy(), score.cpu().numpy(), nms_threshold, max_output)
