"""SA / FP / vote layers with the reference's signatures, on the B200 operators.

Mirrors /root/reference/lib/utils/layers_util.py: vote_layer :12-24, pointnet_sa_module :27-55,
pointnet_sa_module_msg :59-189, pointnet_fp_module :192-225 -- same positional arguments and return values,
torch CUDA tensors instead of TF tensors.  What TF resolved through variable scopes is passed explicitly as
`params` (a dict keyed by the reference's variable names, or a params.PreparedParams).  Inference only:
is_training must be False (training-mode BN / backward ops are SURVEY.md section 8f-3, not built yet).
"""
import torch

from . import config as _cfg
from . import tf_ops
from .params import prepare


def _conv(pp, scope, x, bn=True, relu=True, pool=1, rowmask=None):
    f = pp.conv(scope, bn)
    return tf_ops.linear_bn_relu(x, f.w, f.scale, f.shift, relu=relu, pool=pool, rowmask=rowmask, cin=f.cin)


_CONSTS = {}


def _const(values, device):
    """Small device constants, created once (so that layer calls stay CUDA-graph capturable)."""
    key = (values, str(device))
    if key not in _CONSTS:
        _CONSTS[key] = torch.tensor(values, dtype=torch.float32, device=device)
    return _CONSTS[key]


_SIDE_STREAMS = {}


def _side_stream(device):
    key = str(device)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def _arange_idx(bs, npoint, device):
    return torch.arange(npoint, dtype=torch.int32, device=device).unsqueeze(0).repeat(bs, 1)


def ffps_indices(npoint, xyz, points, mode):
    """F-FPS on concat[xyz, points] (layers_util.py:94-96, :102-104).
    mode 'direct': one kernel, no [B,N,N] tensor, the SAME indices as 'matrix' (each round evaluates the picked point's
                   matrix row on chip with calc_square_dist's arithmetic); falls back to 'matrix' for uncovered shapes;
    mode 'matrix': calc_square_dist + farthest_point_sample_with_distance, the reference's route;
    mode 'fused' : matrix-free -- the generic-c FPS kernel evaluates the feature distance on the fly (no
                   [B,N,N] tensor); identical to the reference's own farthest_point_sample on the features."""
    if mode == "direct":
        c = xyz.shape[2] + points.shape[2]
        # the 132-channel variant walks a 131-long dependent fma chain per round: at layer-3 sizes (512 points) the
        # small matrix is faster, so 'direct' is used where the matrix is the expensive part (c <= 68, layer 2)
        if c <= 68 and tf_ops.ffps_supported(xyz.shape[1], c):
            return tf_ops.farthest_point_sample_features(npoint, xyz, points)
        mode = "matrix"
    feats = torch.cat([xyz, points], dim=-1).contiguous()
    if mode == "matrix":
        return tf_ops.farthest_point_sample_with_distance(npoint, tf_ops.calc_square_dist(feats))
    if mode == "fused":
        return tf_ops.farthest_point_sample(npoint, feats)
    raise ValueError("ffps_mode must be 'direct', 'matrix' or 'fused'")


def pointnet_sa_module_msg(xyz, points, radius_list, nsample_list, mlp_list, is_training, bn_decay, bn,
                           fps_sample_range_list, fps_method_list, npoint_list, former_fps_idx, use_attention, scope,
                           dilated_group, vote_ctr=None, aggregation_channel=None, debugging=False, epsilon=1e-5, *,
                           params, ffps_mode="direct", aggregation=None, return_debug=False, mlp_mode="tc",
                           fuse_scale=True, gather_in_kernel=True, hoist_first=2):
    """PointNet++ SA module with multi-scale grouping; returns (new_xyz, new_points, fps_idx)."""
    if is_training:
        raise NotImplementedError("training-mode BatchNorm / backward ops are out of scope (inference only)")
    if use_attention:
        raise NotImplementedError("query_ball_point_withidx (use_attention) is unused by the shipped 3DSSD configs")
    pp = prepare(params, xyz.device)
    aggregation = _cfg.AGGREGATION_SA_FEATURE if aggregation is None else aggregation
    bs, n, _ = xyz.shape

    cur, last = [], 0
    join_side = False
    for rng, method, npoint in zip(fps_sample_range_list, fps_method_list, npoint_list):
        end = n if rng == -1 else last + rng                      # tf.slice size -1 (:86-87)
        tmp_xyz = xyz[:, last:end].contiguous()
        tmp_points = points[:, last:end]
        if npoint == 0:                                           # :88-90
            last += rng
            continue
        if vote_ctr is not None:                                  # :91-93
            npoint = vote_ctr.shape[1]
            fps_idx = _arange_idx(bs, npoint, xyz.device)
        elif method == "FS":                                      # :94-99 fusion sampling
            # F-FPS and D-FPS are independent latency-bound chains on a handful of SMs each: run them concurrently
            side = _side_stream(xyz.device)
            cur_s = torch.cuda.current_stream()
            side.wait_stream(cur_s)
            tmp_xyz.record_stream(side)
            with torch.cuda.stream(side):
                d_idx = tf_ops.farthest_point_sample(npoint, tmp_xyz)
            f_idx = ffps_indices(npoint, tmp_xyz, tmp_points, ffps_mode)
            cur_s.wait_stream(side)
            d_idx.record_stream(cur_s)
            fps_idx = torch.cat([f_idx, d_idx], dim=-1)
        elif npoint == tmp_xyz.shape[1]:                          # :100-101
            fps_idx = _arange_idx(bs, npoint, xyz.device)
        elif method == "F-FPS":                                   # :102-105
            fps_idx = ffps_indices(npoint, tmp_xyz, tmp_points, ffps_mode)
        elif len(npoint_list) > 1:                                # D-FPS segment next to other segments: side stream
            side = _side_stream(xyz.device)
            cur_s = torch.cuda.current_stream()
            side.wait_stream(cur_s)
            tmp_xyz.record_stream(side)
            with torch.cuda.stream(side):
                fps_idx = tf_ops.farthest_point_sample(npoint, tmp_xyz)
                if last:
                    fps_idx = fps_idx + last
            fps_idx.record_stream(cur_s)
            cur.append(fps_idx)
            last += rng
            join_side = True
            continue
        else:                                                     # D-FPS :106-107
            fps_idx = tf_ops.farthest_point_sample(npoint, tmp_xyz)
        cur.append(fps_idx + last if last else fps_idx)           # :109
        last += rng
    if join_side:
        torch.cuda.current_stream().wait_stream(_side_stream(xyz.device))
    fps_idx = cur[0] if len(cur) == 1 else torch.cat(cur, dim=-1)
    if former_fps_idx is not None:
        fps_idx = torch.cat([fps_idx, former_fps_idx], dim=-1)    # :113-114
    fps_idx = fps_idx.contiguous()
    new_xyz = tf_ops.gather_point(vote_ctr if vote_ctr is not None else xyz, fps_idx)   # :116-119

    debug = {"idx": [], "cnt": []}
    outs = []
    nscale = len(radius_list)
    if nscale:
        min_r = [0.0 if (i == 0 or not dilated_group) else radius_list[i - 1] for i in range(nscale)]   # :137-141
        if nscale <= 4:   # one pass over the candidates for all shells
            idx_list, cnt_list = tf_ops.query_ball_point_multi(min_r, radius_list, nsample_list, xyz, new_xyz,
                                                               dilated_group)
        else:
            idx_list, cnt_list = [], []
            for i in range(nscale):
                if dilated_group:
                    a, c = tf_ops.query_ball_point_dilated(min_r[i], radius_list[i], nsample_list[i], xyz, new_xyz)
                else:
                    a, c = tf_ops.query_ball_point(radius_list[i], nsample_list[i], xyz, new_xyz)
                idx_list.append(a); cnt_list.append(c)
        use_agg = bool(aggregation and aggregation_channel is not None and aggregation_channel != -1)
        tc = mlp_mode == "tc" and all(k in (8, 16, 32, 64, 128) for k in nsample_list)
        if mlp_mode not in ("tc", "fp32"):
            raise ValueError("mlp_mode must be 'tc' or 'fp32'")
        if tc:
            # tensor-core path: every scale pools straight into its slice of the concat buffer (fp32 for the
            # caller, split bf16 for the aggregation conv), activations stay split between layers
            m_q = new_xyz.shape[1]
            ctot = sum(m[-1] for m in mlp_list)
            concat = torch.empty((bs, m_q, ctot), dtype=torch.float32, device=xyz.device)
            cat_hi = cat_lo = None
            if use_agg:
                ldc = tf_ops.round16(ctot)
                mk = torch.zeros if ldc != ctot else torch.empty
                cat_hi = mk((bs, m_q, ldc), dtype=torch.bfloat16, device=xyz.device)
                cat_lo = mk((bs, m_q, ldc), dtype=torch.bfloat16, device=xyz.device)
            off = 0
            c_feat = points.shape[-1]
            stacks = [pp.fused_stack(["%s/conv%d_%d" % (scope, i, j) for j in range(len(mlp_list[i]))], bn, c_feat + 3)
                      if fuse_scale else None for i in range(nscale)]
            # The first conv of every scale with >= 2 convs is hoisted out of the grouped domain: its feature part becomes
            # ONE per-point GEMM for all scales of the layer (z), its xyz part is re-applied per grouped row while the
            # operand of the second conv is built (in the fused kernel's gather, or in the producer warps of
            # ssd3d_linear_tc_hoisted for the scales that run layer by layer).
            hoist = [i for i in range(nscale) if len(mlp_list[i]) >= 2 and (int(hoist_first) >= 2 or stacks[i] is None)] \
                if (gather_in_kernel and hoist_first) else []   # hoist_first: 0 off, 1 layer-by-layer scales only, 2 all
            z = zoffs = wxs = None
            hstacks = {}
            if hoist:
                zconv, wxs, n1s = pp.hoisted(["%s/conv%d_0" % (scope, i) for i in hoist], bn, c_feat)
                p_hi, p_lo = tf_ops.split_rows(points)
                z, _ = tf_ops.linear_tc(p_hi, p_lo, zconv, relu=False, want_f32=True, want_split=False)
                zoffs = [sum(n1s[:t]) for t in range(len(hoist))]
                for t, i in enumerate(hoist):
                    if stacks[i] is not None:    # remaining convs of a fused scale
                        hstacks[i] = pp.fused_stack(["%s/conv%d_%d" % (scope, i, j) for j in range(1, len(mlp_list[i]))], bn, n1s[t])
            for i in range(nscale):
                idx, cnt = idx_list[i], cnt_list[i]
                debug["idx"].append(idx); debug["cnt"].append(cnt)
                nl = len(mlp_list[i])
                stack = stacks[i]
                if stack is not None:                                              # whole scale in one kernel
                    if hstacks.get(i) is not None:
                        t = hoist.index(i)
                        tf_ops.sa_mlp_fused_hoisted(xyz, z, zoffs[t], wxs[t], new_xyz, idx, cnt, hstacks[i], out_f32=(concat, off),
                                                    out_split=(cat_hi, cat_lo, off) if use_agg else None)
                    else:
                        tf_ops.sa_mlp_fused(xyz, points, new_xyz, idx, cnt, stack, out_f32=(concat, off),
                                            out_split=(cat_hi, cat_lo, off) if use_agg else None)
                    off += mlp_list[i][-1]
                    continue
                hi = lo = None
                for j in range(nl):
                    f = pp.conv("%s/conv%d_%d" % (scope, i, j), bn)
                    last_kw = dict(pool=nsample_list[i], rowmask=cnt, out_f32=(concat, off),
                                   out_split=(cat_hi, cat_lo, off) if use_agg else None)   # :167-180 conv+BN+ReLU+max+mask
                    if i in hoist:
                        if j == 0:
                            continue                  # folded into z and into the next conv's operand producer
                        if j == 1:
                            t = hoist.index(i)
                            if nl == 2:
                                tf_ops.linear_tc_hoisted(xyz, z, zoffs[t], wxs[t], new_xyz, idx, f, want_split=False, **last_kw)
                            else:
                                _, (hi, lo) = tf_ops.linear_tc_hoisted(xyz, z, zoffs[t], wxs[t], new_xyz, idx, f)
                            continue
                    if j == 0 and gather_in_kernel:   # :160-165 inside the kernel's operand load (no [B,M,K,C] tensor)
                        if nl == 1:
                            tf_ops.linear_tc_gather(xyz, points, new_xyz, idx, f, want_split=False, **last_kw)
                        else:
                            _, (hi, lo) = tf_ops.linear_tc_gather(xyz, points, new_xyz, idx, f)
                        continue
                    if j == 0:                        # :160-165 fused with the split, materialised once in bf16 hi/lo
                        hi, lo = tf_ops.group_concat_split(xyz, points, new_xyz, idx)
                    if j < nl - 1:
                        _, (hi, lo) = tf_ops.linear_tc(hi, lo, f, want_f32=False, want_split=True)
                    else:
                        tf_ops.linear_tc(hi, lo, f, **last_kw)
                off += mlp_list[i][-1]
            new_points = concat
            if use_agg:                                                            # :183-185
                new_points, _ = tf_ops.linear_tc(cat_hi, cat_lo, pp.conv(scope + "/ensemble", bn))
        else:
            for i in range(nscale):
                idx, cnt = idx_list[i], cnt_list[i]
                # rows with cnt == 0 come back zero-filled, which is what idx * (cnt > 0) produces (:157-159)
                debug["idx"].append(idx); debug["cnt"].append(cnt)
                g = tf_ops.group_concat(xyz, points, new_xyz, idx)                    # :160-165 fused
                nl = len(mlp_list[i])
                for j in range(nl):
                    lastl = j == nl - 1
                    g = _conv(pp, "%s/conv%d_%d" % (scope, i, j), g, bn=bn,
                              pool=nsample_list[i] if lastl else 1, rowmask=cnt if lastl else None)   # :167-180
                outs.append(g)
            new_points = outs[0] if len(outs) == 1 else torch.cat(outs, dim=-1)
            if use_agg:
                new_points = _conv(pp, scope + "/ensemble", new_points, bn=bn)        # :183-185
    else:
        new_points = tf_ops.gather_point(points.contiguous(), fps_idx)           # :186-187
    if return_debug:
        return new_xyz, new_points, fps_idx, debug
    return new_xyz, new_points, fps_idx


def pointnet_sa_module(xyz, points, mlp, is_training, bn_decay, bn, scope, *, params):
    """Global SA layer (layer type SA_Layer_SSG_Last): concat[xyz, points] -> MLP -> max over all points."""
    if is_training:
        raise NotImplementedError("inference only")
    pp = prepare(params, xyz.device)
    g = torch.cat([xyz, points], dim=-1).contiguous()              # xyz FIRST here (:42)
    n = g.shape[1]
    for j in range(len(mlp)):
        g = _conv(pp, "%s/conv%d" % (scope, j), g, bn=bn)
    return g.max(dim=1).values if n else g


def pointnet_fp_module(xyz1, xyz2, points1, points2, mlp, is_training, bn_decay, scope, bn=True, *, params):
    """Feature propagation: inverse-distance interpolation from (xyz2, points2) onto xyz1, then an MLP."""
    if is_training:
        raise NotImplementedError("inference only")
    pp = prepare(params, xyz1.device)
    dist, idx = tf_ops.three_nn(xyz1, xyz2)
    dist = torch.clamp_min(dist, 1e-10)                            # :207
    inv = 1.0 / dist
    weight = inv / inv.sum(dim=2, keepdim=True)                    # :208-210
    x = tf_ops.three_interpolate(points2, idx, weight.contiguous())
    if points1 is not None:
        x = torch.cat([x, points1], dim=2).contiguous()            # :213-214
    for i in range(len(mlp)):
        x = _conv(pp, "%s/conv_%d" % (scope, i), x, bn=bn)
    return x


def vote_layer(xyz, points, mlp_list, is_training, bn_decay, bn, scope, *, params,
               max_translate_range=_cfg.MAX_TRANSLATE_RANGE, mlp_mode="tc"):
    """Vote layer: per-point MLP -> 3 offsets, clamped to +-max_translate_range (layers_util.py:12-24)."""
    if is_training:
        raise NotImplementedError("inference only")
    pp = prepare(params, xyz.device)
    if mlp_mode == "tc":
        hi, lo = tf_ops.split_rows(points)
        for i in range(len(mlp_list)):
            points, (hi, lo) = tf_ops.linear_tc(hi, lo, pp.conv("%s/vote_layer_%d" % (scope, i), bn), want_split=True)
        off, _ = tf_ops.linear_tc(hi, lo, pp.conv(scope + "/vote_offsets", False), relu=False)
    else:
        for i in range(len(mlp_list)):
            points = _conv(pp, "%s/vote_layer_%d" % (scope, i), points, bn=bn)
        off = _conv(pp, scope + "/vote_offsets", points, bn=False, relu=False)
    lo = _const(tuple(max_translate_range), xyz.device).view(1, 1, 3)
    lim = torch.minimum(torch.maximum(off, lo), -lo)
    return xyz + lim, points, off
