"""numpy restatement of the reference's SA / FP / vote layers on top of the CPU oracle ops.

TEST INFRASTRUCTURE ONLY (see ssd3d_oracle.c).  Follows /root/reference/lib/utils/layers_util.py:
vote_layer :12-24, pointnet_sa_module :27-55, pointnet_sa_module_msg :59-189, pointnet_fp_module :192-225,
and the layer dispatch of /root/reference/lib/builder/layer_builder.py:45-101.

Weights are a flat dict keyed by the reference's TF variable names:
  <scope>/conv{i}_{j}/weights [cin,cout], .../biases [cout],
  <scope>/conv{i}_{j}/bn/{gamma,beta,moving_mean,moving_variance} [cout]
"""
import numpy as np

from . import ops


BN_EPS = 1e-3        # tf.contrib.layers.batch_norm default epsilon


def _conv(params, scope, x, bn=True, relu=True, train=None):
    """train = (decay, updates dict) selects training-mode BatchNorm (tf_util.py:424-444 with is_training=True:
    tf.contrib.layers.batch_norm(fused=False, updates_collections=None)): batch moments over every axis but the channel
    (population variance), y = x*inv + (beta - mean*inv) with inv = gamma*rsqrt(var + 0.001), and the moving statistics
    after assign_moving_average (v -= (v - batch) * (1 - decay)) recorded in updates[scope]; all in float64."""
    w = params[scope + "/weights"]
    b = params.get(scope + "/biases")
    if bn and train is not None:
        decay, updates = train
        y = ops.linear_bn_relu(x, w, b, None, False).astype(np.float64)
        flat = y.reshape(-1, y.shape[-1])
        mean = flat.mean(axis=0)
        var = ((flat - mean) ** 2).mean(axis=0)
        g, be = (params[scope + "/bn/" + k].astype(np.float64) for k in ("gamma", "beta"))
        inv = g / np.sqrt(var + BN_EPS)
        out = y * inv + (be - mean * inv)
        if relu:
            out = np.maximum(out, 0.0)
        mm, mv = (params[scope + "/bn/" + k].astype(np.float64) for k in ("moving_mean", "moving_variance"))
        updates[scope] = {"moving_mean": (mm - (mm - mean) * (1.0 - decay)).astype(np.float32),
                          "moving_variance": (mv - (mv - var) * (1.0 - decay)).astype(np.float32),
                          "batch_mean": mean.astype(np.float32), "batch_variance": var.astype(np.float32)}
        return out.astype(np.float32)
    bnp = None
    if bn:
        bnp = tuple(params[scope + "/bn/" + k] for k in ("gamma", "beta", "moving_mean", "moving_variance"))
    return ops.linear_bn_relu(x, w, b, bnp, relu)


def ffps_indices(npoint, xyz, points, mode):
    """F-FPS of layers_util.py:94-96 / :102-104 on concat[xyz, points]."""
    feats = np.concatenate([xyz, points], axis=-1)
    if mode == "matrix":      # faithful route: distance matrix, then farthest_point_sample_with_distance
        return ops.farthest_point_sample_with_distance(npoint, ops.calc_square_dist(feats))
    if mode == "fused":       # matrix-free route == the reference's own generic-c FPS kernel on the features
        return ops.farthest_point_sample(npoint, feats)
    raise ValueError(mode)


def pointnet_sa_module_msg(xyz, points, radius_list, nsample_list, mlp_list, is_training, bn_decay, bn,
                           fps_sample_range_list, fps_method_list, npoint_list, former_fps_idx, use_attention,
                           scope, dilated_group, vote_ctr=None, aggregation_channel=None, *, params,
                           ffps_mode="matrix", aggregation=True, return_debug=False, bn_updates=None):
    """is_training=True: batch-statistics BatchNorm; the updated moving statistics land in bn_updates[scope]."""
    assert not use_attention
    train = None
    if is_training:
        train = (0.9 if bn_decay is None else float(bn_decay), {} if bn_updates is None else bn_updates)
    bs, n, _ = xyz.shape
    cur, last = [], 0
    for rng, method, npoint in zip(fps_sample_range_list, fps_method_list, npoint_list):
        end = n if rng == -1 else last + rng                       # tf.slice size -1 == "to the end" (:86-87)
        tmp_xyz, tmp_points = xyz[:, last:end], points[:, last:end]
        if npoint == 0:                                            # :88-90
            last += rng
            continue
        if vote_ctr is not None:                                   # :91-93
            npoint = vote_ctr.shape[1]
            fps_idx = np.tile(np.arange(npoint, dtype=np.int32)[None], (bs, 1))
        elif method == "FS":                                       # :94-99
            i1 = ffps_indices(npoint, tmp_xyz, tmp_points, ffps_mode)
            i2 = ops.farthest_point_sample(npoint, tmp_xyz)
            fps_idx = np.concatenate([i1, i2], axis=-1)
        elif npoint == tmp_xyz.shape[1]:                           # :100-101
            fps_idx = np.tile(np.arange(npoint, dtype=np.int32)[None], (bs, 1))
        elif method == "F-FPS":                                    # :102-105
            fps_idx = ffps_indices(npoint, tmp_xyz, tmp_points, ffps_mode)
        else:                                                      # D-FPS :106-107
            fps_idx = ops.farthest_point_sample(npoint, tmp_xyz)
        cur.append((fps_idx + last).astype(np.int32))              # :109
        last += rng
    fps_idx = np.concatenate(cur, axis=-1)
    if former_fps_idx is not None:
        fps_idx = np.concatenate([fps_idx, former_fps_idx], axis=-1)
    new_xyz = ops.gather_point(vote_ctr if vote_ctr is not None else xyz, fps_idx)   # :116-119

    debug = {"idx": [], "cnt": []}
    outs = []
    for i, (radius, nsample) in enumerate(zip(radius_list, nsample_list)):
        if dilated_group:                                          # :137-141
            min_r = 0.0 if i == 0 else radius_list[i - 1]
            idx, cnt = ops.query_ball_point_dilated(min_r, radius, nsample, xyz, new_xyz)
        else:
            idx, cnt = ops.query_ball_point(radius, nsample, xyz, new_xyz)
        mask = (cnt > 0).astype(np.int32)
        idx = idx * mask[..., None]                                # :157-159
        debug["idx"].append(idx)
        debug["cnt"].append(cnt)
        g_xyz = ops.group_point(xyz, idx) - new_xyz[:, :, None, :]                # :160-162
        g = np.concatenate([ops.group_point(points, idx), g_xyz], axis=-1)        # :163-165 features first
        for j in range(len(mlp_list[i])):
            g = _conv(params, "%s/conv%d_%d" % (scope, i, j), g, bn=bn, train=train)   # :167-176
        new_points = g.max(axis=2) * mask[..., None].astype(np.float32)           # :178-180
        outs.append(new_points)
    if outs:
        new_points = np.concatenate(outs, axis=-1)
        if aggregation and aggregation_channel is not None and aggregation_channel != -1:   # cfg...AGGREGATION_SA_FEATURE :183-185
            new_points = _conv(params, scope + "/ensemble", new_points, bn=bn, train=train)
    else:
        new_points = ops.gather_point(points, fps_idx)             # :186-187
    if return_debug:
        return new_xyz, new_points, fps_idx, debug
    return new_xyz, new_points, fps_idx


def pointnet_sa_module(xyz, points, mlp, is_training, bn_decay, bn, scope, *, params):
    g = np.concatenate([xyz, points], axis=-1)                     # xyz FIRST here (:42)
    for j in range(len(mlp)):
        g = _conv(params, "%s/conv%d" % (scope, j), g, bn=bn)
    return g.max(axis=1)


def pointnet_fp_module(xyz1, xyz2, points1, points2, mlp, is_training, bn_decay, scope, bn=True, *, params):
    dist, idx = ops.three_nn(xyz1, xyz2)
    dist = np.maximum(dist, np.float32(1e-10))                     # :207
    inv = (np.float32(1.0) / dist).astype(np.float32)
    norm = inv.sum(axis=2, keepdims=True, dtype=np.float32)
    weight = (inv / norm).astype(np.float32)                       # :208-210
    interp = ops.three_interpolate(points2, idx, weight)
    x = np.concatenate([interp, points1], axis=2) if points1 is not None else interp
    for i in range(len(mlp)):
        x = _conv(params, "%s/conv_%d" % (scope, i), x, bn=bn)
    return x


def vote_layer(xyz, points, mlp_list, is_training, bn_decay, bn, scope, *, params,
               max_translate_range=(-3.0, -2.0, -3.0)):
    for i in range(len(mlp_list)):
        points = _conv(params, "%s/vote_layer_%d" % (scope, i), points, bn=bn)
    off = _conv(params, scope + "/vote_offsets", points, bn=False, relu=False)
    lo = np.asarray(max_translate_range, np.float32).reshape(1, 1, 3)
    lim = np.minimum(np.maximum(off, lo), -lo)                     # :22
    return xyz + lim, points, off


def backbone_forward(arch, points_in, params, ffps_mode="matrix", return_debug=False):
    """SingleStageDetector.network_forward's backbone loop
    (/root/reference/lib/modeling/single_stage_detector.py:115-125) over a 16-field layer table."""
    xyz_list = [points_in[..., :3].copy()]
    feat_list = [points_in[..., 3:].copy()]
    fps_list = [None]
    dbg = []
    for spec in arch:
        (xyz_i, feat_i, radius, nsample, mlps, bn, rng, method, npoint, former, attn, ltype, scope, dilated,
         vote_idx, agg) = spec
        former_idx = fps_list[former] if former != -1 else None
        vote_ctr = xyz_list[vote_idx] if vote_idx != -1 else None
        if ltype == "SA_Layer":
            r = pointnet_sa_module_msg(xyz_list[xyz_i[0]], feat_list[feat_i[0]], radius, nsample, mlps, False, None,
                                       bn, rng, method, npoint, former_idx, attn, scope, dilated, vote_ctr, agg,
                                       params=params, ffps_mode=ffps_mode, return_debug=True)
            xyz_list.append(r[0]); feat_list.append(r[1]); fps_list.append(r[2]); dbg.append(r[3])
        elif ltype == "Vote_Layer":
            nx, nf, off = vote_layer(xyz_list[xyz_i[0]], feat_list[feat_i[0]], mlps, False, None, bn, scope, params=params)
            xyz_list.append(nx); feat_list.append(nf); fps_list.append(None); dbg.append({"offsets": off})
        elif ltype == "SA_Layer_SSG_Last":
            xyz_list.append(None)
            feat_list.append(pointnet_sa_module(xyz_list[xyz_i[0]], feat_list[feat_i[0]], mlps, False, None, bn, scope,
                                                params=params))
            fps_list.append(None); dbg.append({})
        elif ltype == "FP_Layer":
            xyz_list.append(xyz_list[xyz_i[0]])
            feat_list.append(pointnet_fp_module(xyz_list[xyz_i[0]], xyz_list[xyz_i[1]], feat_list[feat_i[0]],
                                                feat_list[feat_i[1]], mlps, False, None, scope, bn, params=params))
            fps_list.append(None); dbg.append({})
        else:
            raise ValueError(ltype)
    if return_debug:
        return xyz_list, feat_list, fps_list, dbg
    return xyz_list, feat_list, fps_list
