"""numpy restatement of the detection head, decoding and post-processing of the single-stage 3DSSD detector.

TEST INFRASTRUCTURE ONLY (see ssd3d_oracle.c).  Follows /root/reference/lib/modeling/head_builder.py:81-114,
lib/utils/head_util.py:26-59 (box_regression_head), lib/utils/anchor_decoder.py:6-14, :86-112
(decode_class2angle, decode_dist_anchor_free), lib/modeling/single_stage_detector.py:193-226 (sigmoid scores),
lib/utils/box_3d_utils.py:25-58 (box_3d_to_anchor), lib/utils/anchors_util.py:11-48 (project_to_bev) and
lib/builder/postprocessor.py:52-120 with TensorFlow's greedy non_max_suppression (descending score, a candidate is
dropped when IoU with a kept box is > threshold; ties broken by lower index, which TF 1.4 leaves unspecified).
"""
import numpy as np

from . import ops

ANGLE_CLS_NUM = 12


def _conv(params, scope, x, bn, relu):
    bnp = tuple(params[scope + "/bn/" + k] for k in ("gamma", "beta", "moving_mean", "moving_variance")) if bn else None
    return ops.linear_bn_relu(x, params[scope + "/weights"], params.get(scope + "/biases"), bnp, relu)


def head_forward(xyz, feat, params, mlp=(128,), bn=True, scope=""):
    pre = "" if scope == "" else scope + "/"
    y = feat
    for i in range(len(mlp)):
        y = _conv(params, "%sconv1d_%d" % (pre, i), y, bn, True)
    cls = _conv(params, pre + "pred_cls", _conv(params, pre + "pred_cls_base", y, bn, True), False, False)
    reg = _conv(params, pre + "pred_reg", _conv(params, pre + "pred_reg_base", y, bn, True), False, False)
    off, acls, ares = reg[..., :6], reg[..., 6:6 + ANGLE_CLS_NUM], reg[..., 6 + ANGLE_CLS_NUM:]
    bins = np.argmax(acls, axis=-1)
    res = np.take_along_axis(ares, bins[..., None], axis=-1)[..., 0]
    angle = ((bins.astype(np.float32) + res) * np.float32(2 * np.pi / ANGLE_CLS_NUM)).astype(np.float32)
    ctr = xyz + off[..., :3]
    ctr[..., 1] += off[..., 4]
    lhw = np.maximum(off[..., 3:6] * np.float32(2.0), np.float32(0.1))
    boxes = np.concatenate([ctr, lhw, angle[..., None]], axis=-1).astype(np.float32)
    score = (1.0 / (1.0 + np.exp(-cls.astype(np.float64)))).astype(np.float32)
    return boxes, score[..., 0], {"feat": y, "cls": cls, "reg": reg}


def bev_nms(boxes, scores, iou_threshold=0.1, max_output=100, cls_id=0):
    b, n, _ = boxes.shape
    block = np.zeros((b, max_output, 9), np.float32)
    cnt = np.zeros((b,), np.int32)
    for s in range(b):
        bx, sc = boxes[s], scores[s]
        c, sn = np.abs(np.cos(bx[:, 6])), np.abs(np.sin(bx[:, 6]))
        dimx = bx[:, 3] * c + bx[:, 5] * sn
        dimz = bx[:, 5] * c + bx[:, 3] * sn
        x1, z1 = bx[:, 0] - dimx * np.float32(0.5), bx[:, 2] - dimz * np.float32(0.5)
        x2, z2 = bx[:, 0] + dimx * np.float32(0.5), bx[:, 2] + dimz * np.float32(0.5)
        area = (x2 - x1) * (z2 - z1)
        order = sorted(range(n), key=lambda i: (-sc[i], i))
        kept = []
        for i in order:
            ok = True
            for j in kept:
                if area[i] <= 0 or area[j] <= 0:
                    continue
                iw = min(x2[i], x2[j]) - max(x1[i], x1[j])
                ih = min(z2[i], z2[j]) - max(z1[i], z1[j])
                inter = max(iw, np.float32(0)) * max(ih, np.float32(0))
                if inter / (area[i] + area[j] - inter) > iou_threshold:
                    ok = False
                    break
            if ok:
                kept.append(i)
                if len(kept) == max_output:
                    break
        for k, i in enumerate(kept):
            block[s, k, :7] = bx[i]
            block[s, k, 7] = sc[i]
            block[s, k, 8] = cls_id
        cnt[s] = len(kept)
    return block, cnt
