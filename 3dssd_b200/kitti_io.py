"""Data formats either side of the SA path.

What enters the path: the reference's preprocessed KITTI sample files -- one pickled dict per frame written by
its preprocessing pass and read back in /root/reference/lib/dataset/dataloader/kitti_dataloader.py:103-158 (keys in
lib/dataset/maps_dict.py).  What leaves it: one KITTI result .txt per frame (kitti_dataloader.py:460-492) built from
the post-processed boxes, and the network weights, which the reference keeps in a TF checkpoint whose variable
names follow the scopes in params.py.

Everything here is host-side byte shuffling on a few hundred boxes / a few MB of points per frame; it feeds pinned
host buffers that the backbone copies to HBM, no kernels involved.
"""
import os

import numpy as np

# maps_dict.py:4-36
KEY_POINT_CLOUD = "point_cloud"
KEY_STEREO_CALIB = "stereo_calib"
KEY_SAMPLE_NAME = "sample_name"
KEY_LABEL_BOXES_3D = "label_boxes_3d"
KEY_LABEL_CLASSES = "label_classes"

KITTI_IMG_SHAPE = (375, 1242)       # anchors_util.py:54 default (h, w)


# ------------------------------------------------------------------------------------------------------------------
# input side
# ------------------------------------------------------------------------------------------------------------------
def read_sample(path):
    """One preprocessed frame -> dict.  The file is a 0-d object array holding the dict (kitti_dataloader.py:110)."""
    arr = np.load(path, allow_pickle=True)
    d = arr.item() if isinstance(arr, np.ndarray) and arr.shape == () else arr
    if not isinstance(d, dict) or KEY_POINT_CLOUD not in d:
        raise ValueError("%s: not a 3DSSD sample dict (no '%s' key)" % (path, KEY_POINT_CLOUD))
    return d


def choose_points(num_have, num_want, rng):
    """Indices of the fixed-size point set fed to the network (kitti_dataloader.py:137-147): a random subset without
    replacement when the frame has enough points, otherwise every point once (shuffled) followed by a random refill
    with replacement."""
    if num_have <= 0:
        raise ValueError("choose_points: empty point cloud")
    if num_have >= num_want:
        return rng.choice(num_have, num_want, replace=False)
    first = rng.choice(num_have, num_have, replace=False)
    refill = rng.choice(num_have, num_want - num_have, replace=True)
    return np.concatenate([first, refill], axis=0)


def calib_p2(calib):
    """3x4 camera projection of a sample: the reference stores a calibration object with a .P attribute
    (kitti_dataloader.py:158); a bare array is accepted too."""
    p = getattr(calib, "P", calib)
    p = np.asarray(p, np.float32)
    if p.shape != (3, 4):
        raise ValueError("calibration P must be 3x4, got %s" % (p.shape,))
    return p


def make_batch(samples, num_points=16384, seed=0, pin=True):
    """Stack frames into the [B, num_points, C] fp32 host block the backbone consumes (pinned, so the H2D copy inside
    the timed path is asynchronous).  Returns (points tensor, [P2 per frame], [sample name per frame])."""
    import torch
    rng = np.random.default_rng(seed)
    clouds, calibs, names = [], [], []
    for s in samples:
        pts = np.asarray(s[KEY_POINT_CLOUD], np.float32)
        clouds.append(pts[choose_points(pts.shape[0], num_points, rng)])
        calibs.append(calib_p2(s[KEY_STEREO_CALIB]) if KEY_STEREO_CALIB in s else None)
        names.append(s.get(KEY_SAMPLE_NAME, len(names)))
    block = torch.from_numpy(np.stack(clouds, axis=0))
    if pin and torch.cuda.is_available():
        block = block.pin_memory()
    return block, calibs, names


# ------------------------------------------------------------------------------------------------------------------
# output side
# ------------------------------------------------------------------------------------------------------------------
def box_corners(boxes):
    """[N,7] (x, y, z, l, h, w, ry) camera-frame boxes, origin at the bottom-face centre -> [N,8,3] corners in the order
    of box_3d_utils.py:62-87 (bottom face first, y pointing down so the top face is at y - h)."""
    boxes = np.asarray(boxes, np.float32).reshape(-1, 7)
    l, h, w, ry = boxes[:, 3], boxes[:, 4], boxes[:, 5], boxes[:, 6]
    sx = np.array([1, 1, -1, -1, 1, 1, -1, -1], np.float32) * 0.5
    sz = np.array([1, -1, -1, 1, 1, -1, -1, 1], np.float32) * 0.5
    sy = np.array([0, 0, 0, 0, -1, -1, -1, -1], np.float32)
    cx, cy, cz = l[:, None] * sx, h[:, None] * sy, w[:, None] * sz            # [N,8] box-frame offsets
    c, s = np.cos(ry)[:, None], np.sin(ry)[:, None]
    out = np.empty((boxes.shape[0], 8, 3), np.float32)
    out[:, :, 0] = c * cx + s * cz + boxes[:, 0:1]
    out[:, :, 1] = cy + boxes[:, 1:2]
    out[:, :, 2] = -s * cx + c * cz + boxes[:, 2:3]
    return out


def project_boxes_to_image(boxes, p2, img_shape=KITTI_IMG_SHAPE):
    """[N,7] boxes -> [N,4] (x1, y1, x2, y2) image rectangles clipped to the frame (anchors_util.py:54-91 with
    kitti_util.py:329-348)."""
    corners = box_corners(boxes).reshape(-1, 3)
    hom = np.concatenate([corners, np.ones((corners.shape[0], 1), corners.dtype)], axis=1)
    uvw = hom.astype(np.float64) @ np.asarray(p2, np.float64).T
    uv = (uvw[:, :2] / uvw[:, 2:3]).reshape(-1, 8, 2)
    h, w = img_shape
    rect = np.stack([np.clip(uv[:, :, 0].min(1), 0, w), np.clip(uv[:, :, 1].min(1), 0, h),
                     np.clip(uv[:, :, 0].max(1), 0, w), np.clip(uv[:, :, 1].max(1), 0, h)], axis=-1)
    return rect.astype(np.float32)


def format_kitti_result(boxes, scores, classes, p2, cls_list=("Car",), cls_thresh=0.0, img_shape=KITTI_IMG_SHAPE):
    """KITTI result lines of one frame (kitti_dataloader.py:474-490): type, truncation 0.00, occlusion 0, alpha -10,
    2-D rectangle, dimensions as h w l, location, ry, score -- boxes below cls_thresh are dropped."""
    boxes = np.asarray(boxes, np.float32).reshape(-1, 7)
    scores = np.asarray(scores, np.float32).reshape(-1)
    classes = np.asarray(classes).reshape(-1)
    keep = np.where(scores >= cls_thresh)[0]
    boxes, scores, classes = boxes[keep], scores[keep], classes[keep]
    rect = project_boxes_to_image(boxes, p2, img_shape) if len(boxes) else np.zeros((0, 4), np.float32)
    lines = []
    for b, s, c, r in zip(boxes, scores, classes, rect):
        lines.append("%s %0.2f %d %d %0.2f %0.2f %0.2f %0.2f %0.2f %0.2f %0.2f %0.2f %0.2f %0.2f %0.2f %0.9f\n" % (
            cls_list[int(c)], 0.0, 0, -10, float(r[0]), float(r[1]), float(r[2]), float(r[3]),
            float(b[4]), float(b[5]), float(b[3]), float(b[0]), float(b[1]), float(b[2]), float(b[6]), float(s)))
    return "".join(lines)


def write_kitti_result(out_dir, sample_name, boxes, scores, classes, p2, **kw):
    """One '<out_dir>/kitti_result/%06d.txt' per frame, like save_predictions (kitti_dataloader.py:465,474)."""
    sv_dir = os.path.join(out_dir, "kitti_result")
    os.makedirs(sv_dir, exist_ok=True)
    path = os.path.join(sv_dir, "%06d.txt" % int(sample_name))
    with open(path, "w") as fid:
        fid.write(format_kitti_result(boxes, scores, classes, p2, **kw))
    return path


def write_detections(out_dir, sample_names, det_block, det_count, calibs, **kw):
    """Writes the fixed-size detection block the backbone returns ([B, K, 9] = box7, score, class; counts [B])."""
    det_block = np.asarray(det_block, np.float32)
    paths = []
    for i, name in enumerate(sample_names):
        k = int(det_count[i])
        d = det_block[i, :k]
        paths.append(write_kitti_result(out_dir, name, d[:, :7], d[:, 7], d[:, 8].astype(np.int64), calibs[i], **kw))
    return paths


# ------------------------------------------------------------------------------------------------------------------
# weights
# ------------------------------------------------------------------------------------------------------------------
_SKIP_SUFFIXES = ("/Adam", "/Adam_1", "/Momentum", "/ExponentialMovingAverage")
_SKIP_NAMES = ("global_step", "beta1_power", "beta2_power")


def params_from_tf_variables(variables):
    """name -> array mapping exported from a reference checkpoint (see INTEGRATION.md for the three-line export that a
    TF-equipped machine runs; TF itself is not needed here) -> the params dict of params.py.

    Kernels are squeezed to [cin, cout]: tf_util.conv2d stores [1,1,cin,cout] (tf_util.py:88-112), conv1d
    [1,cin,cout] (tf_util.py:164-190).  Optimiser slots and counters are dropped; ':0' tensor suffixes are stripped.
    """
    out = {}
    for name, val in dict(variables).items():
        name = name[:-2] if name.endswith(":0") else name
        if name in _SKIP_NAMES or name.endswith(_SKIP_SUFFIXES):
            continue
        a = np.asarray(val, np.float32)
        if name.endswith("/weights"):
            if a.ndim < 2 or any(d != 1 for d in a.shape[:-2]):
                raise ValueError("%s: expected a 1x1 kernel, got shape %s" % (name, a.shape))
            a = a.reshape(a.shape[-2], a.shape[-1])
        out[name] = np.ascontiguousarray(a)
    for name in list(out):
        if name.endswith("/weights"):
            scope = name[: -len("/weights")]
            cout = out[name].shape[1]
            for leaf in ("biases", "bn/gamma", "bn/beta", "bn/moving_mean", "bn/moving_variance"):
                v = out.get(scope + "/" + leaf)
                if v is not None and v.shape != (cout,):
                    raise ValueError("%s/%s: shape %s does not match %d output channels" % (scope, leaf, v.shape, cout))
    return out
