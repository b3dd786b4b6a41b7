"""Row f4: the data formats either side of the SA path.  CPU only."""
import importlib
import os

import numpy as np
import pytest

kio = importlib.import_module("3dssd_b200.kitti_io")
params_mod = importlib.import_module("3dssd_b200.params")
config = importlib.import_module("3dssd_b200.config")

GOLD = os.path.join(os.path.dirname(__file__), "golden", "kitti_io.npz")


def test_corners_and_rectangles_match_reference_helpers():
    g = np.load(GOLD)
    corners = kio.box_corners(g["boxes"])
    np.testing.assert_allclose(corners, g["corners"], rtol=0, atol=2e-5)
    rect = kio.project_boxes_to_image(g["boxes"], g["p2"])
    np.testing.assert_allclose(rect, g["rect"], rtol=0, atol=2e-2)       # pixels; the writer prints %0.2f
    assert (rect[:, 0] >= 0).all() and (rect[:, 2] <= 1242).all() and (rect[:, 3] <= 375).all()


def test_result_lines_follow_kitti_layout():
    boxes = np.array([[1.0, 1.5, 20.0, 4.0, 1.5, 1.6, 0.3], [2.0, 1.6, 30.0, 3.5, 1.4, 1.5, -1.0]], np.float32)
    p2 = np.load(GOLD)["p2"]
    txt = kio.format_kitti_result(boxes, [0.9, 0.05], [0, 0], p2, cls_list=("Car",), cls_thresh=0.1)
    lines = txt.splitlines()
    assert len(lines) == 1                      # the second box is under the threshold
    f = lines[0].split(" ")
    assert len(f) == 16 and f[0] == "Car" and f[1] == "0.00" and f[2] == "0" and f[3] == "-10"
    assert [float(v) for v in f[8:11]] == [1.5, 1.6, 4.0]          # h w l
    assert [float(v) for v in f[11:15]] == [1.0, 1.5, 20.0, 0.3]   # x y z ry
    assert f[15] == "0.899999976"                                 # %0.9f of the fp32 score
    rect = kio.project_boxes_to_image(boxes[:1], p2)[0]
    assert [float(v) for v in f[4:8]] == [float("%0.2f" % r) for r in rect]
    assert kio.format_kitti_result(boxes[:0], [], [], p2) == ""


def test_write_detections_roundtrip(tmp_path):
    p2 = np.load(GOLD)["p2"]
    block = np.zeros((2, 5, 9), np.float32)
    block[0, 0] = [1, 1.5, 20, 4, 1.5, 1.6, 0.1, 0.8, 0]
    block[0, 1] = [3, 1.5, 25, 4, 1.5, 1.6, 0.2, 0.7, 0]
    paths = kio.write_detections(str(tmp_path), [7, 123], block, [2, 0], [p2, p2])
    assert [os.path.basename(p) for p in paths] == ["000007.txt", "000123.txt"]
    assert len(open(paths[0]).read().splitlines()) == 2 and open(paths[1]).read() == ""


def test_choose_points_semantics():
    rng = np.random.default_rng(0)
    idx = kio.choose_points(100, 40, rng)
    assert idx.shape == (40,) and len(set(idx.tolist())) == 40
    idx = kio.choose_points(30, 100, rng)
    assert idx.shape == (100,) and sorted(idx[:30].tolist()) == list(range(30)) and idx.max() < 30
    with pytest.raises(ValueError):
        kio.choose_points(0, 10, rng)


def test_read_sample_and_batch(tmp_path):
    class Calib:
        P = np.arange(12, dtype=np.float32).reshape(3, 4)

    rng = np.random.default_rng(1)
    files = []
    for i, n in enumerate((500, 90)):
        d = {kio.KEY_POINT_CLOUD: rng.standard_normal((n, 4)).astype(np.float32), kio.KEY_STEREO_CALIB: Calib.P,
             kio.KEY_SAMPLE_NAME: "%06d" % (i + 5)}
        p = str(tmp_path / ("%d.npy" % i))
        np.save(p, d)
        files.append(p)
    samples = [kio.read_sample(p) for p in files]
    block, calibs, names = kio.make_batch(samples, num_points=128, seed=3, pin=False)
    assert tuple(block.shape) == (2, 128, 4) and names == ["000005", "000006"]
    assert calibs[0].shape == (3, 4) and kio.calib_p2(Calib()).shape == (3, 4)
    src = samples[1][kio.KEY_POINT_CLOUD]
    rows = {tuple(r) for r in src.tolist()}
    assert all(tuple(r) in rows for r in block[1].numpy().tolist())
    np.save(str(tmp_path / "bad.npy"), np.zeros(3))
    with pytest.raises(ValueError):
        kio.read_sample(str(tmp_path / "bad.npy"))


def test_params_from_tf_variables():
    ref = params_mod.init_params(config.ARCH_SINGLE_SA, 1, seed=2)
    tfvars = {}
    for k, v in ref.items():
        if k.endswith("/weights"):
            tfvars[k + ":0"] = v.reshape((1, 1) + v.shape)
            tfvars[k + "/Adam:0"] = np.zeros((1, 1) + v.shape, np.float32)
        else:
            tfvars[k + ":0"] = v
    tfvars["global_step:0"] = np.int64(5)
    got = kio.params_from_tf_variables(tfvars)
    assert sorted(got) == sorted(ref)
    for k in ref:
        np.testing.assert_array_equal(got[k], ref[k])
    bad = dict(tfvars)
    k0 = next(k for k in bad if k.endswith("/weights:0"))
    bad[k0] = np.zeros((3, 3, 4, 8), np.float32)
    with pytest.raises(ValueError):
        kio.params_from_tf_variables(bad)
