"""ImagenetAE localisation pipeline (models/imagenet_forward; reference:
/root/reference/tests/research/ImagenetAE/imagenet_forward/*.py, which has no tests of its own):
box math against brute-force loops, the shot renderer's geometry, the loader's enumeration, both
decision modes, result files and the whole two-stage workflow on a synthetic picture."""
import io
import json
import os
import pickle

import numpy
import pytest

from veles.znicz_b200.core.workflow import DummyLauncher
from veles.znicz_b200.models.imagenet_forward import bbox as B
from veles.znicz_b200.models.imagenet_forward.loader import ForwardLoaderBbox, render_shot
from veles.znicz_b200.models.imagenet_forward.merge import MergeBboxes
from veles.znicz_b200.models.imagenet_forward import writer as W
from veles.znicz_b200.models.imagenet_forward.workflow import ImagenetForward, shard_range

RS = numpy.random.RandomState(4)


def _rand_boxes(n, size=100):
    y0, x0 = RS.randint(0, size - 30, n), RS.randint(0, size - 30, n)
    return numpy.stack([y0, x0, y0 + RS.randint(5, 30, n), x0 + RS.randint(5, 30, n)], 1).astype(float)


def _brute_overlap(a, b):
    ys = set(range(int(a[0]), int(a[2]) + 1)) & set(range(int(b[0]), int(b[2]) + 1))
    xs = set(range(int(a[1]), int(a[3]) + 1)) & set(range(int(b[1]), int(b[3]) + 1))
    return len(ys) * len(xs)


def test_overlap_iou_inclusion_match_pixel_counting():
    boxes = _rand_boxes(25)
    got = B.overlap_area(boxes[:, None, :], boxes[None, :, :])
    iou = B.overlap_ratio(boxes[:, None, :], boxes[None, :, :])
    for i in range(len(boxes)):
        for j in range(len(boxes)):
            inter = _brute_overlap(boxes[i], boxes[j])
            assert got[i, j] == inter
            union = B.areas(boxes[i]) + B.areas(boxes[j]) - inter
            assert abs(iou[i, j] - inter / union) < 1e-12
    incl, a_bigger = B.has_inclusion([0, 0, 49, 49], [10, 10, 19, 19])
    assert incl and a_bigger
    incl, a_bigger = B.has_inclusion([10, 10, 19, 19], [0, 0, 49, 49])
    assert incl and not a_bigger
    assert not B.has_inclusion([0, 0, 9, 9], [5, 5, 14, 14])[0]
    assert B.is_small([0, 0, 9, 30], 20, 20) and not B.is_small([0, 0, 19, 19], 20, 20)
    assert B.is_small([0, 0, 19, 19], 20, 20, min_area=500)
    bb = B.BBox.from_center_view(50, 40, 21, 11)
    assert bb.to_caffe_view() == [35, 40, 45, 60] and bb.area() == 21 * 11
    assert B.BBox.from_json_dict(bb.to_json_dict()).to_caffe_view() == bb.to_caffe_view()
    assert bb.draw_on_pic(numpy.zeros((100, 100, 3), numpy.uint8)).sum() > 0


def test_nms_keeps_best_and_drops_covered():
    boxes = numpy.array([[0, 0, 20, 20], [1, 1, 20, 20], [2, 0, 21, 20], [50, 50, 70, 70],
                         [51, 51, 69, 69], [0, 50, 10, 60]], dtype=float)
    scores = numpy.array([0.5, 0.9, 0.4, 0.3, 0.8, 0.1])
    kept = B.nms_detections(boxes, scores, overlap_thr=0.7)
    assert kept[:, 4].tolist() == [0.9, 0.8, 0.1]                   # 0.5 / 0.4 / 0.3 are covered
    brute = []
    order = list(numpy.argsort(scores))
    while order:
        i = order.pop()
        brute.append(i)
        keep = []
        for j in order:
            w = max(0.0, min(boxes[i, 2], boxes[j, 2]) - max(boxes[i, 0], boxes[j, 0]) + 1)
            h = max(0.0, min(boxes[i, 3], boxes[j, 3]) - max(boxes[i, 1], boxes[j, 1]) + 1)
            area = (boxes[j, 2] - boxes[j, 0]) * (boxes[j, 3] - boxes[j, 1])
            if w * h / area <= 0.7:
                keep.append(j)
        order = keep
    assert kept[:, 4].tolist() == scores[brute].tolist()
    assert B.nms_detections(numpy.zeros((0, 4)), numpy.zeros(0)).shape == (0, 5)


def test_merging_boxes():
    boxes = numpy.array([[10, 10, 49, 49], [12, 12, 51, 51], [100, 100, 139, 139]], dtype=float)
    merged, prob = B.merge_to_one(boxes[:2], numpy.array([0.75, 0.25]), (200, 200), padding_ratio=0)
    assert numpy.allclose(merged, [10.5, 10.5, 49.5, 49.5]) and prob == 0.75
    merged, _ = B.merge_to_one(boxes[:2], numpy.array([0.0, 0.0]), (200, 200), padding_ratio=0)
    assert numpy.allclose(merged, [11, 11, 50, 50])
    merged, _ = B.merge_to_one(boxes[:1], numpy.array([1.0]), (45, 45), padding_ratio=0.5)
    assert merged.tolist() == [0, 0, 44, 44]                       # padded, clipped to the picture
    out_b, out_p = B.merge_by_probs(boxes, numpy.array([0.6, 0.5, 0.9]), (200, 200))
    assert out_p.tolist() == [0.9, 0.6] and len(out_b) == 2         # two clusters, best first
    out_b, out_p = B.merge_by_probs(boxes, numpy.array([0.6, 0.5, 0.9]), (200, 200), max_bboxes=1)
    assert out_p.tolist() == [0.9]
    out_b, out_p = B.merge_by_probs(boxes, numpy.array([0.01, 0.5, 0.9]), (200, 200), primary_thr=0.6)
    assert out_p.tolist() == [0.9]
    res = B.merge_by_dict({(30, 30, 40, 40): [0.1, 0.8], (32, 32, 40, 40): [0.2, 0.7],
                           (120, 120, 40, 40): [0.9, 0.05]}, (200, 200))
    assert [(r[0], round(r[1], 2)) for r in res] == [(0, 0.9), (1, 0.8), (0, 0.2)]
    with pytest.raises(ValueError):
        B.merge_by_dict({(5, 5, 40, 40): [1.0]}, (200, 200))        # sticks out of the picture
    kept = B.remove_inner([(1, 0.9, [0, 0, 99, 99]), (1, 0.8, [10, 10, 29, 29]),
                           (2, 0.7, [10, 10, 29, 29])])
    assert [(k[0], k[1]) for k in kept] == [(1, 0.9), (2, 0.7)]
    post = B.postprocess_same_label([(1, 0.9, (50, 50, 40, 40)), (1, 0.5, (55, 55, 10, 10)),
                                     (2, 0.4, (150, 150, 20, 20)), (1, 0.3, (90, 50, 20, 20))])
    labels = sorted(p[0] for p in post)
    assert labels == [1, 2] and max(p[1] for p in post) == 0.9
    big = [p for p in post if p[0] == 1][0][2]
    assert big[2] > 40                                              # grew towards the absorbed boxes


def test_render_shot_geometry():
    img = RS.randint(0, 255, (60, 80, 3)).astype(numpy.uint8)
    mean = numpy.full((16, 16, 3), 7.0, numpy.float32)
    box = {"x": 40.0, "y": 30.0, "width": 17.0, "height": 17.0}     # odd size: pixel-centred
    # scale 16/17 with centred sampling: compare with direct bilinear evaluation at a few points
    out = render_shot(img, box, 0.0, False, 16, mean)
    assert out.shape == (16, 16, 3) and out.dtype == numpy.float32
    k = 17.0 / 16.0
    for (v, u) in ((0, 0), (7, 8), (15, 15), (3, 12)):
        px, py = 40.0 + (u - 7.5) * k, 30.0 + (v - 7.5) * k
        x0, y0 = int(numpy.floor(px)), int(numpy.floor(py))
        fx, fy = px - x0, py - y0
        ref = (img[y0, x0] * (1 - fx) * (1 - fy) + img[y0, x0 + 1] * fx * (1 - fy) +
               img[y0 + 1, x0] * (1 - fx) * fy + img[y0 + 1, x0 + 1] * fx * fy)
        assert numpy.allclose(out[v, u], ref, atol=1e-3)
    flipped = render_shot(img, box, 0.0, True, 16, mean)
    assert numpy.allclose(flipped, out[:, ::-1], atol=1e-4)
    # a quarter turn of a square box = the un-rotated shot turned by 90 degrees
    rot = render_shot(img, box, numpy.pi / 2, False, 16, mean)
    assert numpy.allclose(rot, numpy.rot90(out, k=-1), atol=1e-3) or \
        numpy.allclose(rot, numpy.rot90(out, k=1), atol=1e-3)
    # 45 degrees: the rotated square fits the aperture, its corners show the mean image
    diag = render_shot(img, box, numpy.pi / 4, False, 16, mean)
    assert numpy.allclose(diag[0, 0], 7.0) and numpy.allclose(diag[15, 15], 7.0)
    assert not numpy.allclose(diag[8, 8], 7.0)
    # a box hanging over the picture edge shows the mean there
    edge = render_shot(img, {"x": 2.0, "y": 30.0, "width": 20.0, "height": 20.0}, 0.0, False, 16, mean)
    assert numpy.allclose(edge[:, 0], 7.0) and not numpy.allclose(edge[:, 15], 7.0)


def _pictures():
    """A grey picture with a bright and a dark square (+ a plain one)."""
    a = numpy.full((100, 120, 1), 110, numpy.uint8)
    a[20:60, 30:70] = 255
    a[60:90, 80:110] = 0
    b = numpy.full((80, 80, 1), 110, numpy.uint8)
    return {"pic_a.npy": a, "pic_b.npy": b}


def _candidates():
    return {
        "pic_a.npy": {"path": "pic_a.npy", "bbxs": [
            {"x": 50.0, "y": 40.0, "width": 40.0, "height": 40.0},
            {"x": 52.0, "y": 41.0, "width": 36.0, "height": 36.0},
            {"x": 48.0, "y": 38.0, "width": 30.0, "height": 30.0},
            {"x": 95.0, "y": 75.0, "width": 28.0, "height": 28.0},
            {"x": 94.0, "y": 74.0, "width": 24.0, "height": 24.0},
            {"x": 20.0, "y": 85.0, "width": 24.0, "height": 24.0},
            {"x": 3.0, "y": 3.0, "width": 4.0, "height": 4.0}]},          # too small: skipped
        "pic_b.npy": {"path": "pic_b.npy", "bbxs": [
            {"x": 40.0, "y": 40.0, "width": 30.0, "height": 30.0}]}}


def test_loader_enumerates_boxes_angles_and_mirrors():
    pics = _pictures()
    ld = ForwardLoaderBbox(DummyLauncher(), bboxes=_candidates(), minibatch_size=8,
                           angle_step=numpy.pi / 2, min_angle=0.0, max_angle=numpy.pi / 2,
                           add_relative_bboxes=False, raw_bboxes_min_size=8,
                           image_reader=lambda p: pics[p], entry_shape=(8, 12, 12, 1),
                           mean=numpy.zeros((12, 12, 1), numpy.float32))
    ld.initialize(device=None)
    assert ld.mode == "merge" and len(ld.angles) == 2
    assert ld.total == 8 * 2 * 2
    shots = []
    while not ld.ended:
        ld.run()
        for i in range(ld.minibatch_size):
            shots.append((ld.minibatch_images[i][0], ld.minibatch_bboxes[i][0]["x"],
                          round(ld.minibatch_bboxes[i][1], 3), ld.minibatch_bboxes[i][2]))
            assert ld.minibatch_images[i][1] == pics[ld.minibatch_images[i][0]].shape[:2]
    assert len(shots) == 7 * 4 and ld.processed == ld.total          # small box counted as done
    assert shots[:4] == [("pic_a.npy", 50.0, 0.0, False), ("pic_a.npy", 50.0, 1.571, False),
                         ("pic_a.npy", 50.0, 0.0, True), ("pic_a.npy", 50.0, 1.571, True)]
    assert shots[-1][0] == "pic_b.npy"
    # the relative probes are added per picture in the merge stage
    ld2 = ForwardLoaderBbox(DummyLauncher(), bboxes=_candidates(), minibatch_size=8, angle_step=1.0,
                            min_angle=0.0, max_angle=0.0, image_reader=lambda p: pics[p],
                            entry_shape=(8, 12, 12, 1), mean=None)
    ld2.initialize(device=None)
    assert ld2.total == (8 + 2 * 7) * 2


class _Probs(object):
    pass


def test_merge_bboxes_accumulates_over_shots_and_decides():
    m = MergeBboxes(DummyLauncher(), ignore_negative=False, probability_threshold=0.6,
                    last_chance_probability_threshold=0.5, use_compatibility=False)
    m.mode = "final"
    box1 = {"x": 50.0, "y": 40.0, "width": 40.0, "height": 40.0}
    box2 = {"x": 95.0, "y": 75.0, "width": 28.0, "height": 28.0}
    m.minibatch_images = [("a", (100, 120))] * 3 + [("b", (120, 120))]
    m.minibatch_bboxes = [(box1, 0.0, False), (box1, 0.1, True), (box2, 0.0, False), (box2, 0.0, False)]
    m.probabilities = numpy.array([[0.5, 0.4, 0.1], [0.2, 0.7, 0.1], [0.3, 0.3, 0.4],
                                   [0.9, 0.05, 0.05]])
    m.minibatch_size = 4
    m.ended = True
    m.initialize()
    m.run()
    assert [w["path"] for w in m.winners] == ["a", "b"]
    a = m.winners[0]["bbxs"]
    assert a == [(1, 0.7, (50.0, 40.0, 40.0, 40.0))]                # max over shots; box2 below both thresholds
    assert m.winners[1]["bbxs"] == []                               # only "nothing here" votes
    # merge mode on the same accumulated state
    m.reset()
    m.mode = "merge"
    m.run()
    labels = [(w[0], round(w[1], 2)) for w in m.winners[0]["bbxs"]]
    assert (1, 0.7) in labels and all(lbl > 0 for lbl, _ in labels)


def test_result_writer_and_converters(tmp_path):
    w = W.ResultWriter(DummyLauncher(), None, str(tmp_path / "res.json"), ignore_negative=False,
                       labels_mapping={1: "n01", 2: "n02"}, image_size_fn=lambda p: (120, 100))
    w.mode = "merge"
    w.winners = [{"path": "/x/pic_a.JPEG", "bbxs": [(1, 0.9, [20.0, 30.0, 59.0, 69.0])]}]
    w.initialize()
    w.run()
    res = json.load(open(tmp_path / "res.json"))
    assert res["pic_a.JPEG"]["bbxs"] == [{"conf": 0.9, "label": "n01", "angle": "0", "x": 50, "y": 40,
                                          "width": 39, "height": 39}]
    assert (res["pic_a.JPEG"]["width"], res["pic_a.JPEG"]["height"]) == (120, 100)
    w.mode = "final"
    w.winners = [{"path": "/x/pic_b.JPEG", "bbxs": [(2, 0.5, (10.0, 12.0, 8.0, 6.0))]}]
    w.run()
    res = json.load(open(tmp_path / "res.json"))
    assert res["pic_b.JPEG"]["bbxs"][0]["label"] == "n02" and len(res) == 2
    out = io.StringIO()
    n = W.convert_det(res, {"pic_a": 7, "pic_b": 9}, {"n01": 3, "n02": 4}, out)
    lines = out.getvalue().splitlines()
    assert n == 2 and lines[0].split() == ["7", "3", "0.900", "31", "21", "70", "60"]
    out = io.StringIO()
    W.convert_cls_loc(res, {"n01": 3}, out, ["pic_a.JPEG", "pic_b.JPEG", "missing.JPEG"])
    lines = out.getvalue().splitlines()
    assert lines[0].split()[0] == "3" and lines[1] == "0 0 1 0 1" and lines[2] == "0 0 1 0 1"
    json.dump({"pic_c.JPEG": {"bbxs": []}}, open(tmp_path / "other.json", "w"))
    merged = W.merge_json([str(tmp_path / "res.json"), str(tmp_path / "other.json")],
                          str(tmp_path / "all.json"))
    assert sorted(merged) == ["pic_a.JPEG", "pic_b.JPEG", "pic_c.JPEG"]


def test_shard_range_and_raw_extract(tmp_path):
    stream = tmp_path / "raw.pickle"
    sizes = [5, 1, 1, 1, 8, 2, 2, 4]
    with open(stream, "wb") as fout:
        for i, n in enumerate(sizes):
            pickle.dump(("img%d" % i, {"path": "img%d" % i, "bbxs": [{}] * n}), fout)
    ranges = [shard_range(str(stream), r, 3) for r in range(3)]
    assert ranges[0][0] == 0 and ranges[-1][1] == len(sizes)
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    loads = [sum(sizes[a:b]) for a, b in ranges]
    assert max(loads) <= 12 and min(loads) >= 4                      # balanced by boxes, not pictures
    key, meta = W.extract_raw_bboxes(str(stream), "img4", str(tmp_path / "one.json"))
    assert key == "img4" and len(meta["bbxs"]) == 8
    assert list(json.load(open(tmp_path / "one.json"))) == ["img4"]


def test_two_stage_pipeline_localises_objects(tmp_path):
    """Hand-set 3-class softmax (nothing / bright / dark by mean brightness) through the whole
    workflow: merge stage over the candidate stream, final stage over its own JSON."""
    from veles.znicz_b200.ops import all2all
    pics = _pictures()
    stream = tmp_path / "raw.pickle"
    with open(stream, "wb") as fout:
        for i, (k, meta) in enumerate(sorted(_candidates().items())):
            pickle.dump((i, meta), fout)
    launcher = DummyLauncher(testing=True)
    holder = launcher
    fc = all2all.All2AllSoftmax(holder, output_sample_shape=3, weights_stddev=0.01)
    wf = ImagenetForward(
        launcher, forwards=[fc], entry_shape=(4, 8, 8, 1), mean=numpy.full((8, 8, 1), 110.0, numpy.float32),
        image_reader=lambda p: pics[p], result_path=str(tmp_path / "result.json"),
        labels_mapping={0: "nothing", 1: "bright", 2: "dark"},
        loader_config={"path_to_bboxes": str(stream), "minibatch_size": 4, "raw_bboxes_min_size": 8,
                       "raw_bboxes_min_area": 64, "add_relative_bboxes": False},
        merge_config={"ignore_negative": False, "use_compatibility": False,
                      "probability_threshold": 0.45, "last_chance_probability_threshold": 0.39})
    wf.initialize(device="numpy")
    fc.weights.map_write()
    fc.bias.map_write()
    fc.weights.mem[...] = 0
    fc.weights.mem[1] = 0.1 / 64                                    # logit = (mean - 160) / 10
    fc.weights.mem[2] = -0.1 / 64                                   # logit = (60 - mean) / 10
    fc.bias.mem[...] = [0.0, -16.0, 6.0]
    final = wf.run_pipeline()
    assert os.path.exists(tmp_path / "result.json.raw")             # the merge stage's own output
    raw = json.load(open(tmp_path / "result.json.raw"))
    assert {b["label"] for b in raw["pic_a.npy"]["bbxs"]} == {"bright", "dark"}
    assert wf.loader.mode == "final"
    dets = {b["label"]: b for b in final["pic_a.npy"]["bbxs"]}
    assert set(dets) == {"bright", "dark"}
    assert abs(dets["bright"]["x"] - 50) <= 4 and abs(dets["bright"]["y"] - 40) <= 4
    assert abs(dets["dark"]["x"] - 95) <= 4 and abs(dets["dark"]["y"] - 75) <= 4
    assert dets["bright"]["conf"] > 0.9
    assert "pic_b.npy" not in final or final["pic_b.npy"]["bbxs"] == []
    assert raw["pic_b.npy"]["bbxs"] == []                           # a plain picture: nothing found
    assert json.load(open(tmp_path / "result.json")) == final


def test_pipeline_from_a_trained_workflow_snapshot(tmp_path):
    """The forward units (and nothing else) are taken over from a snapshot of a trained workflow
    (/root/reference/.../imagenet_forward.py:237-262)."""
    from veles.znicz_b200.core.config import root
    from veles.znicz_b200.models import mnist
    old = root.common.disable.snapshotting
    root.common.disable.snapshotting = False
    try:
        train = mnist.build(
            layers=[{"type": "all2all_tanh", "->": {"output_sample_shape": 6, "weights_stddev": 0.1},
                     "<-": {"learning_rate": 0.05}},
                    {"type": "softmax", "->": {"output_sample_shape": 3, "weights_stddev": 0.1},
                     "<-": {"learning_rate": 0.05}}],
            loader_name="synthetic_image",
            loader_config={"minibatch_size": 10, "n_train": 60, "n_valid": 30, "shape": (8, 8, 1),
                           "n_classes": 3, "normalization_type": "none"},
            decision_config={"max_epochs": 1, "fail_iterations": 5},
            snapshotter_config={"prefix": "trained", "interval": 1, "time_interval": 0,
                                "compression": "gz", "directory": str(tmp_path)})
        train.initialize(device="numpy")
        train.run()
    finally:
        root.common.disable.snapshotting = old
    snap = str(tmp_path / "trained_current.lnk")
    assert os.path.exists(snap)
    pics = _pictures()
    wf = ImagenetForward(
        DummyLauncher(testing=True), trained_workflow=snap, entry_shape=(4, 8, 8, 1),
        mean=numpy.zeros((8, 8, 1), numpy.float32), image_reader=lambda p: pics[p],
        bboxes=_candidates(), result_path=str(tmp_path / "out.json"),
        labels_mapping={0: "a", 1: "b", 2: "c"},
        loader_config={"minibatch_size": 4, "raw_bboxes_min_size": 8, "raw_bboxes_min_area": 64,
                       "add_relative_bboxes": False},
        merge_config={"ignore_negative": True, "use_compatibility": False, "mode": "final",
                      "probability_threshold": 0.0, "last_chance_probability_threshold": 0.0})
    assert [type(f).__name__ for f in wf.forwards] == ["All2AllTanh", "All2AllSoftmax"]
    assert wf.forwards[0].weights.mem.shape == (6, 64)               # the trained parameters
    wf.initialize(device="numpy")
    res = wf.run_stage()
    assert set(res) == {"pic_a.npy", "pic_b.npy"}
    assert all(b["label"] in ("b", "c") for v in res.values() for b in v["bbxs"])
    assert json.load(open(tmp_path / "out.json")) == res


def test_ranks_share_the_candidate_stream(tmp_path):
    """`run_from_config` under torchrun: each rank takes its `shard_range` of the pickled candidate
    stream (the reference printed one command line per slave, distribute_forward.py:40-85) and
    writes `<result>.rank<k>.json`; `merge_json` combines them - every picture exactly once."""
    import socket
    import subprocess
    import sys
    from veles.znicz_b200.core.config import root
    from veles.znicz_b200.models import mnist
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    old = root.common.disable.snapshotting
    root.common.disable.snapshotting = False
    try:
        train = mnist.build(
            layers=[{"type": "softmax", "->": {"output_sample_shape": 3, "weights_stddev": 0.1},
                     "<-": {"learning_rate": 0.05}}],
            loader_name="synthetic_image",
            loader_config={"minibatch_size": 10, "n_train": 30, "n_valid": 10, "shape": (8, 8, 1),
                           "n_classes": 3, "normalization_type": "none"},
            decision_config={"max_epochs": 1, "fail_iterations": 5},
            snapshotter_config={"prefix": "trained", "interval": 1, "time_interval": 0,
                                "compression": "", "directory": str(tmp_path)})
        train.initialize(device="numpy")
        train.run()
    finally:
        root.common.disable.snapshotting = old
    pics = _pictures()
    names = []
    with open(tmp_path / "raw.pickle", "wb") as fout:
        for i in range(4):                                  # four pictures, two kinds
            src = "pic_a.npy" if i % 2 == 0 else "pic_b.npy"
            path = str(tmp_path / ("img%d.npy" % i))
            numpy.save(path, pics[src])
            meta = dict(_candidates()[src], path=path)
            pickle.dump((i, meta), fout)
            names.append(os.path.basename(path))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", str(port),
         os.path.join(repo, "tests", "imagenet_forward_worker.py"), str(tmp_path)],
        env=dict(os.environ, PYTHONPATH=repo, OMP_NUM_THREADS="1"), capture_output=True, text=True,
        timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    parts = [str(tmp_path / ("result.rank%d.json" % k)) for k in range(2)]
    assert all(os.path.exists(p) for p in parts)
    per_rank = [set(json.load(open(p))) for p in parts]
    assert per_rank[0] and per_rank[1] and not (per_rank[0] & per_rank[1])
    merged = W.merge_json(parts, str(tmp_path / "all.json"))
    assert sorted(merged) == sorted(names)
