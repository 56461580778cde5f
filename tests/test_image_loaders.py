"""Image-file loaders on a generated directory tree."""
import os

import numpy
import pytest

cv2 = pytest.importorskip("cv2")

from veles.znicz_b200.core.workflow import DummyWorkflow  # noqa: E402
from veles.znicz_b200.loader import image as limg  # noqa: E402
from veles.znicz_b200.loader.base import UserLoaderRegistry  # noqa: E402


def _make_tree(root, labels=("a", "b", "c"), n=6, size=(20, 16), ext="png"):
    rs = numpy.random.RandomState(5)
    for li, l in enumerate(labels):
        d = os.path.join(root, l)
        os.makedirs(d, exist_ok=True)
        for i in range(n):
            img = (rs.rand(size[1], size[0], 3) * 60).astype(numpy.uint8)
            img[:, :, li % 3] += 150
            cv2.imwrite(os.path.join(d, "img_%d.%s" % (i, ext)), img)
    with open(os.path.join(root, labels[0], "notes.txt"), "w") as f:
        f.write("not an image")


def test_registry_names():
    for name in ("full_batch_file_image", "full_batch_auto_label_file_image",
                 "full_batch_auto_label_file_image_mse", "auto_label_file_image"):
        assert name in UserLoaderRegistry.loaders


def test_full_batch_auto_label(tmp_path):
    tr, va = str(tmp_path / "train"), str(tmp_path / "valid")
    _make_tree(tr, n=6)
    _make_tree(va, n=2)
    wf = DummyWorkflow()
    ld = limg.FullBatchAutoLabelFileImageLoader(
        wf, train_paths=[tr], validation_paths=[va], minibatch_size=5,
        file_subtypes=["png"], color_space="RGB", normalization_type="linear",
        ignored_files=[".*_5.*"], scale=(10, 8))
    ld.initialize(device="numpy")
    assert list(ld.class_lengths) == [0, 6, 15]
    assert ld.original_data.shape == (21, 8, 10, 3)
    assert ld.reversed_labels_mapping == ["a", "b", "c"]
    assert -1.0 <= ld.original_data.mem.min() and ld.original_data.mem.max() <= 1.0
    ld.run()
    assert ld.minibatch_class == 1 and ld.minibatch_size == 5
    # class "a" was written with a bright channel 0 in OpenCV's BGR order = channel 2 in RGB
    lab0 = [i for i, l in enumerate(ld.original_labels) if l == "a"][0]
    assert ld.original_data.mem[lab0].mean(axis=(0, 1)).argmax() == 2


def test_gray_sobel_mirror_rotations(tmp_path):
    tr = str(tmp_path / "train")
    _make_tree(tr, labels=("x", "y"), n=3)
    wf = DummyWorkflow()
    ld = limg.FullBatchAutoLabelFileImageLoader(
        wf, train_paths=[tr], minibatch_size=4, color_space="GRAY", add_sobel=True,
        mirror=True, rotations=(0.0, 0.3), validation_ratio=0.25)
    ld.initialize(device="numpy")
    # 6 files x (orig + rotated) x mirror = 24 samples, 2 channels (gray + sobel)
    assert ld.total_samples == 24 and ld.original_data.shape[1:] == (16, 20, 2)
    assert ld.class_lengths[1] == 6 and ld.class_lengths[2] == 18


def test_label_regexp_and_aspect(tmp_path):
    d = str(tmp_path / "flat")
    os.makedirs(d)
    for i in range(4):
        cv2.imwrite(os.path.join(d, "cat%d_%d.jpeg" % (i % 2, i)),
                    numpy.full((10, 30, 3), 40 * i, numpy.uint8))
    wf = DummyWorkflow()
    ld = limg.FullBatchFileImageLoader(
        wf, train_paths=[d], label_regexp=r"(cat\d)_", minibatch_size=2,
        scale=(16, 16), scale_maintain_aspect_ratio=True, background_color=(255, 0, 0),
        file_subtypes=["jpeg"])
    ld.initialize(device="numpy")
    assert sorted(set(ld.original_labels)) == ["cat0", "cat1"]
    img = ld.original_data.mem[0]
    assert img.shape == (16, 16, 3) and img[0, 0, 0] == 255 and img[0, 0, 1] == 0


def test_mse_targets(tmp_path):
    tr, tg = str(tmp_path / "train"), str(tmp_path / "target")
    _make_tree(tr, labels=("k1", "k2"), n=4)
    os.makedirs(tg)
    for i, l in enumerate(("k1", "k2")):
        cv2.imwrite(os.path.join(tg, l + ".png"), numpy.full((12, 12), 100 * (i + 1),
                                                              numpy.uint8))
    wf = DummyWorkflow()
    ld = limg.FullBatchAutoLabelFileImageLoaderMSE(
        wf, train_paths=[tr], target_paths=[tg], minibatch_size=3, color_space="GRAY",
        normalization_type="linear", target_normalization_type="range_linear",
        targets_shape=(6, 6), validation_ratio=0.25)
    ld.initialize(device="numpy")
    assert ld.class_targets.shape == (2, 6, 6)
    assert ld.original_targets.shape == (8, 6, 6)
    ld.run()
    assert ld.minibatch_targets.shape == (3, 6, 6)
    lab = ld.minibatch_labels.mem[0]
    numpy.testing.assert_allclose(ld.minibatch_targets.mem[0], ld.class_targets.mem[lab])


def test_streaming_loader(tmp_path):
    tr = str(tmp_path / "train")
    _make_tree(tr, n=5)
    wf = DummyWorkflow()
    ld = limg.AutoLabelFileImageLoader(wf, train_paths=[tr], minibatch_size=4,
                                       normalization_type="mean_disp", mirror="random")
    ld.initialize(device="numpy")
    assert ld.class_lengths[2] == 15 and ld.minibatch_data.shape == (4, 16, 20, 3)
    seen = 0
    for _ in range(4):
        ld.run()
        seen += ld.minibatch_size
        assert numpy.isfinite(ld.minibatch_data.mem).all()
    assert seen == 15 and bool(ld.epoch_ended)
