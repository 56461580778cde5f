"""Sample workflows on generated miniature datasets, numpy backend."""
import os

import numpy
import pytest

pytest.importorskip("cv2")

from veles.znicz_b200.core.config import root  # noqa: E402


@pytest.fixture(autouse=True)
def _dirs(tmp_path):
    old = (root.common.dirs.cache, root.common.dirs.snapshots)
    root.common.dirs.cache = str(tmp_path / "cache")
    root.common.dirs.snapshots = str(tmp_path / "snap")
    os.makedirs(root.common.dirs.cache, exist_ok=True)
    os.makedirs(root.common.dirs.snapshots, exist_ok=True)
    yield
    root.common.dirs.cache, root.common.dirs.snapshots = old


def test_kanji_mse_workflow(tmp_path):
    from veles.znicz_b200.models import kanji
    d = kanji.generate_dataset(str(tmp_path / "kanji"), glyphs="ABCD", per_glyph=10)
    layers = [dict(l) for l in kanji.root.kanji.layers]
    for l in layers:
        l["<-"] = dict(l["<-"], learning_rate=0.001)
    loader = dict(root.kanji.loader.to_dict(), minibatch_size=8,
                  train_paths=[os.path.join(d, "train")],
                  target_paths=[os.path.join(d, "target")])
    wf = kanji.build(loader_config=loader, layers=layers,
                     decision_config={"max_epochs": 6, "fail_iterations": 50},
                     image_saver_config={"out_dirs": [str(tmp_path / ("img%d" % i))
                                                      for i in range(3)], "limit": 3},
                     snapshotter_config={"prefix": "kanji_t", "interval": 100,
                                         "time_interval": 1e9})
    wf.initialize(device="numpy")
    first = []
    wf.step_hooks_.append(lambda w: first.append(float(w.decision.epoch_metrics[2][0]))
                          if bool(w.loader.epoch_ended) else None)
    wf.run()
    assert bool(wf.decision.complete)
    assert wf.loader.class_targets.shape == (4, 24, 24)
    assert len(first) >= 2 and first[-1] < first[0]          # train MSE goes down
    # the MSE-mode image saver wrote input/output/target triples
    assert any(f.startswith("target_") for _, _, fs in os.walk(str(tmp_path)) for f in fs)
    # weight injection API of the sample
    w0 = [f.weights.mem.copy() * 0 + 0.01 for f in wf.forwards]
    b0 = [f.bias.mem.copy() * 0 for f in wf.forwards]
    wf2 = kanji.build(loader_config=loader, layers=layers,
                      decision_config={"max_epochs": 1, "fail_iterations": 5})
    wf2.initialize(device="numpy", weights=w0, bias=b0)
    assert numpy.allclose(wf2.forwards[1].weights.mem, 0.01)


def test_lines_mcdnnic_workflow(tmp_path):
    from veles.znicz_b200.models import lines
    d = lines.generate_dataset(str(tmp_path / "lines"), size=32, per_class=(6, 2))
    loader = dict(root.lines.loader.to_dict(), minibatch_size=6,
                  train_paths=[os.path.join(d, "learn")],
                  validation_paths=[os.path.join(d, "test")])
    wf = lines.build(loader_config=loader, mcdnnic_topology="6x32x32-8C4-MP2-8C4-MP3-16N-4N",
                     decision_config={"max_epochs": 2, "fail_iterations": 10},
                     image_saver_config={"out_dirs": [str(tmp_path / ("i%d" % i))
                                                      for i in range(3)]},
                     snapshotter_config={"prefix": "lines_t", "interval": 100,
                                         "time_interval": 1e9})
    types = [type(f).__name__ for f in wf.forwards]
    assert types == ["Conv", "MaxPooling", "Conv", "MaxPooling", "All2All",
                     "All2AllSoftmax"], types
    wf.initialize(device="numpy")
    wf.run()
    assert bool(wf.decision.complete)
    assert wf.table_plotter.rows and wf.multi_hist_plotter[-1].histograms


def test_yale_faces_and_preprocessing(tmp_path):
    from veles.znicz_b200.models import yale_faces as yf
    from veles.znicz_b200.loader.saver import read_minibatches
    d = yf.generate_dataset(str(tmp_path / "CroppedYale"), subjects=4, per_subject=10)
    loader = dict(root.yalefaces.loader.to_dict(), minibatch_size=8, train_paths=[d])
    layers = [dict(l) for l in root.yalefaces.layers]
    layers[1] = dict(layers[1], **{"->": dict(layers[1]["->"], output_sample_shape=4)})
    root.common.disable.publishing = False
    try:
        wf = yf.build(loader_config=loader, layers=layers,
                      decision_config={"max_epochs": 8, "fail_iterations": 20},
                      publisher_config={"backends": {"json": {}},
                                        "directory": str(tmp_path / "reports")},
                      snapshotter_config={"prefix": "yale_t", "interval": 100,
                                          "time_interval": 1e9})
        wf.initialize(device="numpy")
        wf.run()
    finally:
        root.common.disable.publishing = True
    assert wf.loader.total_samples == 40            # Ambient files are ignored
    assert wf.decision.best_n_err_pt[1] < 40.0
    assert wf.publisher.report and os.listdir(str(tmp_path / "reports"))
    # preprocessing variant: loader -> MinibatchesSaver only
    out = str(tmp_path / "mb.dat")
    pre = yf.build_preprocessing(
        loader_config=dict(loader, shuffle_limit=0),
        data_saver_config={"file_name": out, "compression": "gz"})
    pre.initialize(device="numpy")
    pre.run()
    items = list(read_minibatches(out))
    assert sum(len(it[1]) for it in items[1:]) == 40


def test_demo_kohonen(tmp_path):
    from veles.znicz_b200.models import kohonen as km
    path = km.generate_dataset(str(tmp_path / "kohonen.txt.gz"), n=200)
    root.kohonen.loader.minibatch_size = 10
    wf = km.build(dataset_file=path, epochs=4)
    wf.initialize(device="numpy")
    w0 = wf.trainer.weights.mem.copy()
    wf.run()
    assert bool(wf.decision.complete)
    assert numpy.abs(wf.trainer.weights.mem - w0).max() > 1e-3
    assert wf.plotters[0].hits is not None and wf.plotters[0].hits.sum() == 200
    assert wf.plotters[2].link_values is not None


def test_forward_from_snapshot_on_new_images(tmp_path):
    """Train the MNIST FC sample briefly, then classify image files with the restored
    workflow in testing mode (samples/MNIST/mnist_forward.py equivalent)."""
    import glob
    import cv2
    from veles.znicz_b200.models import mnist, mnist_forward
    wf = mnist.build(
        layers=mnist.fc_layers(), loader_name="synthetic_mnist",
        loader_config={"minibatch_size": 20, "n_train": 300, "n_valid": 60, "noise": 0.3,
                       "normalization_type": "linear"},
        decision_config={"max_epochs": 4, "fail_iterations": 10},
        snapshotter_config={"prefix": "mnist_fw", "interval": 1, "time_interval": 0,
                            "compression": "", "directory": str(tmp_path / "snap")})
    wf.initialize(device="numpy")
    wf.run()
    snaps = sorted(f for f in glob.glob(str(tmp_path / "snap" / "mnist_fw*.pickle"))
                   if not os.path.islink(f))
    assert snaps
    # "new pictures": a few validation samples written as PNG files under <label>/ dirs
    ld = wf.loader
    ld.original_data.map_read()
    truth = []
    for k in range(12):
        img = ld.original_data.mem[k, :, :, 0]
        lbl = ld.original_labels[k]
        d = tmp_path / "pics" / str(lbl)
        os.makedirs(d, exist_ok=True)
        u8 = ((img - img.min()) / (img.max() - img.min()) * 255).astype(numpy.uint8)
        cv2.imwrite(str(d / ("p%02d.png" % k)), u8)
        truth.append(lbl)
    out = str(tmp_path / "result.json")
    fwf, results = mnist_forward.forward_from_snapshot(
        snaps[-1], [str(tmp_path / "pics")], result_file=out, device="numpy")
    assert fwf.loader.class_lengths[0] == 12 and bool(fwf.decision.complete)
    assert "Output" in results and os.path.exists(out)
    probs = numpy.asarray(fwf.evaluator.merged_output)
    assert probs.shape == (12, 10)
    numpy.testing.assert_allclose(probs.sum(axis=1), 1.0, rtol=1e-4)
