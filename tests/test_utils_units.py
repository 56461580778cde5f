"""CPU tests for the service / plotting / statistics units (SURVEY §2 aux components)."""
import os
import pickle

import numpy

from veles.znicz_b200.core.config import root
from veles.znicz_b200.core.memory import Array
from veles.znicz_b200.core.workflow import DummyWorkflow


class Holder(object):
    name = "holder"

    def __init__(self, **kw):
        self.__dict__.update(kw)


def test_diff_stats(tmp_path):
    from veles.znicz_b200.utils.diff_stats import DiffStats
    wf = DummyWorkflow()
    h = Holder(weights=Array(numpy.zeros((3, 3), numpy.float32)))
    ds = DiffStats(wf, arrays={h: ("weights",)}, file_name=str(tmp_path / "ds.pickle"))
    ds.initialize()
    ds.run()
    h.weights.mem += 1
    ds.run()
    h.weights.mem += 2
    ds.run()
    st = ds.stats[h]["weights"]
    assert [s["delta"] for s in st] == [9.0, 18.0]
    assert ds.size == 2
    ds.stop()
    with open(tmp_path / "ds.pickle", "rb") as f:
        assert pickle.load(f)["holder"]["weights"][1]["abs"] == 27.0


def test_fix_accumulator():
    from veles.znicz_b200.utils.accumulator import FixAccumulator
    wf = DummyWorkflow()
    acc = FixAccumulator(wf, bars=10, type="tanh")
    acc.input = Array(numpy.array([[-5.0, -1.0, 0.0, 1.0, 5.0, 1.7]], numpy.float32))
    acc.initialize()
    acc.run()
    out = acc.output.mem
    assert out.sum() == 6 and out[0] == 1 and out[11] == 1
    acc.reset_flag <<= False
    acc.run()
    assert acc.output.mem.sum() == 12


def test_range_accumulator():
    from veles.znicz_b200.utils.accumulator import RangeAccumulator
    wf = DummyWorkflow()
    acc = RangeAccumulator(wf, bars=8)
    acc.input = Array(numpy.linspace(0, 1, 100).astype(numpy.float32))
    acc.initialize()
    acc.run()
    assert sum(acc.y) == 100
    acc.input.mem[:] = numpy.linspace(-1, 2, 100)
    acc.run()
    assert sum(acc.y) == 200 and acc.gl_min == -1 and acc.gl_max == 2
    acc.reset_flag <<= True
    acc.run()
    assert sum(acc.y_out) == 200 and sum(acc.y) == 100


def test_labels_printer():
    from veles.znicz_b200.utils.labels_printer import LabelsPrinter
    wf = DummyWorkflow()
    lp = LabelsPrinter(wf, top_number=2)
    lp.input = Array(numpy.array([[0.1, 0.7, 0.2]], numpy.float32))
    lp.reversed_labels_mapping = ["a", "b", "c"]
    lp.initialize()
    lp.run()
    assert [l for _, l in lp.top] == ["b", "c"]


def test_mean_disp_normalizer_numpy():
    from veles.znicz_b200.utils.mean_disp_normalizer import MeanDispNormalizer
    wf = DummyWorkflow()
    u = MeanDispNormalizer(wf)
    x = numpy.random.RandomState(1).rand(4, 5, 5, 3).astype(numpy.float32)
    u.input = Array(x.copy())
    u.mean = Array(x.mean(axis=0))
    u.rdisp = Array(1.0 / (x.max(axis=0) - x.min(axis=0)))
    u.initialize(device="numpy")
    u.run()
    numpy.testing.assert_allclose(u.output.mem, (x - u.mean.mem) * u.rdisp.mem, rtol=1e-6)


def test_image_saver_softmax(tmp_path):
    from veles.znicz_b200.utils.image_saver import ImageSaver
    wf = DummyWorkflow()
    dirs = [str(tmp_path / d) for d in ("t", "v", "tr")]
    s = ImageSaver(wf, out_dirs=dirs, limit=5)
    rs = numpy.random.RandomState(0)
    s.input = Array(rs.rand(4, 8, 8, 3).astype(numpy.float32))
    s.indices = Array(numpy.arange(4, dtype=numpy.int32))
    s.labels = Array(numpy.array([0, 1, 2, 1], numpy.int32))
    s.max_idx = Array(numpy.array([0, 2, 2, 0], numpy.int32))
    s.output = Array(rs.rand(4, 3).astype(numpy.float32))
    s.minibatch_class, s.minibatch_size = 2, 4
    s.initialize()
    s.run()
    files = sorted(os.listdir(dirs[2]))
    assert len(files) == 2 and files[0].startswith("input_image_1_as_")
    # MSE mode
    s2 = ImageSaver(wf, out_dirs=dirs, limit=2)
    s2.input, s2.indices, s2.labels = s.input, s.indices, s.labels
    s2.output = Array(rs.rand(4, 8, 8, 3).astype(numpy.float32))
    s2.target = s.input
    s2.minibatch_class, s2.minibatch_size = 1, 4
    s2.run()
    sub = sorted(os.listdir(dirs[1]))
    assert len(sub) == 2
    assert len(os.listdir(os.path.join(dirs[1], sub[0]))) == 3


def test_weights2d_and_histogram(tmp_path):
    from veles.znicz_b200.utils.nn_plotting_units import Weights2D, MSEHistogram
    wf = DummyWorkflow()
    old = root.common.disable.plotting
    root.common.disable.plotting = False
    old_cache = root.common.dirs.cache
    root.common.dirs.cache = str(tmp_path)
    try:
        w = Weights2D(wf, limit=16)
        w.input = Array(numpy.random.rand(20, 5 * 5 * 3).astype(numpy.float32))
        w.get_shape_from = [5, 5, 3]
        w.initialize()
        w.run()
        assert len(w.pics) == 16 and w.pics[0].shape == (5, 5, 3)
        assert w.last_file and os.path.exists(w.last_file)
        h = MSEHistogram(wf, n_bars=10)
        h.mse = Array(numpy.linspace(0, 1, 50).astype(numpy.float32))
        h.initialize()
        h.run()
        assert h.val_mse.sum() == 50 and os.path.exists(h.last_file)
    finally:
        root.common.disable.plotting = old
        root.common.dirs.cache = old_cache


def test_kohonen_plotters(tmp_path):
    from veles.znicz_b200.utils import nn_plotting_units as npu
    wf = DummyWorkflow()
    old = root.common.disable.plotting
    root.common.disable.plotting = False
    old_cache = root.common.dirs.cache
    root.common.dirs.cache = str(tmp_path)
    try:
        shape = (4, 3)
        wts = Array(numpy.random.rand(12, 5).astype(numpy.float32))
        hits = Array(numpy.arange(12, dtype=numpy.int32))
        kh = npu.KohonenHits(wf)
        kh.input, kh.shape = hits, shape
        kh.run()
        assert kh.hits.shape == (3, 4) and os.path.exists(kh.last_file)
        km = npu.KohonenInputMaps(wf)
        km.input, km.shape = wts, shape
        km.run()
        assert km.maps.shape == (5, 3, 4) and os.path.exists(km.last_file)
        kn = npu.KohonenNeighborMap(wf)
        kn.input, kn.shape = wts, shape
        kn.run()
        # (W-1)*H horizontal + (2W-1)*(H-1) downward links on a hexagonal grid
        assert len(kn.link_values) == 3 * 3 + 7 * 2
        kv = npu.KohonenValidationResults(wf)
        kv.input, kv.shape = hits, shape
        kv.result = [{0, 1, 2}, {3, 4}, {11}]
        kv.fitness, kv.fitness_by_label = 0.5, [0.5, 0.4, 0.9]
        kv.fitness_by_neuron = numpy.linspace(0, 1, 12)
        kv.run()
        assert kv.cells[2][3][0] == 2 and os.path.exists(kv.last_file)
    finally:
        root.common.disable.plotting = old
        root.common.dirs.cache = old_cache


def test_similar_kernels():
    from veles.znicz_b200.utils.diversity import get_similar_kernels
    rs = numpy.random.RandomState(3)
    base = rs.randn(6, 7 * 7 * 3)
    weights = numpy.concatenate([base, base[1:2] + 0.01 * rs.randn(1, 147)])
    sets = get_similar_kernels(weights, 3)
    assert any({1, 6} <= s for s in sets)


def test_publisher_and_shell(tmp_path):
    from veles.znicz_b200.utils.publishing import Publisher
    from veles.znicz_b200.utils.interaction import Shell
    wf = DummyWorkflow()

    class P(object):
        def get_metric_names(self):
            return {"acc"}

        def get_metric_values(self):
            return {"acc": 0.9}
    p = Publisher(wf, directory=str(tmp_path), backends={"json": {}, "markdown": {}})
    p.result_providers.add(P())
    p.initialize()
    root.common.disable.publishing = False
    try:
        p.run()
    finally:
        root.common.disable.publishing = True
    assert p.report["results"]["acc"] == 0.9
    assert len(os.listdir(tmp_path)) == 2
    sh = Shell(wf)
    sh.initialize()
    sh.run()      # no tty → no-op


def test_downloader_present_files(tmp_path):
    from veles.znicz_b200.utils.downloader import Downloader
    wf = DummyWorkflow()
    (tmp_path / "a.bin").write_bytes(b"1")
    d = Downloader(wf, url="http://localhost/x.tar", directory=str(tmp_path),
                   files=["a.bin"])
    d.initialize()
    assert d.missing == []


def test_minibatches_saver_roundtrip(tmp_path):
    from veles.znicz_b200.loader.saver import MinibatchesSaver, MinibatchesLoader
    from veles.znicz_b200.loader.synthetic import SyntheticImageLoader
    wf = DummyWorkflow()
    ld = SyntheticImageLoader(wf, minibatch_size=7, shape=(6, 6, 3), n_classes=4,
                              n_train=30, n_valid=10, n_test=5, shuffle_limit=0)
    ld.initialize(device="numpy")
    sv = MinibatchesSaver(wf, file_name=str(tmp_path / "mb.dat"), compression="gz")
    sv.link_attrs(ld, "shuffle_limit", "minibatch_class", "minibatch_data",
                  "minibatch_labels", "class_lengths", "max_minibatch_size",
                  "has_labels", "labels_mapping", "minibatch_size")
    sv.initialize()
    for _ in range(12):      # > one epoch; the second pass must be ignored
        ld.run()
        sv.run()
    sv.stop()
    wf2 = DummyWorkflow()
    ml = MinibatchesLoader(wf2, file_name=str(tmp_path / "mb.dat"), minibatch_size=7,
                           shuffle_limit=0)
    ml.initialize(device="numpy")
    assert list(ml.class_lengths) == [5, 10, 30]
    ld.original_data.map_read()
    numpy.testing.assert_allclose(ml.original_data.mem, ld.original_data.mem, rtol=1e-6)


def test_model_manifests_and_packaging(tmp_path):
    """manifest.json packaging (/root/reference/samples/Wine/manifest.json)."""
    from veles.znicz_b200.utils import forge
    names = forge.list_models()
    assert {"Wine", "MNIST", "CIFAR10", "Kanji", "Lines", "YaleFaces", "DemoKohonen"} <= set(names)
    for n in names:
        m = forge.load_manifest(n)
        assert m["name"] == n and m["workflow"].endswith(".py")
    pkg = forge.pack("Wine", str(tmp_path / "wine.tar.gz"))
    m = forge.unpack(pkg, str(tmp_path / "out"))
    assert m["name"] == "Wine" and (tmp_path / "out" / "wine.py").is_file()
