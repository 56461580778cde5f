"""Caffe proto codec, LMDB (pure-python env), STL-10 and ImageNet raw loaders."""
import json
import os
import pickle

import numpy
import pytest

from veles.znicz_b200.core.workflow import DummyWorkflow
from veles.znicz_b200.loader import lmdb_mini
from veles.znicz_b200.loader.caffe import protobuf2 as pb


def test_datum_roundtrip_and_google_protobuf_crosscheck():
    img = numpy.random.RandomState(0).randint(0, 255, (5, 7, 3)).astype(numpy.uint8)
    d = pb.Datum.from_hwc(img, label=300)
    raw = d.SerializeToString()
    d2 = pb.Datum.FromString(raw)
    assert (d2.channels, d2.height, d2.width, d2.label) == (3, 5, 7, 300)
    numpy.testing.assert_array_equal(d2.to_hwc(), img)
    fd = pb.Datum.from_hwc(img.astype(numpy.float32) / 3, label=-2)
    fd2 = pb.Datum.FromString(fd.SerializeToString())
    assert fd2.label == -2
    numpy.testing.assert_allclose(fd2.to_hwc(), img.astype(numpy.float32) / 3, rtol=1e-6)
    # cross-check the wire bytes against google.protobuf built from a dynamic descriptor
    gp = pytest.importorskip("google.protobuf")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fdp = descriptor_pb2.FileDescriptorProto(name="datum_t.proto", package="t", syntax="proto2")
    m = fdp.message_type.add(name="Datum")
    T = descriptor_pb2.FieldDescriptorProto
    for name, num, typ, lab in (("channels", 1, T.TYPE_INT32, T.LABEL_OPTIONAL),
                                ("height", 2, T.TYPE_INT32, T.LABEL_OPTIONAL),
                                ("width", 3, T.TYPE_INT32, T.LABEL_OPTIONAL),
                                ("data", 4, T.TYPE_BYTES, T.LABEL_OPTIONAL),
                                ("label", 5, T.TYPE_INT32, T.LABEL_OPTIONAL),
                                ("float_data", 6, T.TYPE_FLOAT, T.LABEL_REPEATED),
                                ("encoded", 7, T.TYPE_BOOL, T.LABEL_OPTIONAL)):
        m.field.add(name=name, number=num, type=typ, label=lab)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    G = message_factory.GetMessageClass(pool.FindMessageTypeByName("t.Datum"))
    g = G()
    g.ParseFromString(raw)
    assert (g.channels, g.height, g.width, g.label) == (3, 5, 7, 300)
    assert g.data == d.data
    g2 = G(channels=1, height=2, width=2, label=-7, float_data=[1.5, 2.5, 3.5, 4.5])
    ours = pb.Datum.FromString(g2.SerializeToString())       # unpacked repeated floats
    assert ours.label == -7 and ours.float_data == [1.5, 2.5, 3.5, 4.5]


def test_blobproto_and_net_messages():
    arr = numpy.arange(24, dtype=numpy.float32).reshape(2, 3, 4)
    b = pb.BlobProto.FromString(pb.BlobProto.from_array(arr).SerializeToString())
    numpy.testing.assert_array_equal(b.to_array(), arr)
    legacy = pb.BlobProto(num=1, channels=3, height=2, width=2, data=list(range(12)))
    assert pb.BlobProto.FromString(legacy.SerializeToString()).to_array().shape == (3, 2, 2)
    layer = pb.LayerParameter(
        name="conv1", type="Convolution", bottom=["data"], top=["conv1"],
        convolution_param=pb.ConvolutionParameter(num_output=32, kernel_size=[5], pad=[2],
                                                  stride=[1]),
        blobs=[pb.BlobProto.from_array(numpy.ones((32, 3, 5, 5)))])
    net = pb.NetParameter(name="n", layer=[layer])
    net2 = pb.NetParameter.FromString(net.SerializeToString())
    assert net2.layer[0].convolution_param.num_output == 32
    assert net2.layer[0].blobs[0].to_array().shape == (32, 3, 5, 5)
    txt = pb.parse_net_text('''
        name: "CIFAR10_quick"  # comment
        layer { name: "conv1" type: "Convolution" bottom: "data" top: "conv1"
                convolution_param { num_output: 32 pad: 2 kernel_size: 5 stride: 1 } }
        layer { name: "pool1" type: "Pooling" pooling_param { pool: MAX kernel_size: 3 } }
    ''')
    assert txt["name"] == "CIFAR10_quick" and len(txt["layer"]) == 2
    assert txt["layer"][0]["convolution_param"]["kernel_size"] == 5
    assert txt["layer"][1]["pooling_param"]["pool"] == "MAX"


def test_lmdb_mini_deep_tree(tmp_path):
    rs = numpy.random.RandomState(1)
    items = [(b"k%07d" % i, rs.bytes(int(rs.choice([3, 40, 200, 700])))) for i in range(5000)]
    lmdb_mini.write_environment(str(tmp_path), items, psize=512)
    env = lmdb_mini.open(str(tmp_path))
    st = env.stat()
    assert st["entries"] == 5000 and st["depth"] >= 3 and st["overflow_pages"] > 0
    assert list(env.cursor()) == sorted(items)
    ref = dict(items)
    for k in (b"k0000000", b"k0002500", b"k0004999"):
        assert env.get(k) == ref[k]
    assert env.get(b"k9999999") is None and env.get(b"a") is None
    env.close()


def _make_lmdb(path, n, seed, shape=(8, 8, 3), n_labels=4):
    rs = numpy.random.RandomState(seed)
    imgs = rs.randint(0, 255, (n,) + shape).astype(numpy.uint8)
    labels = rs.randint(0, n_labels, n)
    items = [(b"%08d" % i, pb.Datum.from_hwc(imgs[i], labels[i]).SerializeToString())
             for i in range(n)]
    lmdb_mini.write_environment(path, items)
    return imgs, labels


def test_lmdb_loader(tmp_path):
    from veles.znicz_b200.loader.loader_lmdb import LMDBLoader, FullBatchLMDBLoader
    timgs, tlabels = _make_lmdb(str(tmp_path / "train"), 23, 0)
    vimgs, vlabels = _make_lmdb(str(tmp_path / "val"), 9, 1)
    wf = DummyWorkflow()
    ld = LMDBLoader(wf, train_path=str(tmp_path / "train"), validation_path=str(tmp_path / "val"),
                    minibatch_size=5, shuffle_limit=0, normalization_type="none")
    ld.initialize(device="numpy")
    assert list(ld.class_lengths) == [0, 9, 23]
    ld.run()
    assert ld.minibatch_class == 1
    numpy.testing.assert_array_equal(ld.minibatch_data.mem[:5], vimgs[:5])
    mapped = [ld.reversed_labels_mapping[i] for i in ld.minibatch_labels.mem[:5]]
    assert mapped == list(vlabels[:5])
    fb = FullBatchLMDBLoader(wf, train_path=str(tmp_path / "train"), minibatch_size=6,
                             normalization_type="linear")
    fb.initialize(device="numpy")
    assert fb.original_data.shape == (23, 8, 8, 3) and fb.class_lengths[2] == 23


def test_stl10_loader(tmp_path):
    from veles.znicz_b200.loader.loader_stl import STL10FullBatchLoader
    rs = numpy.random.RandomState(2)
    size = (6, 6)
    (tmp_path / "class_names.txt").write_text("airplane bird car\n")
    data = {}
    for stem, n in (("train", 7), ("test", 4)):
        x = rs.randint(0, 255, (n, 3, size[0], size[1])).astype(numpy.uint8)
        y = rs.randint(1, 4, n).astype(numpy.uint8)
        x.tofile(str(tmp_path / (stem + "_X.bin")))
        y.tofile(str(tmp_path / (stem + "_y.bin")))
        data[stem] = (x, y)
    wf = DummyWorkflow()
    ld = STL10FullBatchLoader(wf, directory=str(tmp_path), size=size, minibatch_size=3,
                              normalization_type="none")
    ld.initialize(device="numpy")
    assert list(ld.class_lengths) == [0, 4, 7]
    x, y = data["test"]
    numpy.testing.assert_array_equal(ld.original_data.mem[0], x[0].transpose(2, 1, 0))
    assert ld.original_labels[0] == ["airplane", "bird", "car"][y[0] - 1]


def test_imagenet_loaders(tmp_path):
    from veles.znicz_b200.loader.imagenet_loader import ImagenetLoaderBase, ImagenetLoader
    rs = numpy.random.RandomState(3)
    n_val, n_train, sy, sx = 6, 20, 12, 14
    n = n_val + n_train
    samples = rs.randint(0, 255, (n, sy, sx, 3)).astype(numpy.uint8)
    samples.tofile(str(tmp_path / "samples.dat"))
    labels = [("n%04d" % (i % 5), i % 5) for i in range(n)]
    with open(tmp_path / "labels.pickle", "wb") as f:
        pickle.dump(labels, f)
    with open(tmp_path / "count.json", "w") as f:
        json.dump({"test": 0, "val": n_val, "train": n_train}, f)
    mean = samples[n_val:].mean(axis=0)
    rdisp = 1.0 / (samples[n_val:].std(axis=0) + 1.0)
    with open(tmp_path / "matrixes.pickle", "wb") as f:
        pickle.dump([mean, rdisp], f)
    kw = dict(sx=sx, sy=sy, original_labels_filename=str(tmp_path / "labels.pickle"),
              count_samples_filename=str(tmp_path / "count.json"),
              samples_filename=str(tmp_path / "samples.dat"),
              matrixes_filename=str(tmp_path / "matrixes.pickle"), minibatch_size=4)
    wf = DummyWorkflow()
    base = ImagenetLoaderBase(wf, shuffle_limit=0, **kw)
    base.initialize(device="numpy")
    assert list(base.class_lengths) == [0, 6, 20] and base.unique_labels_count == 5
    base.run()
    numpy.testing.assert_array_equal(base.minibatch_data.mem, samples[:4])
    assert list(base.minibatch_labels.mem) == [0, 1, 2, 3]
    assert base.mean.shape == (sy, sx, 3) and base.rdisp.shape == (sy, sx, 3)
    aug = ImagenetLoader(wf, crop_size_sx=8, crop_size_sy=10, mirror=True, **kw)
    aug.initialize(device="numpy")
    assert aug.minibatch_data.shape == (4, 10, 8, 3)
    aug.run()                      # VALID: centre crop, mean subtracted, no mirror
    idx = aug.minibatch_indices.mem[:4]
    want = (samples[idx].astype(numpy.float32) - mean.astype(numpy.float32))[:, 1:11, 3:11]
    numpy.testing.assert_allclose(aug.minibatch_data.mem, want, rtol=1e-5, atol=1e-4)
    aug.run()
    aug.run()                      # TRAIN minibatch: random crops stay in bounds
    assert aug.minibatch_class == 2 and numpy.isfinite(aug.minibatch_data.mem).all()


def test_preparation_imagenet_roundtrip(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from veles.znicz_b200.loader.imagenet_loader import ImagenetLoader
    from veles.znicz_b200.utils import preparation_imagenet as prep
    rs = numpy.random.RandomState(4)
    for split, n in (("train", 5), ("val", 2)):
        for label in ("n01", "n02", "n03"):
            d = tmp_path / "src" / split / label
            os.makedirs(d)
            for k in range(n):
                img = rs.randint(0, 255, (20 + k, 30, 3)).astype(numpy.uint8)
                cv2.imwrite(str(d / ("im%d.png" % k)), img)
    info = prep.prepare(str(tmp_path / "src"), str(tmp_path / "out"), size=16, workers=2)
    assert info["samples"] == 21 and info["labels"] == 3
    wf = DummyWorkflow()
    ld = ImagenetLoader(wf, minibatch_size=4, crop_size_sx=12, crop_size_sy=12, mirror=True,
                        **info["loader_config"])
    ld.initialize(device="numpy")
    assert list(ld.class_lengths) == [0, 6, 15] and ld.unique_labels_count == 3
    assert ld.has_mean_file and ld.mean.shape == (16, 16, 3)
    ld.run()
    assert ld.minibatch_data.shape == (4, 12, 12, 3) and numpy.isfinite(ld.minibatch_data.mem).all()


def test_xmltodict_and_bbox_preparation(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from veles.znicz_b200.external import xmltodict
    from veles.znicz_b200.utils import preparation_imagenet as prep
    doc = ('<annotation v="1"><folder>n01</folder>'
           '<object><name>cat</name><bndbox><xmin>2</xmin><ymin>3</ymin><xmax>12</xmax>'
           '<ymax>13</ymax></bndbox></object>'
           '<object><name>dog</name><bndbox><xmin>20</xmin><ymin>0</ymin><xmax>30</xmax>'
           '<ymax>10</ymax></bndbox></object></annotation>')
    tree = xmltodict.parse(doc)
    assert tree["annotation"]["@v"] == "1" and tree["annotation"]["folder"] == "n01"
    assert [o["name"] for o in tree["annotation"]["object"]] == ["cat", "dog"]
    assert xmltodict.parse(xmltodict.unparse(tree, pretty=True)) == tree
    single = xmltodict.parse("<a><b>1</b></a>", force_list=("b",))
    assert single["a"]["b"] == ["1"]
    # one sample per bounding box, labelled by the object name
    d = tmp_path / "src" / "train" / "n01"
    os.makedirs(d)
    img = numpy.zeros((24, 32, 3), numpy.uint8)
    img[3:13, 2:12] = 200          # the "cat" box is bright, the "dog" box stays dark
    cv2.imwrite(str(d / "im0.png"), img)
    cv2.imwrite(str(d / "im1.png"), img)          # no annotation → whole frame, dir label
    ann = tmp_path / "ann" / "train" / "n01"
    os.makedirs(ann)
    (ann / "im0.xml").write_text(doc)
    info = prep.prepare(str(tmp_path / "src"), str(tmp_path / "out"), size=8, workers=1,
                        annotations=str(tmp_path / "ann"))
    assert info["samples"] == 3 and info["labels"] == 3
    data = numpy.memmap(info["loader_config"]["samples_filename"], dtype=numpy.uint8,
                        mode="r", shape=(3, 8, 8, 3))
    assert data[0].min() == 200 and data[1].max() == 0
    import pickle
    with open(info["loader_config"]["original_labels_filename"], "rb") as f:
        labels = [l for l, _ in pickle.load(f)]
    assert labels == ["cat", "dog", "n01"]
