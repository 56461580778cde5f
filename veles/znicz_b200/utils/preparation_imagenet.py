"""Dataset preparation for the raw ImageNet loaders.

Capability parity with /root/reference/tests/research/utils/preparation_imagenet.py:140-855
(scan a directory tree, resize every picture to a fixed frame, dump one flat ``uint8`` sample
file + labels pickle + counts json + mean / reciprocal-dispersion matrices), re-designed as a
streaming converter: images are decoded by a thread pool, written straight into a memory-mapped
output file (no per-image python list), and the statistics are accumulated in float64 on the
fly, so the tool's memory use is independent of the dataset size.

Input layout:  ``<root>/{test,val,train}/<label>/*.{jpg,jpeg,png,...}`` (any split may be absent).
Outputs (names match ``models/alexnet.py`` / the loader kwargs):
    original_data_<name>_<series>.dat, original_labels_<name>_<series>.pickle,
    count_samples_<name>_<series>.json, matrixes_<name>_<series>.pickle
Bounding boxes (reference :236-262,386-430): with ``annotations=<dir>`` every picture that has
a PASCAL-VOC style ``<annotations>/<split>/<label>/<stem>.xml`` (or ``<stem>.xml`` next to the
picture) is expanded into one sample per ``<object>`` — the crop of its ``bndbox`` labelled with
the object's ``name`` — which is the reference's "DET" series; pictures without an annotation
keep the whole frame and the directory label.
CLI:  python -m veles.znicz_b200.utils.preparation_imagenet <root> <out_dir> [--size 256]
          [--annotations DIR]
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
from concurrent.futures import ThreadPoolExecutor

import numpy

from ..external import xmltodict
from ..loader.image import fit_image, image_subtype, read_image

SPLITS = (("test", "test"), ("val", "val"), ("train", "train"))


def scan(root_dir):
    """→ {split: [(path, label)]} sorted for reproducibility."""
    out = {}
    for split, _ in SPLITS:
        base = os.path.join(root_dir, split)
        items = []
        if os.path.isdir(base):
            for label in sorted(os.listdir(base)):
                d = os.path.join(base, label)
                if not os.path.isdir(d):
                    continue
                for f in sorted(os.listdir(d)):
                    if image_subtype(f) is not None:
                        items.append((os.path.join(d, f), label))
        out[split] = items
    return out


def read_bboxes(xml_path):
    """PASCAL-VOC annotation → [(label, xmin, ymin, xmax, ymax)] (empty when no objects)."""
    with open(xml_path, "r") as f:
        tree = xmltodict.parse(f.read())
    objs = (tree.get("annotation") or {}).get("object")
    if objs is None:
        return []
    if not isinstance(objs, list):
        objs = [objs]
    out = []
    for o in objs:
        bb = o["bndbox"]
        x0, y0, x1, y1 = (int(float(bb[k])) for k in ("xmin", "ymin", "xmax", "ymax"))
        if x1 > x0 and y1 > y0:
            out.append((o["name"], x0, y0, x1, y1))
    return out


def expand_bboxes(files, root_dir, annotations):
    """Replaces (path, label) by (path, label, bbox|None), one entry per annotated object."""
    out = {}
    for split, items in files.items():
        res = []
        for path, label in items:
            stem = os.path.splitext(os.path.basename(path))[0] + ".xml"
            rel = os.path.relpath(os.path.dirname(path), root_dir)
            boxes = []
            for cand in (os.path.join(annotations, rel, stem) if annotations else None,
                         os.path.join(os.path.dirname(path), stem)):
                if cand and os.path.exists(cand):
                    boxes = read_bboxes(cand)
                    break
            if boxes:
                res.extend((path, b[0], b[1:]) for b in boxes)
            else:
                res.append((path, label, None))
        out[split] = res
    return out


def prepare(root_dir, out_dir, size=256, name="imagenet", series="img", color_space="RGB",
            workers=8, maintain_aspect=True, background=(127, 127, 127), annotations=None):
    files = scan(root_dir)
    if annotations is not None:
        files = expand_bboxes(files, root_dir, annotations)
    else:
        files = {k: [(p, l, None) for p, l in v] for k, v in files.items()}
    order = [it for split, _ in SPLITS for it in files[split]]
    if not order:
        raise ValueError("no images under %s/{test,val,train}/<label>/" % root_dir)
    labels = sorted({it[1] for it in order})
    label_id = {l: i for i, l in enumerate(labels)}
    os.makedirs(out_dir, exist_ok=True)
    stem = "%s_%s" % (name, series)
    dat_path = os.path.join(out_dir, "original_data_%s.dat" % stem)
    n = len(order)
    out = numpy.memmap(dat_path, dtype=numpy.uint8, mode="w+", shape=(n, size, size, 3))

    def convert(i):
        img = read_image(order[i][0], color_space)
        box = order[i][2]
        if box is not None:
            x0, y0, x1, y1 = box
            crop = img[max(0, y0):min(img.shape[0], y1), max(0, x0):min(img.shape[1], x1)]
            if crop.size:
                img = crop
        if img.shape[2] == 1:
            img = numpy.repeat(img, 3, axis=2)
        out[i] = fit_image(img[:, :, :3], (size, size), maintain_aspect, background)
        return i

    with ThreadPoolExecutor(max_workers=workers) as pool:
        list(pool.map(convert, range(n)))
    # statistics over the TRAIN part (whole set when there is no train split)
    n_test, n_val = len(files["test"]), len(files["val"])
    first = n_test + n_val if files["train"] else 0
    s1 = numpy.zeros((size, size, 3), numpy.float64)
    s2 = numpy.zeros((size, size, 3), numpy.float64)
    for i in range(first, n):
        x = out[i].astype(numpy.float64)
        s1 += x
        s2 += x * x
    cnt = max(1, n - first)
    mean = s1 / cnt
    disp = numpy.sqrt(numpy.maximum(s2 / cnt - mean * mean, 0.0))
    rdisp = 1.0 / numpy.maximum(disp, 1.0)
    out.flush()
    del out
    with open(os.path.join(out_dir, "original_labels_%s.pickle" % stem), "wb") as f:
        pickle.dump([(it[1], label_id[it[1]]) for it in order], f, protocol=4)
    with open(os.path.join(out_dir, "count_samples_%s.json" % stem), "w") as f:
        json.dump({"test": n_test, "val": n_val, "train": len(files["train"])}, f)
    with open(os.path.join(out_dir, "matrixes_%s.pickle" % stem), "wb") as f:
        pickle.dump([mean.astype(numpy.float32), rdisp.astype(numpy.float32)], f, protocol=4)
    return {"samples": n, "labels": len(labels), "size": size,
            "loader_config": {
                "sx": size, "sy": size, "channels": 3,
                "samples_filename": dat_path,
                "original_labels_filename": os.path.join(
                    out_dir, "original_labels_%s.pickle" % stem),
                "count_samples_filename": os.path.join(out_dir, "count_samples_%s.json" % stem),
                "matrixes_filename": os.path.join(out_dir, "matrixes_%s.pickle" % stem)}}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("root")
    ap.add_argument("out_dir")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--name", default="imagenet")
    ap.add_argument("--series", default="img")
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--annotations", default=None,
                    help="directory with PASCAL-VOC xml files (one sample per bounding box)")
    a = ap.parse_args(argv)
    info = prepare(a.root, a.out_dir, a.size, a.name, a.series, workers=a.workers,
                   annotations=a.annotations)
    print(json.dumps(info, indent=1))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
