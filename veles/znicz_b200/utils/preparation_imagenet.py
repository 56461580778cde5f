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
CLI:  python -m veles.znicz_b200.utils.preparation_imagenet <root> <out_dir> [--size 256]
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
from concurrent.futures import ThreadPoolExecutor

import numpy

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


def prepare(root_dir, out_dir, size=256, name="imagenet", series="img", color_space="RGB",
            workers=8, maintain_aspect=True, background=(127, 127, 127)):
    files = scan(root_dir)
    order = [it for split, _ in SPLITS for it in files[split]]
    if not order:
        raise ValueError("no images under %s/{test,val,train}/<label>/" % root_dir)
    labels = sorted({l for _, l in order})
    label_id = {l: i for i, l in enumerate(labels)}
    os.makedirs(out_dir, exist_ok=True)
    stem = "%s_%s" % (name, series)
    dat_path = os.path.join(out_dir, "original_data_%s.dat" % stem)
    n = len(order)
    out = numpy.memmap(dat_path, dtype=numpy.uint8, mode="w+", shape=(n, size, size, 3))

    def convert(i):
        img = read_image(order[i][0], color_space)
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
        pickle.dump([(l, label_id[l]) for _, l in order], f, protocol=4)
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
    a = ap.parse_args(argv)
    info = prepare(a.root, a.out_dir, a.size, a.name, a.series, workers=a.workers)
    print(json.dumps(info, indent=1))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
