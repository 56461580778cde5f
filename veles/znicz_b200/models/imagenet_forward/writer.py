"""Result files of the localisation pipeline: the JSON writer unit
(/root/reference/tests/research/ImagenetAE/imagenet_forward/forward_json.py:46-119), the ILSVRC
submission converters (json2txt.py:52-200), json merging (merge_json.py) and extraction of one
picture's raw boxes from a pickled stream (raw_bbox_extract.py)."""
from __future__ import annotations

import json
import os
import pickle

import numpy

from ...core.units import Unit
from .bbox import BBox


class ResultWriter(Unit):
    """``winners`` ({"path", "bbxs": [(label index, confidence, box)]}) -> one JSON object per
    picture: {"path", "label", "width", "height", "bbxs": [{"conf", "label", "angle", "x", "y",
    "width", "height"}]}. The box is [ymin, xmin, ymax, xmax] in "merge" mode and
    (x_center, y_center, width, height) in "final" mode."""
    hide_from_registry = True

    def __init__(self, workflow, labels_txt=None, result_path=None, **kwargs):
        super().__init__(workflow, **kwargs)
        self.labels_txt = labels_txt
        self.result_path = result_path
        self.labels_mapping = dict(kwargs.get("labels_mapping", {}))
        self.ignore_negative = kwargs.get("ignore_negative", True)
        self.image_size_fn = kwargs.get("image_size_fn")
        self.mode = kwargs.get("mode", "")
        self.results = {}
        self.demand("winners", "mode")

    def initialize(self, **kwargs):
        self.results = {}
        if self.labels_txt:
            with open(self.labels_txt) as txt:
                values = txt.read().split()
            self.labels_mapping.update(zip(map(int, values[::2]), values[1::2]))

    def _size(self, path):
        if self.image_size_fn is not None:
            return self.image_size_fn(path)
        try:
            from PIL import Image
            return Image.open(path).size
        except Exception:
            return (-1, -1)

    def run(self):
        if self.winners is None:
            return
        shift = 1 if self.ignore_negative else 0
        for win in self.winners:
            out = []
            for label, conf, box in win["bbxs"]:
                if self.mode == "merge":
                    h, w = box[2] - box[0], box[3] - box[1]
                    if w <= 0 or h <= 0:
                        raise ValueError("degenerate box %s for %s" % (box, win["path"]))
                    x, y = (box[3] + box[1]) / 2.0, (box[2] + box[0]) / 2.0
                elif self.mode == "final":
                    x, y, w, h = box
                else:
                    raise ValueError("ResultWriter.mode must be 'merge' or 'final'")
                out.append({"conf": float(conf),
                            "label": self.labels_mapping.get(label + shift, str(label + shift)),
                            "angle": "0", "x": int(numpy.round(x)), "y": int(numpy.round(y)),
                            "width": int(numpy.round(w)), "height": int(numpy.round(h))})
            width, height = self._size(win["path"])
            self.results[os.path.basename(win["path"])] = {
                "path": win["path"], "label": "", "width": int(width), "height": int(height),
                "bbxs": out}
        if self.result_path:
            tmp = self.result_path + ".tmp"
            with open(tmp, "w") as fout:
                json.dump(self.results, fout, indent=4)
            os.replace(tmp, self.result_path)


# ---- submission text formats -------------------------------------------------------------------
def bbox_min_max(bbox, image_wh):
    """(xmin, ymin, xmax, ymax) of an {x, y, width, height} box clipped to the picture."""
    w, h = bbox["width"], bbox["height"]
    if w <= 0 or h <= 0:
        raise ValueError("invalid box size")
    x0, y0 = bbox["x"] - w // 2, bbox["y"] - h // 2
    return (max(x0, 0), max(y0, 0), min(x0 + w, image_wh[0]), min(y0 + h, image_wh[1]))


def _dims(val):
    wh = (val.get("width", -1), val.get("height", -1))
    return (100000, 100000) if wh == (-1, -1) else wh


def convert_det(results, image_index, label_index, out):
    """DET format: ``<image index> <class id> <confidence> <xmin> <ymin> <xmax> <ymax>`` per
    object. ``image_index``: {picture name without extension: int}, ``label_index``: {label: id}
    (the reference reads both from the development kit, json2txt.py:81-120). Returns lines
    written."""
    n = 0
    for key, val in sorted(results.items()):
        for bbox in val["bbxs"]:
            try:
                mm = bbox_min_max(bbox, _dims(val))
                img, lbl = image_index[os.path.splitext(key)[0]], label_index[bbox["label"]]
            except (ValueError, KeyError):
                continue
            out.write("%d %d %.3f %d %d %d %d \n" % ((img, lbl, bbox["conf"]) + tuple(mm)))
            n += 1
    return n


def convert_cls_loc(results, label_index, out, names):
    """CLS-LOC format: one line per picture of ``names`` (in that order) with up to five
    ``<class id> <xmin> <ymin> <xmax> <ymax>`` groups, best confidence first; ``0 0 1 0 1`` for a
    picture without detections (json2txt.py:123-168)."""
    for name in names:
        line = ""
        boxes = [b for b in results.get(name, {}).get("bbxs", ()) if b["label"] in label_index]
        for b in sorted(boxes, key=lambda b: b["conf"], reverse=True)[:5]:
            bb = BBox.from_json_dict(b)
            line += "%d %.0f %.0f %.0f %.0f " % (label_index[b["label"]], bb.xmin, bb.ymin,
                                                  bb.xmax, bb.ymax)
        out.write((line or "0 0 1 0 1") + "\n")


def merge_json(paths, out_path):
    """Union of several result files (one per rank / image range)."""
    merged = {}
    for p in paths:
        with open(p) as fin:
            merged.update(json.load(fin))
    with open(out_path, "w") as fout:
        json.dump(merged, fout, indent=4)
    return merged


def extract_raw_bboxes(stream_path, name_part, out_path=None):
    """First record of a pickled (key, meta) stream whose key contains ``name_part``."""
    with open(stream_path, "rb") as fin:
        while True:
            try:
                key, meta = pickle.load(fin)
            except EOFError:
                raise KeyError(name_part)
            if key.find(name_part) >= 0:
                break
    if out_path:
        with open(out_path, "w") as fout:
            json.dump({key: meta}, fout, indent=4)
    return key, meta
