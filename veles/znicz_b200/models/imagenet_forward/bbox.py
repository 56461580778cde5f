"""Bounding-box geometry, NMS and merging (/root/reference/tests/research/ImagenetAE/
imagenet_forward/forward_bbox.py:43-613).

Boxes are rows ``[ymin, xmin, ymax, xmax]`` with INCLUSIVE pixel bounds (width = xmax - xmin + 1),
the reference's "caffe view"; the "center view" is ``(x_center, y_center, width, height)``.
Every pairwise quantity is computed for whole arrays at once (numpy broadcasting) - the reference
evaluates python-level pair loops.
"""
from __future__ import annotations

import numpy


class BBox(object):
    """Axis-aligned box in inclusive pixel coordinates (forward_bbox.py:43-112)."""
    __slots__ = ("ymin", "xmin", "ymax", "xmax")

    def __init__(self, ymin, xmin, ymax, xmax):
        self.ymin, self.xmin, self.ymax, self.xmax = ymin, xmin, ymax, xmax

    @classmethod
    def from_center_view(cls, x_center, y_center, width, height):
        hw, hh = (width - 1) / 2.0, (height - 1) / 2.0
        return cls(round(y_center - hh), round(x_center - hw), round(y_center + hh),
                   round(x_center + hw))

    @classmethod
    def from_json_dict(cls, d):
        x = d["x_center"] if d.get("x_center") is not None else d["x"]
        y = d["y_center"] if d.get("y_center") is not None else d["y"]
        return cls.from_center_view(x, y, d["width"], d["height"])

    def to_caffe_view(self):
        return [self.ymin, self.xmin, self.ymax, self.xmax]

    def to_dict(self):
        return {"y_min": self.ymin, "x_min": self.xmin, "y_max": self.ymax, "x_max": self.xmax}

    def to_json_dict(self):
        return {"y": (self.ymin + self.ymax) / 2.0, "x": (self.xmin + self.xmax) / 2.0,
                "width": self.xmax - self.xmin + 1, "height": self.ymax - self.ymin + 1,
                "label": None, "angle": 0.0}

    def area(self):
        return (self.xmax - self.xmin + 1) * (self.ymax - self.ymin + 1)

    def draw_on_pic(self, img, color=(255, 255, 255), line_width=1):
        """A copy of ``img`` [H, W, C] with the box outline painted (no OpenCV needed)."""
        out = numpy.array(img, copy=True)
        h, w = out.shape[:2]
        y0, y1 = int(max(0, self.ymin)), int(min(h - 1, self.ymax))
        x0, x1 = int(max(0, self.xmin)), int(min(w - 1, self.xmax))
        lw = max(1, int(line_width))
        out[y0:y0 + lw, x0:x1 + 1] = color
        out[max(y0, y1 - lw + 1):y1 + 1, x0:x1 + 1] = color
        out[y0:y1 + 1, x0:x0 + lw] = color
        out[y0:y1 + 1, max(x0, x1 - lw + 1):x1 + 1] = color
        return out

    def __repr__(self):
        return "BBox(ymin=%g, xmin=%g, ymax=%g, xmax=%g)" % tuple(self.to_caffe_view())


def _arr(b):
    return numpy.asarray(b, dtype=numpy.float64)


def areas(boxes):
    b = _arr(boxes)
    return (b[..., 3] - b[..., 1] + 1) * (b[..., 2] - b[..., 0] + 1)


def overlap_area(a, b):
    """Intersection area of ``a`` and ``b`` (broadcasts: [4] x [N, 4] -> [N], [N, 1, 4] x [M, 4]
    -> [N, M]); 0 where the boxes are disjoint (forward_bbox.py:115-134)."""
    a, b = _arr(a), _arr(b)
    dx = numpy.minimum(a[..., 3], b[..., 3]) - numpy.maximum(a[..., 1], b[..., 1]) + 1
    dy = numpy.minimum(a[..., 2], b[..., 2]) - numpy.maximum(a[..., 0], b[..., 0]) + 1
    return numpy.where((dx > 0) & (dy > 0), dx * dy, 0.0)


def overlap_ratio(a, b):
    """Intersection over union (forward_bbox.py:137-160)."""
    inter = overlap_area(a, b)
    union = areas(a) + areas(b) - inter
    with numpy.errstate(divide="ignore", invalid="ignore"):
        return numpy.where(union == 0, 0.0, inter / union)


def has_inclusion(a, b, area_ratio=0.9):
    """(the smaller box lies - to ``area_ratio`` of its area - inside the other, a is the bigger
    one) (forward_bbox.py:163-180)."""
    aa, ab = areas(a), areas(b)
    return overlap_area(a, b) >= numpy.minimum(aa, ab) * area_ratio, aa > ab


def is_small(boxes, h_min, w_min, min_area=None):
    b = _arr(boxes)
    w, h = b[..., 3] - b[..., 1] + 1, b[..., 2] - b[..., 0] + 1
    res = (w < w_min) | (h < h_min)
    if min_area is not None:
        res = res | (w * h < min_area)
    return res


def nms_detections(bboxes, probs, overlap_thr=0.7):
    """Greedy non-maximum suppression (forward_bbox.py:205-264): take the best-scoring box, drop
    every remaining box of which more than ``overlap_thr`` is covered by it, repeat.
    ``bboxes`` rows are [x1, y1, x2, y2]; returns the kept rows with the score as 5th column, best
    first. The cover matrix is built once ([N, N], vectorised)."""
    b = _arr(bboxes).reshape(-1, 4)
    s = _arr(probs).reshape(-1)
    dets = numpy.concatenate([b, s[:, None]], axis=1)
    n = len(b)
    if n == 0:
        return dets
    w = numpy.maximum(0.0, numpy.minimum(b[:, None, 2], b[None, :, 2]) -
                      numpy.maximum(b[:, None, 0], b[None, :, 0]) + 1)
    h = numpy.maximum(0.0, numpy.minimum(b[:, None, 3], b[None, :, 3]) -
                      numpy.maximum(b[:, None, 1], b[None, :, 1]) + 1)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    with numpy.errstate(divide="ignore", invalid="ignore"):
        cover = (w * h) / area[None, :]           # cover[i, j]: share of box j under box i
    alive = numpy.ones(n, dtype=bool)
    pick = []
    for i in numpy.argsort(s, kind="stable")[::-1]:
        if not alive[i]:
            continue
        pick.append(i)
        alive &= ~(cover[i] > overlap_thr)
        alive[i] = False
    return dets[pick]


def load_synsets(path):
    """``label word, word`` lines -> (ids, descriptions) (forward_bbox.py:267-286)."""
    ids, words = [], []
    with open(path) as fin:
        for line in fin:
            line = line.strip()
            if line:
                k, _, v = line.partition(" ")
                ids.append(k)
                words.append(v)
    return ids, words


def merge_to_one(bboxes, probs, img_size, padding_ratio=0.05):
    """Probability-weighted mean of the boxes (plain mean when all scores are 0), padded by
    ``padding_ratio`` of its size and clipped to the picture (forward_bbox.py:289-333)."""
    b, p = _arr(bboxes).reshape(-1, 4), _arr(probs).reshape(-1)
    if p.min() < 0:
        raise ValueError("negative probability")
    box = b.mean(axis=0) if p.max() == 0 else (b * p[:, None]).sum(axis=0) / p.sum()
    ymin, xmin, ymax, xmax = box
    width, height = xmax - xmin + 1, ymax - ymin + 1
    pic_h, pic_w = img_size
    out = numpy.array((max(0, ymin - height * padding_ratio), max(0, xmin - width * padding_ratio),
                       min(pic_h - 1, ymax + height * padding_ratio),
                       min(pic_w - 1, xmax + width * padding_ratio)))
    return out, float(p.max())


def merge_by_probs(bboxes, probs, img_size, primary_thr=0, secondary_thr=0.02, overlap_thr=0.3,
                   max_bboxes=None, use_inclusions=False):
    """Clusters around the best remaining box (forward_bbox.py:355-430): boxes scoring below
    ``secondary_thr`` or smaller than 20 x 20 never take part; the best remaining box (if it reaches
    ``primary_thr``) absorbs every remaining box with IoU >= ``overlap_thr`` (or, optionally,
    included in / including it) and the group is merged with ``merge_to_one``."""
    b, p = _arr(bboxes).reshape(-1, 4), _arr(probs).reshape(-1)
    order = numpy.argsort(p, kind="stable")
    order = order[(p[order] >= secondary_thr) & ~is_small(b[order], 20, 20)]
    out_b, out_p = [], []
    while len(order):
        if max_bboxes is not None and len(out_b) >= max_bboxes:
            break
        top = order[-1]
        if p[top] < primary_thr:
            break
        near = overlap_ratio(b[top], b[order]) >= overlap_thr
        if use_inclusions:
            near |= has_inclusion(b[top], b[order])[0]
        near[-1] = True
        group = order[near][::-1]              # best first, as the reference enumerates them
        mb, mp = merge_to_one(b[group], p[group], img_size)
        out_b.append(mb)
        out_p.append(mp)
        order = order[~near]
    return numpy.array(out_b).reshape(-1, 4), numpy.array(out_p)


def remove_inner(bboxes_with_probs):
    """Drop the smaller of two same-label boxes when 90 % of it lies inside the other
    (forward_bbox.py:496-512). Entries are (label, prob, [ymin, xmin, ymax, xmax])."""
    items = list(bboxes_with_probs)
    if len(items) <= 1:
        return items
    nested = set()
    for i, (li, _, bi) in enumerate(items):
        if i in nested:
            continue
        for j, (lj, _, bj) in enumerate(items):
            if j in nested or i == j or li != lj:
                continue
            incl, a_bigger = has_inclusion(bi, bj, 0.9)
            if incl:
                nested.add(j if a_bigger else i)
                if not a_bigger:
                    break
    return [it for k, it in enumerate(items) if k not in nested]


def merge_by_dict(bbox_dict, pic_size, primary_thr=0.1, secondary_thr=0.001, max_bboxes=None,
                  use_inclusions=True):
    """{(x_center, y_center, w, h): class-probability vector} -> [(label, prob, [ymin, xmin, ymax,
    xmax])] sorted by probability: per class ``merge_by_probs``, then the best ``max_bboxes`` over
    all classes without nested duplicates (forward_bbox.py:433-493)."""
    if not bbox_dict:
        raise ValueError("no boxes to merge")
    keys = list(bbox_dict)
    boxes = numpy.array([BBox.from_center_view(*k).to_caffe_view() for k in keys], dtype=numpy.float64)
    for (xc, yc, w, h) in keys:
        if not (xc + w / 2 < pic_size[1] + 1 and xc - w / 2 > -1 and
                yc + h / 2 < pic_size[0] + 1 and yc - h / 2 > -1):
            raise ValueError("box %s sticks out of the %s picture" % ((xc, yc, w, h), pic_size))
    probs = numpy.array([bbox_dict[k] for k in keys], dtype=numpy.float64)
    found = []
    for label in range(probs.shape[1]):
        bl, pl = merge_by_probs(boxes, probs[:, label], pic_size, max_bboxes=max_bboxes,
                                use_inclusions=use_inclusions, primary_thr=primary_thr,
                                secondary_thr=secondary_thr)
        found.extend((label, float(pp), [float(v) for v in bb]) for bb, pp in zip(bl, pl))
    found.sort(key=lambda t: t[1], reverse=True)
    if max_bboxes is not None:
        found = found[:max_bboxes]
    return remove_inner(found)


def _cv_to_corners(box):
    return BBox.from_center_view(*box).to_caffe_view()


def _absorb(small, big):
    """Grow ``big`` (center view) so that it contains the centre of ``small``; center view out."""
    ymin, xmin, ymax, xmax = _cv_to_corners(big)
    cy, cx = small[1], small[0]
    ymin, xmin, ymax, xmax = min(ymin, cy), min(xmin, cx), max(ymax, cy), max(xmax, cx)
    return ((xmin + xmax) / 2, (ymin + ymax) / 2, xmax - xmin, ymax - ymin)


def postprocess_same_label(bboxes_with_probs):
    """Final-stage clean-up on (label, prob, (x_center, y_center, w, h)) entries
    (forward_bbox.py:515-613): until nothing changes (1) a box 30 % inside another of the same
    label is absorbed by it, (2) a touching box of the same label whose area is below 75 % of the
    other's is absorbed as well. The survivor keeps the larger probability."""
    items = [tuple(t) for t in bboxes_with_probs]
    if len(items) <= 1:
        return items

    def one_pass(ratio, need_small):
        for i, (li, pi, bi) in enumerate(items):
            for j, (lj, pj, bj) in enumerate(items):
                if i == j or li != lj:
                    continue
                incl, i_bigger = has_inclusion(_cv_to_corners(bi), _cv_to_corners(bj), ratio)
                if not incl:
                    continue
                (s_idx, small), (b_idx, big) = ((j, bj), (i, bi)) if i_bigger else ((i, bi), (j, bj))
                if need_small and (small[2] * small[3]) / max(big[2] * big[3], 1e-12) > 0.75:
                    continue
                items[b_idx] = (li, max(pi, pj), _absorb(small, big))
                del items[s_idx]
                return True
        return False

    changed = True
    while changed:
        changed = False
        while one_pass(0.3, False):
            changed = True
        while one_pass(0.0, True):
            changed = True
    return items
