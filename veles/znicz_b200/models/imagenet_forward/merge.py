"""Per-image accumulation of box probabilities and the two decision modes
(/root/reference/tests/research/ImagenetAE/imagenet_forward/imagenet_forward.py:62-235)."""
from __future__ import annotations

import pickle

import numpy

from ...core.units import Unit
from .bbox import merge_by_dict, postprocess_same_label


class MergeBboxes(Unit):
    """Consumes minibatches of class probabilities for (box, angle, flip) shots. All shots of one
    box are folded into one probability vector (element-wise maximum over angles / mirrors, the
    minimum for the "nothing here" class 0 when it is not ignored); when the picture changes - or
    the loader ends - the picture's boxes are decided:

    * mode "merge": probability-weighted merging per class (``bbox.merge_by_dict``);
    * mode "final": every box votes for its arg-max class; boxes at or above
      ``last_chance_probability_threshold`` are candidates (all boxes if there are none), same-label
      candidates are cleaned up (``bbox.postprocess_same_label``), winners need
      ``probability_threshold`` (the best candidate wins otherwise); optionally winners whose
      label is incompatible with the most central winner are dropped.

    ``winners`` collects {"path": ..., "bbxs": [(label, probability, box)]}."""
    hide_from_registry = True

    def __init__(self, workflow, labels_compatibility=None, **kwargs):
        super().__init__(workflow, **kwargs)
        self.winners = []
        self.max_per_class = kwargs.get("max_per_class", 5)
        self.ignore_negative = kwargs.get("ignore_negative", True)
        self.save_raw = kwargs.get("save_raw_file_name", "")
        self.probability_threshold = kwargs.get("probability_threshold", 0.8)
        self.last_chance_probability_threshold = kwargs.get(
            "last_chance_probability_threshold", 0.7)
        self.labels_compatibility_file_name = labels_compatibility
        self.use_compatibility = bool(kwargs.get("use_compatibility", True) and
                                      labels_compatibility)
        self.mode = kwargs.get("mode", "")
        self.labels_mapping = kwargs.get("labels_mapping", {})
        self.demand("probabilities", "minibatch_bboxes", "minibatch_images", "minibatch_size",
                    "ended", "mode")

    def init_unpickled(self):
        super().init_unpickled()
        self.rawfd_ = None
        self.image_ = None
        self.image_size_ = None
        self.boxes_ = {}

    def initialize(self, **kwargs):
        if self.save_raw:
            self.rawfd_ = open(self.save_raw, "wb")
        if self.use_compatibility:
            with open(self.labels_compatibility_file_name, "rb") as fin:
                self.labels_compatibility, labels_array = pickle.load(fin)
            self.compat_index = {lbl: i for i, lbl in enumerate(labels_array)}
            self.compatibility_threshold = float(numpy.mean(self.labels_compatibility))

    def reset(self):
        self.image_ = None
        self.boxes_ = {}
        del self.winners[:]

    # -- accumulation --------------------------------------------------------------------
    def run(self):
        probs = self.probabilities
        if hasattr(probs, "map_read"):
            probs.map_read()
            probs = probs.mem
        probs = numpy.asarray(probs, dtype=numpy.float64).reshape(len(probs), -1)
        for i in range(int(self.minibatch_size)):
            key, size = self.minibatch_images[i]
            if self.image_ is None:
                self.image_, self.image_size_ = key, size
            elif key != self.image_:
                self._decide()
                self.image_, self.image_size_ = key, size
            self.add_bbox(self.minibatch_bboxes[i][0], probs[i])
        if self.ended and self.boxes_:
            self._decide()
            if self.rawfd_ is not None:
                self.rawfd_.close()
                self.rawfd_ = None

    def add_bbox(self, bbox, probs):
        key = tuple(bbox[k] for k in ("x", "y", "width", "height"))
        cur = self.boxes_.get(key)
        if cur is None:
            self.boxes_[key] = numpy.array(probs, dtype=numpy.float64)
            return
        first = 0 if self.ignore_negative else 1
        cur[first:] = numpy.maximum(cur[first:], probs[first:])
        if not self.ignore_negative:
            cur[0] = min(cur[0], probs[0])

    # -- decisions ---------------------------------------------------------------------------
    def _compatibility(self, a, b):
        if a[0] == b[0]:
            return 1.0
        shift = 1 if self.ignore_negative else 0
        try:
            n1, n2 = (1 + self.compat_index[self.labels_mapping[x[0] + shift]] for x in (a, b))
        except KeyError:
            return 0.0
        return self.labels_compatibility[n1, n2]

    def _remove_incompatible(self, boxes):
        cy, cx = self.image_size_[0] / 2.0, self.image_size_[1] / 2.0
        ordered = sorted(boxes, key=lambda b: numpy.hypot(b[2][0] - cx, b[2][1] - cy))
        best = ordered[0]
        return [b for b in ordered
                if self._compatibility(best, b) > self.compatibility_threshold or
                b[2][2] * b[2][3] >= best[2][2] * best[2][3]]

    def _decide(self):
        boxes = self.boxes_
        if self.rawfd_ is not None and self.mode == "merge":
            pickle.dump({self.image_: boxes}, self.rawfd_, protocol=pickle.HIGHEST_PROTOCOL)
            self.rawfd_.flush()
        if self.ignore_negative:
            boxes = {k: v[1:] for k, v in boxes.items()}
        if self.mode == "merge":
            winners = merge_by_dict(boxes, pic_size=self.image_size_)
            for w in winners:
                if w[2][2] <= w[2][0] or w[2][3] <= w[2][1]:
                    self.error("%s: degenerate merged box %s", self.image_, w)
            if not self.ignore_negative:
                positive = [w for w in winners[:self.max_per_class] if w[0] > 0]
                if not positive:
                    positive = [w for w in winners[self.max_per_class:]
                                if w[0] > 0][:self.max_per_class]
                winners = positive
        elif self.mode == "final":
            votes = []
            for box, probs in sorted(boxes.items()):
                top = int(numpy.argmax(probs))
                if not self.ignore_negative and top == 0:
                    continue
                votes.append((top, float(probs[top]), box))
            candidates = [v for v in votes if v[1] >= self.last_chance_probability_threshold]
            candidates = postprocess_same_label(candidates or votes)
            winners = [c for c in candidates if c[1] >= self.probability_threshold]
            if not winners and candidates:
                winners = [max(candidates, key=lambda c: c[1])]           # last chance
            if self.use_compatibility and len(winners) > 1:
                winners = self._remove_incompatible(winners)
        else:
            raise ValueError("MergeBboxes.mode must be 'merge' or 'final', not %r" % (self.mode,))
        self.winners.append({"path": self.image_, "bbxs": winners})
        self.boxes_ = {}
