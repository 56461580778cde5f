"""SpamKohonen: a SOM over sparse bag-of-lemmas vectors of e-mails, validated against the
spam / ham labels. Parity: /root/reference/tests/research/SpamKohonen/spam_kohonen.py,
spam_kohonen_config.py:43-63 (8x8 map, batch 80, ``lemma:weight`` text format optionally
xz-compressed, ids and/or classes per line, pointwise normalisation, KohonenValidator +
results exporter)."""
from __future__ import annotations

import lzma
import os

import numpy

from ..core.config import root
from ..core.units import Unit
from ..loader.base import TEST, VALID, TRAIN
from ..loader.fullbatch import FullBatchLoader
from ..ops import kohonen
from ..ops.nn_units import NNWorkflow
from ..utils import nn_plotting_units


def _gd(t):
    return 0.002 / (1.0 + t * 0.00002)


def _rd(t):
    return 1.0 / (1.0 + t * 0.00002)


root.spam_kohonen.update({
    "forward": {"shape": (8, 8), "weights_stddev": 0.05, "weights_filling": "uniform"},
    "decision": {"epochs": 200, "snapshot_prefix": "spam_kohonen"},
    "loader": {"minibatch_size": 80, "force_numpy": False, "ids": True, "classes": False,
               "file": os.path.join(str(root.common.dirs.datasets), "spam", "spam.txt.xz"),
               "validation_ratio": 0.0},
    "train": {"gradient_decay": _gd, "radius_decay": _rd},
    "exporter": {"file": "classified_fast4.txt"}})


def generate_dataset(path, n=300, n_lemmas=40, seed=3):
    """``id class lemma:weight ... \\n`` lines: two topics with different lemma usage."""
    rs = numpy.random.RandomState(seed)
    topics = rs.dirichlet(numpy.ones(n_lemmas) * 0.2, 2)
    lines = []
    for k in range(n):
        cls = int(rs.randint(0, 2))
        counts = rs.multinomial(30, topics[cls])
        fields = ["msg%05d" % k, str(cls)]
        fields += ["%d:%.4f" % (l, c / 30.0) for l, c in enumerate(counts) if c]
        lines.append(" ".join(fields) + " \n")
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    opener = lzma.open if path.endswith(".xz") else open
    with opener(path, "wb") as f:
        f.write("".join(lines).encode())
    return path


class SpamKohonenLoader(FullBatchLoader):
    MAPPING = "spam_kohonen_loader"

    def __init__(self, workflow, **kwargs):
        kwargs["normalization_type"] = "pointwise"
        super().__init__(workflow, **kwargs)
        self.file = kwargs.get("file")
        self.has_ids = kwargs.get("ids", False)
        self.has_classes = kwargs.get("classes", True)
        self.lemmas_map = {}
        self.kohonen_labels_mapping = []
        self.samples_by_label = {}
        self.ids = []

    def load_data(self):
        opener = lzma.open if self.file.endswith(".xz") else open
        with opener(self.file, "rb") as fin:
            lines = fin.readlines()
        rows, labels, lemmas = [], [], set()
        self.ids = []
        for line in lines:
            fields = line.split(b" ")
            off = 0
            if self.has_ids:
                self.ids.append(fields[off].decode("charmap"))
                off += 1
            if self.has_classes:
                labels.append(int(fields[off]))
                off += 1
            row = []
            for field in fields[off:]:
                if b":" not in field:
                    continue
                lemma, weight = field.split(b":")
                row.append((int(lemma), float(weight)))
                lemmas.add(int(lemma))
            rows.append(row)
        self.lemmas_map = {l: i for i, l in enumerate(sorted(lemmas))}
        data = numpy.zeros((len(rows), len(self.lemmas_map)), dtype=self.dtype)
        for r, row in enumerate(rows):
            for lemma, weight in row:
                data[r, self.lemmas_map[lemma]] = weight
        self.original_data.reset(data)
        distinct = sorted(set(labels)) if self.has_classes else [0]
        self.kohonen_labels_mapping = distinct
        rev = {l: i for i, l in enumerate(distinct)}
        self.samples_by_label = {i: set() for i in range(len(distinct))}
        for i, l in enumerate(labels if self.has_classes else [0] * len(rows)):
            self.samples_by_label[rev[l]].add(i)
        self.labels_mapping = dict(rev)
        self.reversed_labels_mapping = list(distinct)
        self.class_lengths[TEST] = self.class_lengths[VALID] = 0
        self.class_lengths[TRAIN] = len(rows)


class ResultsExporter(Unit):
    """Writes ``id  winner-neuron`` lines once training is complete."""

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.file_name = kwargs["file_name"]
        self.demand("total", "ids", "shuffled_indices")

    def initialize(self, **kwargs):
        pass

    def run(self):
        self.total.map_read()
        idx = self.shuffled_indices.mem if hasattr(self.shuffled_indices, "mem") \
            else self.shuffled_indices
        with open(self.file_name, "w") as f:
            for pos, winner in enumerate(self.total.mem[:len(idx)]):
                sid = self.ids[idx[pos]] if self.ids else str(idx[pos])
                f.write("%s %d\n" % (sid, int(winner)))


class SpamKohonenWorkflow(NNWorkflow):
    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("name", "SpamKohonen")
        super().__init__(workflow, **kwargs)
        cfg = root.spam_kohonen
        self.repeater.link_from(self.start_point)
        self.loader = SpamKohonenLoader(
            self, minibatch_size=kwargs.get("minibatch_size", cfg.loader.minibatch_size),
            file=kwargs.get("file", cfg.loader.file), ids=kwargs.get("ids", cfg.loader.ids),
            classes=kwargs.get("classes", cfg.loader.classes),
            force_numpy=cfg.loader.force_numpy)
        self.loader.link_from(self.repeater)
        self.trainer = kohonen.KohonenTrainer(
            self, shape=cfg.forward.shape, weights_filling=cfg.forward.weights_filling,
            weights_stddev=cfg.forward.weights_stddev,
            gradient_decay=cfg.train.gradient_decay, radius_decay=cfg.train.radius_decay)
        self.trainer.link_from(self.loader)
        self.trainer.link_attrs(self.loader, ("input", "minibatch_data"))
        self.forward = kohonen.KohonenForward(self, total=True)
        self.forward.link_from(self.trainer)
        self.forward.link_attrs(self.loader, ("input", "minibatch_data"), "minibatch_offset",
                                "minibatch_size", ("batch_size", "total_samples"))
        self.forward.link_attrs(self.trainer, "weights", "argmins")
        self.validator = kohonen.KohonenValidator(self)
        self.validator.link_attrs(self.trainer, "shape")
        self.validator.link_attrs(self.forward, ("input", "output"))
        self.validator.link_attrs(self.loader, "minibatch_indices", "minibatch_size",
                                  "samples_by_label", "labels_mapping",
                                  "reversed_labels_mapping")
        self.validator.link_from(self.forward)
        self.decision = kohonen.KohonenDecision(
            self, max_epochs=kwargs.get("epochs", cfg.decision.epochs))
        self.decision.link_from(self.validator)
        self.decision.link_attrs(self.loader, "minibatch_class", "last_minibatch",
                                 "class_lengths", "epoch_ended", "epoch_number")
        self.decision.link_attrs(self.trainer, "weights", "winners")
        self.repeater.link_from(self.decision)
        self.repeater.gate_block = self.decision.complete
        self.loader.gate_block = self.decision.complete
        self.exporter = ResultsExporter(
            self, file_name=kwargs.get("export_file", os.path.join(
                str(root.common.dirs.cache), cfg.exporter.file)))
        self.exporter.link_from(self.decision)
        self.exporter.link_attrs(self.forward, "total")
        self.exporter.link_attrs(self.loader, "ids", "shuffled_indices")
        self.exporter.gate_block = ~self.decision.complete
        self.end_point.link_from(self.exporter)
        self.end_point.gate_block = ~self.decision.complete
        self.plotters = [nn_plotting_units.KohonenHits(self),
                         nn_plotting_units.KohonenNeighborMap(self)]
        for p, src in zip(self.plotters, ("winners_mem", "weights_mem")):
            p.link_attrs(self.trainer, "shape").link_from(self.decision)
            p.link_attrs(self.decision, ("input", src))
            p.gate_block = ~self.decision.epoch_ended
        if self.loader.has_classes:
            vp = nn_plotting_units.KohonenValidationResults(self)
            vp.link_attrs(self.trainer, "shape").link_from(self.decision)
            vp.link_attrs(self.decision, ("input", "winners_mem"))
            vp.link_attrs(self.validator, "result", "fitness", "fitness_by_label",
                          "fitness_by_neuron")
            vp.gate_block = ~self.decision.epoch_ended
            self.plotters.append(vp)


def build(launcher=None, **kwargs):
    from ..core.workflow import DummyLauncher
    return SpamKohonenWorkflow(launcher or DummyLauncher(), **kwargs)


def run(load, main):
    load(SpamKohonenWorkflow)
    main()
