"""Inference with a trained workflow on new image files ("forward propagation" sample).

Parity: /root/reference/samples/MNIST/mnist_forward.py:58-112 — restore a trained workflow
from a snapshot, switch the launcher to *testing* mode, replace the loader by an image-file
loader over new pictures (re-using the training loader's normaliser and label mapping), run
one pass and write the per-sample class probabilities with ``write_results``.

    python -m veles.znicz_b200.models.mnist_forward <snapshot> <images dir> [result.json]
"""
from __future__ import annotations

import os
import sys

from ..core.workflow import DummyLauncher
from ..loader.base import UserLoaderRegistry


def create_forward(workflow, normalizer, labels_mapping, loader_config,
                   loader_name="full_batch_auto_label_file_image"):
    """Swap ``workflow.loader`` for a test-only loader and relink its consumers."""
    old = workflow.loader
    new_loader = UserLoaderRegistry.get_factory(loader_name, **loader_config)(workflow)
    new_loader.link_from(workflow.repeater)
    for dst in list(old.links_to):
        dst.unlink_from(old)
        dst.link_from(new_loader)
    old.unlink_all()
    workflow.del_ref(old)
    workflow.loader = new_loader
    workflow.forwards[0].link_attrs(new_loader, ("input", "minibatch_data"))
    workflow.evaluator.link_attrs(
        new_loader, ("batch_size", "minibatch_size"), ("labels", "minibatch_labels"),
        ("max_samples_per_epoch", "total_samples"), "class_lengths",
        ("offset", "minibatch_offset"))
    if hasattr(new_loader, "class_keys"):
        workflow.evaluator.link_attrs(new_loader, "class_keys")
    workflow.decision.link_attrs(
        new_loader, "minibatch_class", "last_minibatch", "minibatch_size", "class_lengths",
        "epoch_ended", "epoch_number")
    for f in workflow.forwards:                     # dropout etc. follow the class
        if f.has_linked_attr("minibatch_class"):
            f.link_attrs(new_loader, "minibatch_class")
    workflow.repeater.gate_block = workflow.decision.complete
    new_loader.gate_block = workflow.decision.complete
    new_loader._normalizer = normalizer                      # same statistics as in training
    workflow.evaluator.labels_mapping = list(
        sorted(labels_mapping, key=labels_mapping.get)) if isinstance(labels_mapping, dict) \
        else labels_mapping
    return new_loader


def forward_from_snapshot(snapshot, test_paths, loader_config=None, result_file=None,
                          device="auto"):
    """→ (workflow, {"Output": ...}) after one testing pass over ``test_paths``."""
    from ..core.snapshotter import SnapshotterToFile
    wf = SnapshotterToFile.import_file(snapshot)
    launcher = DummyLauncher(testing=True)
    wf.workflow = launcher
    old = wf.loader
    cfg = {"minibatch_size": 10, "color_space": getattr(old, "color_space", "GRAY"),
           "normalization_type": old.normalization_type, "test_paths": list(test_paths)}
    shape = tuple(old.minibatch_data.shape[1:3]) if old.minibatch_data else None
    if shape:
        cfg["scale"] = (shape[1], shape[0])
    cfg.update(loader_config or {})
    create_forward(wf, old.normalizer, old.labels_mapping, cfg)
    wf.decision.max_epochs = 1
    wf.decision.complete <<= False
    wf.initialize(device=device, snapshot=False)
    wf.run()
    results = wf.gather_results()
    if result_file:
        wf.write_results(result_file)
    return wf, results


if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    _wf, _res = forward_from_snapshot(
        sys.argv[1], [sys.argv[2]],
        result_file=sys.argv[3] if len(sys.argv) > 3 else os.path.join(sys.argv[2], "result.json"))
    print({k: (v if not hasattr(v, "shape") else "array%s" % (v.shape,)) for k, v in _res.items()})
