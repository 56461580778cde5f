"""GPU data-parallel worker (torchrun, NCCL bootstrap only): trains the CIFAR caffe net for a
few minibatches with the fused peer-memory reduce+update and reports per-rank checksums."""
import json
import os
import sys

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from veles.znicz_b200.core import prng  # noqa: E402
from veles.znicz_b200.core.config import root  # noqa: E402
from veles.znicz_b200.models import cifar  # noqa: E402


def main():
    out_dir, compute, graphs = sys.argv[1], sys.argv[2], sys.argv[3] == "graphs"
    rank = int(os.environ.get("RANK", "0"))
    root.common.disable.snapshotting = True
    root.common.engine.compute_type = compute
    prng.get(1).seed(100 + rank)          # different init per rank: broadcast must fix it
    prng.get(2).seed(5678)
    layers = cifar.caffe_layers()
    for l in layers:
        if "<-" in l:
            l["<-"].update(learning_rate=0.02, learning_rate_bias=0.02)
        if l["type"] == "conv":
            l["->"]["weights_stddev"] = 0.05
    wf = cifar.build(
        layers=layers, use_graphs=graphs,
        loader_config={"minibatch_size": 20, "n_train": 400, "n_valid": 80,
                       "normalization_type": "internal_mean", "noise": 0.3,
                       "on_device": True},
        decision_config={"max_epochs": 3, "fail_iterations": 10})
    wf.initialize(device="cuda")
    wf.run()
    import torch
    torch.cuda.synchronize()
    ws = []
    for f in wf.forwards:
        if getattr(f, "weights", None):
            f.weights.map_read()
            f.bias.map_read()
            ws.append(f.weights.mem.astype(numpy.float64))
            ws.append(f.bias.mem.astype(numpy.float64))
    res = {"rank": rank, "world": wf.dp_.world_size if wf.dp_ is not None else 1,
           "fused_symm": bool(wf.dp_ is not None and wf.dp_.symm is not None),
           "step_launches": wf.fused_step_.launches if wf.fused_step_ else 0,
           "checksum": [float(w.sum()) for w in ws],
           "abssum": [float(numpy.abs(w).sum()) for w in ws],
           "finite": bool(all(numpy.isfinite(w).all() for w in ws)),
           "train_len": int(wf.loader.class_lengths[2]),
           "epoch_n_err": [int(x) if x is not None else None for x in wf.decision.epoch_n_err],
           "best_valid_err_pt": wf.decision.best_n_err_pt[1]}
    with open(os.path.join(out_dir, "gpu_rank%d.json" % rank), "w") as f:
        json.dump(res, f)


if __name__ == "__main__":
    main()
