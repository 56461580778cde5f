"""Worker for tests/test_data_parallel.py (launched by torch.distributed.run, gloo)."""
import json
import os
import sys

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from veles.znicz_b200.core import prng  # noqa: E402
from veles.znicz_b200.core.config import root  # noqa: E402
from veles.znicz_b200.models import mnist  # noqa: E402


def main():
    out_dir = sys.argv[1]
    rank = int(os.environ.get("RANK", "0"))
    root.common.disable.snapshotting = True
    prng.get(1).seed(100 + rank)          # different init per rank: broadcast must fix it
    prng.get(2).seed(5678)
    wf = mnist.build(
        layers=mnist.fc_layers(), loader_name="synthetic_mnist",
        loader_config={"minibatch_size": 10, "n_train": 80, "n_valid": 40, "noise": 0.3,
                       "normalization_type": "linear"},
        decision_config={"max_epochs": 3, "fail_iterations": 10})
    wf.initialize(device="numpy")
    wf.run()
    ws = [f.weights.mem for f in wf.forwards if getattr(f, "weights", None)]
    res = {"rank": rank, "world": wf.dp_.world_size if wf.dp_ is not None else 1,
           "train_len": int(wf.loader.class_lengths[2]),
           "valid_len": int(wf.loader.class_lengths[1]),
           "checksum": [float(numpy.float64(w.astype(numpy.float64).sum())) for w in ws],
           "absmax": [float(numpy.abs(w).max()) for w in ws],
           "best_valid_err_pt": wf.decision.best_n_err_pt[1],
           "epoch_n_err": [int(x) if x is not None else None for x in wf.decision.epoch_n_err]}
    with open(os.path.join(out_dir, "rank%d.json" % rank), "w") as f:
        json.dump(res, f)


if __name__ == "__main__":
    main()
