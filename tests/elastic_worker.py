"""Worker for tests/test_elastic_restart.py (torch.distributed.run --max-restarts 1, gloo): on the
first attempt rank 1 dies after the second epoch; the restarted group resumes from the newest
rank-0 snapshot."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from veles.znicz_b200.core import prng  # noqa: E402
from veles.znicz_b200.core.config import root  # noqa: E402
from veles.znicz_b200.launcher import Launcher  # noqa: E402
from veles.znicz_b200.models import mnist  # noqa: E402


def main():
    out_dir = sys.argv[1]
    rank = int(os.environ.get("RANK", "0"))
    attempt = int(os.environ.get("TORCHELASTIC_RESTART_COUNT", "0"))
    root.common.dirs.snapshots = os.path.join(out_dir, "snapshots")
    root.common.disable.snapshotting = False
    prng.get(1).seed(100)
    prng.get(2).seed(5678)
    launcher = Launcher(backend="numpy", snapshot="latest:elastic")

    wf, restored = launcher.load(
        mnist.MnistWorkflow, layers=mnist.fc_layers(), loader_name="synthetic_mnist",
        loader_config={"minibatch_size": 10, "n_train": 80, "n_valid": 40, "noise": 0.3,
                       "normalization_type": "linear"},
        decision_config={"max_epochs": 5, "fail_iterations": 50},
        snapshotter_config={"prefix": "elastic", "interval": 1, "time_interval": 0,
                            "compression": ""},
        loss_function="softmax", lr_adjuster_config=root.mnistr.lr_adjuster)
    start_epoch = int(wf.loader.epoch_number)
    if attempt == 0 and rank == 1:
        # die in the middle of training: after epoch 2 has been snapshotted by rank 0
        def bomb(w):
            if int(w.loader.epoch_number) >= 2:
                os._exit(17)
        wf.step_hooks_.append(bomb)
    launcher.main()
    res = {"rank": rank, "attempt": attempt, "restored": bool(restored), "start_epoch": start_epoch,
           "end_epoch": int(wf.loader.epoch_number), "complete": bool(wf.decision.complete),
           "best": wf.decision.best_n_err_pt[1]}
    with open(os.path.join(out_dir, "elastic_rank%d.json" % rank), "w") as f:
        json.dump(res, f)


if __name__ == "__main__":
    main()
