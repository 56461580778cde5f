"""Data-parallel equivalence worker: K fp32 training steps of a small conv net at GLOBAL batch
80 (per-rank batch 80 / world), no shuffling, so that the union of the ranks' k-th minibatches
is exactly the k-th minibatch of the single-process run. Rank 0 saves the final weights."""
import json
import os
import sys

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from veles.znicz_b200.core import prng  # noqa: E402
from veles.znicz_b200.core.config import root  # noqa: E402
from veles.znicz_b200.models import cifar  # noqa: E402

GLOBAL_BATCH = 80


def main():
    out_dir, tag = sys.argv[1], sys.argv[2]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    root.common.disable.snapshotting = True
    root.common.engine.compute_type = "fp32"
    prng.get(1).seed(4242)
    prng.get(2).seed(5678)
    layers = cifar.caffe_layers()
    for l in layers:
        if "<-" in l:
            l["<-"].update(learning_rate=0.05, learning_rate_bias=0.05)
        if l["type"] == "conv":
            l["->"]["weights_stddev"] = 0.05
    wf = cifar.build(
        layers=layers, use_graphs=False,
        loader_config={"minibatch_size": GLOBAL_BATCH // world, "n_train": 320, "n_valid": 0,
                       "n_test": 0, "normalization_type": "internal_mean", "noise": 0.3,
                       "on_device": True, "shuffle_limit": 0},
        decision_config={"max_epochs": 1000, "fail_iterations": 1000})
    wf.initialize(device="cuda")
    wf.run(iterations=6)
    import torch
    torch.cuda.synchronize()
    ws = []
    for f in wf.forwards:
        if getattr(f, "weights", None):
            f.weights.map_read()
            f.bias.map_read()
            ws.append(numpy.array(f.weights.mem, dtype=numpy.float32).ravel())
            ws.append(numpy.array(f.bias.mem, dtype=numpy.float32).ravel())
    flat = numpy.concatenate(ws)
    fs = wf.fused_step_
    res = {"rank": rank, "world": world, "tag": tag,
           "algo": getattr(fs, "algo_name", None) if fs is not None else None,
           "mc": bool(getattr(fs, "mc_red", 0)) if fs is not None else False,
           "sha": __import__("hashlib").sha1(flat.tobytes()).hexdigest(),
           "finite": bool(numpy.isfinite(flat).all()), "absmax": float(numpy.abs(flat).max())}
    with open(os.path.join(out_dir, "%s_rank%d.json" % (tag, rank)), "w") as f:
        json.dump(res, f)
    if rank == 0:
        numpy.save(os.path.join(out_dir, "%s_weights.npy" % tag), flat)


if __name__ == "__main__":
    main()
