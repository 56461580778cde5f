"""``python -m veles.znicz_b200.models.imagenet_forward [config.py] [root.path=value ...]``

Runs the two-stage localisation pipeline configured under ``root.imagenet_forward`` (trained
workflow snapshot, candidate-box stream, result path, thresholds); under ``torchrun`` every rank
processes its share of the candidate stream and writes ``<result>.rank<k>.json`` (combine them with
``writer.merge_json``). Reference entry point:
/root/reference/tests/research/ImagenetAE/imagenet_forward/imagenet_forward.py:365-388."""
import json
import sys

from ...launcher import apply_config_file, _set_by_path, _parse_value
from .workflow import run_from_config


def main(argv):
    args = list(argv)
    device = "auto"
    if "--backend" in args:
        i = args.index("--backend")
        device = args[i + 1]
        del args[i:i + 2]
    for a in args:
        if "=" in a and not a.endswith(".py"):
            k, _, v = a.partition("=")
            _set_by_path(k.strip(), _parse_value(v.strip()))
        else:
            apply_config_file(a)
    wf, results = run_from_config(device=device)
    print(json.dumps({"pictures": len(results),
                      "detections": sum(len(v["bbxs"]) for v in results.values()),
                      "result_path": wf.result_path}))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
