"""Worker for tests/test_imagenet_forward.py::test_ranks_share_the_candidate_stream: every rank
runs the pipeline over its `shard_range` of the candidate stream."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from veles.znicz_b200.core.config import root  # noqa: E402
from veles.znicz_b200.models.imagenet_forward.workflow import run_from_config  # noqa: E402


def main():
    d = sys.argv[1]
    root.imagenet_forward.update({
        "trained_workflow": os.path.join(d, "trained_current.lnk"),
        "result_path": os.path.join(d, "result.json"),
        "loader": {"path_to_bboxes": os.path.join(d, "raw.pickle"), "minibatch_size": 4,
                   "raw_bboxes_min_size": 8, "raw_bboxes_min_area": 64, "raw_bboxes_min_area_ratio": 0,
                   "raw_bboxes_min_size_ratio": 0, "add_relative_bboxes": False,
                   "angle_step_final": 1.0, "min_angle_final": 0.0, "max_angle_final": 0.0},
        "mergebboxes": {"ignore_negative": True, "use_compatibility": False, "mode": "",
                        "labels_compatibility": "", "probability_threshold": 0.0,
                        "last_chance_probability_threshold": 0.0, "raw_path": ""}})
    wf, results = run_from_config(device="numpy")
    print(json.dumps({"rank": int(os.environ.get("RANK", "0")), "pictures": sorted(results),
                      "range": [wf.loader.min_index, wf.loader.max_index]}))


if __name__ == "__main__":
    main()
