"""Failure detection + restart (SURVEY §5 "failure detection / elastic"): a rank dies mid-training,
the survivor's epoch-end health check names the failure instead of hanging, torchrun restarts the
group and the job resumes from the newest rank-0 snapshot (``--snapshot latest`` semantics)."""
import json
import os
import socket
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_rank_failure_restart_resumes_from_latest_snapshot(tmp_path):
    env = dict(os.environ, PYTHONPATH=REPO, OMP_NUM_THREADS="1", ZNICZ_DP_TIMEOUT_S="20")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--max-restarts=1", "--rdzv-backend=c10d", "--rdzv-endpoint=127.0.0.1:%d" % _free_port(),
           "--local-addr", "127.0.0.1",
           os.path.join(REPO, "tests", "elastic_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res = [json.load(open(tmp_path / ("elastic_rank%d.json" % i))) for i in range(2)]
    for x in res:
        assert x["attempt"] == 1                       # written by the restarted group only
        assert x["restored"] and x["start_epoch"] >= 1     # resumed, not started over
        assert x["complete"] and x["end_epoch"] >= 5
    assert res[0]["best"] == res[1]["best"]
    snaps = os.listdir(tmp_path / "snapshots")
    assert any(n.startswith("elastic_") and ".pickle" in n for n in snaps)
    # the survivor did not hang: the failure surfaced as an error of the first attempt
    assert "RankFailure" in r.stderr or "exitcode" in r.stderr or "failed" in r.stderr.lower()


def test_latest_snapshot_resolution(tmp_path):
    from veles.znicz_b200.core.config import root
    from veles.znicz_b200.launcher import Launcher
    old = root.common.dirs.snapshots
    root.common.dirs.snapshots = str(tmp_path)
    try:
        assert Launcher.resolve_snapshot("latest") is None          # nothing yet: fresh start
        assert Launcher.resolve_snapshot("/x/y.pickle") == "/x/y.pickle"
        for i, name in enumerate(("a_1.4.pickle", "b_7.4.pickle.gz", "a_2.4.pickle")):
            p = tmp_path / name
            p.write_bytes(b"x")
            os.utime(p, (1000 + i, 1000 + i))
        assert Launcher.resolve_snapshot("latest").endswith("a_2.4.pickle")
        assert Launcher.resolve_snapshot("latest:b").endswith("b_7.4.pickle.gz")
        os.symlink("a_1.4.pickle", tmp_path / "a_current.lnk")
        assert Launcher.resolve_snapshot("latest:a").endswith("a_current.lnk")
        assert Launcher.resolve_snapshot("latest:zzz") is None
    finally:
        root.common.dirs.snapshots = old
