"""CLI launcher (``python -m veles.znicz_b200``) and the Range/genetics search."""
import glob
import json
import os
import subprocess
import sys

import numpy

from veles.znicz_b200.core import genetics
from veles.znicz_b200.core.config import Config

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_range_markers_and_fix_config():
    cfg = Config("t")
    cfg.update({"a": {"lr": genetics.Range(0.1, 0.001, 1.0), "n": genetics.Range(8, 2, 64)},
                "layers": [{"type": "x", "->": {"k": genetics.Range("tanh", "relu", "sigmoid")}}],
                "plain": 5})
    markers = genetics.process_config(cfg)
    assert sorted(".".join(map(str, m[0])) for m in markers) == [
        "a.lr", "a.n", "layers.0.->.k"]
    genetics.apply_values(cfg, markers, [m[1].max_value for m in markers])
    by = {".".join(map(str, m[0])): m for m in markers}
    assert cfg.a.n == 64 and cfg.a.lr == 1.0
    assert cfg.layers[0]["->"]["k"] in ("relu", "sigmoid")
    cfg2 = Config("u")
    cfg2.update({"x": genetics.Range(3, 1, 5), "l": [genetics.Range(0.5, 0.0, 1.0)]})
    genetics.fix_config(cfg2)
    assert cfg2.x == 3 and cfg2.l == [0.5]
    assert by["a.n"][1].is_int and not by["a.lr"][1].is_int


def test_genetics_optimizer_finds_maximum():
    markers = [(("x",), genetics.Range(0.0, -5.0, 5.0)), (("n",), genetics.Range(1, 1, 20))]

    def fitness(values):
        x, n = markers[0][1].decode(values[0]), markers[1][1].decode(values[1])
        return -(x - 2.0) ** 2 - (n - 13) ** 2
    opt = genetics.GeneticsOptimizer(markers, fitness, population_size=24, generations=15,
                                     seed=3)
    best, fit = opt.run()
    assert abs(best[0] - 2.0) < 0.5 and abs(best[1] - 13) <= 1
    assert opt.history[-1][1] >= opt.history[0][1]
    opt2 = genetics.GeneticsOptimizer(markers, fitness, population_size=24, generations=15,
                                      seed=3)
    assert opt2.run() == (best, fit)          # deterministic for a seed


def _run_cli(args, cwd):
    env = dict(os.environ, PYTHONPATH=REPO)
    return subprocess.run([sys.executable, "-m", "veles.znicz_b200"] + args, cwd=cwd, env=env,
                          capture_output=True, text=True, timeout=300)


def test_cli_wine_train_snapshot_resume(tmp_path):
    cfg = tmp_path / "cfg.py"
    cfg.write_text(
        "from veles.config import root\n"
        "root.common.dirs.snapshots = %r\n"
        "root.wine.decision.max_epochs = 6\n"
        "root.wine.snapshotter.interval = 1\n" % str(tmp_path))
    res = tmp_path / "res.json"
    r = _run_cli(["-b", "numpy", "--seed", "5", "--result-file", str(res), "wine", str(cfg),
                  "root.wine.learning_rate=0.25"], str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    metrics = json.loads(r.stdout.strip().splitlines()[-1])
    assert metrics and os.path.exists(res)
    snaps = [f for f in glob.glob(str(tmp_path / "wine*.pickle*")) if not os.path.islink(f)]
    assert snaps
    r2 = _run_cli(["-b", "numpy", "-s", sorted(snaps)[-1], "wine", "-",
                   "root.wine.decision.max_epochs=8"], str(tmp_path))
    assert r2.returncode == 0, r2.stderr[-2000:]
    r3 = _run_cli(["-b", "numpy", "--dry-run", "init", "wine"], str(tmp_path))
    assert r3.returncode == 0, r3.stderr[-2000:]


def test_cli_optimize(tmp_path):
    cfg = tmp_path / "cfg.py"
    cfg.write_text(
        "root.common.dirs.snapshots = %r\n"
        "root.wine.decision.max_epochs = 3\n"
        "root.wine.learning_rate = Range(0.3, 0.01, 1.0)\n" % str(tmp_path))
    r = _run_cli(["-b", "numpy", "--seed", "2", "--optimize", "3:2", "-v", "warning", "wine",
                  str(cfg)], str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert "wine.learning_rate" in out["best"] and numpy.isfinite(out["best_fitness"])
