"""Functional test of the Wine sample (north-star config 1: CPU/numpy, world 1).
Mirrors /root/reference/tests/functional/test_wine.py:67-85 (train, assert
convergence, snapshot, resume)."""
import glob
import os

import numpy

from veles.znicz_b200.core.config import root
from veles.znicz_b200.core.workflow import DummyLauncher
from veles.znicz_b200.core.snapshotter import SnapshotterToFile
from veles.znicz_b200.models.wine import WineWorkflow


def test_wine_trains_and_resumes():
    root.wine.decision.max_epochs = 60
    root.wine.snapshotter.interval = 1
    wf = WineWorkflow(DummyLauncher(), layers=[8, 3])
    wf.initialize(device="numpy")
    wf.run()
    dec = wf.decision
    assert bool(dec.complete)
    assert dec.best_n_err_pt[2] is not None
    assert dec.best_n_err_pt[2] < 3.0, dec.best_n_err_pt   # reference: 0.56 %
    snaps = glob.glob(os.path.join(root.common.dirs.snapshots, "wine_*.pickle"))
    assert snaps, "no snapshot written"
    # resume from the snapshot and train on
    wf2 = SnapshotterToFile.import_file(sorted(snaps, key=os.path.getmtime)[-1])
    wf2.workflow = DummyLauncher()
    wf2.decision.max_epochs = wf2.loader.epoch_number + 3
    wf2.decision.complete <<= False
    e0 = wf2.loader.epoch_number
    wf2.initialize(device="numpy", snapshot=True)
    wf2.run()
    assert wf2.loader.epoch_number >= e0 + 1
    assert wf2.decision.best_n_err_pt[2] < 3.0
