"""cProfile of the host side of the streaming (end-to-end) training loop: where the Python time of
one step goes. Usage (GPU box): PYTHONPATH=. python tools/profile_host.py [steps]"""
import cProfile
import pstats
import sys
import types

import torch

import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
args = types.SimpleNamespace(dtype="bf16", no_graphs=False, n_train=50000, model="cifar_caffe")
wf = bench.build_workflow(True, args.dtype, True, args.n_train, args.model)
wf.initialize(device="cuda")
from veles.znicz_b200.utils.step_reader import StepResultReader
reader = StepResultReader(wf.evaluator)
wf.step_hooks_.append(reader)
wf.run(iterations=bench.GRAPH_CAPTURE_STEPS + 50)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
wf.run(iterations=steps)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
