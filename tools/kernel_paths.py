"""Which kernel family every unit's launches used (SIMT fallbacks stand out):
PYTHONPATH=. python tools/kernel_paths.py [model]"""
import sys
import types

import torch

import bench
from veles.znicz_b200.kernels import api

model = sys.argv[1] if len(sys.argv) > 1 else "alexnet"
wf = bench.build_workflow(False, "bf16", False, 2048, model)
wf.initialize(device="cuda")
wf.run(iterations=3)
torch.cuda.synchronize()
for name, tags in sorted(api.paths.items()):
    print(name, sorted(tags))
print("launches", api.counters)
