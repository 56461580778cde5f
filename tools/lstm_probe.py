"""Per-step clock stamps of the persistent LSTM forward kernel (CTA 0): where a time step goes."""
import torch
from veles.znicz_b200.kernels import load_extension

ext = load_extension(required=True)
T, B, I, H = 32, 128, 128, 256
dev = torch.device("cuda")
xh = torch.randn(T + 1, B, I + H, device=dev).bfloat16()
w = (torch.randn(4 * H, I + H, device=dev) * 0.05).bfloat16()
bias = torch.zeros(4 * H, device=dev)
state = torch.empty(int(ext.lstm_state_floats(T, B, I, H)), device=dev)
dbg = torch.zeros(T, 8, device=dev, dtype=torch.int64)
for _ in range(3):
    r = ext.lstm_fwd_persist(xh, w, bias, state, H, dbg)
torch.cuda.synchronize()
assert r == 0, r
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ext.lstm_fwd_persist(xh, w, bias, state, H, None)
e1.record()
torch.cuda.synchronize()
print("forward kernel: %.1f us per launch (T=%d)" % (e0.elapsed_time(e1) * 1000 / 20, T))
# backward kernel timing
err = torch.randn(B, T, H, device=dev).bfloat16()
dz = torch.empty(T, B, 4 * H, device=dev, dtype=torch.bfloat16)
whp = torch.empty(H, 4 * H, device=dev, dtype=torch.bfloat16)
part = torch.empty(int(ext.lstm_part_floats(B, H)), device=dev)
for _ in range(3):
    r = ext.lstm_bwd_persist(err, True, state, part, dz, w, whp, I)
assert r == 0, r
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    ext.lstm_bwd_persist(err, True, state, part, dz, w, whp, I)
e1.record()
torch.cuda.synchronize()
print("backward kernel (+ W_h permutation): %.1f us per launch" % (e0.elapsed_time(e1) * 1000 / 20))
d = dbg.cpu().numpy()
names = {2: "mma: h_ready seen", 3: "epi: acc_full seen", 4: "epi: math done", 5: "epi: h stored, fenced, arrived",
         6: "epi: state stores issued"}
import numpy
per = numpy.diff(d[2:T - 1, 3])
print("cycles per step (acc_full to acc_full): mean %.0f => %.2f us at 1.9 GHz" % (per.mean(), per.mean() / 1900.0))
for k in (4, 5, 6):
    print(names[k], "+%.0f cycles after acc_full" % float((d[2:T - 1, k] - d[2:T - 1, 3]).mean()))
print(names[2], "+%.0f cycles after the previous step's acc_full" % float((d[3:T - 1, 2] - d[2:T - 2, 3]).mean()))
print("acc_full after h_ready: +%.0f cycles" % float((d[3:T - 1, 3] - d[3:T - 1, 2]).mean()))
