"""One culled engine step, kernel by kernel (run under ncu --metrics gpu__time_duration.sum)."""
import sys
sys.path.insert(0, ".")
import torch
from reart_b200.engine import RelaxationEngine, tau_schedule
from reart_b200.synth import make_sequence
T, N, P = 64, 16384, 15
seq = make_sequence(T, N, P, seed=2)
dev = torch.device("cuda")
eng = RelaxationEngine(torch.from_numpy(seq["cano"]).to(dev), torch.from_numpy(seq["frames"]).to(dev), num_parts=P, use_graph=False, cull=True)
for i in range(12):
    eng.step(tau_schedule(i, 15000, 5.0, 1.0))
torch.cuda.synchronize()
