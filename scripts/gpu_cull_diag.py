"""Why does the culling fade late in a fit?  Per 256-row chunk: max vs high percentiles of the seed bounds ub_i."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from reart_b200.engine import RelaxationEngine, tau_schedule
from reart_b200.synth import make_sequence

dev = torch.device("cuda")
T, N, P = 16, 16384, 15
seq = make_sequence(T, N, P, seed=2)
eng = RelaxationEngine(torch.from_numpy(seq["cano"]).to(dev), torch.from_numpy(seq["frames"]).to(dev), num_parts=P, use_graph=False, cull=True)
for steps, label in ((20, "early (20 steps, tau~5)"), (600, "late (600 steps, tau=1)")):
    while eng.iteration < steps:
        eng.step(tau_schedule(eng.iteration, 600, 5.0, 1.0))
    b = eng._nat
    nn_rows, nn_cols = b["nn_rows"].long().clone(), b["nn_cols"].long().clone()
    W_prev = b["W"].argmax(1).clone()
    eng.step(tau_schedule(eng.iteration, 600, 5.0, 1.0))                 # the next step's cloud, seeded by nn_* above
    sk, fr = eng.skinned, eng.frames
    flipped = (b["W"].argmax(1) != W_prev).float().mean().item()
    ub = (sk - torch.gather(fr, 1, nn_rows[:, :, None].expand(-1, -1, 3))).square().sum(-1)          # [T,N]
    cand = [ub]
    for d in (-2, -1, 1, 2):
        j = torch.roll(nn_rows, d, dims=1)
        cand.append((sk - torch.gather(fr, 1, j[:, :, None].expand(-1, -1, 3))).square().sum(-1))
    ub5 = torch.stack(cand).min(0).values
    true_d = torch.cdist(sk[:2], fr[:2]).min(-1).values.square()
    for name, u in (("own seed", ub), ("5 seeds", ub5)):
        ch = u.reshape(T, N // 256, 256).sqrt()
        srt = ch.sort(dim=-1).values
        print(f"{label} | {name} | flipped {flipped:.3f} | sqrt(ub): median of rows {srt[..., 128].median():.4f} | chunk max (median over chunks) {srt[..., 255].median():.4f} | "
              f"249th {srt[..., 248].median():.4f} | 240th {srt[..., 239].median():.4f} | 224th {srt[..., 223].median():.4f} | true NN dist median {true_d.sqrt().median():.4f}, p99 {true_d.sqrt().quantile(0.99):.4f}", flush=True)
    ubc = (fr - torch.gather(sk, 1, nn_cols[:, :, None].expand(-1, -1, 3))).square().sum(-1)
    chc = ubc.reshape(T, N // 32, 32).sqrt().sort(dim=-1).values
    print(f"{label} | columns own seed | chunk max median {chchc if False else chc[..., 31].median():.4f} | 30th {chc[..., 29].median():.4f} | median {chc[..., 16].median():.4f}", flush=True)
