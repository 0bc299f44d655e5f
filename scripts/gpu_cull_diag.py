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

# ---- how much could be culled at this (late) state?  fraction of (256-row chunk, 32-target chunk) blocks evaluated under
# different bounds, with the actual chunk boxes
def boxes(x, chunk):
    c = x.reshape(x.shape[0], -1, chunk, 3)
    return c.amin(2), c.amax(2)
def fraction(sk, fr, rb, cb):
    rlo, rhi = boxes(sk, 256); clo, chi = boxes(fr, 32)
    tot = ev = 0
    for t in range(sk.shape[0]):
        g = torch.clamp(torch.maximum(rlo[t][:, None] - chi[t][None], clo[t][None] - rhi[t][:, None]), min=0).square().sum(-1)   # [R,C]
        e = (g <= rb[t][:, None]) | (g <= cb[t][None])
        tot += e.numel(); ev += int(e.sum())
    return ev / tot
sk, fr = eng.skinned, eng.frames
d2 = torch.stack([torch.cdist(sk[t:t+1], fr[t:t+1])[0].square() for t in range(T)])          # [T,N,N]
true_r, true_c = d2.min(2).values, d2.min(1).values
rb_true = true_r.reshape(T, -1, 256).amax(-1); cb_true = true_c.reshape(T, -1, 32).amax(-1)
rb_seed = ub5.reshape(T, -1, 256).amax(-1); cb_seed = ubc.reshape(T, -1, 32).amax(-1)
print("late: fraction evaluated with the seed bounds (5 row seeds, 1 column seed): %.3f" % fraction(sk, fr, rb_seed, cb_seed))
print("late: fraction evaluated with EXACT bounds (true NN distances, the floor for these boxes): %.3f" % fraction(sk, fr, rb_true, cb_true))
q = 0.97
rb_q = true_r.reshape(T, -1, 256).sort(-1).values[..., int(256 * q) - 1]
print("late: ... exact bounds but ignoring the worst 3 %% of rows per chunk: %.3f" % fraction(sk, fr, rb_q, cb_true))
# rows regrouped per frame by posed position (k-d order of the POSED cloud): the floor with compact row boxes
from reart_b200 import ops
fr_f = []
for t in range(T):
    perm = ops.kd_order(sk[t:t+1], 8)[0]
    fr_f.append(perm)
perm = torch.stack(fr_f)
sk2 = torch.gather(sk, 1, perm[:, :, None].expand(-1, -1, 3)); tr2 = torch.gather(true_r, 1, perm)
print("late: exact bounds, rows regrouped per frame on the posed cloud: %.3f" % fraction(sk2, fr, tr2.reshape(T, -1, 256).amax(-1), cb_true))
