"""Time every kernel family of the path in isolation (CUDA events, warm, 10 reps) and print a markdown table with the
algorithmic work per launch: pairs/s for the searches, GB/s of ALGORITHMIC bytes for the memory-bound ones."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from reart_b200 import _lib, ops
from reart_b200.chamfer import ChamferDistance, knn_points, _ChamferBidir
from reart_b200.flow_utils import FlowReference, blend_anchor_motion_batched
from reart_b200.synth import make_sequence, make_flow_reference, kinematic_init
from reart_b200.kinematic import FlatTree

dev = torch.device("cuda")
T, N, P = 64, 16384, 15
seq = make_sequence(T, N, P, seed=2)
cano = torch.from_numpy(seq["cano"]).to(dev); frames = torch.from_numpy(seq["frames"]).to(dev)
W = torch.eye(P, device=dev)[torch.from_numpy(seq["part"]).to(dev)]
R = torch.from_numpy(np.ascontiguousarray(seq["pose"][:, :, :3, :3])).to(dev); tr = torch.from_numpy(np.ascontiguousarray(seq["pose"][:, :, :3, 3])).to(dev)
packed = ops.pack_cloud(frames)
skinned = ops.skin(cano, W, R, tr)
rows = []

def timeit(name, fn, work, unit, note=""):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    rate = work / (ms * 1e-3)
    rows.append(f"| {name} | {ms*1e3:9.1f} | {rate/1e9:10.1f} {unit} | {note} |")
    print(rows[-1], flush=True)

pairs = 2.0 * T * N * N
timeit("bidirectional search + index recovery (`reart_chamfer_bidir_fwd`: 2 packs, search, finalize)", lambda: _ChamferBidir.apply(skinned, frames), pairs, "Gpairs/s", "FP32 issue bound; 8 FLOP/pair => x8 for GFLOP/s")
timeit("one-direction search + index recovery (`reart_knn1_fwd`)", lambda: knn_points(skinned, frames, K=1), pairs / 2, "Gpairs/s", "cap 57 % of FP32 peak (6+1 slots per pair)")
d_f, d_b, i_f, i_b = _ChamferBidir.apply(skinned, frames)
L = _lib.lib()
gs = torch.empty_like(skinned); gt = torch.empty_like(frames); ones = torch.ones(T, N, device=dev)
timeit("Chamfer backward, both directions (`reart_chamfer_bidir_bwd`)", lambda: _lib.check(L.reart_chamfer_bidir_bwd(_lib.ptr(skinned), _lib.ptr(frames), _lib.ptr(i_f), _lib.ptr(i_b), _lib.ptr(ones), _lib.ptr(ones), T, N, N, _lib.ptr(gs), _lib.ptr(gt), _lib.stream_ptr()), "b"),
       T * N * 2 * (12 + 12 + 8 + 4 + 24), "GB/s", "gather + REDG scatter; 60 B per point per direction")
timeit("skinning forward (`reart_skin_fwd`)", lambda: ops.skin(cano, W, R, tr), 12.0 * N + 4.0 * N * P + 48.0 * T * P + 12.0 * T * N, "GB/s", "reads 12N+4NP+48TP, writes 12TN")
g = torch.randn(T, N, 3, device=dev)
Wg, Rg, tg = W.clone().requires_grad_(True), R.clone().requires_grad_(True), tr.clone().requires_grad_(True)
nb_ws = L.reart_skin_bwd_workspace_bytes(T, N, P); bws = torch.empty(nb_ws, dtype=torch.uint8, device=dev)
gW_, gR_, gtr_ = torch.empty_like(W), torch.empty_like(R), torch.empty_like(tr)
def skin_bwd():
    _lib.check(L.reart_skin_bwd(_lib.ptr(cano), _lib.ptr(W), _lib.ptr(R), _lib.ptr(tr), _lib.ptr(g), T, N, P, _lib.ptr(gW_), _lib.ptr(gR_), _lib.ptr(gtr_), _lib.ptr(bws), nb_ws, _lib.stream_ptr()), "sb")
timeit("skinning backward (`reart_skin_bwd`: fused pass + fixed-order reduce)", skin_bwd, 12.0 * T * N + 12.0 * N + 4.0 * N * P * 2 + 48.0 * T * P * 2, "GB/s", "g read once; deterministic")
def fused():
    Wt, Rt, tt = W.clone().requires_grad_(True), R.clone().requires_grad_(True), tr.clone().requires_grad_(True)
    loss, _ = ops.skinned_chamfer_loss(cano, Wt, Rt, tt, frames, packed, unit_grad=True)
timeit("whole fused energy fwd+bwd (`reart_skinned_chamfer_fwd_bwd`)", fused, pairs, "Gpairs/s", "skin_fwd_sorted + search + energy rows/cols + skin_bwd")
d6 = torch.randn(T * P, 6, device=dev)
timeit("6D -> R forward (`reart_rot6d_fwd`, T*P = 960 elements)", lambda: ops.rot6d(d6), T * P * 60.0, "GB/s", "launch-latency bound")
kw = kinematic_init(seq)
tree = FlatTree(kw["paths_to_base"], kw["reverse_topo"], kw["edge_index"], device=dev)
ax, mo, th = (torch.from_numpy(kw[k]).to(dev) for k in ("axis_list", "moment_list", "theta_list"))
timeit("tree FK forward (`reart_fk_fwd`, T=64, P=15)", lambda: ops.fk_flat(ax, mo, th, None, tree.order, tree.parent, tree.edge, None), T * P * 64.0, "GB/s", "one launch instead of ~5k ATen ops; latency bound")
axg, mog, thg = ax.clone().requires_grad_(True), mo.clone().requires_grad_(True), th.clone().requires_grad_(True)
def fk_fb():
    out = ops.fk_flat(axg, mog, thg, None, tree.order, tree.parent, tree.edge, None)
    out.backward(torch.ones_like(out))
timeit("tree FK forward + backward through autograd", fk_fb, T * P * 128.0, "GB/s", "2 launches + memsets")
refs, flows = make_flow_reference(seq, cano_idx=0, n_ref=N // 4)
ref = FlowReference([torch.from_numpy(r).to(dev) for r in refs], [torch.from_numpy(f).to(dev) for f in flows])
q = torch.cat((cano[None], skinned), 0)[:-1].contiguous()
nref = float(sum(r.shape[0] for r in refs))
timeit("flow blend k=3 for all 64 pairs (`reart_knn3_blend`)", lambda: blend_anchor_motion_batched(q, ref), N * nref, "Gpairs/s", "one launch; scalar 6+~2 slots per pair")
xyz = frames[:, :, :].contiguous()
timeit("FPS 16384 -> 1024 for 64 clouds (`reart_fps`)", lambda: ops.fps(xyz, 1024), 64 * 1024 * 16384.0, "Gpairs/s", "sequential in samples: 1024 rounds of block arg-max")
w0 = torch.randn(128, 3, device=dev); b0 = torch.randn(128, device=dev); w2 = torch.randn(P, 128, device=dev)
timeit("seg MLP forward (`reart_segmlp_fwd`)", lambda: ops.seg_mlp(cano, w0, b0, w2), N * (128 * 4 + 128 * P) * 2.0, "GFLOP/s", "latency bound (16k points)")
print("\n".join(rows))
