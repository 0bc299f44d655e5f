import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from reart_b200 import _lib, ops
from conftest import synthetic_sequence
import oracle
dev = torch.device("cuda")
L = _lib.lib()
T, N, P = 4, 1500, 6
seq = synthetic_sequence(T, N, P, seed=3)
rng = np.random.default_rng(0)
cano, frames = seq["cano"], seq["frames"]
W = np.eye(P, dtype=np.float32)[seq["part"]]
W[::5] = rng.random((len(W[::5]), P)).astype(np.float32)
R = np.ascontiguousarray(seq["pose"][:, :, :3, :3]).astype(np.float32)
tr = np.ascontiguousarray(seq["pose"][:, :, :3, 3]).astype(np.float32)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d_cano, d_W, d_R, d_tr, d_frames = cu(cano), cu(W), cu(R), cu(tr), cu(frames)
# 1) skin bwd alone
g = rng.standard_normal((T, N, 3)).astype(np.float32)
gW_ref, gR_ref, gt_ref = oracle.skin_bwd(cano, W, R, tr, g)
d_g = cu(g)
gW, gR, gt = torch.empty(N, P, device=dev), torch.empty(T, P, 3, 3, device=dev), torch.empty(T, P, 3, device=dev)
nb = L.reart_skin_bwd_workspace_bytes(T, N, P); ws = torch.empty(nb, dtype=torch.uint8, device=dev)
_lib.check(L.reart_skin_bwd(_lib.ptr(d_cano), _lib.ptr(d_W), _lib.ptr(d_R), _lib.ptr(d_tr), _lib.ptr(d_g), T, N, P,
                            _lib.ptr(gW), _lib.ptr(gR), _lib.ptr(gt), _lib.ptr(ws), nb, _lib.stream_ptr()), "sb")
torch.cuda.synchronize()
for name, a, b in (("gW", gW, gW_ref), ("gR", gR, gR_ref), ("gt", gt, gt_ref)):
    a = a.cpu().numpy()
    print("skin_bwd", name, "max err", np.abs(a - b).max(), "max ref", np.abs(b).max(), flush=True)
# 2) energy g_src
sk_ref = oracle.skin_fwd(cano, W, R, tr)
ch = oracle.chamfer_bidir_fwd_bwd(sk_ref, frames)
skinned = torch.empty(T, N, 3, device=dev); loss = torch.empty(1, dtype=torch.float64, device=dev)
gs = torch.empty(T, N, 3, device=dev)
gW2, flat = torch.empty(N, P, device=dev), torch.empty(T * P * 12, device=dev)
nbytes = L.reart_energy_workspace_bytes(T, N, N); ws2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
packed = ops.pack_cloud(d_frames)
dF = torch.empty(T, N, device=dev); iF = torch.empty(T, N, dtype=torch.int64, device=dev)
dB = torch.empty(T, N, device=dev); iB = torch.empty(T, N, dtype=torch.int64, device=dev)
_lib.check(L.reart_skinned_chamfer_fwd_bwd_ex(_lib.ptr(d_cano), _lib.ptr(d_W), _lib.ptr(d_R), _lib.ptr(d_tr), _lib.ptr(d_frames),
           _lib.ptr(packed), T, N, N, P, _lib.ptr(skinned), _lib.ptr(loss), _lib.ptr(gW2), _lib.ptr(flat[:T*P*9]), _lib.ptr(flat[T*P*9:]),
           _lib.ptr(gs), 1, _lib.ptr(dF), _lib.ptr(iF), _lib.ptr(dB), _lib.ptr(iB), _lib.ptr(ws2), nbytes, _lib.stream_ptr()), "fused")
torch.cuda.synchronize()
print("skinned max err", np.abs(skinned.cpu().numpy() - sk_ref).max())
print("loss", loss.item(), ch["loss"])
d = np.abs(gs.cpu().numpy() - ch["grad_src"])
print("g_src max err", d.max(), "max ref", np.abs(ch["grad_src"]).max(), "bad rows", int((d.max(-1) > 1e-4).sum()), "of", T * N)
bad = np.argwhere(d.max(-1) > 1e-4)[:5]
for t, n in bad:
    print(" row", t, n, gs[t, n].cpu().numpy(), ch["grad_src"][t, n], "indeg", int((ch["i_bwd"][t] == n).sum()))

print("i_fwd mismatches", int((iF.cpu().numpy() != ch["i_fwd"]).sum()), "i_bwd mismatches", int((iB.cpu().numpy() != ch["i_bwd"]).sum()))
print("d_fwd equal", np.array_equal(dF.cpu().numpy(), ch["d_fwd"]), "d_bwd equal", np.array_equal(dB.cpu().numpy(), ch["d_bwd"]))
ib = iB.cpu().numpy(); mm = np.argwhere(ib != ch["i_bwd"])[:8]
for t, j in mm: print("  col", t, j, "ours", ib[t, j], "oracle", ch["i_bwd"][t, j])
# ---- inspect the workspace: psrc | ka | kb | (gs skipped: caller buffer) | perm | xq
al = lambda v: (v + 255) // 256 * 256
n_pad = (N + 255) // 256 * 256
off = 0
psrc = ws2[off:off + T * n_pad * 12].view(torch.float32).cpu().numpy().reshape(T, n_pad // 4, 3, 4); off += al(T * n_pad * 12)
off += al(T * N * 8) * 2
perm = ws2[off:off + T * n_pad].cpu().numpy().reshape(T, n_pad); off += al(T * n_pad)
xq = ws2[off:off + T * (n_pad // 256) * 64].view(torch.float32).cpu().numpy().reshape(T, n_pad // 256, 16)
sk = skinned.cpu().numpy()
xs = psrc[:, :, 0, :].reshape(T, n_pad); ys = psrc[:, :, 1, :].reshape(T, n_pad)
for t in range(1):
    for c in range(n_pad // 256):
        x = xs[t, 256 * c:256 * (c + 1)]
        print("chunk", c, "sorted", bool((np.diff(x) >= 0).all()), "perm is permutation", len(set(perm[t, 256*c:256*(c+1)].tolist())) == 256,
              "xq ok", np.array_equal(xq[t, c], x[15::16]))
        idx = 256 * c + perm[t, 256 * c:256 * (c + 1)].astype(np.int64)
        ok = idx < N
        print("   values match AoS", np.array_equal(x[ok], sk[t, idx[ok], 0]), np.array_equal(ys[t, 256*c:256*(c+1)][ok], sk[t, idx[ok], 1]), "first x", x[:4], "perm", perm[t, 256*c:256*c+4])
print("perm[0,:64]", perm[0, :64].tolist())
print("expected argsort", np.argsort(sk[0, :256, 0], kind="stable")[:64].tolist())
