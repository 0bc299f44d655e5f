"""Host logic of reart_b200/retarget.py (the kernels it drives are covered by the -m gpu parity tests).

A small pure-torch articulated chain stands in for KinematicModel so that the batching claim can be checked on CPU:
fitting S novel states stacked on the frame axis follows the same Adam/AMSGrad trajectories as S separate fits
(utils/kinematic_utils.py:201-266 runs them one by one)."""
import numpy as np
import torch

from reart_b200 import retarget as rt


class ChainModel(torch.nn.Module):
    """E revolute joints in a chain; part p is moved by joints 0..p-1; one label per input point (nearest anchor)."""

    def __init__(self, E=3, seed=0):
        super().__init__()
        gen = torch.Generator().manual_seed(seed)
        self.axis_list = torch.nn.Parameter(torch.nn.functional.normalize(torch.randn(E, 3, generator=gen), dim=-1))
        self.points = 0.3 * torch.randn(E, 3, generator=gen)
        self.anchors = torch.randn(E + 1, 3, generator=gen)

    def _joint(self, e, theta):                                           # (S,) -> (S,4,4) rotation about a fixed line
        a = self.axis_list[e]
        K = torch.zeros(3, 3).index_put((torch.tensor([0, 0, 1, 1, 2, 2]), torch.tensor([1, 2, 0, 2, 0, 1])),
                                        torch.stack([-a[2], a[1], a[2], -a[0], -a[1], a[0]]))
        R = torch.eye(3) + torch.sin(theta)[:, None, None] * K + (1 - torch.cos(theta))[:, None, None] * (K @ K)
        t = self.points[e] - R @ self.points[e]
        M = torch.eye(4).repeat(theta.shape[0], 1, 1)
        M[:, :3, :3], M[:, :3, 3] = R, t
        return M

    def forward(self, input_pc, theta_list=None):
        S, E = theta_list.shape
        trans = [torch.eye(4).repeat(S, 1, 1)]
        for e in range(E):
            trans.append(trans[-1] @ self._joint(e, theta_list[:, e]))
        trans = torch.stack(trans, dim=1)                                # (S,P,4,4)
        seg = torch.cdist(input_pc, self.anchors).argmin(dim=1)
        per_point = trans[:, seg]                                        # (S,n,4,4)
        out = torch.einsum("snij,nj->sni", per_point[..., :3, :3], input_pc) + per_point[..., :3, 3]
        return out, seg, trans


def test_batched_retarget_equals_one_fit_per_state():
    torch.manual_seed(0)
    model = ChainModel()
    sparse = model.anchors + 0.05 * torch.randn(4, 3)
    truth = torch.tensor([[0.4, -0.3, 0.2], [-0.5, 0.1, 0.6], [0.05, 0.7, -0.4]])
    with torch.no_grad():
        target = model(sparse, theta_list=truth)[0]
    batched = rt.retarget(model, sparse, target, n_iter=120)["theta_list"]
    for s in range(3):
        single = rt.retarget(model, sparse, target[s], n_iter=120)["theta_list"]
        torch.testing.assert_close(batched[s:s + 1], single, rtol=1e-5, atol=1e-6)
    assert not batched.requires_grad
    dense = torch.randn(200, 3)
    with torch.no_grad():
        novel = model(dense, theta_list=truth)[0]
    err, posed, seg = rt.retarget_error(model, dense, novel, {"theta_list": batched})
    assert err.shape == (3,) and posed.shape == (3, 200, 3) and seg.shape == (200,)
    assert float(err.max()) < 1.0                                        # 100 x metres: under a centimetre
    start, _, _ = rt.retarget_error(model, dense, novel, {"theta_list": torch.full((3, 3), 1e-6)})
    assert float(err.mean()) < 0.05 * float(start.mean())


def test_unknowns_follow_the_reference_start_values():
    class Relax:
        num_parts = 4
        proposal_6d = proposal_t = None

    u = rt.init_unknowns(Relax(), 2, "cpu")
    assert u["proposal_6d"].shape == (2, 4, 6) and u["proposal_t"].shape == (2, 4, 3)
    assert u["proposal_6d"][1, 3].tolist() == [1, 0, 0, 0, 1, 0] and float(u["proposal_t"].detach().abs().sum()) == 0
    k = rt.init_unknowns(ChainModel(E=5), 3, "cpu")
    assert k["theta_list"].shape == (3, 5) and torch.allclose(k["theta_list"], torch.tensor(1e-6))
    assert all(v.requires_grad for v in list(u.values()) + list(k.values()))


def test_ik_entry_point_with_a_dataset_like_object():
    torch.manual_seed(1)
    model = ChainModel()
    dense = np.random.default_rng(0).normal(size=(120, 3)).astype(np.float32)
    sparse = (model.anchors + 0.05 * torch.randn(4, 3)).numpy()
    thetas = [torch.tensor([[0.3, 0.2, -0.4]]), torch.tensor([[-0.2, 0.5, 0.1]])]

    class Data:
        cano_idx = 0
        pose_list = ["cano"]
        novel_pose_list = [0, 1]

        def __getitem__(self, i):
            return {"cano_pc": dense, "gt_cano_part": np.zeros(120, np.int64)}

    def sampler(cano_pc, gt_part, cano_pose, novel_pose, sparse_sample_per_part=1):
        assert cano_pose == "cano" and sparse_sample_per_part == 1
        with torch.no_grad():
            return {"sparse_cano_pc": sparse.astype(np.float64),
                    "sparse_novel_pc": model(torch.from_numpy(sparse), theta_list=thetas[novel_pose])[0][0].numpy(),
                    "novel_pc": model(torch.from_numpy(cano_pc), theta_list=thetas[novel_pose])[0][0].numpy()}

    err = rt.ik(Data(), model, "cpu", verbose=False, vis=False, sampler=sampler, n_iter=150)
    assert isinstance(err, float) and err < 1.0
