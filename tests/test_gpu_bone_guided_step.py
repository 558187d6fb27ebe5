"""The bone-guided training step (train.BoneGuidedStep; train_funcs.py:128-392: three passes, six loss terms, one backward)
against the oracle's restatement of the same lines on the small fixture: every loss term, the total, every parameter
gradient; then the CUDA-graph replay against the eager step."""
import numpy as np
import pytest
import torch

from oracle import spiral_oracle as so
from tests.golden.constants import KPS_INDEX_LIST, PART_LIST
from tests.helpers import TOL_BF16, TOL_F32, golden, params_from_golden, ref_args, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SKL_LIST = [[15, 12], [15, 12], [12, 9], [6, 0], [0, 1, 2], [1, 4], [4, 7], [7, 10], [2, 5], [5, 8], [8, 11], [16, 18],
            [18, 20], [20, 22], [17, 19], [19, 21], [21, 23]]  # configure/cfgs.py:18-20 (one bone per part, full keypoint ids)
KPS_KEEP = [i for i in range(35) if i not in (3, 13, 14)]      # train_funcs.py: kps_keep
FE = [[3, 8, 8, 16, 16], [[], [], [], [], []]]
FD = [[16, 16, 8, 8, 8], [[], [], [], [], 3]]


def _setup(shb):
    g = golden("golden_multiz_small")
    h, sizes, ssz, spirals, D, U = ref_args("small")
    vdict = {n: g["part_" + n] for n in PART_LIST}
    model = shb.SpiralAutoencoder_multiz_partkps(KPS_INDEX_LIST, vdict, FE, FD, latent_size=8, part_kps_latent_size=8,
                                                 sizes=sizes, spiral_sizes=ssz, spirals=[s.to(DEV) for s in spirals],
                                                 D=[d.to(DEV) for d in D], U=[u.to(DEV) for u in U], device=DEV).to(DEV)
    model.load_state_dict(params_from_golden(g), strict=True)
    gen = torch.Generator().manual_seed(11)
    V = sizes[0]
    # level-0 vertex lists of the 17 parts (contiguous slabs of the template) and a keypoint regressor with positive rows
    order = np.argsort(np.asarray(h.verts0)[:, 1], kind="stable")
    parts = [np.sort(c) for c in np.array_split(order, len(PART_LIST))]
    J = torch.rand(35, V, generator=gen) ** 8
    J = J / J.sum(1, keepdim=True)
    from tests.golden.synthetic import synthetic_meshes

    batches = [synthetic_meshes(h.verts0, 4, seed=20 + i, noise=0.05) for i in range(3)]
    measure = torch.rand(4, 16, generator=gen) * 0.8 + 0.2
    P, Q = [int(v) for v in g["P"]], [int(v) for v in g["Q"]]
    return g, (sizes, ssz, spirals, D, U), model, parts, J, batches, measure, P, Q


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bone_guided_step_matches_oracle(dtype):
    import semantichuman_b200 as shb
    from semantichuman_b200.train import BoneGuidedStep

    g, (sizes, ssz, spirals, D, U), model, parts, J, batches, measure, P, Q = _setup(shb)
    model.set_compute_dtype(dtype)
    step = BoneGuidedStep(model, J, KPS_KEEP, parts, SKL_LIST, P, Q, optimizer=False)
    loss = step._eager(*[b.to(DEV) for b in batches], measure.to(DEV), factor=0.93)
    params = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.named_parameters()}
    part_idx_coarse = [g["part_" + n] for n in PART_LIST]
    ref, terms = so.bone_guided_step_loss(params, *batches, measure, J, KPS_KEEP, KPS_INDEX_LIST, part_idx_coarse,
                                          SKL_LIST, P, Q, 0.93, FE, FD, sizes, spirals, D, U, step.weights,
                                          part_index_lists_fine=parts)
    ref.backward()
    tol = TOL_F32 if dtype == torch.float32 else TOL_BF16
    for k, v in terms.items():
        assert abs(step.terms[k].item() - v.item()) <= tol * max(1.0, abs(v.item())), (k, step.terms[k].item(), v.item())
    assert abs(loss.item() - ref.item()) <= tol * max(1.0, abs(ref.item()))
    for n, p in model.named_parameters():
        assert relerr(p.grad, params[n].grad) < (3 * tol if dtype == torch.float32 else tol), n


def test_bone_guided_step_graph_replay_matches_eager():
    import semantichuman_b200 as shb
    from semantichuman_b200.train import BoneGuidedStep

    outs = []
    for graph in (False, True):
        g, _, model, parts, J, batches, measure, P, Q = _setup(shb)
        model.set_compute_dtype(torch.bfloat16)
        step = BoneGuidedStep(model, J, KPS_KEEP, parts, SKL_LIST, P, Q, factor=(0.0, 0.9), graph=graph)
        dev = [b.to(DEV) for b in batches] + [measure.to(DEV)]
        if graph:
            step.capture(*dev)
            assert step.launches_per_step > 60
        else:
            for _ in range(3):
                step(*dev)
        outs.append([step(*dev).item() for _ in range(4)])
    assert all(abs(a - b) <= 1e-5 * abs(b) + 1e-7 for a, b in zip(outs[1], outs[0])), outs
