"""Own Adam (semantichuman_b200/optim.py -> shb_adam_step) against torch.optim.Adam, and the bf16 weight shadows."""
import pytest
import torch

from tests.helpers import relerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _params(seed, shapes):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g).to(DEV).requires_grad_(True) for s in shapes]


@pytest.mark.parametrize("wd", [0.0, 5e-5])
def test_adam_trajectory_matches_torch(wd):
    from semantichuman_b200.optim import Adam

    shapes = [(7,), (33, 5), (256, 1031), (3,), (64, 64)] + [(5, 3)] * 40  # > 32 tensors: two kernel launches
    mine, ref = _params(1, shapes), _params(1, shapes)
    shadows = {mine[2]: [torch.empty_like(mine[2], dtype=torch.bfloat16), -1]}
    opt = Adam(mine, lr=1e-3, weight_decay=wd, shadows=shadows)
    topt = torch.optim.Adam(ref, lr=1e-3, weight_decay=wd)
    g = torch.Generator().manual_seed(2)
    for step in range(25):
        for a, b in zip(mine, ref):
            gr = torch.randn(a.shape, generator=g).to(DEV) * (1.0 + step)
            a.grad, b.grad = gr.clone(), gr.clone()
        v0 = mine[2]._version
        opt.step()
        topt.step()
        assert mine[2]._version > v0  # raw-pointer update is made visible to autograd's version counter
    for a, b in zip(mine, ref):
        assert relerr(a, b) < 1e-6
    assert torch.equal(shadows[mine[2]][0], mine[2].detach().bfloat16())
    sd = opt.state_dict()
    tsd = topt.state_dict()
    assert float(sd["state"][0]["step"]) == float(tsd["state"][0]["step"]) == 25.0
    assert relerr(sd["state"][2]["exp_avg_sq"], tsd["state"][2]["exp_avg_sq"]) < 1e-6
    # state round trip through torch's own optimizer object and back
    topt2 = torch.optim.Adam(mine, lr=1e-3, weight_decay=wd)
    topt2.load_state_dict(sd)
    opt2 = Adam(mine, lr=1e-3, weight_decay=wd)
    opt2.load_state_dict(topt2.state_dict())
    assert float(opt2.step_count) == 25.0 and torch.equal(opt2.exp_avg[2], opt.exp_avg[2])


def test_adam_takes_bf16_gradients_from_a_grad_map():
    """The data-parallel gradient sinks (dp.GradSync) hand the optimizer bf16 tensors that are not p.grad."""
    from semantichuman_b200.optim import Adam

    shapes = [(256, 1031), (33, 5), (8200,)]
    mine, ref = _params(5, shapes), _params(5, shapes)
    opt, topt = Adam(mine, lr=1e-3, weight_decay=5e-5), torch.optim.Adam(ref, lr=1e-3, weight_decay=5e-5)
    g = torch.Generator().manual_seed(6)
    for _ in range(10):
        gmap = {}
        for i, (a, b) in enumerate(zip(mine, ref)):
            gr = torch.randn(a.shape, generator=g).to(DEV)
            if i != 1:   # tensors 0 and 2 arrive in bf16 through the map, tensor 1 through p.grad in fp32
                gmap[a] = gr.bfloat16()
                a.grad, b.grad = None, gr.bfloat16().float()
            else:
                a.grad, b.grad = gr.clone(), gr.clone()
        opt.step(grads=gmap)
        topt.step()
    for a, b in zip(mine, ref):
        assert relerr(a, b) < 1e-6


def test_adam_step_replays_in_a_cuda_graph():
    from semantichuman_b200.optim import Adam

    shapes = [(1000,), (17, 9)]
    mine, ref = _params(3, shapes), _params(3, shapes)
    opt, topt = Adam(mine, lr=1e-2), torch.optim.Adam(ref, lr=1e-2)
    grads = [torch.randn(s, device=DEV) for s in shapes]
    for a, b, g in zip(mine, ref, grads):
        a.grad, b.grad = g, g.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        opt.step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        opt.step()
    for _ in range(5):
        graph.replay()
    for _ in range(6):  # one eager step + five replays (the capture itself only records)
        topt.step()
    torch.cuda.synchronize()
    for a, b in zip(mine, ref):
        assert relerr(a, b) < 1e-6


def test_cpu_parameters_raise():
    from semantichuman_b200.optim import Adam

    with pytest.raises(TypeError):
        Adam([torch.zeros(3, requires_grad=True)])
