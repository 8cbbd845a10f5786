"""Step epilogue: FlatAdam vs the oracle's restatement of torch.optim.Adam (exp_runner.py:115, 263).

CPU: the oracle is pinned to the live torch.optim.Adam; FlatAdam's host logic (flat views, checkpoint format both ways,
loud failure without CUDA).  GPU: the kernel against the oracle and against torch.optim.Adam on the same gradients.
"""
import numpy as np
import pytest
import torch

from oracle import rnb_oracle as O
from conftest import rel_l2


def _params(device, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(256, 39), (256,), (256, 1), (217, 256), (3,), (1,), (257, 256), ()]
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(device)) for s in shapes]


def _grads(params, k, scale=1e-3):
    g = torch.Generator().manual_seed(100 + k)
    return [(torch.randn(p.shape, generator=g) * scale * (1 + i)).to(p.device) for i, p in enumerate(params)]


def test_oracle_adam_matches_torch_adam():
    ps = _params("cpu")
    ref = [torch.nn.Parameter(p.detach().double().clone()) for p in ps]
    opt = torch.optim.Adam(ref, lr=5e-4)
    st = [(p.detach().double().numpy().copy(), np.zeros(p.shape), np.zeros(p.shape)) for p in ps]
    for k in range(1, 8):
        lr = O.learning_rate(k * 700)
        for grp in opt.param_groups:
            grp["lr"] = lr
        gs = _grads(ps, k)
        for r, g in zip(ref, gs):
            r.grad = g.double()
        opt.step()
        st = [O.adam_step(p, g.double().numpy(), m, v, k, lr) for (p, m, v), g in zip(st, gs)]
    for r, (p, m, v) in zip(ref, st):
        np.testing.assert_allclose(p, r.detach().numpy(), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(m, opt.state[r]["exp_avg"].numpy(), rtol=1e-12, atol=1e-18)


def test_learning_rate_schedule():
    assert O.learning_rate(0) == 0.0 and abs(O.learning_rate(2500) - 2.5e-4) < 1e-12
    assert abs(O.learning_rate(5000) - 5e-4) < 1e-12 and abs(O.learning_rate(300000) - 2.5e-5) < 1e-12


def test_flat_adam_host_logic_and_checkpoint_format():
    from rnb_b200.optim import FlatAdam
    ps = _params("cpu")
    before = [p.detach().clone() for p in ps]
    opt = FlatAdam(ps, lr=5e-4)
    for p, b in zip(ps, before):                               # parameters moved into the flat buffer unchanged
        assert torch.equal(p.detach(), b)
        assert p.data_ptr() >= opt.flat_param.data_ptr() and p.grad is not None and p.data_ptr() % 16 == 0
    assert opt.param_groups[0]["lr"] == 5e-4
    loss = sum((p ** 2).sum() for p in ps)
    loss.backward()                                            # autograd accumulates into the flat gradient buffer
    assert abs(float(opt.reducer.flat.sum()) - float(sum((2 * p).sum() for p in ps))) < 1e-2
    opt.zero_grad()                                            # torch semantics (set_to_none): nothing launched
    assert all(p.grad is None for p in ps)
    opt.reducer.collect()                                      # no gradient arrived: the flat buffer reads zero
    assert float(opt.reducer.flat.abs().sum()) == 0.0 and all(p.grad is v for p, v in zip(ps, opt.reducer.views))
    opt.zero_grad()
    (ps[0] ** 2).sum().backward()                              # autograd hands the gradient over without a kernel ...
    assert ps[0].grad is not opt.reducer.views[0]
    opt.reducer.collect()                                      # ... and one multi-tensor copy gathers it
    assert ps[0].grad is opt.reducer.views[0] and torch.equal(ps[0].grad, 2 * ps[0].detach())
    assert float(opt.reducer.views[1].abs().sum()) == 0.0
    opt.zero_grad()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        opt.step()
    with pytest.raises(NotImplementedError):
        FlatAdam(_params("cpu"), weight_decay=0.1)

    # a torch.optim.Adam checkpoint loads into FlatAdam, and back
    ref = [torch.nn.Parameter(b.clone()) for b in before]
    tadam = torch.optim.Adam(ref, lr=5e-4)
    for k in range(1, 4):
        for r, g in zip(ref, _grads(ref, k)):
            r.grad = g
        tadam.step()
    opt.load_state_dict(tadam.state_dict())
    assert opt._step == 3
    for p, r in zip(ps, ref):
        assert torch.equal(opt.state[p]["exp_avg"], tadam.state[r]["exp_avg"])
        assert opt.state[p]["exp_avg"].data_ptr() >= opt.flat_m.data_ptr()
    tadam2 = torch.optim.Adam([torch.nn.Parameter(b.clone()) for b in before], lr=1.0)
    tadam2.load_state_dict(opt.state_dict())
    assert tadam2.param_groups[0]["lr"] == 5e-4
    for q, r in zip(tadam2.param_groups[0]["params"], ref):
        assert torch.equal(tadam2.state[q]["exp_avg_sq"], tadam.state[r]["exp_avg_sq"])
        assert float(tadam2.state[q]["step"]) == 3.0


@pytest.mark.gpu
def test_flat_adam_kernel_matches_oracle_and_torch():
    from rnb_b200.optim import FlatAdam
    ps = _params("cuda")
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    st = [(p.detach().double().cpu().numpy(), np.zeros(p.shape), np.zeros(p.shape)) for p in ps]
    opt = FlatAdam(ps, lr=5e-4)
    tadam = torch.optim.Adam(ref, lr=5e-4)
    for k in range(1, 21):
        lr = O.learning_rate(k * 300)
        for o in (opt, tadam):
            for grp in o.param_groups:
                grp["lr"] = lr
        gs = _grads(ps, k)
        opt.zero_grad()
        for p, r, g in zip(ps, ref, gs):
            p.grad = g.clone()                                 # what autograd does after zero_grad()
            r.grad = g.clone()
        opt.step()
        tadam.step()
        st = [O.adam_step(p, g.double().cpu().numpy(), m, v, k, lr) for (p, m, v), g in zip(st, gs)]
    for p, r, (po, mo, vo) in zip(ps, ref, st):
        # float32 arithmetic against the float64 oracle: the update itself (p - p0) to 1e-5, parameters to fp32 rounding
        assert rel_l2(p.detach().cpu().numpy(), po) < 2e-7
        assert rel_l2(opt.state[p]["exp_avg"].cpu().numpy(), mo) < 1e-6
        assert rel_l2(opt.state[p]["exp_avg_sq"].cpu().numpy(), vo) < 1e-6
        assert rel_l2(p.detach().cpu().numpy(), r.detach().cpu().numpy()) < 2e-7
    # grad_scale folds the 1/world of the all-reduce into the same launch: scale 0.5 on doubled gradients = same update
    a, b = _params("cuda", 3), _params("cuda", 3)
    oa, ob = FlatAdam(a, lr=1e-3), FlatAdam(b, lr=1e-3, grad_scale=0.5)
    for p, q, g in zip(a, b, _grads(a, 1)):
        p.grad.add_(g)                                         # views attached at construction: in-place accumulation
        q.grad.add_(2 * g)
    v0 = a[0]._version
    oa.step()
    ob.step()
    assert torch.equal(oa.flat_param, ob.flat_param)
    assert a[0]._version > v0                                  # caches keyed on version counters see the update
    # gradients a caller assigned itself (not the attached views) and missing gradients are folded in
    c = _params("cuda", 4)
    oc = FlatAdam(c, lr=1e-3)
    c0 = [p.detach().clone() for p in c]
    c[0].grad = torch.ones_like(c[0])
    c[1].grad = None
    oc.step()
    assert torch.allclose(c[0].detach(), c0[0] - 1e-3, atol=1e-6) and torch.equal(c[1].detach(), c0[1])


@pytest.mark.gpu
def test_flat_adam_trains_like_torch_adam():
    """A short train_rnb run (exp_runner.py:259-263) with FlatAdam follows the same loss curve as torch.optim.Adam."""
    from rnb_b200 import synth
    from rnb_b200.optim import FlatAdam
    from test_gpu_e2e import loss_fn, make_renderer
    curves = []
    for make in (lambda ps: torch.optim.Adam(ps, lr=5e-4), lambda ps: FlatAdam(ps, lr=5e-4)):
        renderer, sdf, var, col = make_renderer(False)
        params = [p for m in (sdf, var, col) for p in m.parameters()]
        opt = make(params)
        b = {k: v.cuda() for k, v in synth.make_batch(256, 3, True, 1).items()}
        losses = []
        for it in range(6):
            opt.zero_grad()
            torch.manual_seed(7 + it)
            out = renderer.render_rnb_warmup(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"],
                                             cos_anneal_ratio=1.0)
            loss = loss_fn(out, b["true_rgb"], b["mask"], 0.1)
            loss.backward()
            opt.step()
            losses.append(float(loss))
        curves.append(losses)
    assert curves[0][-1] < curves[0][0]
    np.testing.assert_allclose(curves[1], curves[0], rtol=2e-3)
