"""B200 (`-m gpu`): fused clip_grad_norm_ / AdamW kernels (simxns_b200/optim.py over simx_grad_sq_norm, simx_clip_grads,
simx_adamw_step) against the float64 oracle (oracle/optim_numpy.py, pinned to torch in tests/test_optim_cpu.py) and
against torch's own CUDA optimizer.  fp32 arithmetic: 2e-6 relative on parameters after several steps."""
import numpy as np
import pytest
import torch

from oracle import optim_numpy as on
from simxns_b200 import _lib
from simxns_b200 import optim as sx

pytestmark = pytest.mark.gpu

# ragged sizes: below one float4, not a multiple of 4, exactly one chunk, several chunks, a bias-like vector
SHAPES = [(3,), (8192,), (8193,), (77, 13), (3072, 768), (768,), (30522, 64)]


@pytest.fixture(scope="module", autouse=True)
def _device():
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    _lib.require_device()


def _params(seed):
    rng = np.random.RandomState(seed)
    return [rng.standard_normal(s).astype(np.float32) for s in SHAPES]


def _close(got, want, tol, what):
    want = np.asarray(want, np.float64)
    err = np.abs(got.double().cpu().numpy() - want).max() / (np.abs(want).max() + 1e-30)
    assert err < tol, (what, err)


@pytest.mark.parametrize("scale,max_norm", [(1.0, 1.0), (1e-4, 1.0), (30.0, 2.0)])
def test_clip_grad_norm_vs_oracle(scale, max_norm):
    grads = [g * scale for g in _params(1)]
    ps = [torch.nn.Parameter(torch.zeros(g.shape, device="cuda")) for g in grads]
    for p, g in zip(ps, grads):
        p.grad = torch.from_numpy(g).cuda()
    ps.append(torch.nn.Parameter(torch.zeros(5, device="cuda")))  # no gradient: skipped
    total = sx.clip_grad_norm_(ps, max_norm)
    t, clipped = on.clip_grad_norm(grads, max_norm)
    assert abs(total.item() - t) < 2e-6 * t
    for p, c in zip(ps, clipped):
        _close(p.grad, c, 2e-6, "clipped grad")
    assert ps[-1].grad is None


@pytest.mark.parametrize("flavor,wd,correct_bias", [("hf", 0.0, True), ("hf", 0.01, True), ("hf", 0.01, False),
                                                    ("torch", 0.01, True), ("torch", 0.0, True)])
def test_adamw_steps_vs_oracle(flavor, wd, correct_bias):
    p0 = _params(2)
    ps = [torch.nn.Parameter(torch.from_numpy(p.copy()).cuda()) for p in p0]
    # the reference's two groups (co_training_marco_train.py:57-65): decay / no decay, one lr schedule
    groups = [dict(params=ps[:4], weight_decay=wd), dict(params=ps[4:], weight_decay=0.0)]
    opt = sx.AdamW(groups, lr=2e-3, eps=1e-6, correct_bias=correct_bias, flavor=flavor)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))  # get_linear_schedule_* is a LambdaLR too
    state = [(p.astype(np.float64), np.zeros(p.shape), np.zeros(p.shape)) for p in p0]
    fn = on.adamw_hf if flavor == "hf" else on.adamw_torch
    for step in range(1, 5):
        grads = [g * 0.1 for g in _params(20 + step)]
        for p, g in zip(ps, grads):
            p.grad = torch.from_numpy(g).cuda()
        lr = opt.param_groups[0]["lr"]
        v0 = ps[0]._version
        opt.step()
        sched.step()
        assert ps[0]._version > v0  # the encoder's bf16 weight cache is keyed on _version
        new = []
        for i, ((p, m, v), g) in enumerate(zip(state, grads)):
            kw = dict(eps=1e-6, weight_decay=wd if i < 4 else 0.0)
            if flavor == "hf":
                kw["correct_bias"] = correct_bias
            new.append(fn(p, g, m, v, step, lr, **kw))
        state = new
    for p, (q, m, v) in zip(ps, state):
        _close(p.data, q, 2e-6, "param")
        _close(opt.state[p]["exp_avg"], m, 2e-6, "exp_avg")
        _close(opt.state[p]["exp_avg_sq"], v, 2e-6, "exp_avg_sq")


def test_fused_clip_inside_step_equals_clip_then_step_and_torch_cuda():
    p0, grads = _params(3), [g * 5.0 for g in _params(4)]
    def make():
        ps = [torch.nn.Parameter(torch.from_numpy(p.copy()).cuda()) for p in p0]
        for p, g in zip(ps, grads):
            p.grad = torch.from_numpy(g.copy()).cuda()
        return ps
    a, b, c = make(), make(), make()
    oa = sx.AdamW(a, lr=1e-3, eps=1e-8, weight_decay=0.01, flavor="torch")
    oa.step(max_grad_norm=1.0)                      # norm + update: two launches, clip coefficient never leaves the GPU
    ob = sx.AdamW(b, lr=1e-3, eps=1e-8, weight_decay=0.01, flavor="torch")
    nb = sx.clip_grad_norm_(b, 1.0)
    ob.step()
    oc = torch.optim.AdamW(c, lr=1e-3, eps=1e-8, weight_decay=0.01)
    nc = torch.nn.utils.clip_grad_norm_(c, 1.0)
    oc.step()
    assert abs(oa.last_grad_norm.item() - nc.item()) < 2e-6 * nc.item() and abs(nb.item() - nc.item()) < 2e-6 * nc.item()
    for x, y, z in zip(a, b, c):
        assert torch.allclose(x, z, rtol=2e-6, atol=2e-7) and torch.allclose(y, z, rtol=2e-6, atol=2e-7)
        assert torch.allclose(x.grad, z.grad, rtol=2e-6, atol=1e-9)  # the clipped gradient is left in .grad


def test_cpu_tensors_raise():
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sx.AdamW([p]).step()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sx.clip_grad_norm_([p], 1.0)
