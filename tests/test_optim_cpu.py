"""CPU: the optimizer oracle (oracle/optim_numpy.py) pinned against the installed torch — torch.optim.AdamW is LEAD's
optimizer (LEAD/util.py:901), torch.nn.utils.clip_grad_norm_ is every driver's clip.  transformers.AdamW (SimANS/PROD) no
longer exists in the installed transformers: its restatement is tied to adamw_torch through the relation the two
published algorithms imply."""
import numpy as np
import torch

from oracle import optim_numpy as on


def _tensors(seed, shapes=((7, 5), (33,), (4, 8, 3))):
    rng = np.random.RandomState(seed)
    return [rng.standard_normal(s) for s in shapes]


def test_clip_grad_norm_matches_torch():
    for scale, max_norm in ((1.0, 1.0), (1e-3, 1.0), (50.0, 2.0)):
        grads = [g * scale for g in _tensors(1)]
        ps = [torch.nn.Parameter(torch.zeros(g.shape, dtype=torch.float64)) for g in grads]
        for p, g in zip(ps, grads):
            p.grad = torch.from_numpy(g.copy())
        total = torch.nn.utils.clip_grad_norm_(ps, max_norm)
        t, clipped = on.clip_grad_norm(grads, max_norm)
        assert abs(total.item() - t) < 1e-12 * max(1.0, t)
        for p, c in zip(ps, clipped):
            assert np.abs(p.grad.numpy() - c).max() < 1e-12


def test_adamw_torch_oracle_matches_torch_optim():
    ps0, lr, wd, eps = _tensors(2), 3e-3, 0.01, 1e-8
    ps = [torch.nn.Parameter(torch.from_numpy(p.copy())) for p in ps0]
    opt = torch.optim.AdamW(ps, lr=lr, eps=eps, weight_decay=wd)
    state = [(p.copy(), np.zeros_like(p), np.zeros_like(p)) for p in ps0]
    for step in range(1, 6):
        grads = _tensors(10 + step)
        for p, g in zip(ps, grads):
            p.grad = torch.from_numpy(g.copy())
        opt.step()
        state = [on.adamw_torch(p, g, m, v, step, lr, eps=eps, weight_decay=wd) for (p, m, v), g in zip(state, grads)]
        for p, (q, _, _) in zip(ps, state):
            assert np.abs(p.detach().numpy() - q).max() < 1e-12


def test_adamw_hf_restatement_relations():
    p0, g = _tensors(3)[0], _tensors(4)[0]
    m = v = np.zeros_like(p0)
    # eps = 0, no decay: both published algorithms take the same step
    a = on.adamw_torch(p0, g, m, v, 3, 1e-3, eps=0.0, weight_decay=0.0)
    b = on.adamw_hf(p0, g, m, v, 3, 1e-3, eps=0.0, weight_decay=0.0)
    for x, y in zip(a, b):
        assert np.abs(x - y).max() < 1e-15
    # first step, correct_bias, eps -> 0: |delta p| = lr (sign of g); decay multiplies the UPDATED parameter
    p1, _, _ = on.adamw_hf(p0, g, m, v, 1, 1e-2, eps=0.0, weight_decay=0.1)
    want = (p0 - 1e-2 * np.sign(g)) * (1 - 1e-2 * 0.1)
    assert np.abs(p1 - want).max() < 1e-12
    # correct_bias=False: plain lr on the uncorrected moments
    p2, m2, v2 = on.adamw_hf(p0, g, m, v, 1, 1e-2, eps=1e-6, correct_bias=False)
    assert np.abs(p2 - (p0 - 1e-2 * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-6))).max() < 1e-12
