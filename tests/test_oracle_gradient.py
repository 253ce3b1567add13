"""The gradient / KFAC oracle (oracle/gradient.py) against independent definitions: finite differences of the loss, and the
Kronecker factors recomputed from per-sample autograd."""
import math

import pytest
import torch

from oracle import gradient as og, model as om

SMALL = dict(n_iterations=2, n_hidden_one_el=[16, 16], n_hidden_two_el=[4], emb_dim=8, n_dets=3)


def _case(B=5):
    d = om.ModelDims(n_el=4, n_up=2, n_ion=2, Z_max=3, **SMALL)
    p = om.init_params(d, seed=5, bias_scale=0.1, envelope_jitter=0.5)
    R = torch.tensor([[0.0, 0, 0], [3.0, 0, 0]], dtype=torch.float64)
    g = torch.Generator().manual_seed(1)
    r = R[torch.tensor([0, 1, 0, 0])][None] + torch.randn(B, 4, 3, generator=g, dtype=torch.float64)
    return d, p, R, [3, 1], r


def test_gradient_matches_finite_differences():
    d, p, R, Z, r = _case()
    c = torch.tensor([0.3, -0.2, 0.5, -0.7, 0.1], dtype=torch.float64)
    grads = og.param_gradient(p, d, r, R, Z, c)
    f = lambda q: float((om.log_psi_sqr(q, d, r, R, Z)[1] * c).sum())
    gen = torch.Generator().manual_seed(0)
    for m, leaves in p.items():
        for k, v in leaves.items():
            idx = tuple(int(torch.randint(0, s, (1,), generator=gen)) for s in v.shape)
            q = {mm: {kk: vv.clone() for kk, vv in ll.items()} for mm, ll in p.items()}
            h = 1e-6
            q[m][k][idx] += h
            up = f(q)
            q[m][k][idx] -= 2 * h
            fd = (up - f(q)) / (2 * h)
            assert abs(fd - float(grads[m][k][idx])) <= 1e-6 * max(1.0, abs(fd)), (m, k, fd, float(grads[m][k][idx]))


def test_kfac_factors_definition():
    """A = [x, 1]^T [x, 1] / B', G = dy^T dy / B' with dy = (1 / sqrt 2) d log psi^2 / dy per sample; the tiled forward pass used to
    record them equals the plain one; the Kronecker product of the factors reproduces the sum of per-row gradients in expectation
    form for a layer applied once per walker (B' = B): trace(A) trace(G) bounds, symmetry, positive semi-definiteness."""
    d, p, R, Z, r = _case(B=6)
    fac = og.kfac_factors(p, d, r, R, Z)
    names = set(fac)
    assert f"{om.EMB}/h_el_0/linear_0" in names and f"{om.ORB}/bf_up/linear_0" in names and "wf/~/input/h_ion" in names
    assert len(names) == 5 * d.n_iterations + 3 * (d.n_iterations - 1) + 2 + 1
    for name, (A, G, rpw) in fac.items():
        assert torch.allclose(A, A.T) and torch.allclose(G, G.T)
        assert torch.linalg.eigvalsh(A).min() > -1e-10 and torch.linalg.eigvalsh(G).min() > -1e-12
    A, G, rpw = fac[f"{om.EMB}/h_el_1/linear_0"]
    assert rpw == d.n_el and A.shape == (3 * 16 + 8 + 4 + 1,) * 2 and abs(float(A[-1, -1]) - 1.0) < 1e-12      # ones column: mean of 1
    A, G, rpw = fac[f"{om.EMB}/symm_features_0/convolutional_features/w_same/linear_0"]
    assert rpw == 2 * 2 + 2 * 2                                                                                 # U^2 + D^2 same-spin pairs
    A, G, rpw = fac["wf/~/input/h_ion"]
    assert rpw == 2 and torch.allclose(A, torch.diag(torch.tensor([0.5, 0.0, 0.5], dtype=torch.float64)))       # one-hot rows: Z = 3 and Z = 1
    # G of a once-per-walker ... per-row layer against per-sample autograd: bf_up, rows = (walker, spin-up electron)
    name = f"{om.ORB}/bf_up/linear_0"
    pl = {m: {k: v.clone().requires_grad_(m == name) for k, v in l.items()} for m, l in p.items()}
    rows = []
    for b in range(r.shape[0]):
        lp = om.log_psi_sqr(pl, d, r[b:b + 1], R, Z)[1].sum() / math.sqrt(2.0)
        (gw,) = torch.autograd.grad(lp, pl[name]["w"])
        rows.append(gw)
    # sum over rows of x_row^T dy_row = dW per walker; the factors satisfy  sum_b ||dW_b||_F^2 <= B'^2 trace(A) trace(G)
    A, G, rpw = fac[name]
    total = sum(float((g ** 2).sum()) for g in rows)
    assert total <= (r.shape[0] * rpw) ** 2 * float(torch.trace(A)) * float(torch.trace(G)) * (1 + 1e-9)
