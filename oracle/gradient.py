"""ORACLE (test infrastructure, never the product path): parameter gradient of the VMC loss and the KFAC statistics
of the optimisation step, restated with torch autograd in float64.

Reference (paths relative to /root/reference/src/deeperwin/ and /root/reference/custom_kfac_jax/kfac_jax/_src/):
  * gradient: optimization/loss_function.py:112-154 -- total_energy's custom jvp:
        d loss = (1 / B) sum_b (E_clipped_b - mean(E_clipped)) * d log psi^2_b          (real wavefunctions)
    i.e. the gradient is the backward pass of  sum_b c_b log psi^2_b  with the per-walker cotangent c_b = (E_c,b - E_c,mean) / B.
    `param_gradient` takes the cotangents c_b; tests/test_reference_pin.py checks it against the reference's own
    jax.value_and_grad(total_energy) executed under tests/ref_shim.
  * KFAC statistics (PARITY UNPINNED numerically: kfac_jax needs jax; restated from its source): the loss is registered as a normal
    predictive distribution on 1/2 log psi^2 with variance 1/2 (loss_function.py:148-150, loss_functions.py:1191-1197), estimation
    mode "fisher_exact" (optimizers.py:183-216): for the single output coordinate the cotangent on the registered mean is
    1 / sqrt(variance) (loss_functions.py NormalMeanNegativeLogProbLoss.multiply_fisher_factor_replicated_one_hot,
    curvature_estimator.py:1469-1510), so every dense layer sees  dy = sqrt(2) * d(1/2 log psi^2)/dy = (1/sqrt 2) d log psi^2 / dy
    per sample.  Dense blocks (curvature_blocks.py:1594-1624), with the repeated-dense folding of the electron / pair / ion axes
    into the batch (curvature_tags_and_blocks.py:41-64):
        A = [x, 1]^T [x, 1] / B'     (the ones column only if the layer has a bias),     G = dy^T dy / B',     B' = rows of x.
    The reference tiles R and Z over the walkers (input_features.py:121-125), so the ion-level layers (hk.Embed lookup, h_ion_map)
    also see one row per (walker, ion); `tile_ions=True` below reproduces that.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import model as om


def _leaf_params(params):
    return {m: {k: v.clone().requires_grad_(True) for k, v in l.items()} for m, l in params.items()}


def param_gradient(params, d: om.ModelDims, r, R, Z, cotangent, tao=None) -> Dict[str, Dict[str, torch.Tensor]]:
    """d/d params of sum_b cotangent_b * log psi^2_b  (loss_function.py:143-154 with cotangent = (E_c - mean E_c) / B).
    TAO models (`tao` = the geometry cache): the gradient with respect to the embedding parameters; the cache is held fixed."""
    p = _leaf_params(params)
    lp = om.log_psi_sqr(p, d, r, R, Z, tao)[1]
    (lp * torch.as_tensor(cotangent, dtype=lp.dtype)).sum().backward()
    return {m: {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in l.items()} for m, l in p.items()}


def loss_gradient_from_energies(params, d, r, R, Z, E_clipped, tao=None):
    """The full gradient of total_energy for given clipped local energies (loss_function.py:143-154)."""
    E = torch.as_tensor(E_clipped, dtype=torch.float64)
    return param_gradient(params, d, r, R, Z, (E - E.mean()) / E.numel(), tao)


class _Recorder:
    """Records (x, y) of every dense layer of one forward pass (oracle.model._lin is routed through it)."""
    def __init__(self):
        self.layers = {}

    def lin(self, params, name, x):
        p = params[name]
        y = x @ p["w"]
        if "b" in p:
            y = y + p["b"]
        y.retain_grad()
        self.layers[name] = (x, y, "b" in p)
        return y


def kfac_factors(params, d: om.ModelDims, r, R, Z, tao=None):
    """{layer: (A, G, rows_per_walker)} of every dense layer of the dpe4 model, and the per-walker forward pass they come from."""
    rec = _Recorder()
    p = _leaf_params(params)
    B = r.shape[0]
    saved = om._lin
    om._lin = rec.lin
    try:
        # ion-level layers see one row per (walker, ion), as in the reference where Z is tiled over the batch
        emb_tab = p["wf/~/input/h_ion"]["embeddings"].clone().requires_grad_(True)
        onehot = torch.nn.functional.one_hot(torch.as_tensor(Z).long() - d.Z_min, d.Z_max - d.Z_min + 1).to(r.dtype)       # [I, V]
        onehot_b = onehot[None].expand(B, -1, -1)
        h_ion_b = onehot_b @ emb_tab                                                                                        # [B, I, F]
        h_ion_b.retain_grad()
        lp = _log_psi_sqr_tiled(p, d, r, R, h_ion_b, tao)
    finally:
        om._lin = saved
    (lp.sum() / math.sqrt(2.0)).backward()
    out = {}
    for name, (x, y, has_bias) in rec.layers.items():
        x2 = x.reshape(-1, x.shape[-1]).detach()
        dy = y.grad.reshape(-1, y.shape[-1])
        rows = x2.shape[0]
        if has_bias:
            x2 = torch.cat([x2, torch.ones(rows, 1, dtype=x2.dtype)], 1)
        out[name] = (x2.T @ x2 / rows, dy.T @ dy / rows, rows // B)
    # hk.Embed lookup (ONE_HOT style: a dense layer without bias on the one-hot rows)
    x2 = onehot_b.reshape(-1, onehot.shape[-1])
    dy = h_ion_b.grad.reshape(-1, h_ion_b.shape[-1])
    out["wf/~/input/h_ion"] = (x2.T @ x2 / x2.shape[0], dy.T @ dy / x2.shape[0], x2.shape[0] // B)
    return out


def _log_psi_sqr_tiled(params, d, r, R, h_ion_b, tao=None):
    """oracle.model.embedding / orbitals / sum_of_determinants with per-walker ion features h_ion_b [B, I, F]."""
    U, D, N = d.n_up, d.n_dn, d.n_el
    b = r.shape[:-2]
    _, dist_ee, diff_eI, dist_eI = om.distances(r, R)
    h_eI = torch.cat([dist_eI[..., None], diff_eI], -1)
    h_one = h_eI.reshape(b + (N, 4 * d.n_ion))
    same, diff = om._split_same_diff(dist_ee[..., None], U)
    EMB = om.EMB
    for it in range(d.n_iterations):
        cf = f"{EMB}/symm_features_{it}/convolutional_features"
        w_s = torch.tanh(om._lin(params, f"{cf}/w_same/linear_0", same))
        w_d = torch.tanh(om._lin(params, f"{cf}/w_diff/linear_0", diff))
        w_uu = w_s[..., : U * U, :].reshape(b + (U, U, -1))
        w_dd = w_s[..., U * U:, :].reshape(b + (D, D, -1))
        w_ud = w_d[..., : U * D, :].reshape(b + (U, D, -1))
        w_du = w_d[..., U * D:, :].reshape(b + (D, U, -1))
        hm = torch.tanh(om._lin(params, f"{cf}/h_map/linear_0", h_one))
        h_u, h_d = hm[..., None, :U, :], hm[..., None, U:, :]
        conv_ee = torch.cat([(w_uu * h_u).sum(-2) + (w_ud * h_d).sum(-2), (w_du * h_u).sum(-2) + (w_dd * h_d).sum(-2)], -2)
        him = torch.tanh(om._lin(params, f"{cf}/h_ion_map/linear_0", h_ion_b))                  # [B, I, dE]
        conv_eI = (h_eI * him[..., None, :, :]).sum(-2)
        mean_up = h_one[..., :U, :].mean(-2, keepdim=True).expand(b + (N, -1))
        mean_dn = h_one[..., U:, :].mean(-2, keepdim=True).expand(b + (N, -1))
        f = torch.cat([h_one, mean_up, mean_dn, conv_ee, conv_eI], -1)
        h_one = om._res(torch.tanh(om._lin(params, f"{EMB}/h_el_{it}/linear_0", f)), f)
        if it == d.n_iterations - 1:
            break
        same = om._res(torch.tanh(om._lin(params, f"{EMB}/h_same_{it}/linear_0", same)), same)
        diff = om._res(torch.tanh(om._lin(params, f"{EMB}/h_diff_{it}/linear_0", diff)), diff)
        h_eI = om._res(torch.tanh(om._lin(params, f"{EMB}/h_el_ion_{it}/linear_0", h_eI)), h_eI)
    if d.use_taos:
        return om.sum_of_determinants(om.orbitals_tao(tao, d, h_one, dist_eI))[1]
    ORB = om.ORB
    p = params[ORB]
    nd = d.n_dets

    def block(h, dist, name, alpha, weights):
        bf = om._lin(params, f"{ORB}/{name}/linear_0", h)
        env = (weights * torch.exp(-torch.nn.functional.softplus(alpha) * dist[..., None])).sum(-2)
        return (env * bf).reshape(b + (h.shape[-2], nd, N)).transpose(-3, -2)

    A = torch.cat([block(h_one[..., :U, :], dist_eI[..., :U, :], "bf_up", p["alpha_up"], p["weights_up"]),
                   block(h_one[..., U:, :], dist_eI[..., U:, :], "bf_dn", p["alpha_dn"], p["weights_dn"])], -2)
    return om.sum_of_determinants(A)[1]
