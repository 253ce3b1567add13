"""ORACLE (test infrastructure, never the product path): CPU restatement in torch of DeepErwin's
default `dpe4` wavefunction, its local energy and its Metropolis step.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (deeperwin_b200/) never does and fails loudly without its CUDA library.

PARITY UNPINNED (except the RNG, see oracle/threefry.py): the reference (jax 0.4.23, dm-haiku 0.0.13,
folx 0.2.12; /root/reference/uv.lock) cannot be imported in this image and its tests hold no numeric
golden vector for log psi^2 / E_loc / MCMC.  What pins this file instead:
(1) three independent Laplacians agree (torch.func Hessian trace in fp64, the reference's own
jvp-loop definition hamiltonian.py:234-267, and the explicit forward-Laplacian below);
(2) fermionic antisymmetry; (3) closed-form E_pot; (4) the analytic He-like test where
E_loc = -Z^2 + 1/r12 exactly (tests/test_oracle_model.py).

Every function cites the reference file:line (relative to /root/reference/src/deeperwin/) it restates.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np
import torch

LOG_EPSILON = 1e-8  # model/wavefunction.py:64
SQRT2 = math.sqrt(2.0)


# ----------------------------------------------------------------------------- dims / params
@dataclass
class ModelDims:
    """The subset of ModelConfigDeepErwin4 (configuration.py:422-440, 634-674, 783-806, 875-890)
    that fixes tensor shapes of the default model."""
    n_el: int
    n_up: int
    n_ion: int
    Z_max: int
    Z_min: int = 1
    n_iterations: int = 4
    n_hidden_one_el: List[int] = field(default_factory=lambda: [256] * 4)
    n_hidden_two_el: List[int] = field(default_factory=lambda: [32] * 3)
    emb_dim: int = 32
    n_ion_features: int = 32
    n_dets: int = 32
    # orbitals: False = envelope orbitals (dpe4 default); True = transferable atomic orbitals evaluated from a per-geometry
    # cache (sample_configs/pre_training_basemodel/config_bm_hfcoeff.yml:45-77: n_determinants 4, envelope_orbitals null)
    use_taos: bool = False

    def __post_init__(self):
        if isinstance(self.n_hidden_one_el, int):
            self.n_hidden_one_el = [self.n_hidden_one_el] * self.n_iterations
        if isinstance(self.n_hidden_two_el, int):
            self.n_hidden_two_el = [self.n_hidden_two_el] * (self.n_iterations - 1)
        assert len(self.n_hidden_one_el) == self.n_iterations
        assert len(self.n_hidden_two_el) == self.n_iterations - 1

    @property
    def n_dn(self):
        return self.n_el - self.n_up

    # feature widths entering iteration `it`
    def d_one_in(self, it):   # h_one
        return 4 * self.n_ion if it == 0 else self.n_hidden_one_el[it - 1]

    def d_pair_in(self, it):  # same / diff stream
        return 1 if it == 0 else self.n_hidden_two_el[it - 1]

    def d_eion_in(self, it):  # el-ion stream
        return 4 if it == 0 else self.n_hidden_two_el[it - 1]


EMB = "wf/fermi_net_embedding"
ORB = "wf/~/orbitals/envelope_orbitals"


def param_shapes(d: ModelDims) -> Dict[str, Dict[str, Tuple[int, ...]]]:
    """haiku parameter tree of the default model (names per SURVEY.md section 8b)."""
    shapes: Dict[str, Dict[str, Tuple[int, ...]]] = {}
    shapes["wf/~/input/h_ion"] = {"embeddings": (d.Z_max - d.Z_min + 1, d.n_ion_features)}
    for it in range(d.n_iterations):
        cf = f"{EMB}/symm_features_{it}/convolutional_features"
        for nm in ("w_same", "w_diff"):
            shapes[f"{cf}/{nm}/linear_0"] = {"w": (d.d_pair_in(it), d.emb_dim), "b": (d.emb_dim,)}
        shapes[f"{cf}/h_map/linear_0"] = {"w": (d.d_one_in(it), d.emb_dim), "b": (d.emb_dim,)}
        shapes[f"{cf}/h_ion_map/linear_0"] = {"w": (d.n_ion_features, d.d_eion_in(it)), "b": (d.d_eion_in(it),)}
        d_in = 3 * d.d_one_in(it) + d.emb_dim + d.d_eion_in(it)
        shapes[f"{EMB}/h_el_{it}/linear_0"] = {"w": (d_in, d.n_hidden_one_el[it]), "b": (d.n_hidden_one_el[it],)}
        if it < d.n_iterations - 1:
            for nm, din in (("h_same", d.d_pair_in(it)), ("h_diff", d.d_pair_in(it)), ("h_el_ion", d.d_eion_in(it))):
                shapes[f"{EMB}/{nm}_{it}/linear_0"] = {"w": (din, d.n_hidden_two_el[it]), "b": (d.n_hidden_two_el[it],)}
    if d.use_taos:   # the TAO heads have no per-walker parameters: backflows / exponents come from the geometry cache
        return shapes
    n_orb_tot = d.n_dets * d.n_el  # full_det: every spin block has n_el orbitals
    d_emb = d.n_hidden_one_el[-1]
    shapes[f"{ORB}/bf_up/linear_0"] = {"w": (d_emb, n_orb_tot)}
    shapes[f"{ORB}/bf_dn/linear_0"] = {"w": (d_emb, n_orb_tot)}
    shapes[ORB] = {k: (d.n_ion, n_orb_tot) for k in ("alpha_up", "alpha_dn", "weights_up", "weights_dn")}
    return shapes


def init_params(d: ModelDims, seed: int = 1234, bias_scale: float = 0.0, envelope_jitter: float = 0.0,
                dtype=torch.float64) -> Dict[str, Dict[str, torch.Tensor]]:
    """Random-init weights with the reference's distributions: VarianceScaling(1.0, fan_avg, uniform)
    weights (mlp.py:42), zero biases (mlp.py:43, init_bias_scale=0), unit envelopes
    (envelope_orbitals.py:34-37), hk.Embed default TruncatedNormal(1) embeddings.
    `bias_scale` / `envelope_jitter` > 0 give non-degenerate values for parity tests.
    The random stream is torch's, not haiku's: weights are INPUTS to the parity check."""
    g = torch.Generator().manual_seed(seed)
    params = {}
    for mod, leaves in param_shapes(d).items():
        params[mod] = {}
        for name, shape in leaves.items():
            if name == "w":
                lim = math.sqrt(3.0 / (0.5 * (shape[0] + shape[1])))
                t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim
            elif name == "b":
                t = torch.randn(shape, generator=g, dtype=torch.float64) * bias_scale
            elif name == "embeddings":
                t = torch.randn(shape, generator=g, dtype=torch.float64).clamp(-2, 2)
            else:  # alpha_*, weights_*
                t = 1.0 + envelope_jitter * (torch.rand(shape, generator=g, dtype=torch.float64) - 0.5)
            params[mod][name] = t.to(dtype)
    return params


def cast_params(params, dtype):
    return {m: {k: v.to(dtype) for k, v in leaves.items()} for m, leaves in params.items()}


# ----------------------------------------------------------------------------- arithmetic of the dense layers
# north_star permits two arithmetics for the batched dense layers: true fp32, or "fp32-accurate 3xTF32" (x = xh + xl, w = wh + wl
# with 11-bit tf32 halves; x w ~ xh wh + xh wl + xl wh, fp32 accumulation), which is ~4x noisier per product than an fp32 FMA
# (the xl wl term and the rounding of the lo halves are each ~2^-22 relative).  `arithmetic("3xtf32")` makes the fp32 evaluation of
# this oracle use that arithmetic in exactly the layers the CUDA path runs on the tensor cores (h_map, h_el, backflow / TAO
# projection) -- the parity rule's fp32 floor covers both (oracle/parity_rule.py).  fp64 evaluations are never affected.
_ARITH = "fp32"
DENSE_TC_LAYERS = ("h_map", "h_el_", "bf_up", "bf_dn")


class arithmetic:
    def __init__(self, mode):
        assert mode in ("fp32", "3xtf32")
        self.mode = mode

    def __enter__(self):
        global _ARITH
        self.prev, _ARITH = _ARITH, self.mode

    def __exit__(self, *exc):
        global _ARITH
        _ARITH = self.prev


def _tf32_rna(x):
    """cvt.rna.tf32.f32: round to 10 explicit mantissa bits, ties away from zero."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _mm(x, w, tc=False):
    if not (tc and _ARITH == "3xtf32" and x.dtype == torch.float32):
        return x @ w
    xh, wh = _tf32_rna(x), _tf32_rna(w)
    xl, wl = _tf32_rna(x - xh), _tf32_rna(w - wh)
    return (xh @ wl + xl @ wh) + xh @ wh          # small terms first, as the kernels issue them


def _is_tc_layer(name):
    return any(t in name for t in DENSE_TC_LAYERS)


# ----------------------------------------------------------------------------- plain forward
def _lin(params, name, x):
    p = params[name]
    y = _mm(x, p["w"], _is_tc_layer(name))
    return y + p["b"] if "b" in p else y


def _res(y, x):
    """mlp.py:13-16 residual_update."""
    return (x + y) / SQRT2 if x.shape == y.shape else y


def distances(r, R):
    """utils/utils.py:262-305. r [...,N,3], R [I,3]."""
    N = r.shape[-2]
    eye = torch.eye(N, dtype=r.dtype)
    diff_ee = r[..., None, :, :] - r[..., :, None, :]            # [i,j] = r_j - r_i   (utils.py:296)
    dist_ee = torch.linalg.norm(diff_ee + eye[..., None], dim=-1) * (1 - eye)   # utils.py:299-300
    diff_eI = r[..., :, None, :] - R[None, :, :]
    dist_eI = torch.linalg.norm(diff_eI, dim=-1)
    return diff_ee, dist_ee, diff_eI, dist_eI


def _split_same_diff(x, U):
    """ferminet_embedding.py:197-205. x [...,N,N,f] -> same [...,U^2+D^2,f], diff [...,2UD,f]."""
    b = x.shape[:-3]
    uu = x[..., :U, :U, :].reshape(b + (-1, x.shape[-1]))
    ud = x[..., :U, U:, :].reshape(b + (-1, x.shape[-1]))
    du = x[..., U:, :U, :].reshape(b + (-1, x.shape[-1]))
    dd = x[..., U:, U:, :].reshape(b + (-1, x.shape[-1]))
    return torch.cat([uu, dd], -2), torch.cat([ud, du], -2)


def embedding(params, d: ModelDims, r, R, Z):
    """input_features.py:112-303 + ferminet_embedding.py:207-267 for dpe4 defaults. Returns (h_el, dist_eI)."""
    U, D, N = d.n_up, d.n_dn, d.n_el
    b = r.shape[:-2]
    _, dist_ee, diff_eI, dist_eI = distances(r, R)
    h_eI = torch.cat([dist_eI[..., None], diff_eI], -1)                  # input_features.py:56-65 (dist first)
    h_one = h_eI.reshape(b + (N, 4 * d.n_ion))                           # input_features.py:237-239
    h_ion = params["wf/~/input/h_ion"]["embeddings"][(torch.as_tensor(Z).long() - d.Z_min)]  # :184-191
    same, diff = _split_same_diff(dist_ee[..., None], U)
    for it in range(d.n_iterations):
        cf = f"{EMB}/symm_features_{it}/convolutional_features"
        w_s = torch.tanh(_lin(params, f"{cf}/w_same/linear_0", same))   # ferminet_embedding.py:136-150
        w_d = torch.tanh(_lin(params, f"{cf}/w_diff/linear_0", diff))
        w_uu = w_s[..., : U * U, :].reshape(b + (U, U, -1))
        w_dd = w_s[..., U * U:, :].reshape(b + (D, D, -1))
        w_ud = w_d[..., : U * D, :].reshape(b + (U, D, -1))
        w_du = w_d[..., U * D:, :].reshape(b + (D, U, -1))
        hm = torch.tanh(_lin(params, f"{cf}/h_map/linear_0", h_one))    # :156-158
        h_u, h_d = hm[..., None, :U, :], hm[..., None, U:, :]
        emb_up = (w_uu * h_u).sum(-2) + (w_ud * h_d).sum(-2)            # :162-168
        emb_dn = (w_du * h_u).sum(-2) + (w_dd * h_d).sum(-2)
        conv_ee = torch.cat([emb_up, emb_dn], -2)
        him = torch.tanh(_lin(params, f"{cf}/h_ion_map/linear_0", h_ion))  # :171-175
        conv_eI = (h_eI * him).sum(-2)
        mean_up = h_one[..., :U, :].mean(-2, keepdim=True).expand(b + (N, -1))   # :60-72
        mean_dn = h_one[..., U:, :].mean(-2, keepdim=True).expand(b + (N, -1))
        f = torch.cat([h_one, mean_up, mean_dn, conv_ee, conv_eI], -1)
        h_one = _res(torch.tanh(_lin(params, f"{EMB}/h_el_{it}/linear_0", f)), f)   # :233-241
        if it == d.n_iterations - 1:
            break
        same = _res(torch.tanh(_lin(params, f"{EMB}/h_same_{it}/linear_0", same)), same)     # :247-253
        diff = _res(torch.tanh(_lin(params, f"{EMB}/h_diff_{it}/linear_0", diff)), diff)
        h_eI = _res(torch.tanh(_lin(params, f"{EMB}/h_el_ion_{it}/linear_0", h_eI)), h_eI)  # :259-262
    return h_one, dist_eI


def orbitals(params, d: ModelDims, h_el, dist_eI):
    """orbitals/envelope_orbitals.py:39-127 (full_det). Returns A [...,n_det,N,N] (rows = electrons)."""
    U, N, nd = d.n_up, d.n_el, d.n_dets
    b = h_el.shape[:-2]
    p = params[ORB]

    def block(h, dist, w_bf, alpha, weights):
        bf = _mm(h, w_bf, True)                                         # [...,n,nd*N]
        env = (weights * torch.exp(-torch.nn.functional.softplus(alpha) * dist[..., None])).sum(-2)  # :110-116
        mo = (env * bf).reshape(b + (h.shape[-2], nd, N))
        return mo.transpose(-3, -2)                                     # [...,nd,n,N]

    mo_up = block(h_el[..., :U, :], dist_eI[..., :U, :], params[f"{ORB}/bf_up/linear_0"]["w"], p["alpha_up"], p["weights_up"])
    mo_dn = block(h_el[..., U:, :], dist_eI[..., U:, :], params[f"{ORB}/bf_dn/linear_0"]["w"], p["alpha_dn"], p["weights_dn"])
    return torch.cat([mo_up, mo_dn], -2)                                # wavefunction.py:68


def make_tao_cache(d: ModelDims, seed: int = 11, dtype=torch.float64):
    """Synthetic stand-in for Wavefunction._calculate_cache (wavefunction.py:164-209): per spin
    backflows [I, n_orb, 2, n_det, emb] and exponents [I, n_orb, 2, n_det] (transferable_atomic_orbitals.py:142, 159 comments).
    The geometry-only nets that produce them (TAOBackflow / TAOExponents on orbital descriptors) are outside the hot path;
    values here are random with the scale of their outputs (exponents positive so that the orbitals decay)."""
    g = torch.Generator().manual_seed(seed)
    emb = d.n_hidden_one_el[-1]
    bfs, exs = [], []
    for n_orb in (d.n_up, d.n_dn):
        bfs.append((torch.randn(d.n_ion, n_orb, 2, d.n_dets, emb, generator=g, dtype=torch.float64) / math.sqrt(emb)).to(dtype))
        exs.append((0.5 + torch.rand(d.n_ion, n_orb, 2, d.n_dets, generator=g, dtype=torch.float64)).to(dtype))
    return {"backflows": bfs, "exponents": exs}


def cast_tao_cache(tao, dtype):
    return {k: [t.to(dtype) for t in v] for k, v in tao.items()}


def orbitals_tao(tao, d: ModelDims, h_el, dist_eI):
    """TransferableAtomicOrbitals.__call__ with a cache (transferable_atomic_orbitals.py:287-349), defaults
    use_el_ion_embedding=False, use_separate_ion_sum_for_envelopes=False, use_exponentials=True, full_det.
    Returns A [..., n_det, N, N] (rows = electrons; columns = [spin-up orbitals | spin-down orbitals])."""
    U, N = d.n_up, d.n_el
    mos = []
    for spin, (ex, bf) in enumerate(zip(tao["exponents"], tao["backflows"])):
        sl_same, sl_diff = (slice(None, U), slice(U, None)) if spin == 0 else (slice(U, None), slice(None, U))   # :316-319
        b_same = bf[..., :, :, 0, :, :]                                              # :234
        # :255-260 -- without el-ion embedding BOTH products use b_same (b_diff, :235, stays unused)
        I_, k_, d_, a_ = b_same.shape
        w_tao = b_same.permute(3, 2, 0, 1).reshape(a_, d_ * I_ * k_)                     # [emb, (det, ion, orb)]
        proj = lambda h: _mm(h, w_tao, True).reshape(h.shape[:-1] + (d_, I_, k_)).movedim(-3, -4)   # "Ikda,...ia->...diIk"
        mo_same, mo_diff = proj(h_el[..., sl_same, :]), proj(h_el[..., sl_diff, :])
        # :271-284: exponent[..., spin-type, det] * dist -> [el, ion, orb, det] -> moveaxis(-1, -4) -> [det, el, ion, orb]
        e_same = torch.exp(-ex[None, :, :, 0, :] * dist_eI[..., sl_same, :, None, None]).movedim(-1, -4)
        e_diff = torch.exp(-ex[None, :, :, 1, :] * dist_eI[..., sl_diff, :, None, None]).movedim(-1, -4)
        mos.append(((mo_same * e_same).sum(-2), (mo_diff * e_diff).sum(-2)))          # :322-329 (sum over ions)
    mo_up = torch.cat([mos[0][0], mos[1][1]], -1)                                     # :342-345
    mo_dn = torch.cat([mos[0][1], mos[1][0]], -1)
    return torch.cat([mo_up, mo_dn], -2)                                              # wavefunction.py:68


def sum_of_determinants(A):
    """model/wavefunction.py:63-83. Returns (phase, log_psi_sqr, sign_total)."""
    sign, logdet = torch.linalg.slogdet(A)
    shift = logdet.max(-1, keepdim=True).values
    psi = (torch.exp(logdet - shift) * sign).sum(-1)
    log_psi_sqr = 2 * (torch.log(psi.abs() + LOG_EPSILON) + shift.squeeze(-1))
    phase = torch.where(psi < 0, torch.full_like(psi, math.pi), torch.zeros_like(psi))  # jnp.angle of a real
    return phase, log_psi_sqr


def log_psi_sqr(params, d: ModelDims, r, R, Z, tao=None):
    """model/wavefunction.py:118-134, 293: (phase, log psi^2) for r [...,N,3].  `tao` = fixed_params["cache"]["taos"]
    (orbital_net.py:84-95) when the model uses transferable atomic orbitals."""
    h_el, dist_eI = embedding(params, d, r, R, Z)
    A = orbitals_tao(tao, d, h_el, dist_eI) if d.use_taos else orbitals(params, d, h_el, dist_eI)
    return sum_of_determinants(A)


# ----------------------------------------------------------------------------- energies
def potential_energy(r, R, Z):
    """hamiltonian.py:17-39."""
    Zf = torch.as_tensor(Z, dtype=r.dtype)
    _, _, _, dist_eI = distances(r, R)
    e_ei = -(Zf / dist_eI).sum((-2, -1))
    N = r.shape[-2]
    iu = torch.triu_indices(N, N, 1)
    dee = torch.linalg.norm(r[..., iu[0], :] - r[..., iu[1], :], dim=-1)
    e_ee = (1.0 / dee).sum(-1)
    I = R.shape[0]
    e_ii = r.new_zeros(())
    if I > 1:
        ju = torch.triu_indices(I, I, 1)
        dII = torch.linalg.norm(R[ju[0]] - R[ju[1]], dim=-1)
        e_ii = (Zf[ju[0]] * Zf[ju[1]] / dII).sum()
    return e_ee + e_ei + e_ii


def potential_energy_scale(r, R, Z):
    """Sum of the magnitudes of the terms of E_pot (the scale an fp32 evaluation's error is relative to)."""
    Zf = torch.as_tensor(Z, dtype=r.dtype)
    _, _, _, dist_eI = distances(r, R)
    N = r.shape[-2]
    iu = torch.triu_indices(N, N, 1)
    dee = torch.linalg.norm(r[..., iu[0], :] - r[..., iu[1], :], dim=-1)
    s = (Zf / dist_eI).sum((-2, -1)) + (1.0 / dee).sum(-1)
    I = R.shape[0]
    if I > 1:
        ju = torch.triu_indices(I, I, 1)
        s = s + (Zf[ju[0]] * Zf[ju[1]] / torch.linalg.norm(R[ju[0]] - R[ju[1]], dim=-1)).sum()
    return s


def kinetic_energy_hessian(params, d, r, R, Z, tao=None):
    """Definition: E_kin = -1/2 (1/2 lap L + 1/4 |grad L|^2), L = log psi^2 (hamiltonian.py:216),
    with lap/grad from torch.func (exact autodiff). r [B,N,3]. Returns (E_kin, grad[B,3N], lap[B])."""
    from torch.func import grad, hessian, vmap

    def f(x):
        return log_psi_sqr(params, d, x.reshape(d.n_el, 3), R, Z, tao)[1]

    x = r.reshape(r.shape[0], -1)
    g = vmap(grad(f))(x)
    H = vmap(hessian(f))(x)
    lap = torch.diagonal(H, dim1=-2, dim2=-1).sum(-1)
    return -0.5 * (0.5 * lap + 0.25 * (g * g).sum(-1)), g, lap


def kinetic_energy_jvp_loop(params, d, r, R, Z):
    """The reference's fallback branch restated literally (hamiltonian.py:234-267): linearize grad,
    loop over 3N unit vectors summing H_kk."""
    from torch.func import grad, jvp

    out = []
    for b in range(r.shape[0]):
        def f(x):
            return log_psi_sqr(params, d, x.reshape(d.n_el, 3), R, Z)[1]
        x = r[b].reshape(-1)
        gfun = grad(f)
        gval = gfun(x)
        lap = x.new_zeros(())
        eye = torch.eye(x.numel(), dtype=x.dtype)
        for i in range(x.numel()):
            lap = lap + jvp(gfun, (x,), (eye[i],))[1][i]
        out.append(-0.5 * (0.25 * (gval * gval).sum() + 0.5 * lap))
    return torch.stack(out)


def local_energy_hessian(params, d, r, R, Z, tao=None):
    """hamiltonian.py:281-289 with the autodiff-definition kinetic energy."""
    ek, _, _ = kinetic_energy_hessian(params, d, r, R, Z, tao)
    return ek + potential_energy(r, R, Z)


# ----------------------------------------------------------------------------- explicit forward Laplacian
# Restates what folx.forward_laplacian (folx 0.2.12, un-vendored; call site hamiltonian.py:209-214)
# computes for this network: value, dense Jacobian over the 3N coordinates, Laplacian; rules of
# SURVEY.md Appendix B.  Tensors carry a channel axis C = 3N + 2: [value, d/dx_0..d/dx_{3N-1}, laplacian].
# The pair stream and the el-ion stream use their structural sparsity (functions of the scalar
# distance / of r_i only), which is what folx's sparse Jacobians exploit.

def _tanh_rule(z, n_t):
    """z [..., C, f] with C = 1 + n_t + 1. Returns tanh with tangent/Laplacian channels."""
    y = torch.tanh(z[..., 0, :])
    d1 = 1 - y * y
    zt = z[..., 1:1 + n_t, :]
    s = (zt * zt).sum(-2)
    out = torch.empty_like(z)
    out[..., 0, :] = y
    out[..., 1:1 + n_t, :] = d1[..., None, :] * zt
    out[..., 1 + n_t, :] = d1 * z[..., 1 + n_t, :] - 2 * y * d1 * s
    return out


def _lin_rule(params, name, x):
    """x [..., C, f]: bias only on the value channel."""
    p = params[name]
    y = _mm(x, p["w"], _is_tc_layer(name))
    if "b" in p:
        y[..., 0, :] = y[..., 0, :] + p["b"]
    return y


def forward_laplacian(params, d: ModelDims, r, R, Z, return_intermediates=False, tao=None):
    """Returns dict(logpsi2[B], phase[B], grad[B,3N], lap[B], E_kin[B], E_pot[B], E_loc[B])."""
    U, D, N, I = d.n_up, d.n_dn, d.n_el, d.n_ion
    B = r.shape[0]
    K = 3 * N
    C = K + 2
    dt = r.dtype
    inter = {}
    diff_ee, dist_ee, diff_eI, dist_eI = distances(r, R)
    eyeN = torch.eye(N, dtype=dt)

    # ---- el-ion stream: channels [val, d/dr_i(x,y,z), lap]
    h_eI = torch.zeros(B, N, I, 5, 4, dtype=dt)
    h_eI[..., 0, 0] = dist_eI
    h_eI[..., 0, 1:] = diff_eI
    h_eI[..., 1:4, 0] = diff_eI / dist_eI[..., None]
    for a in range(3):
        h_eI[..., 1 + a, 1 + a] = 1.0
    h_eI[..., 4, 0] = 2.0 / dist_eI
    h_ion = params["wf/~/input/h_ion"]["embeddings"][(torch.as_tensor(Z).long() - d.Z_min)]

    # ---- pair stream as a function of the scalar distance: channels [f, f', f'']; diagonal has f'=f''=0
    offd = (1 - eyeN)
    pair = torch.zeros(B, N, N, 3, 1, dtype=dt)
    pair[..., 0, 0] = dist_ee
    pair[..., 1, 0] = offd
    same_mask = torch.zeros(N, N, dtype=torch.bool)
    same_mask[:U, :U] = True
    same_mask[U:, U:] = True
    u_ee = diff_ee / (dist_ee + eyeN)[..., None]                       # unit vector (r_j - r_i)/d, 0 on diagonal
    inv_d = offd / (dist_ee + eyeN)

    # ---- one-electron stream: [B,N,C,f]
    h_one = torch.zeros(B, N, C, 4 * I, dtype=dt)
    h_one[:, :, 0, :] = h_eI[..., 0, :].reshape(B, N, 4 * I)
    for i in range(N):
        h_one[:, i, 1 + 3 * i:4 + 3 * i, :] = h_eI[:, i, :, 1:4, :].permute(0, 2, 1, 3).reshape(B, 3, 4 * I)
    h_one[:, :, C - 1, :] = h_eI[..., 4, :].reshape(B, N, 4 * I)

    def pair_layer(name_same, name_diff, x):
        zs = _lin_rule(params, name_same, x)
        zd = _lin_rule(params, name_diff, x)
        z = torch.where(same_mask[None, :, :, None, None], zs, zd)
        return _tanh_rule(z, 1)

    for it in range(d.n_iterations):
        cf = f"{EMB}/symm_features_{it}/convolutional_features"
        w = pair_layer(f"{cf}/w_same/linear_0", f"{cf}/w_diff/linear_0", pair)       # [B,N,N,3,emb]
        hm = _tanh_rule(_lin_rule(params, f"{cf}/h_map/linear_0", h_one), K)           # [B,N,C,emb]
        w0, w1, w2 = w[..., 0, :], w[..., 1, :], w[..., 2, :]
        # conv_ee with the product rule
        cee = torch.einsum("bije,bjce->bice", w0, hm)                                  # w * (all channels of hm)
        hm0 = hm[:, :, 0, :]
        t = w1[..., None, :] * u_ee[..., :, None]                                      # [B,i,j,3,e] = w' u_a
        # d/d r_j of w_ij -> +w' u ; d/d r_i -> -w' u
        contrib_j = t * hm0[:, None, :, None, :]                                       # [B,i,j,3,e]
        cee_t = cee[:, :, 1:1 + K, :].view(B, N, N, 3, -1)                          # view [B,i,j',a,e]
        cee_t.add_(contrib_j)                                                          # in place: cee_t is a view of cee
        idx = torch.arange(N)
        cee_t[:, idx, idx] -= contrib_j.sum(2)
        lapw = 2 * w2 + 4 * w1 * inv_d[..., None]
        hm_t = hm[:, :, 1:1 + K, :].reshape(B, N, N, 3, -1)                            # [B,j,e',a,emb]
        hm_jj = hm_t[:, idx, idx]                                                      # [B,j,a,emb] d hm_j / d r_j
        cross = (u_ee[..., None] * (hm_jj[:, None, :, :, :] - hm_t.permute(0, 2, 1, 3, 4))).sum(-2)  # [B,i,j,emb]
        cee[:, :, C - 1, :] += (lapw * hm0[:, None, :, :] + 2 * w1 * cross).sum(2)
        # conv_eI
        him = torch.tanh(_lin(params, f"{cf}/h_ion_map/linear_0", h_ion))             # [I,dE]
        ceI5 = (h_eI * him[None, None, :, None, :]).sum(2)                             # [B,N,5,dE]
        ceI = torch.zeros(B, N, C, ceI5.shape[-1], dtype=dt)
        ceI[:, :, 0] = ceI5[:, :, 0]
        ceI[:, :, C - 1] = ceI5[:, :, 4]
        for i in range(N):
            ceI[:, i, 1 + 3 * i:4 + 3 * i] = ceI5[:, i, 1:4]
        mean_up = h_one[:, :U].mean(1, keepdim=True).expand(B, N, C, -1)
        mean_dn = h_one[:, U:].mean(1, keepdim=True).expand(B, N, C, -1)
        f = torch.cat([h_one, mean_up, mean_dn, cee, ceI], -1)
        if return_intermediates:
            inter[f"f_{it}"] = f
            inter[f"w_{it}"] = w
            inter[f"hm_{it}"] = hm
        y = _tanh_rule(_lin_rule(params, f"{EMB}/h_el_{it}/linear_0", f), K)
        h_one = _res(y, f)
        if return_intermediates:
            inter[f"h_{it}"] = h_one
        if it == d.n_iterations - 1:
            break
        pair = _res(pair_layer(f"{EMB}/h_same_{it}/linear_0", f"{EMB}/h_diff_{it}/linear_0", pair), pair)
        h_eI = _res(_tanh_rule(_lin_rule(params, f"{EMB}/h_el_ion_{it}/linear_0", h_eI), 3), h_eI)

    # ---- orbitals
    nd = d.n_dets
    mo = torch.empty(B, N, C, nd * N, dtype=dt)
    if d.use_taos:
        # TAOs (transferable_atomic_orbitals.py:287-349): mo[i, d, k] = sum_I (h_i . b[I, k, d]) exp(-x[I, k, s(i,k), d] |r_i - R_I|);
        # product rule with e = exp(-x dist): grad_i e = -x e (r_i - R_I)/dist, lap e = e (x^2 - 2 x / dist)
        for spin, (ex, bf) in enumerate(zip(tao["exponents"], tao["backflows"])):
            k0, n_orb = (0, U) if spin == 0 else (U, D)
            for sl, st in ((slice(0, U), 0 if spin == 0 else 1), (slice(U, N), 1 if spin == 0 else 0)):   # st: 0 same / 1 diff
                I_, k_, _, d_, a_ = bf.shape
                w_tao = bf[:, :, 0].permute(3, 0, 2, 1).reshape(a_, I_ * d_ * k_)                # [emb, (ion, det, orb)]
                g = _mm(h_one[:, sl], w_tao, True).reshape(h_one[:, sl].shape[:-1] + (I_, d_, k_))   # "Ikda,bnca->bncIdk"
                x = ex[:, :, st, :].permute(0, 2, 1)                                             # [I,nd,n_orb]
                dist = dist_eI[:, sl]                                                            # [B,n,I]
                e = torch.exp(-x * dist[..., None, None])                                        # [B,n,I,nd,n_orb]
                unit = diff_eI[:, sl] / dist[..., None]                                          # [B,n,I,3]
                e_t = torch.einsum("bnIx,bnIdk->bnxIdk", unit, -x * e)                           # [B,n,3,I,nd,n_orb]
                e_l = e * (x * x - 2 * x / dist[..., None, None])
                m = (g * e[:, :, None]).sum(3)                                                   # [B,n,C,nd,n_orb]
                for n_loc, i in enumerate(range(N)[sl]):
                    m[:, n_loc, 1 + 3 * i:4 + 3 * i] += (e_t[:, n_loc] * g[:, n_loc, 0:1]).sum(2)
                    m[:, n_loc, C - 1] += (e_l[:, n_loc] * g[:, n_loc, 0]).sum(1) + 2 * (e_t[:, n_loc] * g[:, n_loc, 1 + 3 * i:4 + 3 * i]).sum((1, 2))
                mo.view(B, N, C, nd, N)[:, sl, :, :, k0:k0 + n_orb] = m
    else:
      p = params[ORB]
      for sl, wname, an, wn in ((slice(0, U), "bf_up", "alpha_up", "weights_up"), (slice(U, N), "bf_dn", "alpha_dn", "weights_dn")):
        bf = _mm(h_one[:, sl], params[f"{ORB}/{wname}/linear_0"]["w"], True)         # [B,n,C,nd*N]
        a = torch.nn.functional.softplus(p[an])                                        # [I,cols]
        dist = dist_eI[:, sl]                                                          # [B,n,I]
        e = p[wn] * torch.exp(-a * dist[..., None])                                    # [B,n,I,cols]
        env = e.sum(2)
        env_t = torch.einsum("bnia,bnic->bnac", diff_eI[:, sl] / dist[..., None], -a * e)   # [B,n,3,cols]
        env_l = (e * (a * a - 2 * a / dist[..., None])).sum(2)
        m = env[:, :, None, :] * bf
        els = range(N)[sl]
        for n_loc, i in enumerate(els):
            m[:, n_loc, 1 + 3 * i:4 + 3 * i] += env_t[:, n_loc] * bf[:, n_loc, 0:1]
            m[:, n_loc, C - 1] += env_l[:, n_loc] * bf[:, n_loc, 0] + 2 * (env_t[:, n_loc] * bf[:, n_loc, 1 + 3 * i:4 + 3 * i]).sum(1)
        mo[:, sl] = m
    A = mo.reshape(B, N, C, nd, N).permute(0, 3, 2, 1, 4)                              # [B,nd,C,i,orb]
    A0 = A[:, :, 0]
    sign, logdet = torch.linalg.slogdet(A0)
    Ainv = torch.linalg.inv(A0)                                                        # [B,nd,orb,i]
    P = torch.einsum("bdoi,bdkip->bdkop", Ainv, A[:, :, 1:1 + K])                       # A^-1 dA_k
    g_d = torch.diagonal(P, dim1=-2, dim2=-1).sum(-1)                                  # [B,nd,K]
    trP2 = torch.einsum("bdkop,bdkpo->bd", P, P)
    lap_d = torch.einsum("bdoi,bdio->bd", Ainv, A[:, :, C - 1]) - trP2
    # ---- signed log-sum-exp incl. the epsilon of wavefunction.py:81
    shift, m_idx = logdet.max(-1, keepdim=True)
    q = sign * torch.exp(logdet - shift)
    psi = q.sum(-1)
    wgt = q / psi[:, None]
    G = torch.einsum("bd,bdk->bk", wgt, g_d)
    Gkk = torch.einsum("bd,bdk->bk", wgt, g_d * g_d).sum(-1) + (wgt * lap_d).sum(-1) - (G * G).sum(-1)
    g_m = torch.gather(g_d, 1, m_idx[..., None].expand(B, 1, K)).squeeze(1)
    l_m = torch.gather(lap_d, 1, m_idx).squeeze(1)
    rho = psi.abs() / (psi.abs() + LOG_EPSILON)
    F_k = rho[:, None] * (G - g_m) + g_m
    F_lap = (rho * (1 - rho))[:, None].mul((G - g_m) ** 2).sum(-1) + rho * (Gkk - l_m) + l_m
    logpsi2 = 2 * (torch.log(psi.abs() + LOG_EPSILON) + shift.squeeze(-1))
    grad = 2 * F_k
    lap = 2 * F_lap
    e_kin = -0.5 * (0.5 * lap + 0.25 * (grad * grad).sum(-1))
    e_pot = potential_energy(r, R, Z)
    # conditioning of the walker: every determinant enters through its inverse (error amplification cond_2(A_d)) and the signed sum
    # over determinants amplifies by sum|q_d| / |sum q_d|  ->  cond_eff = sum_d |q_d| cond(A_d) / |sum_d q_d|
    cond_eff = (q.abs() * torch.linalg.cond(A0.double())).sum(-1) / psi.abs().double().clamp_min(1e-300)
    out = dict(logpsi2=logpsi2, phase=torch.where(psi < 0, torch.full_like(psi, math.pi), torch.zeros_like(psi)), grad=grad, lap=lap,
               E_kin=e_kin, E_pot=e_pot, E_pot_scale=potential_energy_scale(r, R, Z), E_loc=e_kin + e_pot, sign_d=sign, logdet_d=logdet,
               cond=cond_eff)
    if return_intermediates:
        inter.update(mo=mo, g_d=g_d, lap_d=lap_d)
        out["inter"] = inter
    return out


def local_energy(params, d, r, R, Z, max_batch_size=64, tao=None):
    """build_local_energy(..., forward_lap=True, max_batch_size) (hamiltonian.py:272-291): sequential
    chunks of <= max_batch_size walkers, as folx.batched_vmap does."""
    outs = [forward_laplacian(params, d, r[s:s + max_batch_size], R, Z, tao=tao)["E_loc"] for s in range(0, r.shape[0], max_batch_size)]
    return torch.cat(outs)
