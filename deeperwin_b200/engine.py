"""Host-side owner of one dpe_model handle: parameter flattening, geometry, workspace, stream plumbing.
PyTorch is used for device memory and streams only; all arithmetic happens in libdpe_b200.so."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import DpeDims, DpeMcmcConfig, DpeMcmcState, MODE_FORWARD, MODE_LAPLACIAN, check

EMB = "wf/fermi_net_embedding"
ORB = "wf/~/orbitals/envelope_orbitals"


def canonical_leaves(n_iterations: int, use_taos: bool = False) -> List[Tuple[str, str]]:
    """(haiku module path, leaf name) in the flat order documented in include/dpe_b200.h."""
    leaves = [("wf/~/input/h_ion", "embeddings")]
    for it in range(n_iterations):
        cf = f"{EMB}/symm_features_{it}/convolutional_features"
        for nm in ("w_same", "w_diff", "h_map", "h_ion_map"):
            leaves += [(f"{cf}/{nm}/linear_0", "w"), (f"{cf}/{nm}/linear_0", "b")]
        leaves += [(f"{EMB}/h_el_{it}/linear_0", "w"), (f"{EMB}/h_el_{it}/linear_0", "b")]
        if it < n_iterations - 1:
            for nm in ("h_same", "h_diff", "h_el_ion"):
                leaves += [(f"{EMB}/{nm}_{it}/linear_0", "w"), (f"{EMB}/{nm}_{it}/linear_0", "b")]
    if use_taos:      # TAO heads: nothing per walker is trainable here, the cache carries backflows / exponents
        return leaves
    leaves += [(f"{ORB}/bf_up/linear_0", "w"), (f"{ORB}/bf_dn/linear_0", "w")]
    leaves += [(ORB, k) for k in ("alpha_up", "alpha_dn", "weights_up", "weights_dn")]
    return leaves


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, *, n_el, n_up, n_ion, n_iterations, n_hidden_one_el, n_hidden_two_el, emb_dim, n_ion_features,
                 n_dets, z_min, z_max, use_taos: bool = False, device="cuda:0", workspace_gb: float = 48.0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("deeperwin_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        d = DpeDims()
        d.n_el, d.n_up, d.n_ion, d.n_iterations = n_el, n_up, n_ion, n_iterations
        for i, v in enumerate(n_hidden_one_el):
            d.n_hidden_one_el[i] = v
        for i, v in enumerate(n_hidden_two_el):
            d.n_hidden_two_el[i] = v
        d.emb_dim, d.n_ion_features, d.n_dets, d.z_min, d.z_max = emb_dim, n_ion_features, n_dets, z_min, z_max
        d.use_taos = int(bool(use_taos))
        self.use_taos = bool(use_taos)
        self.n_dets, self.d_last = n_dets, list(n_hidden_one_el)[n_iterations - 1]
        self._tao_keep = None
        self.dims = d
        self.n_el, self.n_up, self.n_ion = n_el, n_up, n_ion
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.dpe_model_create(C.byref(d), C.byref(self.handle)), "dpe_model_create")
        self.n_params = self.lib.dpe_param_count(self.handle)
        self.leaves = canonical_leaves(n_iterations, self.use_taos)
        assert len(self.leaves) == self.lib.dpe_param_leaf_count(self.handle)
        self.leaf_shapes = []
        for i in range(len(self.leaves)):
            off, size, rows, cols = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
            check(self.lib.dpe_param_leaf(self.handle, i, C.byref(off), C.byref(size), C.byref(rows), C.byref(cols)), "dpe_param_leaf")
            self.leaf_shapes.append((off.value, size.value, rows.value, cols.value))
        self._flat = torch.empty(self.n_params, dtype=torch.float32, device=self.device)
        self._flat_views = [self._flat[off:off + size] for off, size, _, _ in self.leaf_shapes]
        self._geom_keep = None
        self._ws: Optional[torch.Tensor] = None
        self._ws_need: Dict[Tuple[int, int], int] = {}
        self._workspace_cap = int(workspace_gb * 2 ** 30)
        if self.use_taos:
            # TAO head: the sum over ions of the projected orbitals amplifies the rounding of the embedding 3x more than the envelope
            # head does; the tensor core's round-toward-zero accumulation then puts single walkers at > 16x their fp32 floor
            # (tools/parity_table.py), so these models keep their dense layers on the FP32 SIMT GEMM unless asked otherwise
            self.set_gemm_path(0)
        if os.environ.get("DPE_GEMM_PATH", "") != "":       # 0 = FP32 SIMT GEMM, 1 = tcgen05 3xTF32 GEMM
            self.set_gemm_path(int(os.environ["DPE_GEMM_PATH"]))

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.dpe_model_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    @property
    def workspace_cap(self) -> int:
        return self._workspace_cap

    @workspace_cap.setter
    def workspace_cap(self, nbytes: int):
        self._workspace_cap = int(nbytes)
        self._ws_need.clear()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_params(self, params: Dict[str, Dict[str, torch.Tensor]]):
        """`params` of log_psi_sqr(params, ...).  The weights are uploaded on EVERY call (one multi-tensor copy into the flat
        vector + the library's re-split, a few MB device-to-device): freshness is never inferred from tensor identity, so in-place
        updates through `.data`, raw-pointer writes by an optimiser kernel or a re-allocated tensor at the same address are all seen."""
        with torch.cuda.device(self.device):
            srcs = []
            for (mod, name), (off, size, rows, cols) in zip(self.leaves, self.leaf_shapes):
                t = params[mod][name]
                if t.numel() != size:
                    raise ValueError(f"parameter {mod}/{name} has {t.numel()} values, expected {size} ({rows}x{cols})")
                if t.device != self.device or t.dtype != torch.float32:
                    t = t.to(device=self.device, dtype=torch.float32)
                srcs.append(t.detach().reshape(-1))
            torch._foreach_copy_(self._flat_views, srcs)
            check(self.lib.dpe_model_set_params(self.handle, _ptr(self._flat), self.n_params, self._stream()), "dpe_model_set_params")

    def set_geometry(self, R, Z):
        """`R, Z` of log_psi_sqr(params, n_up, n_dn, r, R, Z, fixed_params).  CUDA tensors (MCMCState.R / .Z) are handed to the
        library as device pointers -- no host copy, no synchronisation; host arrays go through the library's pinned staging ring."""
        if isinstance(R, torch.Tensor) and R.is_cuda:
            Rd = R.detach().to(device=self.device, dtype=torch.float32)
            Zd = (Z.detach() if isinstance(Z, torch.Tensor) else torch.as_tensor(np.asarray(Z))).to(device=self.device, dtype=torch.int32)
            if Rd.ndim == 3:      # tiled over a device/batch axis by the caller: all copies are equal
                Rd, Zd = Rd[0], Zd[0]
            if tuple(Rd.shape) != (self.n_ion, 3) or tuple(Zd.shape) != (self.n_ion,):
                raise ValueError(f"R/Z shapes {tuple(Rd.shape)}/{tuple(Zd.shape)} do not match n_ion={self.n_ion}")
            Rd, Zd = Rd.contiguous(), Zd.contiguous()
            with torch.cuda.device(self.device):
                check(self.lib.dpe_model_set_geometry_dev(self.handle, _ptr(Rd), _ptr(Zd), self._stream()), "dpe_model_set_geometry_dev")
            self._geom_keep = (Rd, Zd)       # the copy kernel runs asynchronously on the stream
            return
        Rn = np.ascontiguousarray(np.asarray(R.detach().cpu() if isinstance(R, torch.Tensor) else R, dtype=np.float32))
        Zn = np.ascontiguousarray(np.asarray(Z.detach().cpu() if isinstance(Z, torch.Tensor) else Z).astype(np.int32))
        if Rn.ndim == 3:
            Rn, Zn = Rn[0], Zn[0]
        if Rn.shape != (self.n_ion, 3) or Zn.shape != (self.n_ion,):
            raise ValueError(f"R/Z shapes {Rn.shape}/{Zn.shape} do not match n_ion={self.n_ion}")
        with torch.cuda.device(self.device):
            check(self.lib.dpe_model_set_geometry(self.handle, Rn.ctypes.data_as(C.POINTER(C.c_float)),
                                                  Zn.ctypes.data_as(C.POINTER(C.c_int32)), self._stream()), "dpe_model_set_geometry")

    def geometry_status(self):
        """Raises if a device-side set_geometry ever saw a nuclear charge outside the embedding vocabulary (synchronises)."""
        with torch.cuda.device(self.device):
            check(self.lib.dpe_model_geometry_status(self.handle, self._stream()), "dpe_model_geometry_status")

    def set_tao_cache(self, cache: Optional[Dict]):
        """fixed_params["cache"]["taos"] of the reference (wavefunction.py:164-209, orbital_net.py:84-95):
        {"backflows": [up, dn], "exponents": [up, dn]} with shapes [I, n_orb, 2, n_det, emb] / [I, n_orb, 2, n_det]."""
        if not self.use_taos:
            return
        if not cache or "backflows" not in cache or "exponents" not in cache:
            raise NotImplementedError('this model evaluates transferable atomic orbitals from fixed_params["cache"]["taos"] '
                                      "(backflows, exponents); the geometry-only nets that fill the cache are not part of the hot path")
        bfs, exs = list(cache["backflows"]), list(cache["exponents"])     # re-packed on every call, like the weights (set_params)
        n_dn = self.n_el - self.n_up
        dev = []
        for t, n_orb, tail in ((bfs[0], self.n_up, (self.n_dets, self.d_last)), (bfs[1], n_dn, (self.n_dets, self.d_last)),
                               (exs[0], self.n_up, (self.n_dets,)), (exs[1], n_dn, (self.n_dets,))):
            want = (self.n_ion, n_orb, 2) + tail
            if tuple(t.shape) != want:
                raise ValueError(f"TAO cache entry has shape {tuple(t.shape)}, expected {want}")
            dev.append(t.to(device=self.device, dtype=torch.float32).contiguous())
        with torch.cuda.device(self.device):
            check(self.lib.dpe_model_set_tao_cache(self.handle, _ptr(dev[0]), _ptr(dev[1]), _ptr(dev[2]), _ptr(dev[3]), self._stream()),
                  "dpe_model_set_tao_cache")
        self._tao_keep = dev       # the pack kernel runs asynchronously on the stream

    def workspace(self, n_walkers: int, mode: int) -> torch.Tensor:
        key = (n_walkers, mode)
        need = self._ws_need.get(key)
        if need is None:
            need = self._ws_need[key] = min(self.lib.dpe_workspace_bytes(self.handle, n_walkers, mode), self.workspace_cap)
        have = 0 if self._ws is None else self._ws.numel()
        if have < need:            # grow only: the device is queried for free memory when (and only when) a larger buffer is needed
            free, _ = torch.cuda.mem_get_info(self.device)
            need = min(need, have + int(free * 0.9))
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _r(self, r: torch.Tensor) -> torch.Tensor:
        r = r.to(device=self.device, dtype=torch.float32).contiguous()
        if r.shape[-2:] != (self.n_el, 3):
            raise ValueError(f"r has shape {tuple(r.shape)}, expected [..., {self.n_el}, 3]")
        return r

    # ------------------------------------------------------------------ hot path
    def log_psi_sqr(self, r: torch.Tensor):
        r = self._r(r)
        batch = r.shape[:-2]
        B = int(np.prod(batch)) if len(batch) else 1
        phase = torch.empty(B, dtype=torch.float32, device=self.device)
        lp = torch.empty(B, dtype=torch.float32, device=self.device)
        ws = self.workspace(B, MODE_FORWARD)
        with torch.cuda.device(self.device):
            check(self.lib.dpe_log_psi_sqr(self.handle, _ptr(r), B, _ptr(phase), _ptr(lp), _ptr(ws), ws.numel(), self._stream()), "dpe_log_psi_sqr")
        return phase.reshape(batch), lp.reshape(batch)

    def local_energy(self, r: torch.Tensor, with_aux: bool = False):
        r = self._r(r)
        batch = r.shape[:-2]
        B = int(np.prod(batch)) if len(batch) else 1
        e_loc = torch.empty(B, dtype=torch.float32, device=self.device)
        lp = grad = ekin = epot = None
        if with_aux:
            lp = torch.empty(B, dtype=torch.float32, device=self.device)
            grad = torch.empty(B, 3 * self.n_el, dtype=torch.float32, device=self.device)
            ekin = torch.empty(B, dtype=torch.float32, device=self.device)
            epot = torch.empty(B, dtype=torch.float32, device=self.device)
        ws = self.workspace(B, MODE_LAPLACIAN)
        with torch.cuda.device(self.device):
            check(self.lib.dpe_local_energy(self.handle, _ptr(r), B, _ptr(e_loc), _ptr(lp), _ptr(grad), _ptr(ekin), _ptr(epot),
                                            _ptr(ws), ws.numel(), self._stream()), "dpe_local_energy")
        if with_aux:
            return e_loc.reshape(batch), dict(log_psi_sqr=lp.reshape(batch), grad=grad.reshape(batch + (-1,)),
                                              E_kin=ekin.reshape(batch), E_pot=epot.reshape(batch))
        return e_loc.reshape(batch)

    # ------------------------------------------------------------------ optimisation step
    def kfac_layers(self):
        """[(haiku module, din, dout, has_bias, rows_per_walker, a_offset, g_offset)] of the dense layers (dpe_kfac_layer)."""
        if getattr(self, "_kfac_layers", None) is None:
            out = []
            for i in range(self.lib.dpe_kfac_layer_count(self.handle)):
                name = C.create_string_buffer(128)
                v = [C.c_int32() for _ in range(4)]
                o = [C.c_int64() for _ in range(2)]
                check(self.lib.dpe_kfac_layer(self.handle, i, name, 128, *[C.byref(x) for x in v], *[C.byref(x) for x in o]), "dpe_kfac_layer")
                out.append((name.value.decode(), v[0].value, v[1].value, v[2].value, v[3].value, o[0].value, o[1].value))
            self._kfac_layers = out
        return self._kfac_layers

    def param_gradient(self, r: torch.Tensor, cotangent: Optional[torch.Tensor], with_kfac: bool = False, out: Optional[torch.Tensor] = None):
        """Backward pass of log psi^2 (dpe_param_gradient).  Returns (flat, log_psi_sqr): `flat` = [gradient (n_params) | KFAC factors] in ONE
        buffer, so that a multi-GPU caller reduces both with a single all-reduce."""
        r = self._r(r).reshape(-1, self.n_el, 3)
        B = r.shape[0]
        n_k = int(self.lib.dpe_kfac_floats(self.handle)) if with_kfac else 0
        n_g = self.n_params if cotangent is not None else 0
        if out is None:
            out = torch.empty(n_g + n_k, dtype=torch.float32, device=self.device)
        lp = torch.empty(B, dtype=torch.float32, device=self.device)
        cot = None if cotangent is None else cotangent.to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
        need = min(self.lib.dpe_gradient_workspace_bytes(self.handle, B), self.workspace_cap)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        grad_ptr = _ptr(out) if n_g else None
        kfac_ptr = C.c_void_p(out.data_ptr() + 4 * n_g) if n_k else None
        with torch.cuda.device(self.device):
            check(self.lib.dpe_param_gradient(self.handle, _ptr(r), B, _ptr(cot), grad_ptr, kfac_ptr, _ptr(lp), _ptr(self._ws), self._ws.numel(),
                                              self._stream()), "dpe_param_gradient")
        return out, lp

    def set_mcmc_graph(self, on: bool):
        """Replay repeated mcmc_steps calls from a captured CUDA graph (default) or launch every kernel eagerly."""
        check(self.lib.dpe_set_mcmc_graph(self.handle, 1 if on else 0), "dpe_set_mcmc_graph")

    def mcmc_steps(self, state_struct: DpeMcmcState, n_walkers: int, n_steps: int, cfg: DpeMcmcConfig, recompute: bool,
                   run_controller: bool, counts: torch.Tensor):
        ws = self.workspace(n_walkers, MODE_FORWARD)
        with torch.cuda.device(self.device):
            check(self.lib.dpe_mcmc_steps(self.handle, C.byref(state_struct), n_walkers, n_steps, C.byref(cfg), int(recompute),
                                          int(run_controller), _ptr(counts), _ptr(ws), ws.numel(), self._stream()), "dpe_mcmc_steps")

    def mcmc_controller(self, state_struct: DpeMcmcState, counts: torch.Tensor, n_steps: int, n_total: int, cfg: DpeMcmcConfig):
        with torch.cuda.device(self.device):
            check(self.lib.dpe_mcmc_controller(C.byref(state_struct), _ptr(counts), n_steps, n_total, C.byref(cfg), self._stream()),
                  "dpe_mcmc_controller")

    # ------------------------------------------------------------------ debug / accounting
    def ws_view(self, name: str, n_walkers: int, mode: int, shape: Sequence[int]) -> torch.Tensor:
        off = self.lib.dpe_debug_ws_offset(self.handle, n_walkers, mode, name.encode())
        if off < 0:
            raise KeyError(name)
        n = int(np.prod(shape))
        return self._ws[off:off + 4 * n].view(torch.float32).reshape(*shape)

    def ldx(self, n_walkers: int, mode: int) -> int:
        return int(self.lib.dpe_debug_ws_offset(self.handle, n_walkers, mode, b"ldx"))

    def launch_count(self) -> int:
        return int(self.lib.dpe_launch_count(self.handle))

    def set_gemm_path(self, path: int):
        check(self.lib.dpe_set_gemm_path(self.handle, path), "dpe_set_gemm_path")

    def set_det_path(self, generic: bool = False, simt: bool = False):
        """Test knob (dpe_set_det_path): generic block-per-matrix determinant kernel also for n_el <= 16 / CUDA-core tangent stage."""
        check(self.lib.dpe_set_det_path(self.handle, int(generic) | (int(simt) << 1)), "dpe_set_det_path")

    GEMM_CLASSES = {0: "k_gemm_simt<128,128,8,8>", 1: "k_gemm_simt<128,64,8,4>", 2: "k_gemm_simt<256,32,8,4>", 3: "k_gemm_tc_3xtf32",
                    4: "k_gemm_tc2_3xtf32"}

    def profile_stages(self, fn):
        """Runs fn() once with per-stage CUDA-event timing (dpe_profile_stages): {stage: (ms, timed launch groups)}."""
        from ._lib import STAGE_NAMES
        fn()
        torch.cuda.synchronize(self.device)
        check(self.lib.dpe_profile_enable(self.handle, 1), "dpe_profile_enable")
        fn()
        torch.cuda.synchronize(self.device)
        check(self.lib.dpe_profile_enable(self.handle, 0), "dpe_profile_enable")
        n = len(STAGE_NAMES)
        ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
        rc = self.lib.dpe_profile_stages(self.handle, ms, cnt, n)
        if rc < 0:
            check(rc, "dpe_profile_stages")
        for klass in self.GEMM_CLASSES:            # drop the per-GEMM records of the same pass
            a, b, c = C.c_double(), C.c_int64(), C.c_double()
            self.lib.dpe_profile_collect(self.handle, klass, C.byref(a), C.byref(b), C.byref(c))
        return {STAGE_NAMES[i]: (ms[i], cnt[i]) for i in range(n) if cnt[i]}

    def profile_gemms(self, fn):
        """Runs fn() once with per-launch CUDA-event timing of the dense-layer GEMMs (bench.py roofline).
        Returns the kernel class with the largest summed time."""
        fn()
        torch.cuda.synchronize(self.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()                      # un-instrumented duration of the whole pass
        e1.record()
        torch.cuda.synchronize(self.device)
        check(self.lib.dpe_profile_enable(self.handle, 1), "dpe_profile_enable")
        fn()
        torch.cuda.synchronize(self.device)
        check(self.lib.dpe_profile_enable(self.handle, 0), "dpe_profile_enable")
        from ._lib import STAGE_NAMES
        n_st = len(STAGE_NAMES)
        st_ms, st_cnt = (C.c_double * n_st)(), (C.c_int64 * n_st)()
        self.lib.dpe_profile_stages(self.handle, st_ms, st_cnt, n_st)
        stages = {STAGE_NAMES[i]: round(st_ms[i], 4) for i in range(n_st) if st_cnt[i]}
        best = None
        launches = {}
        for klass, name in self.GEMM_CLASSES.items():
            cap = 256
            ms_arr, fl_arr, n = (C.c_double * cap)(), (C.c_double * cap)(), C.c_int32()
            check(self.lib.dpe_profile_launches(self.handle, klass, ms_arr, fl_arr, cap, C.byref(n)), "dpe_profile_launches")
            launches[name] = [(ms_arr[i], fl_arr[i]) for i in range(n.value)]
        for klass, name in self.GEMM_CLASSES.items():
            ms, cnt, fl = C.c_double(), C.c_int64(), C.c_double()
            check(self.lib.dpe_profile_collect(self.handle, klass, C.byref(ms), C.byref(cnt), C.byref(fl)), "dpe_profile_collect")
            if cnt.value and (best is None or ms.value > best["ms"]):
                best = dict(kernel=name, ms=ms.value, count=cnt.value, flops=fl.value)
        if best:
            best["total_ms"] = e0.elapsed_time(e1)
            best["class_ms"], best["class_count"], best["class_flops"] = best["ms"], best["count"], best["flops"]
            # the dominant launch SHAPE of that class: the launches with the largest algorithmic FLOP count
            per = launches[best["kernel"]]
            fmax = max(f for _, f in per)
            top = [(t, f) for t, f in per if f >= 0.999 * fmax]
            best.update(ms=sum(t for t, _ in top), count=len(top), flops=sum(f for _, f in top))
            best["stages_ms"] = stages
        return best
