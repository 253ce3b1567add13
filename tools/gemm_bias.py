"""Characterises the rounding of the tcgen05 3xTF32 dense-layer GEMM against FP32 SIMT and fp64: signed bias and spread of the
relative error for same-sign and mixed-sign sums, as a function of K (number of accumulation steps)."""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deeperwin_b200 import _lib
from deeperwin_b200._lib import DpeDims

lib = _lib.load()
d = DpeDims()
d.n_el, d.n_up, d.n_ion, d.n_iterations = 4, 2, 2, 1
d.n_hidden_one_el[0] = 16
d.emb_dim, d.n_ion_features, d.n_dets, d.z_min, d.z_max = 8, 32, 2, 1, 3
h = C.c_void_p()
_lib.check(lib.dpe_model_create(C.byref(d), C.byref(h)), "create")
p = lambda t: C.c_void_p(t.data_ptr())
g = torch.Generator(device="cuda").manual_seed(0)
M, N = 4096, 256
for K in (16, 64, 128, 256, 320):
    for kind in ("positive", "mixed"):
        A = torch.rand(M, K, device="cuda", generator=g) + 0.5 if kind == "positive" else torch.randn(M, K, device="cuda", generator=g)
        W = (torch.rand(K, N, device="cuda", generator=g) + 0.5) * 0.1 if kind == "positive" else (torch.rand(K, N, device="cuda", generator=g) * 2 - 1) * 0.1
        ref = A.double() @ W.double()
        row = []
        for path in (0, 1):
            Cm = torch.empty(M, N, device="cuda")
            _lib.check(lib.dpe_debug_gemm(h, path, p(A), K, p(W), p(Cm), N, M, N, K, 0, 0, 0, 0, 0, 0, None))
            torch.cuda.synchronize()
            err = Cm.double() - ref
            scale = ref.abs().max()
            if kind == "positive":
                rel = err / ref
                row.append(f"path {path}: mean rel {rel.mean().item():+.2e} std {rel.std().item():.2e}")
            else:
                row.append(f"path {path}: mean {(err / scale).mean().item():+.2e} rms {((err / scale) ** 2).mean().sqrt().item():.2e} max {(err.abs().max() / scale).item():.2e}"
                           f" sign-corr {(err * ref.sign()).mean().item() / scale.item():+.2e}")
        print(f"K={K:4d} {kind:9s} " + " | ".join(row))
