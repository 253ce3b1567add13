"""GPU check of the dense-layer GEMM paths (0 = FP32 SIMT, 1 = tcgen05 3xTF32) against torch fp64, plus timing."""
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


def run(path, A, W, M, N, K, lda, ldc, seg_len=0, a_ss=0, a_so=0, c_ss=0, c_so=0, c_rows=None, col_off=0):
    Cm = torch.full((c_rows or M, ldc), float("nan"), device="cuda")
    _lib.check(lib.dpe_debug_gemm(h, path, p(A), lda, p(W), p(Cm), ldc, M, N, K, seg_len, a_ss, a_so, c_ss, c_so, col_off, None), f"gemm path {path}")
    torch.cuda.synchronize()
    return Cm


def check(name, M, N, K, lda, seg=None):
    g = torch.Generator(device="cuda").manual_seed(0)
    W = (torch.rand(K, N, device="cuda", generator=g) * 2 - 1) * 0.1
    if seg is None:
        A = torch.randn(M, lda, device="cuda", generator=g)
        ref = (A[:, :K].double() @ W.double())
        outs = {pth: run(pth, A, W, M, N, K, lda, N)[:, :N] for pth in (0, 1)}
    else:
        n_seg, seg_len, stride, off = seg
        A = torch.randn(n_seg * stride, lda, device="cuda", generator=g)
        rows = (torch.arange(n_seg, device="cuda")[:, None] * stride + off + torch.arange(seg_len, device="cuda")[None]).reshape(-1)
        ref_rows = A[rows][:, :K].double() @ W.double()
        outs = {}
        for pth in (0, 1):
            Cm = run(pth, A, W, n_seg * seg_len, N, K, lda, N, seg_len, stride, off, stride, off, c_rows=n_seg * stride)
            outs[pth] = Cm[rows][:, :N]
            untouched = torch.ones(n_seg * stride, dtype=torch.bool, device="cuda"); untouched[rows] = False
            assert torch.isnan(Cm[untouched]).all(), "wrote outside its segment"
        ref = ref_rows
    scale = ref.abs().max()
    errs = {pth: ((o.double() - ref).abs().max() / scale).item() for pth, o in outs.items()}
    print(f"{name:34s} M={M} N={N} K={K}: rel err simt {errs[0]:.2e}  tc {errs[1]:.2e}  nan_tc={int(torch.isnan(outs[1]).sum())}")
    return errs


check("tiny", 64, 256, 32, 32)
check("one tile", 256, 256, 320, 320)
check("ragged rows", 1000, 256, 320, 320)
check("K tail (44 of lda 320)", 500, 256, 44, 320)
check("narrow N=64", 700, 64, 256, 256)
check("bf-like N=448", 3000, 448, 256, 320)
check("segmented up (308 of 616)", 0, 448, 256, 320, seg=(9, 308, 616, 0))
check("segmented dn (308 of 616)", 0, 448, 256, 320, seg=(9, 308, 616, 308))
check("narrow N=32 K=256", 100000, 32, 256, 320)
check("narrow N=32 K=8", 3000, 32, 8, 44)
check("narrow N=8 K=16", 777, 8, 16, 28)
check("narrow N=32 K=320 ragged", 1001, 32, 320, 320)
check("packed segs 7 of 14", 0, 448, 256, 320, seg=(1000, 7, 14, 0))
check("packed segs 7 of 14 (dn)", 0, 448, 256, 320, seg=(1000, 7, 14, 7))
check("packed segs 21 of 42", 0, 1344, 256, 320, seg=(333, 21, 42, 21))
check("many tiles", 200000, 256, 320, 320)

# timing
for M, N, K, lda in ((2_523_136, 256, 320, 320), (1_261_568, 448, 256, 320), (2_523_136, 32, 256, 320)):
    A = torch.randn(M, lda, device="cuda")
    W = torch.randn(K, N, device="cuda") * 0.05
    for pth in (0, 1):
        Cm = torch.empty(M, N, device="cuda")
        args = (h, pth, p(A), lda, p(W), p(Cm), N, M, N, K, 0, 0, 0, 0, 0, 0, None)
        _lib.check(lib.dpe_debug_gemm(*args))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            _lib.check(lib.dpe_debug_gemm(*args))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"timing M={M} N={N} K={K} path {pth}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
