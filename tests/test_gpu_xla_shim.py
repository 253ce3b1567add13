"""The XLA custom-call entry points (include/dpe_b200.h, csrc/xla_shim.cu) called exactly as XLA's GPU runtime calls a legacy custom
call -- fn(stream, void **buffers, opaque, opaque_len) with a hand-built buffer table -- and compared bit for bit with the C-ABI
entry points the ctypes binding uses."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _table(*tensors):
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


def test_xla_custom_calls_match_the_c_abi():
    import deeperwin_b200 as dpe
    from deeperwin_b200 import _lib
    from deeperwin_b200._lib import DpeXlaDescriptor, MODE_LAPLACIAN
    lib = _lib.load()
    cfg = dpe.Configuration(physical=dict(name="LiH"))
    phys = cfg.physical
    f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=3, device="cuda:0")
    eng = f.engine
    st = dpe.MCMCState.initialize_around_nuclei(96, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(9), device="cuda:0")
    phase0, lp0 = f(params, 2, 2, st.r, st.R, st.Z, fixed)               # also uploads parameters and geometry into the handle
    e0 = eng.local_energy(st.r)
    B = 96
    ws = eng.workspace(B, MODE_LAPLACIAN)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    desc = DpeXlaDescriptor(model=eng.handle.value, workspace_bytes=ws.numel(), n_walkers=B, n_steps=0, recompute_log_psi=0, run_controller=1)
    opaque = bytes(desc)

    phase, lp = torch.empty(B, device="cuda"), torch.empty(B, device="cuda")
    lib.dpe_xla_log_psi_sqr(stream, _table(st.r, ws, phase, lp), opaque, len(opaque))
    assert lib.dpe_xla_last_status() == 0
    assert torch.equal(lp, lp0) and torch.equal(phase, phase0)

    e = torch.empty(B, device="cuda")
    lib.dpe_xla_local_energy(stream, _table(st.r, ws, e), opaque, len(opaque))
    assert lib.dpe_xla_last_status() == 0 and torch.equal(e, e0)

    # Metropolis steps: distinct operand / result buffers as XLA allocates them
    mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=5, stepsize_update_interval=2, initialization="gaussian"))
    ref = mc.run_inter_steps(f, st, params, 2, 2, fixed)
    desc = DpeXlaDescriptor(model=eng.handle.value, workspace_bytes=ws.numel(), n_walkers=B, n_steps=5, recompute_log_psi=1, run_controller=1, mcmc=mc._cfg)
    opaque = bytes(desc)
    ins = [st.r.contiguous(), st.log_psi_sqr.contiguous(), st.walker_age.contiguous(), st.rng_state.contiguous(), st.stepsize.reshape(1).clone(),
           st.step_nr.reshape(1).clone(), st.acc_rate.reshape(1).clone()]
    keep = [t.clone() for t in ins]
    outs = [torch.empty_like(t) for t in ins] + [torch.zeros(5, dtype=torch.int32, device="cuda")]
    lib.dpe_xla_mcmc_steps(stream, _table(*ins, ws, *outs), opaque, len(opaque))
    torch.cuda.synchronize()
    assert lib.dpe_xla_last_status() == 0
    for a, b in zip(ins, keep):                                           # operands untouched (XLA buffers are immutable values)
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    assert torch.equal(outs[0], ref.r) and torch.equal(outs[1], ref.log_psi_sqr) and torch.equal(outs[2], ref.walker_age)
    assert torch.equal(outs[3].view(torch.int32), ref.rng_state.view(torch.int32))
    assert outs[4].item() == ref.stepsize.item() and int(outs[5]) == int(ref.step_nr) == 5 and outs[6].item() == ref.acc_rate.item()
    assert torch.equal(outs[7], mc.last_accept_counts)

    # a descriptor of the wrong size is refused and latched; a failing call poisons its results with NaN
    lib.dpe_xla_local_energy(stream, _table(st.r, ws, e), opaque[:-4], len(opaque) - 4)
    assert lib.dpe_xla_last_status() == -1 and lib.dpe_xla_last_status() == 0
    bad = DpeXlaDescriptor(model=eng.handle.value, workspace_bytes=64, n_walkers=B)        # workspace too small for one walker
    lib.dpe_xla_local_energy(stream, _table(st.r, ws, e), bytes(bad), len(bytes(bad)))
    torch.cuda.synchronize()
    assert lib.dpe_xla_last_status() == -3 and torch.isnan(e).all()
