"""Writes tests/golden/reference_*.npz: OUTPUTS OF THE REFERENCE ITSELF (the unmodified /root/reference/src/deeperwin modules executed
under tests/ref_shim in float64) for fp32 weights / walker positions, so that the GPU box -- where /root/reference does not exist --
can compare the CUDA path with the reference's own numbers (tests/test_gpu_parity.py::test_reference_fixtures).

    python tests/golden/make_reference_golden.py          (in the container that has /root/reference)

Per system: weights = oracle.model.init_params(dims, seed, bias_scale, envelope_jitter) cast to fp32 (a checksum is stored),
walkers r (fp32), and from the reference: log_psi_sqr + phase (model/wavefunction.py Wavefunction via hk.multi_transform),
E_loc (hamiltonian.py build_local_energy, forward_lap=False branch), E_pot (get_potential_energy); for LiH also a Metropolis chain
(mcmc.py MetropolisHastingsMonteCarlo._run_mcmc_steps, 8 steps) started from the reference's own initialize_around_nuclei."""
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parents[1]))
import ref_shim  # noqa: E402

ref_shim.install()
import haiku as hk  # noqa: E402
import deeperwin.configuration as rcfg  # noqa: E402
import deeperwin.hamiltonian as rham  # noqa: E402
import deeperwin.mcmc as rmcmc  # noqa: E402
import deeperwin.model.wavefunction as rwf  # noqa: E402
from deeperwin.model.definitions import WavefunctionDefinition  # noqa: E402

import deeperwin_b200 as dpe  # noqa: E402
from oracle import model as om, threefry  # noqa: E402

SMALL = dict(n_iterations=2, n_hidden_one_el=[16, 16], n_hidden_two_el=[4], emb_dim=8, n_dets=3)
SMALL_CFG = dict(embedding=dict(n_iterations=2, n_hidden_one_el=[16, 16], n_hidden_two_el=[4], n_hidden_el_ions=[4], emb_dim=8),
                 orbitals=dict(n_determinants=3))


def checksum(p32):
    return float(sum(v.double().abs().sum() for l in p32.values() for v in l.values()))


def build(phys, small):
    cfg = rcfg.ModelConfigDeepErwin4(**(SMALL_CFG if small else {}))
    model = hk.multi_transform(lambda: rwf.Wavefunction(cfg, WavefunctionDefinition(Z_max=max(phys.Z), Z_min=1)).init_for_multitransform())
    return lambda params, n_up, n_dn, *batch: model.apply[0](params, None, n_up, n_dn, *batch)


for name, small, B, seed in (("LiH", False, 16, 21), ("N2", False, 8, 22), ("LiH", True, 16, 23), ("B", True, 12, 24), ("Ethene", True, 6, 25)):
    phys = dpe.PhysicalConfig(name=name)
    d = om.ModelDims(n_el=phys.n_electrons, n_up=phys.n_up, n_ion=len(phys.Z), Z_max=max(phys.Z), **(SMALL if small else {}))
    p32 = om.cast_params(om.init_params(d, seed=seed, bias_scale=0.1, envelope_jitter=0.5), torch.float32)
    p64 = om.cast_params(p32, torch.float64)
    R32 = torch.tensor(phys.R, dtype=torch.float32)
    g = torch.Generator().manual_seed(seed + 100)
    r32 = (R32[torch.tensor(phys.el_ion_mapping)][None] + torch.randn(B, d.n_el, 3, generator=g)).float()
    r, R, Z = r32.double(), R32.double(), torch.tensor(phys.Z)
    log_psi_sqr = build(phys, small)
    phase, lp = log_psi_sqr(p64, phys.n_up, phys.n_dn, r, R, Z, {})
    e_loc = rham.build_local_energy(log_psi_sqr, forward_lap=False)(p64, (phys.n_up, phys.n_dn), r, R, Z.double(), {})
    e_pot = rham.get_potential_energy(r, R, Z.double())
    out = dict(name=name, small=small, seed=seed, bias_scale=0.1, envelope_jitter=0.5, param_checksum=checksum(p32), r=r32.numpy(),
               R=R32.numpy(), Z=np.array(phys.Z), n_up=phys.n_up, logpsi2=lp.numpy(), phase=phase.numpy(), E_loc=e_loc.numpy(), E_pot=e_pot.numpy())
    if name == "LiH" and not small:
        ref_phys = rcfg.PhysicalConfig(name=None, R=phys.R, Z=phys.Z, n_electrons=4, n_up=2, el_ion_mapping=phys.el_ion_mapping)
        st0 = rmcmc.MCMCState.initialize_around_nuclei(32, ref_phys, "gaussian", "el_ion_mapping", threefry.prng_key(4321))
        st0.r = st0.r.float().double()
        mc = rmcmc.MetropolisHastingsMonteCarlo(rcfg.MCMCConfigOptimization(n_inter_steps=8, max_age=3, stepsize_update_interval=4))
        st0.stepsize = torch.tensor(np.float32(0.3)).double()
        import copy
        st1 = mc._run_mcmc_steps(log_psi_sqr, copy.copy(st0), p64, 2, 2, {}, 8)
        out.update(mcmc_seed=4321, mcmc_r0=st0.r.float().numpy(), mcmc_keys0=np.asarray(st0.rng_state), mcmc_r=st1.r.numpy(), mcmc_keys=np.asarray(st1.rng_state),
                   mcmc_age=st1.walker_age.numpy(), mcmc_logpsi2=st1.log_psi_sqr.numpy(), mcmc_stepsize=float(st1.stepsize), mcmc_acc_rate=float(st1.acc_rate),
                   mcmc_step_nr=int(st1.step_nr))
    fn = HERE / f"reference_{name}{'_small' if small else ''}.npz"
    np.savez_compressed(fn, **out)
    print(fn.name, "logpsi2", lp[:3].numpy(), "E_loc", e_loc[:3].numpy())
