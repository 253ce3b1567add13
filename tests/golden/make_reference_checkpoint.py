"""Writes tests/golden/reference_chkpt.zip WITH THE REFERENCE'S OWN save_run (checkpoints.py:49-66, executed under tests/ref_shim):
a LiH checkpoint (small model: 2 iterations, widths 16 / 4 / 8, 3 determinants) holding params (haiku tree of numpy arrays, as loggers.py:507 stores them), an MCMCState of the reference's dataclass,
clipping state, metadata, history and summary.  config.yml is absent: the reference writes it with ruamel.yaml, which is not installed here
(reading the reference's YAML files is covered by tests/test_reference_pin.py::test_reference_yaml_files_load).

    python tests/golden/make_reference_checkpoint.py          (in the container that has /root/reference)"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parents[1]))
import ref_shim  # noqa: E402

ref_shim.install()
import deeperwin.checkpoints as rchk  # noqa: E402
import deeperwin.configuration as rcfg  # noqa: E402
import deeperwin.mcmc as rmcmc  # noqa: E402

import deeperwin_b200 as dpe  # noqa: E402
from oracle import model as om, threefry  # noqa: E402

phys = dpe.PhysicalConfig(name="LiH")
SMALL = dict(n_iterations=2, n_hidden_one_el=[16, 16], n_hidden_two_el=[4], emb_dim=8, n_dets=3)      # the small model of tests/test_gpu_parity.py
d = om.ModelDims(n_el=4, n_up=2, n_ion=2, Z_max=3, **SMALL)
p32 = om.cast_params(om.init_params(d, seed=31, bias_scale=0.1, envelope_jitter=0.5), torch.float32)
params = {m: {k: v.numpy() for k, v in l.items()} for m, l in p32.items()}
ref_phys = rcfg.PhysicalConfig(name=None, R=phys.R, Z=phys.Z, n_electrons=4, n_up=2, el_ion_mapping=phys.el_ion_mapping)
st = rmcmc.MCMCState.initialize_around_nuclei(24, ref_phys, "gaussian", "el_ion_mapping", threefry.prng_key(99))
state = rmcmc.MCMCState(r=st.r.float().numpy(), R=st.R.float().numpy(), Z=np.asarray(st.Z, np.int32), log_psi_sqr=st.log_psi_sqr.float().numpy(),
                        walker_age=np.arange(24, dtype=np.int32) % 3, rng_state=np.asarray(st.rng_state, np.uint32),
                        stepsize=np.float32(0.37), step_nr=np.int32(120), acc_rate=np.float32(0.52))
data = rchk.RunData(config=None, metadata=dict(n_epochs=120, code_version="fixture"), history=[dict(opt_epoch=0, opt_E_mean=-7.9), dict(opt_epoch=1, opt_E_mean=-8.0)],
                    summary=dict(E_mean=-8.05, E_mean_sigma=0.002), params=params, fixed_params={}, mcmc_state=state,
                    clipping_state=(np.float32(-8.01), np.float32(0.7)))
rchk.save_run(str(HERE / "reference_chkpt.zip"), data)
print("written", HERE / "reference_chkpt.zip")
