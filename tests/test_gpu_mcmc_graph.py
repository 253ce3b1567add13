"""Graph replay of repeated dpe_mcmc_steps calls gives the same chain, bit for bit, as plain launches."""
import pytest
import torch

pytestmark = pytest.mark.gpu
from test_gpu_parity import _mcmc_setup  # noqa: E402


def _chain(graph, n_calls=6):
    dpe, phys, f, params, fixed, state = _mcmc_setup(B=128)
    f.engine.set_mcmc_graph(graph)
    mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=5, initialization="gaussian"))
    out = []
    for _ in range(n_calls):                  # the allocator hands the same two sets of state buffers back and forth: calls repeat from the third on
        state = mc.run_inter_steps(f, state, params, phys.n_up, phys.n_dn, fixed)
        out.append((state.r.clone(), state.log_psi_sqr.clone(), state.walker_age.clone(), state.rng_state.clone(), float(state.stepsize), int(state.step_nr),
                    float(state.acc_rate), mc.last_accept_counts.clone()))
    return out, f.engine


def test_graph_replay_is_bit_identical_to_eager_launches():
    eager, _ = _chain(False)
    graph, eng = _chain(True)
    assert eng.lib.dpe_get_mcmc_graph(eng.handle) == 1            # capture did not fail (a failure switches the model back to eager)
    for a, b in zip(eager, graph):
        for x, y in zip(a, b):
            assert torch.equal(x, y) if isinstance(x, torch.Tensor) else x == y


def test_in_place_state_replays_and_counts_launches():
    """The bench's pattern: one resident state advanced in place by repeated identical calls."""
    dpe, phys, f, params, fixed, state = _mcmc_setup(B=64)
    eng = f.engine
    mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=4, initialization="gaussian"))
    state = mc.run_inter_steps(f, state, params, phys.n_up, phys.n_dn, fixed)

    def run(graph):
        eng.set_mcmc_graph(graph)
        r, lp, age, rng = state.r.clone(), state.log_psi_sqr.clone(), state.walker_age.clone(), state.rng_state.clone()
        ss, nr, acc = state.stepsize.clone().reshape(1), state.step_nr.clone().reshape(1).int(), state.acc_rate.clone().reshape(1)
        from deeperwin_b200._lib import DpeMcmcState
        s = DpeMcmcState(*[t.data_ptr() for t in (r, lp, age, rng, ss, nr, acc)])
        counts = torch.zeros(4, dtype=torch.int32, device="cuda")
        launches = []
        for _ in range(4):
            l0 = eng.launch_count()
            eng.mcmc_steps(s, 64, 4, mc._cfg, False, True, counts)
            launches.append(eng.launch_count() - l0)
        torch.cuda.synchronize()
        return (r, lp, age, rng, ss, nr, acc, counts), launches

    a, la = run(False)
    b, lb = run(True)
    assert eng.lib.dpe_get_mcmc_graph(eng.handle) == 1            # the capture succeeded (a failure switches the model back to plain launches)
    assert la == lb and len(set(la)) == 1 and la[0] > 100
    for x, y in zip(a, b):
        assert torch.equal(x, y)
