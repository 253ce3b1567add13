"""Oracle Metropolis step and energy statistics (mcmc.py:345-387, loss_function.py:12-109)."""
from pathlib import Path

import numpy as np

from oracle import mcmc as omc, threefry

GOLD = Path(__file__).parent / "golden"
f32 = np.float32


def gaussian_logp(r):
    return (-np.sum(r.astype(f32) ** 2, axis=(1, 2))).astype(f32)


def test_golden_chain():
    g = np.load(GOLD / "mcmc_gaussian.npz")
    B = g["r0"].shape[0]
    st = omc.OracleMCMCState(r=g["r0"], R=np.zeros((1, 3), f32), Z=np.array([4]), log_psi_sqr=-np.ones(B, f32) * 1000,
                             walker_age=np.zeros(B, np.int32), rng_state=g["keys0"], stepsize=f32(0.3))
    out = omc.run_mcmc_steps(gaussian_logp, st, 7, max_age=2, stepsize_update_interval=3)
    assert np.array_equal(out.rng_state, g["keys"]) and np.array_equal(out.walker_age, g["age"])
    assert np.array_equal(out.r, g["r"]) and np.array_equal(out.log_psi_sqr, g["log_psi_sqr"])
    assert out.stepsize == g["stepsize"] and out.acc_rate == g["acc_rate"] and out.step_nr == 7


def test_step_bookkeeping():
    B, N = 32, 3
    keys = threefry.split(threefry.prng_key(3), B)
    r0 = threefry.normal(threefry.prng_key(4), (B, N, 3))
    st = omc.OracleMCMCState(r=r0, R=np.zeros((1, 3), f32), Z=np.array([3]), log_psi_sqr=gaussian_logp(r0),
                             walker_age=np.full(B, 1, np.int32), rng_state=keys, stepsize=f32(0.5))
    new, mask = omc.make_mcmc_step(gaussian_logp, st, max_age=2, stepsize_update_interval=1, return_mask=True)
    assert np.array_equal(new.rng_state, np.stack([threefry.split(k)[0] for k in keys]))
    assert np.array_equal(new.walker_age, np.where(mask, 0, 2))
    assert np.array_equal(new.r[~mask], st.r[~mask]) and not np.array_equal(new.r[mask], st.r[mask])
    assert new.step_nr == 1 and new.acc_rate == f32(f32(0.1) * f32(mask.mean()))
    assert new.stepsize == f32(np.clip(f32(0.5) / f32(1.05), 0.01, 1.0))   # pre-update acc_rate 0 < 0.5 -> shrink
    # forced accept once age >= max_age
    st2 = omc.OracleMCMCState(**{**st.__dict__, "walker_age": np.full(B, 2, np.int32)})
    _, mask2 = omc.make_mcmc_step(gaussian_logp, st2, max_age=2, return_mask=True)
    assert mask2.all()


def test_energy_statistics():
    rng = np.random.default_rng(0)
    E = rng.normal(-10, 1, 257).astype(f32)
    E[5] = np.nan
    loss, (c, w), aux = omc.energy_statistics(E, omc.init_clipping_state())
    assert np.isclose(aux["E_mean"], np.nanmean(E)) and np.isclose(aux["E_var"], np.nanvar(E), rtol=1e-5)
    assert np.isclose(loss, aux["E_mean"], rtol=1e-6)          # width 1e12: tanh clipping is the identity
    assert np.isclose(c, aux["E_mean_clipped"]) and np.isclose(w, 5 * np.sqrt(aux["E_var_clipped"]), rtol=1e-6)
    E2 = E.copy(); E2[7] = 1e4
    _, _, aux2 = omc.energy_statistics(E2, (c, w))
    assert np.nanmax(aux2["E_loc_clipped"]) <= c + w + 1e-3
    _, _, aux3 = omc.energy_statistics(E2, (c, w), name="hard")
    assert np.nanmax(aux3["E_loc_clipped"]) == f32(c + w)


def test_exponential_radial_initialisation():
    """The reference's default walker initialisation (orbitals.py:854-928, utils/utils.py:387-506), restated."""
    # Slater-rule helpers on hand-checked cases
    assert omc.get_electron_configuration(7) == {1: 2, 2: 5}
    assert omc.get_electron_configuration(21) == {1: 2, 2: 8, 3: 9, 4: 2}            # 4s fills before 3d
    assert abs(omc.get_effective_charge(7, 1, 1.0, 0.7) - 6.3) < 1e-12
    assert abs(omc.get_effective_charge(7, 2, 1.0, 0.7) - 2.2) < 1e-12
    assert omc.get_effective_charge(1, 1) == 1
    # radial pdf r^2 exp(-k r): mean radius 3 / k, isotropic directions
    r = omc.generate_exp_distributed(threefry.prng_key(3), [40000], k=2.0)
    rad = np.linalg.norm(r, axis=-1)
    assert abs(rad.mean() - 1.5) < 0.02 and abs((rad ** 2).mean() - 12 / 4.0) < 0.08
    assert np.abs(r.mean(0)).max() < 0.02
    # electrons end up around their mapped nucleus, spin-up block first
    R, Z, mapping = [[0, 0, 0], [2.068, 0, 0]], [7, 7], [0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 1, 1, 1, 1]
    st = omc.initialize_around_nuclei(512, R, Z, mapping, 1234, "exponential", n_up=7)
    assert st.r.shape == (512, 14, 3) and st.r.dtype == np.float32
    d = np.stack([np.linalg.norm(st.r - np.asarray(Rk, np.float32), axis=-1).mean(0) for Rk in R])    # [ion, electron]
    assert (d.argmin(0) == np.asarray(mapping)).all()
    # the 1s electrons sit much closer than the n = 2 shell (exponents 2 * 6.3 vs 2.2)
    assert d[0, 0] < 0.4 < d[0, 2]
    # same seed -> same walkers; keys / ages as for the gaussian initialisation
    st2 = omc.initialize_around_nuclei(512, R, Z, mapping, 1234, "exponential", n_up=7)
    assert np.array_equal(st.r, st2.r) and np.array_equal(st.rng_state, omc.initialize_around_nuclei(512, R, Z, mapping, 1234).rng_state)


def test_proposal_variants():
    """mcmc.py:183-201: cauchy (heavy tails, same shapes / key usage) and normal_one_el (only electron step_nr % n_el moves)."""
    B, N = 64, 4
    keys = threefry.split(threefry.prng_key(5), B)
    nk0, noise0, thr0 = threefry.mcmc_step_randoms(keys, N)
    nk1, noise1, thr1 = threefry.mcmc_step_randoms(keys, N, "cauchy")
    nk2, noise2, thr2 = threefry.mcmc_step_randoms(keys, N, "normal_one_el")
    assert np.array_equal(nk0, nk1) and np.array_equal(nk0, nk2) and np.array_equal(thr0, thr1) and np.array_equal(thr0, thr2)
    assert noise1.shape == (B, N, 3) and noise2.shape == (B, 3)
    # cauchy = tan(pi (u - 1/2)) of the same uniforms that feed the normal's erf_inv: same sign, median |x| = 1
    assert (np.sign(noise1) == np.sign(noise0)).mean() > 0.99
    big = threefry.cauchy(threefry.prng_key(1), (20000,))
    assert abs(np.median(np.abs(big)) - 1.0) < 0.03 and np.abs(big).max() > 100
    # normal(sub, [3]) is NOT a prefix of normal(sub, [N, 3]): jax lays bits out in halves
    assert np.array_equal(noise2[0], threefry.normal(threefry.split(keys[0], 2)[1], (3,)))
    func = lambda r: (-np.sum(r.astype(np.float32) ** 2, axis=(1, 2))).astype(np.float32)
    r0 = threefry.normal(threefry.prng_key(7), (B, N, 3))
    st = omc.OracleMCMCState(r=r0, R=np.zeros((1, 3), np.float32), Z=np.array([4]), log_psi_sqr=func(r0), walker_age=np.zeros(B, np.int32),
                             rng_state=keys, stepsize=np.float32(0.3), step_nr=6)
    out = omc.make_mcmc_step(func, st, proposal="normal_one_el")
    moved = np.abs(out.r - st.r).max(axis=(0, 2)) > 0
    assert moved.tolist() == [False, False, True, False]          # 6 % 4 == 2
    assert out.step_nr == 7
