"""Pins oracle/threefry.py against published known answers (tests/golden/threefry_kat.json)."""
import json
from pathlib import Path

import numpy as np

from oracle import threefry

KAT = json.loads((Path(__file__).parent / "golden" / "threefry_kat.json").read_text())


def test_threefry2x32_random123_kats():
    for v in KAT["threefry2x32"]:
        y0, y1 = threefry.threefry2x32(v["key"][0], v["key"][1], v["ctr"][0], v["ctr"][1])
        assert [int(y0), int(y1)] == v["out"]


def test_jax_facts():
    j = KAT["jax"]
    assert threefry.split(threefry.prng_key(0)).tolist() == j["split_PRNGKey0"]
    assert float(threefry.uniform(threefry.prng_key(0))) == np.float32(j["uniform_PRNGKey0"])
    assert abs(float(threefry.normal(threefry.prng_key(0))) - j["normal_PRNGKey0"]) < 1e-7
    assert abs(float(threefry.normal(threefry.prng_key(42))) - j["normal_PRNGKey42"]) < 1e-7


def test_bits_layout_odd_and_even():
    key = threefry.prng_key(99)
    for n in (1, 2, 5, 12, 13):
        bits = threefry.random_bits(key, n)
        h = (n + 1) // 2
        for p in range(h):
            x1 = h + p if h + p < n else 0
            y0, y1 = threefry.threefry2x32(key[0], key[1], p, x1)
            assert bits[p] == y0
            if h + p < n:
                assert bits[h + p] == y1


def test_mcmc_randoms_reuse_same_subkey():
    """mcmc.py:348 + :360: the subkey that drives the Gaussian proposal also drives the acceptance threshold."""
    keys = threefry.split(threefry.prng_key(5), 3)
    nk, noise, thr = threefry.mcmc_step_randoms(keys, 4)
    for b in range(3):
        new, sub = threefry.split(keys[b], 2)
        assert (nk[b] == new).all()
        assert np.array_equal(noise[b], threefry.normal(sub, (4, 3)))
        assert thr[b] == threefry.uniform(sub, ())
    assert noise.dtype == np.float32 and np.isfinite(noise).all()
