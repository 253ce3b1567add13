"""ORACLE (test infrastructure, never the product path): numpy restatement of the
jax 0.4.23 threefry2x32 PRNG as DeepErwin's MCMC uses it.

The algorithm lives in a third-party dependency that is NOT under /root/reference
(jax==0.4.23 / jaxlib==0.4.23, pinned in /root/reference/uv.lock:492-494,513-515).
Reference call sites: src/deeperwin/mcmc.py:43,90,171,178-179,360-361 and
src/deeperwin/utils/utils.py:115 (batch_rng_split).

Parity status: PINNED by known-answer vectors (Random123 threefry2x32-20 KATs and
upstream-JAX facts listed in SURVEY.md section 8 row a19); see
tests/golden/threefry_kat.json and tests/test_oracle_threefry.py.  The float `normal`
goes through XLA's f32 erf_inv polynomial (Giles); that polynomial is restated from
memory and only pinned by normal(PRNGKey(0))=-0.20584226, normal(PRNGKey(42))=-0.18471177.
"""
import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_U32 = np.uint32


def _rotl(x, d):
    return ((x << _U32(d)) | (x >> _U32(32 - d))).astype(np.uint32)


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds. All args uint32 arrays (broadcastable). Returns (y0, y1)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=np.uint32)
        k1 = np.asarray(k1, dtype=np.uint32)
        x0 = np.asarray(x0, dtype=np.uint32).copy()
        x1 = np.asarray(x1, dtype=np.uint32).copy()
        ks = (k0, k1, (k0 ^ k1 ^ _U32(0x1BD11BDA)).astype(np.uint32))
        x0 = (x0 + ks[0]).astype(np.uint32)
        x1 = (x1 + ks[1]).astype(np.uint32)
        for g in range(5):
            for rot in _ROT[g % 2]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = _rotl(x1, rot)
                x1 = (x1 ^ x0).astype(np.uint32)
            x0 = (x0 + ks[(g + 1) % 3]).astype(np.uint32)
            x1 = (x1 + ks[(g + 2) % 3] + _U32(g + 1)).astype(np.uint32)
    return x0, x1


def prng_key(seed):
    """jax.random.PRNGKey(seed) -> uint32[2] = [hi32, lo32]."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=np.uint32)


def random_bits(key, n):
    """jax `_threefry_random_bits` for 32-bit output: counters iota(n), odd n padded with one 0,
    first half -> x0, second half -> x1, outputs concatenated and truncated to n."""
    key = np.asarray(key, dtype=np.uint32)
    c = np.arange(n, dtype=np.uint32)
    if n % 2:
        c = np.concatenate([c, np.zeros(1, np.uint32)])
    h = len(c) // 2
    y0, y1 = threefry2x32(key[0], key[1], c[:h], c[h:])
    return np.concatenate([y0, y1])[:n]


def split(key, num=2):
    """jax.random.split(key, num) -> uint32[num, 2]."""
    return random_bits(key, 2 * num).reshape(num, 2)


def _bits_to_unit_float(bits):
    return ((bits >> _U32(9)) | _U32(0x3F800000)).astype(np.uint32).view(np.float32) - np.float32(1.0)


def uniform(key, shape=(), minval=0.0, maxval=1.0):
    """jax.random.uniform(key, shape, float32, minval, maxval)."""
    n = int(np.prod(shape)) if len(shape) else 1
    f = _bits_to_unit_float(random_bits(key, n))
    lo, hi = np.float32(minval), np.float32(maxval)
    out = np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)
    return out.reshape(shape)


_ERFINV_SMALL = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
                 -0.00125372503, -0.00417768164, 0.246640727, 1.50140941]
_ERFINV_LARGE = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
                 -0.0076224613, 0.00943887047, 1.00167406, 2.83297682]


def erf_inv_f32(x):
    """XLA's float32 ErfInv (Giles' single-precision polynomial), all arithmetic in float32."""
    x = np.asarray(x, dtype=np.float32)
    w = (-np.log1p((-x * x).astype(np.float32))).astype(np.float32)
    small = w < np.float32(5.0)
    ws = (w - np.float32(2.5)).astype(np.float32)
    wl = (np.sqrt(np.maximum(w, np.float32(5.0))) - np.float32(3.0)).astype(np.float32)
    ww = np.where(small, ws, wl).astype(np.float32)
    p = np.where(small, np.float32(_ERFINV_SMALL[0]), np.float32(_ERFINV_LARGE[0])).astype(np.float32)
    for cs, cl in zip(_ERFINV_SMALL[1:], _ERFINV_LARGE[1:]):
        p = (np.where(small, np.float32(cs), np.float32(cl)) + p * ww).astype(np.float32)
    return (p * x).astype(np.float32)


def normal(key, shape=()):
    """jax.random.normal(key, shape, float32) = sqrt(2) * erf_inv(uniform(lo=nextafter(-1,0), hi=1))."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = uniform(key, shape, lo, 1.0)
    return (np.float32(np.sqrt(2.0)) * erf_inv_f32(u)).astype(np.float32)


def cauchy(key, shape=()):
    """jax.random.cauchy(key, shape, float32) (jax 0.4.23 `_cauchy`): tan(pi * (uniform(minval=eps, maxval=1) - 0.5)), float32."""
    eps = np.finfo(np.float32).eps
    u = uniform(key, shape, eps, 1.0)
    return np.tan((np.float32(np.pi) * (u - np.float32(0.5))).astype(np.float32)).astype(np.float32)


def mcmc_step_randoms(keys, n_el, proposal="normal"):
    """Per-walker randoms of one Metropolis step, exactly as mcmc.py:175-180 + :360-361 consume them.

    keys: uint32[B, 2].  Returns (new_keys[B,2], noise[B,n_el,3] f32, thr[B] f32).
    `split(key)` -> (new_key, sub); noise = normal(sub,[n_el,3]); thr = uniform(sub,()) -- the SAME subkey.
    proposal "cauchy" (mcmc.py:196-201): cauchy noise of the same shape; "normal_one_el" (mcmc.py:183-193) and
    "local_one_el" (:231-253): noise[B, 3] only; "local" (:212-228) and "langevin" (:256-284): normal noise of the full shape.
    """
    keys = np.asarray(keys, dtype=np.uint32)
    B = keys.shape[0]
    new_keys = np.empty((B, 2), np.uint32)
    shape = (3,) if proposal in ("normal_one_el", "local_one_el") else (n_el, 3)
    noise = np.empty((B,) + shape, np.float32)
    thr = np.empty((B,), np.float32)
    for b in range(B):
        ks = split(keys[b], 2)
        new_keys[b] = ks[0]
        noise[b] = cauchy(ks[1], shape) if proposal == "cauchy" else normal(ks[1], shape)
        thr[b] = uniform(ks[1], ())
    return new_keys, noise, thr
