"""TEST INFRASTRUCTURE: runs the UNMODIFIED reference modules (/root/reference/src/deeperwin) without jax.

The reference is pure `jax.numpy` + `haiku` Python; jax / jaxlib / dm-haiku / chex / folx / kfac_jax / pyscf are not installed
in this image and cannot be installed (no network).  `install()` registers minimal stand-ins for those packages in
`sys.modules`, backed by torch in float64 on the CPU:

  jax.numpy   -> the numpy-style functions the reference's hot path calls, on torch tensors (module _jnp below)
  jax         -> grad / jvp / linearize / value_and_grad via torch.func, vmap / pmap as Python loops over the mapped axis,
                 lax.cond / fori_loop as Python control flow, lax.pmean / psum as the identity (one device),
                 random.* on the repository's threefry restatement (oracle/threefry.py, pinned by Random123 / JAX known answers)
  haiku       -> Module name scoping ("wf/~/input/h_ion", ...), Linear, Embed, get_parameter, transform / multi_transform
  chex, folx, kfac_jax, optax, pyscf, ruamel.yaml, h5py, ... -> the few names the import graph touches

Nothing of the reference is copied or edited: the modules are imported from where they lie and executed as they are, so
`log_psi_sqr`, `get_potential_energy`, `get_kinetic_energy`, `MetropolisHastingsMonteCarlo.make_mcmc_step`, `_clip_energies`
... computed under this shim are outputs of the reference's own code.  tests/test_reference_pin.py compares oracle/ against
them and tests/golden/make_reference_golden.py stores such outputs as fixtures for the GPU box (where /root/reference
does not exist).  Only tests import this package.
"""
from __future__ import annotations

import dataclasses
import functools
import importlib
import importlib.abc
import importlib.machinery
import math
import re
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
REFERENCE_SRC = Path("/root/reference/src")
F64 = torch.float64


def available() -> bool:
    return (REFERENCE_SRC / "deeperwin" / "mcmc.py").exists()


# =============================================================================================== pytrees
def _is_dataclass_instance(x):
    return dataclasses.is_dataclass(x) and not isinstance(x, type)


def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def tree_flatten(tree):
    """Leaves + a rebuild function. Containers: dict, list, tuple, namedtuple, dataclass; None is an empty subtree (as in jax)."""
    if tree is None:
        return [], lambda leaves: None
    if isinstance(tree, dict):
        keys = list(tree.keys())
        subs = [tree_flatten(tree[k]) for k in keys]
    elif _is_namedtuple(tree):
        keys, subs = None, [tree_flatten(v) for v in tree]
    elif isinstance(tree, (list, tuple)):
        keys, subs = None, [tree_flatten(v) for v in tree]
    elif _is_dataclass_instance(tree):
        keys = [f.name for f in dataclasses.fields(tree)]
        subs = [tree_flatten(getattr(tree, k)) for k in keys]
    else:
        return [tree], lambda leaves: leaves[0]
    counts = [len(s[0]) for s in subs]
    leaves = [l for s in subs for l in s[0]]

    def rebuild(new_leaves):
        out, pos = [], 0
        for (_, rb), n in zip(subs, counts):
            out.append(rb(new_leaves[pos:pos + n]))
            pos += n
        if isinstance(tree, dict):
            return {k: v for k, v in zip(keys, out)}
        if _is_namedtuple(tree):
            return type(tree)(*out)
        if isinstance(tree, list):
            return out
        if isinstance(tree, tuple):
            return tuple(out)
        return type(tree)(**dict(zip(keys, out)))

    return leaves, rebuild


def tree_map(f, tree, *rest):
    leaves, rebuild = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return rebuild([f(l, *[o[i] for o in others]) for i, l in enumerate(leaves)])


def tree_leaves(tree):
    return tree_flatten(tree)[0]


def tree_reduce(f, tree, initializer=None):
    leaves = tree_leaves(tree)
    return functools.reduce(f, leaves) if initializer is None else functools.reduce(f, leaves, initializer)


# =============================================================================================== jax.numpy on torch
def _t(x, dtype=None):
    """Anything -> torch tensor (float64 for floats)."""
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, np.ndarray) and x.dtype == np.uint32:
        return x                                        # PRNG keys stay numpy
    if isinstance(x, (list, tuple)) and any(isinstance(v, torch.Tensor) for v in x):
        return torch.stack([_t(v, dtype) for v in x])
    a = np.asarray(x)
    if dtype is None:
        dtype = F64 if a.dtype.kind == "f" else (torch.int64 if a.dtype.kind in "iu" else (torch.bool if a.dtype.kind == "b" else
                                                                                               (torch.complex128 if a.dtype.kind == "c" else F64)))
    return torch.as_tensor(a).to(dtype)


def _dtype(d):
    if d is None:
        return None
    if isinstance(d, torch.dtype):
        return d
    if d is int or d in (np.int32, np.int64) or str(d) in ("int32", "int64"):
        return torch.int64
    if d is bool or d is np.bool_:
        return torch.bool
    if d is complex or str(d).startswith("complex"):
        return torch.complex128
    return F64                                            # float, float32, float64, jnp.float32 ... -> float64


def _axis_kw(axis):
    if axis is None:
        return {}
    return {"dim": tuple(axis) if isinstance(axis, (list, tuple)) else axis}


def _reduce(fn):
    def f(x, axis=None, keepdims=False, dtype=None):
        x = _t(x)
        if axis is None:
            return fn(x)
        return fn(x, dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis, keepdim=keepdims)
    return f


def _minmax(fn_all, fn_dim):
    def f(x, axis=None, keepdims=False):
        x = _t(x)
        if axis is None:
            return fn_all(x)
        return fn_dim(x, dim=axis, keepdim=keepdims)
    return f


def _unary(fn):
    return lambda x: fn(_t(x))


def _nan_reduce(kind):
    def f(x, axis=None, keepdims=False):
        x = _t(x)
        if kind == "mean":
            return torch.nanmean(x) if axis is None else torch.nanmean(x, dim=axis, keepdim=keepdims)
        if axis is None:
            v = x.flatten()
            v = v[~torch.isnan(v)]
            return torch.quantile(v, 0.5) if v.numel() else x.new_tensor(float("nan"))
        raise NotImplementedError
    return f


class _Linalg:
    @staticmethod
    def norm(x, ord=None, axis=None, keepdims=False):
        x = _t(x)
        if axis is None:
            return torch.linalg.norm(x.flatten())
        return torch.linalg.norm(x, dim=axis, keepdim=keepdims)

    @staticmethod
    def slogdet(a):
        s, l = torch.linalg.slogdet(_t(a))
        return s, l

    inv = staticmethod(lambda a: torch.linalg.inv(_t(a)))
    det = staticmethod(lambda a: torch.linalg.det(_t(a)))


def _make_jnp():
    m = types.ModuleType("jax.numpy")
    m.ndarray = torch.Tensor
    m.float32 = m.float64 = F64
    m.int32 = m.int64 = torch.int64
    m.complex64 = m.complex128 = torch.complex128
    m.bool_ = torch.bool
    m.pi, m.newaxis, m.inf, m.nan, m.e = math.pi, None, math.inf, math.nan, math.e
    m.linalg = _Linalg
    m.array = m.asarray = lambda x, dtype=None: _t(x, _dtype(dtype))
    m.zeros = lambda shape, dtype=None: torch.zeros(tuple(shape) if isinstance(shape, (list, tuple, torch.Size)) else (shape,), dtype=_dtype(dtype) or F64)
    m.ones = lambda shape, dtype=None: torch.ones(tuple(shape) if isinstance(shape, (list, tuple, torch.Size)) else (shape,), dtype=_dtype(dtype) or F64)
    m.zeros_like = lambda x, dtype=None: torch.zeros_like(_t(x), dtype=_dtype(dtype))
    m.ones_like = lambda x, dtype=None: torch.ones_like(_t(x), dtype=_dtype(dtype))
    m.eye = lambda n, dtype=None: torch.eye(n, dtype=_dtype(dtype) or F64)
    m.arange = lambda *a, dtype=None: torch.arange(*a, dtype=_dtype(dtype))
    m.linspace = lambda a, b, n: torch.linspace(a, b, n, dtype=F64)
    _mean = _reduce(torch.mean)
    m.sum, m.prod = _reduce(torch.sum), _reduce(torch.prod)
    m.mean = lambda x, axis=None, keepdims=False, dtype=None: _mean(_t(x) if _t(x).is_floating_point() or _t(x).is_complex() else _t(x).to(F64), axis, keepdims)
    m.any, m.all = _reduce(torch.any), _reduce(torch.all)
    m.max = _minmax(torch.max, torch.amax)
    m.min = _minmax(torch.min, torch.amin)
    m.argmax = lambda x, axis=None: torch.argmax(_t(x)) if axis is None else torch.argmax(_t(x), dim=axis)
    m.argmin = lambda x, axis=None: torch.argmin(_t(x)) if axis is None else torch.argmin(_t(x), dim=axis)
    m.nanmean, m.nanmedian = _nan_reduce("mean"), _nan_reduce("median")
    for name in ("exp", "log", "tanh", "sqrt", "sin", "cos", "abs", "sign", "log1p", "isnan", "square", "conj", "real", "imag", "angle", "floor"):
        setattr(m, name, _unary(getattr(torch, name)))
    m.concatenate = lambda xs, axis=0: torch.cat([_t(x) for x in xs], dim=axis)
    m.stack = lambda xs, axis=0: torch.stack([_t(x) for x in xs], dim=axis)
    m.append = lambda a, b: torch.cat([_t(a).flatten(), _t(b).flatten()])
    m.tile = lambda x, reps: torch.tile(_t(x), tuple(int(r) for r in reps) if isinstance(reps, (list, tuple, torch.Size)) else (int(reps),))
    m.reshape = lambda x, shape: _t(x).reshape(tuple(shape))
    m.expand_dims = lambda x, axis: torch.unsqueeze(_t(x), axis)
    m.squeeze = lambda x, axis=None: torch.squeeze(_t(x)) if axis is None else torch.squeeze(_t(x), axis)
    m.swapaxes = lambda x, a, b: torch.swapaxes(_t(x), a, b)
    m.moveaxis = lambda x, a, b: torch.moveaxis(_t(x), a, b)
    m.broadcast_to = lambda x, shape: torch.broadcast_to(_t(x), tuple(shape))
    m.shape = lambda x: tuple(_t(x).shape)
    m.clip = lambda x, a_min=None, a_max=None: torch.clamp(_t(x), min=a_min, max=a_max)
    m.where = lambda c, a, b: torch.where(_t(c), _t(a) if not isinstance(a, (int, float)) else a, _t(b) if not isinstance(b, (int, float)) else b)
    m.logical_or = lambda a, b: torch.logical_or(_t(a), _t(b))
    m.logical_and = lambda a, b: torch.logical_and(_t(a), _t(b))
    m.add = lambda a, b: a + b
    m.maximum = lambda a, b: torch.maximum(_t(a), _t(b))
    m.minimum = lambda a, b: torch.minimum(_t(a), _t(b))
    m.dot = lambda a, b: _t(a) @ _t(b)
    m.matmul = lambda a, b: _t(a) @ _t(b)
    m.einsum = lambda eq, *ops: torch.einsum(eq, *[_t(o) for o in ops])
    m.triu = lambda x, k=0: torch.triu(_t(x), diagonal=k)
    m.diag = lambda x: torch.diag(_t(x))
    m.diag_indices = lambda n: (torch.arange(n), torch.arange(n))
    m.iscomplexobj = lambda x: isinstance(x, complex) or (isinstance(x, torch.Tensor) and x.is_complex())
    m.power = lambda a, b: torch.pow(_t(a), b)
    m.cumsum = lambda x, axis=None: torch.cumsum(_t(x), dim=axis)

    def interp(x, xp, fp):
        return torch.from_numpy(np.interp(np.asarray(x), np.asarray(xp), np.asarray(fp)))
    m.interp = interp
    return m


class _At:
    """x.at[idx].set / add / get of jax arrays: functional updates on a clone."""
    def __init__(self, x):
        self.x = x

    def __getitem__(self, idx):
        x = self.x

        class _Ref:
            def set(self, v):
                y = x.clone()
                y[idx] = v
                return y

            def add(self, v):
                y = x.clone()
                y[idx] = y[idx] + v
                return y

            def get(self):
                return x[idx]
        return _Ref()


def _patch_tensor():
    # jax arrays are immutable: `x += y` REBINDS x.  torch's augmented assignment mutates, which would leak through the shallow
    # copies the reference takes of its state (mcmc.py:177 `copy.copy(state)` then `new_state.r += ...`).  Inside this test process
    # the augmented operators therefore return new tensors (explicit in-place methods such as add_ are untouched).
    torch.Tensor.__iadd__ = lambda self, o: self + o
    torch.Tensor.__isub__ = lambda self, o: self - o
    torch.Tensor.__imul__ = lambda self, o: self * o
    torch.Tensor.__itruediv__ = lambda self, o: self / o
    torch.Tensor.at = property(lambda self: _At(self))
    torch.Tensor.astype = lambda self, d: self.to(_dtype(d))
    if not hasattr(torch.Tensor, "block_until_ready"):
        torch.Tensor.block_until_ready = lambda self: self


# =============================================================================================== jax transforms
def _index_axis(x, axis, i):
    if axis is None:
        return x
    if isinstance(x, np.ndarray):
        return np.take(x, i, axis=axis)
    return torch.select(x, axis, i)


def _stack_axis(xs, axis):
    if isinstance(xs[0], np.ndarray):
        return np.stack(xs, axis=axis)
    xs = [x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=F64 if isinstance(x, float) else None) for x in xs]
    return torch.stack(xs, dim=axis)


def _broadcast_axes(axes, tree):
    """in_axes / out_axes spec -> one axis per leaf of `tree`."""
    leaves, _ = tree_flatten(tree)
    if axes is None or isinstance(axes, int):
        return [axes] * len(leaves)
    # a pytree prefix: containers of the spec are matched against the value's containers
    out = []

    def rec(spec, val):
        if spec is None or isinstance(spec, int):
            out.extend([spec] * len(tree_flatten(val)[0]))
        elif isinstance(spec, dict):
            for k in val:
                rec(spec[k], val[k])
        elif _is_dataclass_instance(spec):
            for f in dataclasses.fields(spec):
                rec(getattr(spec, f.name), getattr(val, f.name))
        else:
            for s, v in zip(spec, val):
                rec(s, v)
    rec(axes, tree)
    return out


def vmap(fun, in_axes=0, out_axes=0, axis_name=None, **_):
    """jax.vmap as a Python loop: slices every mapped leaf, calls `fun`, stacks the results."""
    @functools.wraps(fun)
    def mapped(*args, **kwargs):                   # keyword arguments are broadcast (jax maps them over axis 0; the reference passes configs only)
        axes_per_arg = in_axes if isinstance(in_axes, (tuple, list)) and len(in_axes) == len(args) and not _is_dataclass_instance(in_axes) else [in_axes] * len(args)
        flat, rebuilds, leaf_axes = [], [], []
        for a, ax in zip(args, axes_per_arg):
            leaves, rb = tree_flatten(a)
            flat.append(leaves)
            rebuilds.append(rb)
            leaf_axes.append(_broadcast_axes(ax, a))
        n = None
        for leaves, axs in zip(flat, leaf_axes):
            for l, ax in zip(leaves, axs):
                if ax is not None:
                    n = l.shape[ax]
                    break
            if n is not None:
                break
        outs = []
        for i in range(n):
            call_args = [rb([_index_axis(l, ax, i) for l, ax in zip(leaves, axs)]) for leaves, axs, rb in zip(flat, leaf_axes, rebuilds)]
            outs.append(fun(*call_args, **kwargs))
        out_leaves0, out_rb = tree_flatten(outs[0])
        o_axes = _broadcast_axes(out_axes, outs[0]) if not (isinstance(out_axes, (tuple, list)) and isinstance(outs[0], tuple) and len(out_axes) == len(outs[0])) else \
            [ax for spec, val in zip(out_axes, outs[0]) for ax in _broadcast_axes(spec, val)]
        all_leaves = [tree_flatten(o)[0] for o in outs]
        stacked = []
        for j, ax in enumerate(o_axes):
            stacked.append(all_leaves[0][j] if ax is None else _stack_axis([al[j] for al in all_leaves], ax))
        return out_rb(stacked)
    return mapped


def pmap(fun=None, axis_name=None, static_broadcasted_argnums=(), **_):
    """One device: maps over the leading axis (of size 1) of every non-static argument."""
    if fun is None:
        return functools.partial(pmap, axis_name=axis_name, static_broadcasted_argnums=static_broadcasted_argnums)
    static = set(static_broadcasted_argnums if isinstance(static_broadcasted_argnums, (tuple, list)) else (static_broadcasted_argnums,))

    @functools.wraps(fun)
    def mapped(*args):
        in_axes = tuple(None if i in static else 0 for i in range(len(args)))
        return vmap(fun, in_axes=in_axes)(*args)
    return mapped


def _grad(fun, argnums=0, has_aux=False):
    def g(*args):
        def wrapped(x):
            a = list(args)
            a[argnums] = x
            return fun(*a)
        return torch.func.grad(wrapped, has_aux=has_aux)(args[argnums])
    return g


def _value_and_grad(fun, argnums=0, has_aux=False):
    if isinstance(fun, _custom_jvp):
        # the value from the primal function, the gradient by transposing the recorded jvp rule (linear in the tangent):
        # grad = d/dt rule(primals, t)[tangent of the loss] at t = 0
        def g(*args):
            out = fun.fun(*args)
            nondiff = fun.nondiff_argnums
            diff_idx = [i for i in range(len(args)) if i not in nondiff]
            primals = tuple(args[i] for i in diff_idx)
            leaves, rebuild = tree_flatten(args[argnums])

            def lin(*t_leaves):
                tangents = tuple(rebuild(list(t_leaves)) if i == argnums else None for i in diff_idx)
                _, tangents_out = fun.jvp_rule(*[args[i] for i in nondiff], primals, tangents)
                return tangents_out[0] if has_aux else tangents_out
            grads = torch.func.grad(lin, argnums=tuple(range(len(leaves))))(*[torch.zeros_like(l) for l in leaves])
            return out, rebuild(list(grads))
        g.fun = fun
        return g

    def g(*args):
        def wrapped(x):
            a = list(args)
            a[argnums] = x
            out = fun(*a)
            return (out[0], out) if has_aux else (out, out)
        gr, val = torch.func.grad(wrapped, has_aux=True)(args[argnums])
        return val, gr
    g.fun = fun
    return g


def _jvp(fun, primals, tangents, has_aux=False):
    return torch.func.jvp(fun, tuple(primals), tuple(tangents), has_aux=has_aux)


def _linearize(fun, *primals):
    return fun(*primals), lambda *tangents: torch.func.jvp(fun, tuple(primals), tuple(tangents))[1]


class _custom_jvp:
    """jax.custom_jvp: the primal function, with the jvp rule recorded (never differentiated here)."""
    def __init__(self, fun, nondiff_argnums=()):
        self.fun, self.jvp_rule, self.nondiff_argnums = fun, None, tuple(nondiff_argnums)
        functools.update_wrapper(self, fun)

    def defjvp(self, rule):
        self.jvp_rule = rule
        return rule

    def __call__(self, *a, **k):
        return self.fun(*a, **k)


def _make_random():
    from oracle import threefry
    m = types.ModuleType("jax.random")
    key = lambda k: np.asarray(k, dtype=np.uint32).reshape(2)
    m.PRNGKey = lambda seed: threefry.prng_key(int(seed))
    m.key = m.PRNGKey
    m.split = lambda k, num=2: threefry.split(key(k), int(num))
    m.normal = lambda k, shape=(), dtype=None: torch.from_numpy(np.asarray(threefry.normal(key(k), tuple(shape)), dtype=np.float64))
    m.uniform = lambda k, shape=(), dtype=None, minval=0.0, maxval=1.0: torch.from_numpy(
        np.asarray(threefry.uniform(key(k), tuple(shape), minval, maxval), dtype=np.float64).reshape(tuple(shape)))
    m.cauchy = lambda k, shape=(), dtype=None: torch.from_numpy(np.asarray(threefry.cauchy(key(k), tuple(shape)), dtype=np.float64))

    def _unsupported(*a, **k):
        raise NotImplementedError("jax.random function outside the pinned subset of tests/ref_shim")
    m.randint = m.permutation = m.choice = m.bernoulli = _unsupported
    return m


def _make_jax(jnp):
    jax = types.ModuleType("jax")
    jax.__path__ = []
    jax.numpy = jnp
    jax.Array = torch.Tensor
    jax.vmap, jax.pmap = vmap, pmap
    jax.jit = lambda f=None, **k: (f if f is not None else (lambda g: g))
    jax.grad, jax.value_and_grad, jax.jvp, jax.linearize, jax.custom_jvp = _grad, _value_and_grad, _jvp, _linearize, _custom_jvp
    jax.device_count = jax.local_device_count = jax.process_count = lambda *a: 1
    jax.process_index = lambda *a: 0
    jax.devices = jax.local_devices = lambda *a: ["cpu:0"]
    jax.tree_map, jax.tree_leaves = tree_map, tree_leaves
    tu = types.ModuleType("jax.tree_util")
    tu.tree_map, tu.tree_leaves, tu.tree_reduce = tree_map, tree_leaves, tree_reduce
    tu.tree_flatten = lambda t: (lambda lv_rb: (lv_rb[0], lv_rb[1]))(tree_flatten(t))
    tu.tree_unflatten = lambda rb, leaves: rb(leaves)
    tu.register_pytree_node = lambda *a, **k: None
    tu.register_pytree_node_class = lambda c: c
    jax.tree_util = tu
    lax = types.ModuleType("jax.lax")
    lax.cond = lambda pred, tf, ff, *ops: tf(*ops) if bool(pred) else ff(*ops)

    def fori_loop(lo, hi, body, init):
        val = init
        for i in range(int(lo), int(hi)):
            val = body(i, val)
        return val
    lax.fori_loop = fori_loop
    lax.pmean = lax.psum = lambda x, axis_name=None: x
    lax.axis_index = lambda name: 0
    lax.stop_gradient = lambda x: tree_map(lambda t: t.detach() if isinstance(t, torch.Tensor) else t, x)
    jax.lax = lax
    nn = types.ModuleType("jax.nn")
    nn.softplus = lambda x: torch.nn.functional.softplus(_t(x), beta=1.0, threshold=1e9)
    nn.silu, nn.elu, nn.relu = (lambda x: torch.nn.functional.silu(_t(x))), (lambda x: torch.nn.functional.elu(_t(x))), (lambda x: torch.relu(_t(x)))
    nn.gelu = lambda x, approximate=True: torch.nn.functional.gelu(_t(x), approximate="tanh" if approximate else "none")
    nn.softmax = lambda x, axis=-1: torch.softmax(_t(x), dim=axis)
    nn.one_hot = lambda x, n, dtype=None: torch.nn.functional.one_hot(_t(x).long(), n).to(F64)
    nn.sigmoid = lambda x: torch.sigmoid(_t(x))
    jax.nn = nn
    sp = types.ModuleType("jax.scipy")
    sp.__path__ = []
    special = types.ModuleType("jax.scipy.special")
    special.erfc = lambda x: torch.special.erfc(_t(x))
    sp.special = special
    sp.linalg = types.ModuleType("jax.scipy.linalg")
    jax.scipy = sp
    jax.random = _make_random()
    cfg = types.SimpleNamespace(update=lambda *a, **k: None)
    jax.config = cfg

    def _getattr(k):                                       # names used only in annotations / out-of-path code (jax.core, ...)
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(f"jax.{k}")
    jax.__getattr__ = _getattr
    return jax, {"jax.tree_util": tu, "jax.lax": lax, "jax.nn": nn, "jax.scipy": sp, "jax.scipy.special": special,
                 "jax.scipy.linalg": sp.linalg, "jax.random": jax.random}


# =============================================================================================== haiku
class _Frame:
    def __init__(self, params, rng, init):
        self.params, self.rng, self.init = params, rng, init
        self.module_stack = []           # (module, method name)
        self.counters = {}               # scope -> {base name -> next index}


_frames = []


def _snake(name):
    s = re.sub(r"(.)([A-Z][a-z]+)", r"\1_\2", name)
    return re.sub(r"([a-z0-9])([A-Z])", r"\1_\2", s).lower()


def _wrap_method(name, fn):
    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        if not _frames:
            return fn(self, *a, **k)
        fr = _frames[-1]
        fr.module_stack.append((self, getattr(fn, "_hk_name_like", name)))
        try:
            return fn(self, *a, **k)
        finally:
            fr.module_stack.pop()
    return wrapped


class _ModuleMeta(type):
    def __new__(mcs, cls_name, bases, ns):
        for k, v in list(ns.items()):
            if callable(v) and isinstance(v, types.FunctionType) and (not k.startswith("__") or k == "__call__"):
                ns[k] = _wrap_method(k, v)
        return super().__new__(mcs, cls_name, bases, ns)

    def __call__(cls, *args, **kwargs):
        obj = cls.__new__(cls)
        fr = _frames[-1] if _frames else None
        if fr is not None:
            fr.module_stack.append((obj, "__init__"))
        try:
            cls.__init__(obj, *args, **kwargs)
        finally:
            if fr is not None:
                fr.module_stack.pop()
        if not hasattr(obj, "module_name"):
            raise ValueError(f"{cls.__name__}.__init__ did not call super().__init__()")
        return obj


class Module(metaclass=_ModuleMeta):
    """hk.Module: the full name is fixed at construction time from the module whose method is executing
    (`parent/child`; `parent/~/child` when constructed inside the parent's __init__, `parent/~method/child` inside another
    method unless that method is `name_like("__call__")`), de-duplicated per scope with _1, _2, ..."""

    def __init__(self, name=None):
        base = name if name is not None else _snake(type(self).__name__)
        fr = _frames[-1] if _frames else None
        prefix = ""
        if fr is not None:
            # the innermost executing module other than the one being constructed
            for mod, method in reversed(fr.module_stack):
                if mod is self:
                    continue
                prefix = mod.module_name
                if method == "__init__":
                    prefix += "/~"
                elif method != "__call__":
                    prefix += "/~" + method
                prefix += "/"
                break
            counters = fr.counters.setdefault(prefix, {})
            idx = counters.get(base, 0)
            counters[base] = idx + 1
            if idx and name is None:
                base = f"{base}_{idx}"
        self.module_name = prefix + base
        self.name = base


def get_parameter(name, shape, dtype=None, init=None):
    fr = _frames[-1]
    mod = None
    for m, _ in reversed(fr.module_stack):
        mod = m
        break
    scope = mod.module_name if mod is not None else "~"
    bucket = fr.params.setdefault(scope, {}) if fr.init else fr.params.get(scope, {})
    if name not in bucket:
        if not fr.init:
            raise KeyError(f"parameter {scope}/{name} is missing (have: {sorted(fr.params)})")
        bucket[name] = _t(init(tuple(shape), dtype))
    p = bucket[name]
    if tuple(p.shape) != tuple(shape):
        raise ValueError(f"parameter {scope}/{name}: shape {tuple(p.shape)} != requested {tuple(shape)}")
    return p


class _Init:
    _gen = torch.Generator().manual_seed(0)


def _variance_scaling(scale=1.0, mode="fan_in", distribution="truncated_normal"):
    def init(shape, dtype=None):
        fan_in, fan_out = (shape[0], shape[1]) if len(shape) > 1 else (shape[0], shape[0])
        n = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": 0.5 * (fan_in + fan_out)}[mode]
        if distribution == "uniform":
            return (torch.rand(shape, generator=_Init._gen, dtype=F64) * 2 - 1) * math.sqrt(3.0 * scale / n)
        return torch.randn(shape, generator=_Init._gen, dtype=F64) * math.sqrt(scale / n)
    return init


class Linear(Module):
    def __init__(self, output_size, with_bias=True, w_init=None, b_init=None, name=None):
        super().__init__(name=name)
        self.output_size, self.with_bias, self.w_init, self.b_init = output_size, with_bias, w_init, b_init

    def __call__(self, x):
        x = _t(x)
        w = get_parameter("w", (x.shape[-1], self.output_size), init=self.w_init or _variance_scaling(1.0, "fan_in", "truncated_normal"))
        y = x @ w
        if self.with_bias:
            y = y + get_parameter("b", (self.output_size,), init=self.b_init or (lambda s, d=None: torch.zeros(s, dtype=F64)))
        return y


class EmbedLookupStyle:
    ARRAY_INDEX, ONE_HOT = 1, 2


class Embed(Module):
    def __init__(self, vocab_size=None, embed_dim=None, embedding_matrix=None, w_init=None, lookup_style="ARRAY_INDEX", name=None):
        super().__init__(name=name)
        self.vocab_size, self.embed_dim = vocab_size, embed_dim

    def __call__(self, ids):
        emb = get_parameter("embeddings", (self.vocab_size, self.embed_dim), init=lambda s, d=None: torch.randn(s, generator=_Init._gen, dtype=F64).clamp(-2, 2))
        return torch.nn.functional.one_hot(_t(ids).long(), self.vocab_size).to(F64) @ emb      # lookup_style ONE_HOT


class LayerNorm(Module):
    def __init__(self, *a, name=None, **k):
        super().__init__(name=name)

    def __call__(self, x):
        raise NotImplementedError("hk.LayerNorm is outside the hot path (mlp.use_layer_norm = False)")


class _Transformed:
    def __init__(self, f, multi):
        self.f, self.multi = f, multi

    def _run(self, params, rng, init, select, args, kwargs):
        fr = _Frame(params, rng, init)
        _frames.append(fr)
        try:
            out = self.f()
            if self.multi:
                template, fns = out
                fn = template if select is None else fns[select]
                return fn(*args, **kwargs)
            return out if select is None and not callable(out) else out
        finally:
            _frames.pop()

    def init(self, rng, *args, **kwargs):
        params = {}
        if self.multi:
            self._run(params, rng, True, None, args, kwargs)
        else:
            fr = _Frame(params, rng, True)
            _frames.append(fr)
            try:
                self.f(*args, **kwargs)
            finally:
                _frames.pop()
        return params

    @property
    def apply(self):
        if self.multi:
            class _Apply:
                def __getitem__(_, i):
                    return lambda params, rng, *a, **k: self._run(params, rng, False, i, a, k)
            return _Apply()

        def apply(params, rng, *a, **k):
            fr = _Frame(params, rng, False)
            _frames.append(fr)
            try:
                return self.f(*a, **k)
            finally:
                _frames.pop()
        return apply


def _name_like(method_name):
    def deco(fn):
        fn._hk_name_like = method_name
        return fn
    return deco


def _make_haiku():
    hk = types.ModuleType("haiku")
    hk.__path__ = []
    hk.Module, hk.Linear, hk.Embed, hk.EmbedLookupStyle, hk.LayerNorm, hk.get_parameter = Module, Linear, Embed, EmbedLookupStyle, LayerNorm, get_parameter
    hk.transform = lambda f: _Transformed(f, False)
    hk.multi_transform = lambda f: _Transformed(f, True)
    hk.without_apply_rng = lambda t: t
    hk.next_rng_key = lambda: np.zeros(2, np.uint32)
    init = types.ModuleType("haiku.initializers")
    init.VarianceScaling = _variance_scaling
    init.TruncatedNormal = lambda stddev=1.0, mean=0.0: (lambda s, d=None: mean + stddev * torch.randn(s, generator=_Init._gen, dtype=F64).clamp(-2, 2))
    init.RandomNormal = lambda stddev=1.0, mean=0.0: (lambda s, d=None: mean + stddev * torch.randn(s, generator=_Init._gen, dtype=F64))
    init.Constant = lambda c: (lambda s, d=None: torch.full(s, float(c), dtype=F64))
    init.Initializer = object
    hk.initializers = init
    ds = types.ModuleType("haiku.data_structures")
    ds.tree_size = lambda p: sum(int(l.numel()) for l in tree_leaves(p))
    ds.traverse = lambda p: ((m, n, v) for m, leaves in p.items() for n, v in leaves.items())

    def partition(pred, params):
        a, b = {}, {}
        for m, leaves in params.items():
            for n, v in leaves.items():
                (a if pred(m, n, v) else b).setdefault(m, {})[n] = v
        return a, b
    ds.partition = partition
    ds.merge = lambda *ps: {m: {**{k: v for p in ps for k, v in p.get(m, {}).items()}} for m in {m for p in ps for m in p}}
    hk.data_structures = ds
    exp = types.ModuleType("haiku.experimental")
    exp.name_like = _name_like
    hk.experimental = exp

    def _getattr(k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(f"haiku.{k}")
    hk.__getattr__ = _getattr
    return hk, {"haiku.initializers": init, "haiku.data_structures": ds, "haiku.experimental": exp}


# =============================================================================================== small stand-ins
class _Anything:
    """Attribute / call / subscript sink for packages the import graph names but the hot path never executes."""
    def __init__(self, name="stub"):
        self.__name__ = name

    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(f"{self.__name__}.{k}")

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k and isinstance(a[0], (types.FunctionType, type)):
            return a[0]                      # used as a decorator
        return _Anything(self.__name__ + "()")

    def __getitem__(self, k):
        return _Anything(self.__name__ + "[]")

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())

    def _binop(self, other):
        return _Anything(self.__name__)
    __add__ = __radd__ = __mul__ = __rmul__ = __sub__ = __rsub__ = __or__ = __ror__ = __truediv__ = __rtruediv__ = _binop


def _stub_module(name):
    m = types.ModuleType(name)
    m.__path__ = []
    def _getattr(k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(f"{name}.{k}")
    m.__getattr__ = _getattr
    return m


def _make_chex():
    chex = types.ModuleType("chex")

    def dataclass(cls=None, **kw):
        def wrap(c):
            c = dataclasses.dataclass(c)
            c.replace = lambda self, **ch: dataclasses.replace(self, **ch)
            return c
        return wrap(cls) if cls is not None else wrap
    chex.dataclass = dataclass
    chex.assert_rank = lambda *a, **k: None
    chex.Array = chex.ArrayTree = chex.PRNGKey = object
    return chex


def _make_folx():
    folx = types.ModuleType("folx")

    def forward_laplacian(f, *a, **k):
        raise NotImplementedError("folx is not installed: the reference's forward_lap=False branch (same quantity) is the one run under the shim")

    def batched_vmap(fun=None, in_axes=0, max_batch_size=64, **k):
        if fun is None:
            return functools.partial(batched_vmap, in_axes=in_axes, max_batch_size=max_batch_size)
        return vmap(fun, in_axes=in_axes)
    folx.forward_laplacian, folx.batched_vmap = forward_laplacian, batched_vmap
    return folx


def _make_kfac():
    k = types.ModuleType("kfac_jax")
    k.__path__ = []
    k.register_scale_and_shift = lambda y, x, scale=None, shift=None: y
    k.register_normal_predictive_distribution = lambda *a, **kw: None
    k.register_dense = lambda y, *a, **kw: y
    def _getattr(name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything(f"kfac_jax.{name}")
    k.__getattr__ = _getattr
    return k


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Any (sub)module of these packages imports as an attribute sink: they configure set-up time chemistry / logging / plotting
    that the hot path never executes."""
    ROOTS = ("pyscf", "optax", "ruamel", "h5py", "wandb", "e3nn_jax", "ase", "distrax", "tensorflow_probability",
             "jax", "jaxlib", "haiku", "kfac_jax", "chex", "folx")          # unknown submodules of the stand-ins too

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _stub_module(spec.name)

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Registers the stand-ins and puts the reference's source tree on sys.path. Idempotent."""
    global _installed
    if _installed:
        return
    if str(ROOT) not in sys.path:
        sys.path.insert(0, str(ROOT))
    _patch_tensor()
    jnp = _make_jnp()
    jax, jax_subs = _make_jax(jnp)
    hk, hk_subs = _make_haiku()
    mods = {"jax": jax, "jax.numpy": jnp, **jax_subs, "haiku": hk, **hk_subs, "chex": _make_chex(), "folx": _make_folx(),
            "kfac_jax": _make_kfac()}
    for name in ("kfac_jax._src", "kfac_jax._src.curvature_blocks", "kfac_jax._src.utils", "kfac_jax._src.layers_and_loss_tags",
                 "jax.experimental", "jax.example_libraries", "jax.flatten_util", "jax._src", "jax.interpreters", "wandb"):
        mods[name] = _stub_module(name)
    sys.meta_path.append(_StubFinder())          # after the real finders: installed packages win
    for name, m in mods.items():
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            if parent in mods and not hasattr(mods[parent], child):
                try:
                    setattr(mods[parent], child, m)
                except Exception:
                    pass
    if str(REFERENCE_SRC) not in sys.path:
        sys.path.insert(0, str(REFERENCE_SRC))
    _installed = True


def to_reference_params(params):
    """{module: {leaf: torch tensor}} -> float64 CPU tensors, the reference's haiku tree layout (names are already haiku's)."""
    return {m: {k: v.detach().double().cpu() for k, v in leaves.items()} for m, leaves in params.items()}
