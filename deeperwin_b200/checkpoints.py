"""Reading and writing the reference's checkpoint files (checkpoints.py:25-93, loggers.py:492-520): a BZIP2 zip with `config.yml`,
`history.csv`, `summary.csv` and one pickle per RunData field (`params.pkl`, `mcmc_state.pkl`, `clipping_state.pkl`, ...).

Interop in both directions without the reference (or jax) being installed:
  * `load_run` unpickles with a restricted Unpickler: numpy / builtins pass through, `deeperwin.mcmc.MCMCState` becomes this package's
    MCMCState (torch tensors), pickled jax arrays (`jax._src.array._reconstruct_array`) become numpy arrays, and every other class of the
    reference's environment (kfac_jax optimiser state, haiku containers, configuration classes inside fixed_params) is kept as an `Opaque`
    record instead of being executed;
  * `save_run` writes `mcmc_state.pkl` under the class path `deeperwin.mcmc.MCMCState` with numpy fields, parameters as numpy trees, so that
    the reference's own `load_run` / `load_data_for_reuse` resume from it (tests/test_reference_pin.py reads one back with the reference's code).
Host-side setup code: nothing here is on the hot path."""
from __future__ import annotations

import contextlib
import io
import os
import pickle
import sys
import types
import zipfile
from dataclasses import dataclass, fields
from typing import Any, List, Optional, Union

import numpy as np
import torch
import yaml

from .configuration import Configuration
from .mcmc import MCMCState

_STATE_FIELDS = ("r", "R", "Z", "log_psi_sqr", "walker_age", "rng_state", "stepsize", "step_nr", "acc_rate")


@dataclass
class RunData:
    """checkpoints.py:25-36."""
    config: Optional[Union[Configuration, dict]] = None
    history: Optional[List[dict]] = None
    summary: Optional[dict] = None
    metadata: Optional[dict] = None
    params: Optional[dict] = None
    ema_params: Optional[dict] = None
    fixed_params: Optional[dict] = None
    opt_state: Optional[Any] = None
    mcmc_state: Optional[MCMCState] = None
    clipping_state: Optional[Any] = None


# ------------------------------------------------------------------------------------------------ reading
class Opaque:
    """A class of the reference's environment that is not needed on this path (kfac_jax state, haiku containers, ...): constructor
    arguments and state are kept, nothing is executed."""
    _path = ("", "")

    def __init__(self, *args, **kwargs):
        self.args, self.kwargs, self.state = args, kwargs, None

    def __setstate__(self, state):
        self.state = state

    def __repr__(self):
        return f"Opaque<{'.'.join(self._path)}>"


class _RefMCMCState:
    """Landing pad for deeperwin.mcmc.MCMCState (a plain dataclass: pickled as class + __dict__)."""


def _reconstruct_jax_array(fun, args, arr_state, aval_state):
    """jax._src.array._reconstruct_array without jax: the payload is a pickled numpy array."""
    arr = fun(*args)
    arr.__setstate__(arr_state)
    return arr


_SAFE_BUILTINS = {"dict", "list", "tuple", "set", "frozenset", "int", "float", "complex", "bool", "str", "bytes", "bytearray", "slice", "range", "object"}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        root = module.split(".")[0]
        if (module, name) == ("deeperwin.mcmc", "MCMCState"):
            return _RefMCMCState
        if (module, name) == ("jax._src.array", "_reconstruct_array"):
            return _reconstruct_jax_array
        if root == "numpy" or (module, name) in (("collections", "OrderedDict"), ("copyreg", "_reconstructor"), ("copyreg", "__newobj__")):
            return super().find_class(module, name)
        if module == "builtins" and name in _SAFE_BUILTINS:
            return super().find_class(module, name)
        return type(name, (Opaque,), {"_path": (module, name)})


def _to_numpy(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu()
        return x.view(torch.int32).numpy().view(np.uint32) if x.dtype == torch.uint32 else x.numpy()
    return np.asarray(x)


def _tree_map(f, tree):
    if isinstance(tree, dict):
        return {k: _tree_map(f, v) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(_tree_map(f, v) for v in tree)
    return f(tree)


def mcmc_state_from_reference(obj, device="cuda") -> MCMCState:
    """The reference's pickled MCMCState (numpy / former jax arrays; possibly split across devices, mcmc.py:105-140) -> MCMCState."""
    device = torch.device(device)
    v = {k: getattr(obj, k, None) for k in _STATE_FIELDS}
    r = np.asarray(v["r"], np.float32)
    if r.ndim == 4:                                   # split across devices [n_dev, B / n_dev, ...]: merge (mcmc.py:131-147)
        flat = lambda a: None if a is None else np.asarray(a).reshape((-1,) + np.asarray(a).shape[2:])
        first = lambda a: None if a is None else np.asarray(a)[0]
        v = {"r": flat(v["r"]), "log_psi_sqr": flat(v["log_psi_sqr"]), "walker_age": flat(v["walker_age"]), "rng_state": flat(v["rng_state"]),
             "R": first(v["R"]), "Z": first(v["Z"]), "stepsize": first(v["stepsize"]), "step_nr": first(v["step_nr"]), "acc_rate": first(v["acc_rate"])}

    def t(a, dtype):
        return None if a is None else torch.as_tensor(np.array(a, dtype=dtype, order="C")).to(device)      # (np.ascontiguousarray would turn 0-d into 1-d)

    keys = None
    if v["rng_state"] is not None:
        keys = torch.as_tensor(np.array(v["rng_state"], dtype=np.uint32, order="C").view(np.int32)).to(device).view(torch.uint32)
    return MCMCState(r=t(v["r"], np.float32), R=t(v["R"], np.float32), Z=t(v["Z"], np.int32), log_psi_sqr=t(v["log_psi_sqr"], np.float32),
                     walker_age=t(v["walker_age"], np.int32), rng_state=keys, stepsize=t(v["stepsize"], np.float32),
                     step_nr=t(v["step_nr"], np.int32), acc_rate=t(v["acc_rate"], np.float32))


def load_run(fname, parse_config=True, parse_csv=False, load_pkl=True, device="cuda") -> RunData:
    """checkpoints.py:69-93.  `params` / `ema_params` come back as numpy trees (as the reference stores them), `mcmc_state` as an
    MCMCState on `device`, `clipping_state` as a tuple of floats."""
    data = RunData()
    with zipfile.ZipFile(fname, "r") as zf:
        names = zf.namelist()
        if "config.yml" in names:
            with zf.open("config.yml", "r") as f:
                raw = yaml.safe_load(io.TextIOWrapper(f, encoding="utf-8")) or {}
            data.config = Configuration.model_validate(raw) if parse_config else raw
        for field_name in names:
            key, ext = os.path.splitext(field_name)
            if ext == ".pkl" and load_pkl:
                with zf.open(field_name, "r") as f:
                    value = _Unpickler(f).load()
                if key == "mcmc_state" and isinstance(value, _RefMCMCState):
                    value = mcmc_state_from_reference(value, device)
                elif key in ("params", "ema_params"):
                    value = _tree_map(np.asarray, value) if value is not None else None
                elif key == "clipping_state" and value is not None:
                    value = tuple(float(np.asarray(x)) for x in value)
                setattr(data, key, value)
            elif ext == ".csv" and parse_csv:
                import pandas as pd
                with zf.open(field_name, "r") as f:
                    setattr(data, key, pd.read_csv(f, sep=";"))
    return data


def params_to_torch(params, device="cuda"):
    """A loaded numpy parameter tree -> the torch tree the callables of this package take."""
    return _tree_map(lambda a: torch.as_tensor(np.asarray(a, np.float32)).to(device), params)


# ------------------------------------------------------------------------------------------------ writing
@contextlib.contextmanager
def _reference_mcmc_class():
    """Yields the class object that pickles as `deeperwin.mcmc.MCMCState`: the real one if the reference is importable, else a stand-in
    registered under that path for the duration of the dump (pickle stores only module + name; instances are rebuilt from __dict__)."""
    created = []
    try:
        real = sys.modules.get("deeperwin.mcmc")
        cls = getattr(real, "MCMCState", None) if real is not None else None
        if cls is None:
            for name in ("deeperwin", "deeperwin.mcmc"):
                if name not in sys.modules:
                    sys.modules[name] = types.ModuleType(name)
                    created.append(name)
            cls = type("MCMCState", (), {})
            cls.__module__, cls.__qualname__ = "deeperwin.mcmc", "MCMCState"
            sys.modules["deeperwin.mcmc"].MCMCState = cls
            created.append(("attr", "deeperwin.mcmc", "MCMCState"))
        yield cls
    finally:
        for item in reversed(created):
            if isinstance(item, tuple):
                if item[1] in sys.modules:
                    sys.modules[item[1]].__dict__.pop(item[2], None)
            else:
                sys.modules.pop(item, None)


def write_history(f, history, delim=";"):
    """checkpoints.py:39-46."""
    keys = []
    for h in history:
        keys.extend(k for k in h.keys() if k not in keys)
    f.write((delim.join(str(k) for k in keys) + "\n").encode("utf-8"))
    for h in history:
        f.write((delim.join(str(h.get(k, "")) for k in keys) + "\n").encode("utf-8"))


def save_run(fname, data: RunData):
    """checkpoints.py:49-66, readable by the reference's load_run."""
    with zipfile.ZipFile(fname, "w", zipfile.ZIP_BZIP2) as zf:
        if data.config is not None:
            raw = data.config.model_dump() if hasattr(data.config, "model_dump") else data.config
            with zf.open("config.yml", "w", force_zip64=True) as f:
                f.write(yaml.safe_dump(raw, sort_keys=False).encode("utf-8"))
        if data.history is not None:
            with zf.open("history.csv", "w", force_zip64=True) as f:
                write_history(f, data.history)
        if data.summary is not None:
            with zf.open("summary.csv", "w", force_zip64=True) as f:
                f.write("\n".join(f"{k};{v}" for k, v in data.summary.items()).encode("utf-8"))
        for fld in fields(RunData):
            key, value = fld.name, getattr(data, fld.name)
            if value is None or key in ("config", "history", "summary"):
                continue
            with zf.open(key + ".pkl", "w", force_zip64=True) as f:
                if key == "mcmc_state":
                    state = {k: (None if getattr(value, k) is None else _to_numpy(getattr(value, k))) for k in _STATE_FIELDS}
                    with _reference_mcmc_class() as cls:
                        obj = cls.__new__(cls)                    # rebuilt by the reader as  copyreg.__newobj__(deeperwin.mcmc.MCMCState) + __dict__
                        obj.__dict__.update(state)
                        pickle.dump(obj, f, protocol=4)
                elif key in ("params", "ema_params", "fixed_params"):
                    pickle.dump(_tree_map(lambda x: _to_numpy(x) if isinstance(x, torch.Tensor) else x, value), f, protocol=4)
                elif key == "clipping_state":
                    pickle.dump(tuple(np.float32(float(x)) for x in value), f, protocol=4)
                else:
                    if isinstance(value, Opaque) or any(isinstance(v, Opaque) for v in (value if isinstance(value, (list, tuple)) else [value])):
                        raise NotImplementedError(f"{key}: records of classes that were skipped while loading (Opaque) cannot be written back")
                    pickle.dump(value, f, protocol=4)
