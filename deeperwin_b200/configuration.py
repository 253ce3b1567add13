"""Config surface of the hot path: the subset of DeepErwin's pydantic tree that the VMC inner loop reads.

Mirrors src/deeperwin/configuration.py of the reference (field names, defaults, `extra="forbid"`,
YAML in/out): MLPConfig (:215-226), EmbeddingConfigDeepErwin4 (:277-375, 422-440),
InputFeatureConfigDPE4 (:557-674), EnvelopeOrbitalsConfig / OrbitalsConfigFermiNet (:688-700, 783-806),
ModelConfigDeepErwin4 (:875-890), MCMCConfig* (:952-1049), ClippingConfig (:1052-1058),
PhysicalConfig (:1537-1740), ComputationConfig (:1880-1900).  Options of the reference that select a
different architecture are accepted only at the value the CUDA path implements; anything else raises.
"""
from __future__ import annotations

import pathlib
from typing import List, Literal, Optional, Union

import numpy as np
import yaml
from pydantic import BaseModel, ConfigDict, model_validator


class ConfigBaseclass(BaseModel):
    """configuration.py:40-95: unknown keys raise (extra='forbid')."""
    model_config = ConfigDict(extra="forbid", validate_assignment=False, arbitrary_types_allowed=True)

    def as_dict(self):
        return self.model_dump()

    def save(self, fname):
        with open(fname, "w") as f:
            yaml.safe_dump(self.model_dump(), f, sort_keys=False)


def _only(value, allowed, what):
    if value != allowed:
        raise NotImplementedError(f"{what}={value!r}: the B200 hot path implements only {allowed!r}")


class MLPConfig(ConfigBaseclass):
    activation: Literal["tanh"] = "tanh"
    init_bias_scale: float = 0.0
    init_weights_scale: Literal["fan_in", "fan_out", "fan_avg"] = "fan_avg"
    init_weights_distribution: Literal["normal", "truncated_normal", "uniform"] = "uniform"
    use_residual: bool = False
    use_layer_norm: bool = False

    @model_validator(mode="after")
    def _check(self):
        _only(self.use_layer_norm, False, "mlp.use_layer_norm")
        return self


class InputFeatureConfigDPE4(ConfigBaseclass):
    name: Literal["dpe4"] = "dpe4"
    use_rbf_features: bool = False
    n_rbf_features: int = 0
    use_distance_features: bool = True
    use_el_ion_differences: bool = True
    use_el_el_differences: bool = False
    concatenate_el_ion_features: bool = True
    coordinates: Literal["cartesian"] = "cartesian"
    full_el_el_distance_matrix: bool = True
    n_ion_ion_rbf_features: int = 32
    ion_embed_type: Optional[Literal["lookup"]] = "lookup"
    n_ion_features: int = 32
    log_scale_distances: bool = False
    use_el_spin: bool = False

    @model_validator(mode="after")
    def _check(self):
        _only(self.use_rbf_features, False, "features.use_rbf_features")
        _only(self.use_distance_features, True, "features.use_distance_features")
        _only(self.use_el_ion_differences, True, "features.use_el_ion_differences")
        _only(self.use_el_el_differences, False, "features.use_el_el_differences")
        _only(self.concatenate_el_ion_features, True, "features.concatenate_el_ion_features")
        _only(self.log_scale_distances, False, "features.log_scale_distances")
        _only(self.use_el_spin, False, "features.use_el_spin")
        return self


class EmbeddingConfigDeepErwin4(ConfigBaseclass):
    name: Literal["dpe4"] = "dpe4"
    n_iterations: int = 4
    n_hidden_one_el: Union[List[int], int] = 256
    n_hidden_two_el: Union[List[int], int] = 32
    n_hidden_el_ions: Union[List[int], int] = 32
    use_el_ion_stream: bool = True
    use_h_two_same_diff: bool = True
    emb_dim: int = 32
    use_w_mapping: bool = True
    use_schnet_features: bool = True
    sum_schnet_features: bool = False
    use_average_h_one: bool = True
    use_average_h_two: bool = False
    use_h_one: bool = True
    use_h_one_same_diff: bool = False
    use_linear_out: bool = False
    use_schnet_bias_feat: bool = True
    schnet_aggregation: Literal["sum"] = "sum"
    neighbor_normalization: Literal["sum", "sqrt", "mean"] = "mean"
    use_h_one_mlp: bool = True
    h_one_correlation: int = 0
    use_deep_schnet_feat: bool = False
    use_ln_aft_act: bool = False
    use_ln_bef_act: bool = False

    @model_validator(mode="after")
    def _set_n_hidden(self):
        # configuration.py:352-364
        if isinstance(self.n_hidden_one_el, int):
            self.n_hidden_one_el = [self.n_hidden_one_el] * self.n_iterations
        if isinstance(self.n_hidden_two_el, int):
            self.n_hidden_two_el = [self.n_hidden_two_el] * (self.n_iterations - 1)
        if isinstance(self.n_hidden_el_ions, int):
            self.n_hidden_el_ions = [self.n_hidden_el_ions] * (self.n_iterations - 1)
        if len(self.n_hidden_one_el) != len(self.n_hidden_two_el) + 1:
            raise ValueError("Number of layers for 1-el-stream must be one more than nr of layers in 2-el-stream")
        for k, v in dict(use_el_ion_stream=True, use_h_two_same_diff=True, use_w_mapping=True, use_schnet_features=True,
                         use_average_h_one=True, use_average_h_two=False, use_h_one=True, use_h_one_same_diff=False,
                         use_linear_out=False, use_schnet_bias_feat=True, use_h_one_mlp=True, h_one_correlation=0,
                         use_deep_schnet_feat=False, use_ln_aft_act=False, use_ln_bef_act=False).items():
            _only(getattr(self, k), v, f"embedding.{k}")
        return self


class EnvelopeOrbitalsConfig(ConfigBaseclass):
    envelope_type: Literal["isotropic_exp"] = "isotropic_exp"
    n_hidden: List[int] = []
    use_bias: bool = False
    initialization: Literal["constant", "analytical"] = "constant"

    @model_validator(mode="after")
    def _check(self):
        _only(list(self.n_hidden), [], "orbitals.envelope_orbitals.n_hidden")
        _only(self.use_bias, False, "orbitals.envelope_orbitals.use_bias")
        return self


class TransferableAtomicOrbitalsConfig(ConfigBaseclass):
    """configuration.py:732-775 of the reference.  The hot path evaluates the orbitals from the per-geometry cache
    (backflows, exponents); the widths / depths of the geometry-only nets are accepted and unused here."""
    name: Literal["taos"] = "taos"
    envelope_width: int = 64
    envelope_depth: int = 2
    backflow_width: int = 256
    backflow_depth: int = 2
    symmetrize_exponent_mlp: bool = False
    antisymmetrize_backflow_mlp: bool = False
    use_prefactors: bool = False
    use_exponentials: bool = True
    use_el_ion_embedding: bool = False
    use_squared_envelope_input: bool = False
    use_separate_ion_sum_for_envelopes: bool = False

    @model_validator(mode="after")
    def _supported(self):
        _only(self.use_exponentials, True, "orbitals.transferable_atomic_orbitals.use_exponentials")
        _only(self.use_el_ion_embedding, False, "orbitals.transferable_atomic_orbitals.use_el_ion_embedding")
        _only(self.use_separate_ion_sum_for_envelopes, False, "orbitals.transferable_atomic_orbitals.use_separate_ion_sum_for_envelopes")
        _only(self.use_prefactors, False, "orbitals.transferable_atomic_orbitals.use_prefactors")
        return self


class OrbitalsConfigFermiNet(ConfigBaseclass):
    envelope_orbitals: Optional[EnvelopeOrbitalsConfig] = EnvelopeOrbitalsConfig()
    transferable_atomic_orbitals: Optional[TransferableAtomicOrbitalsConfig] = None
    n_determinants: int = 32

    @model_validator(mode="after")
    def _one_head(self):
        # orbital_net.py:69-96 multiplies the enabled heads; the path implements one at a time
        if (self.envelope_orbitals is None) == (self.transferable_atomic_orbitals is None):
            raise NotImplementedError("exactly one of orbitals.envelope_orbitals / orbitals.transferable_atomic_orbitals must be set")
        return self
    determinant_schema: Literal["full_det"] = "full_det"


class ModelConfigDeepErwin4(ConfigBaseclass):
    name: Literal["dpe4"] = "dpe4"
    features: InputFeatureConfigDPE4 = InputFeatureConfigDPE4()
    embedding: EmbeddingConfigDeepErwin4 = EmbeddingConfigDeepErwin4()
    orbitals: OrbitalsConfigFermiNet = OrbitalsConfigFermiNet()
    mlp: MLPConfig = MLPConfig()
    jastrow: None = None
    use_el_el_cusp_correction: bool = False
    Z_max: Optional[int] = None
    Z_min: Optional[int] = 1
    use_cache: bool = True
    complex_wf: bool = False

    @model_validator(mode="after")
    def _check(self):
        _only(self.use_el_el_cusp_correction, False, "model.use_el_el_cusp_correction")
        _only(self.complex_wf, False, "model.complex_wf")
        return self


class MCMCSimpleProposalConfig(ConfigBaseclass):
    name: Literal["normal", "cauchy", "normal_one_el"] = "normal"     # configuration.py:952-953


class MCMCConfig(ConfigBaseclass):
    n_inter_steps: int
    n_burn_in: int
    max_age: int
    stepsize_update_interval: int
    n_walkers: int = 2048
    spin_initialization: Literal["el_ion_mapping"] = "el_ion_mapping"
    initialization: Literal["gaussian", "exponential"] = "exponential"
    target_acceptance_rate: float = 0.5
    min_stepsize_scale: float = 1e-2
    max_stepsize_scale: float = 1.0
    proposal: MCMCSimpleProposalConfig = MCMCSimpleProposalConfig()
    p_spin_swap: float = 0.0
    p_spin_flip: float = 0.0

    @model_validator(mode="after")
    def _check(self):
        _only(self.p_spin_swap, 0.0, "mcmc.p_spin_swap")
        _only(self.p_spin_flip, 0.0, "mcmc.p_spin_flip")
        return self


class MCMCConfigPreTrain(MCMCConfig):
    n_inter_steps: int = 1
    n_burn_in: int = 0
    stepsize_update_interval: int = 1000
    max_age: int = 20


class MCMCConfigOptimization(MCMCConfig):
    n_inter_steps: int = 20
    n_burn_in: int = 1000
    stepsize_update_interval: int = 100
    max_age: int = 20


class MCMCConfigEvaluation(MCMCConfig):
    n_inter_steps: int = 20
    n_burn_in: int = 500
    stepsize_update_interval: int = 100
    max_age: int = 100


class ClippingConfig(ConfigBaseclass):
    name: Literal["hard", "tanh"] = "tanh"
    width_metric: Literal["std", "mae"] = "std"
    center: Literal["mean", "median"] = "mean"
    from_previous_step: bool = True
    clip_by: float = 5.0
    clip_imag_around_0: bool = False      # complex wavefunctions only (loss_function.py:37-38); unused on the real path


class OptimizationConfig(ConfigBaseclass):
    mcmc: MCMCConfigOptimization = MCMCConfigOptimization()
    clipping: ClippingConfig = ClippingConfig()
    n_epochs: int = 60_000
    forward_lap: bool = True
    max_batch_size: int = 64
    stop_on_nan: bool = True


class EvaluationConfig(ConfigBaseclass):
    mcmc: MCMCConfigEvaluation = MCMCConfigEvaluation()
    n_epochs: int = 0
    forward_lap: bool = True
    max_batch_size: int = 64


class ComputationConfig(ConfigBaseclass):
    float_precision: Literal["float32"] = "float32"
    rng_seed: Optional[int] = None
    n_local_devices: Optional[int] = None
    n_nodes: int = 1
    disable_jit: bool = False
    workspace_gb: float = 48.0
    """(B200 addition) upper bound of the scratch workspace per GPU; batches that need more run in chunks."""


_PERIODIC_TABLE = {k: i + 1 for i, k in enumerate(
    "H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr".split(" "))}
_DEFAULT_MOLECULES = yaml.safe_load(pathlib.Path(__file__).parent.joinpath("molecules.yaml").read_text())


def _spin_from_hunds_rule(Z):
    # configuration.py:1541-1569
    n_orbitals = [1, 1, 3, 1, 3, 1, 5, 3, 1, 5, 3, 1, 7, 5]
    n_electrons, n_up, n_dn = Z, 0, 0
    for n_in_orb in n_orbitals:
        n_up += min(n_in_orb, n_electrons)
        n_electrons -= min(n_in_orb, n_electrons)
        n_dn += min(n_in_orb, n_electrons)
        n_electrons -= min(n_in_orb, n_electrons)
        if n_electrons == 0:
            break
    return n_up - n_dn


class PhysicalConfig(ConfigBaseclass):
    name: Optional[str] = None
    R: Optional[List[List[float]]] = None
    Z: Optional[List[int]] = None
    n_electrons: Optional[int] = None
    n_up: Optional[int] = None
    el_ion_mapping: Optional[List[int]] = None
    E_ref: Optional[float] = None
    E_ref_source: Optional[str] = None
    comment: Optional[str] = None
    periodic: None = None

    @model_validator(mode="after")
    def populate_physical_config_from_name(self):
        # configuration.py:1717-1750
        mol = {}
        if self.name:
            if self.name in _PERIODIC_TABLE:
                Z = _PERIODIC_TABLE[self.name]
                mol = dict(Z=[Z], R=[[0.0, 0.0, 0.0]], spin=_spin_from_hunds_rule(Z))
            elif self.name in _DEFAULT_MOLECULES:
                mol = _DEFAULT_MOLECULES[self.name]
        if self.Z is None:
            self.Z = mol.get("Z")
        if self.R is None:
            self.R = mol.get("R")
        if self.n_electrons is None and self.Z is not None:
            self.n_electrons = sum(self.Z) - mol.get("charge", 0)
        if self.n_up is None and self.n_electrons is not None:
            self.n_up = (self.n_electrons + mol.get("spin", 0) + 1) // 2
        if self.el_ion_mapping is None:
            if "el_ion_mapping" in mol:
                self.el_ion_mapping = list(mol["el_ion_mapping"])
            elif self.Z is not None and self.n_electrons == sum(self.Z):
                # neutral default: fill ions in order, spin-up block first (the reference's greedy
                # local-spin balancing, configuration.py:1571-1615, is setup-time host code)
                per_ion = [[(z + 1) // 2, z // 2] for z in self.Z]
                self.el_ion_mapping = [i for s in range(2) for i, n in enumerate(per_ion) for _ in range(n[s])]
        if self.E_ref is None:
            self.E_ref = mol.get("E_ref")
        if self.E_ref_source is None:
            self.E_ref_source = mol.get("E_ref_source")
        return self

    def get_basic_params(self):
        return self.n_electrons, self.n_up, np.array(self.R), np.array(self.Z)

    @property
    def n_dn(self):
        return self.n_electrons - self.n_up

    @property
    def n_ions(self):
        return len(self.Z)


class Configuration(ConfigBaseclass):
    """Root config (configuration.py:1934-1992), restricted to the sections the hot path reads."""
    physical: Optional[PhysicalConfig] = None
    model: ModelConfigDeepErwin4 = ModelConfigDeepErwin4()
    optimization: OptimizationConfig = OptimizationConfig()
    evaluation: EvaluationConfig = EvaluationConfig()
    computation: ComputationConfig = ComputationConfig()
    experiment_name: Optional[str] = "deeperwin_experiment"
    comment: Optional[str] = None

    @classmethod
    def load_configuration_file(cls, config_file):
        """configuration.py:1988-1992."""
        with open(config_file) as f:
            return cls.model_validate(yaml.safe_load(f) or {})

    @classmethod
    def load(cls, config_file):
        return cls.load_configuration_file(config_file)
