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
from typing import Any, Dict, List, Literal, Optional, Union

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
        # mlp.py:39: the MLPs that do not pass `residual=` explicitly (w_same, w_diff, h_map, h_ion_map, bf_up, bf_dn) take this
        # switch, and with equal in/out widths (32 -> 32 from the second iteration on) it changes the wavefunction
        _only(self.use_residual, False, "mlp.use_residual")
        return self


class InitializationConfig(ConfigBaseclass):
    """configuration.py (embedding.initialization): kept for YAML round trips, the dpe4 layers initialise from `mlp`."""
    bias_scale: float = 0.0
    weight_scale: str = "glorot"
    weight_distribution: str = "uniform"


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
    # the remaining fields of InputFeatureConfig (configuration.py:557-627); they only matter for other feature variants
    r_cut_bessel: float = 5.0
    n_ion_ion_mlp_features: int = 0
    use_el_ion_convolution: bool = False
    init_as_zeros: bool = False
    n_el_el_features: int = 32
    n_el_el_layers: int = 2
    el_el_gating_operation: Literal["rbf", "gauss", "none"] = "none"
    exp_decay_el_el_edge: bool = False
    init_with_el_el_feat: bool = False
    n_el_ion_features: int = 32
    n_el_ion_layers: int = 2
    el_ion_gating_operation: Literal["rbf", "gauss", "none"] = "none"
    exp_decay_el_ion_edge: bool = False
    init_with_el_ion_feat: bool = False
    rmax: int = 5
    max_scale_gauss: float = 8.0
    include_twist: Optional[List[str]] = None
    use_el_el_spin: bool = False

    @model_validator(mode="after")
    def _check(self):
        for k in ("use_el_ion_convolution", "exp_decay_el_el_edge", "init_with_el_el_feat", "exp_decay_el_ion_edge",
                  "init_with_el_ion_feat", "use_el_el_spin"):
            _only(getattr(self, k), False, f"features.{k}")
        _only(self.include_twist, None, "features.include_twist")
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
    initialization: InitializationConfig = InitializationConfig()
    use_layer_norm: bool = False
    use_symmetric_product: bool = True      # these three only act when h_one_correlation > 0 (ferminet_embedding.py:228-236)
    downmap_during_product: bool = True
    one_el_skip_conn: bool = True

    @model_validator(mode="after")
    def _set_n_hidden(self):
        # configuration.py:352-364.  As in the reference the lists may be longer than n_iterations (tests/test_training.yaml has
        # four widths and n_iterations = 1): the embedding loop reads entry i of each list (ferminet_embedding.py:217-262).
        if isinstance(self.n_hidden_one_el, int):
            self.n_hidden_one_el = [self.n_hidden_one_el] * self.n_iterations
        if isinstance(self.n_hidden_two_el, int):
            self.n_hidden_two_el = [self.n_hidden_two_el] * (self.n_iterations - 1)
        if isinstance(self.n_hidden_el_ions, int):
            self.n_hidden_el_ions = [self.n_hidden_el_ions] * (self.n_iterations - 1)
        if len(self.n_hidden_one_el) != len(self.n_hidden_two_el) + 1:
            raise ValueError("Number of layers for 1-el-stream must be one more than nr of layers in 2-el-stream")
        if len(self.n_hidden_one_el) < self.n_iterations:
            raise ValueError("n_hidden_one_el has fewer entries than n_iterations")
        _only(self.use_layer_norm, False, "embedding.use_layer_norm")
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
        # 'analytical' fits the exponents to a PySCF baseline at set-up time (wavefunction.py:329-360): out of the hot-path scope
        _only(self.initialization, "constant", "orbitals.envelope_orbitals.initialization")
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
    periodic_orbitals: None = None
    use_bloch_envelopes: bool = False


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
    disable_determinant: bool = False
    max_n_up_orbitals: Optional[int] = None
    max_n_dn_orbitals: Optional[int] = None
    max_n_ions: Optional[int] = None
    kfac_register_complex: bool = False

    @model_validator(mode="after")
    def _check(self):
        _only(self.use_el_el_cusp_correction, False, "model.use_el_el_cusp_correction")
        _only(self.complex_wf, False, "model.complex_wf")
        _only(self.disable_determinant, False, "model.disable_determinant")
        return self


class MCMCSimpleProposalConfig(ConfigBaseclass):
    name: Literal["normal", "cauchy", "normal_one_el"] = "normal"     # configuration.py:952-953


class MCMCLangevinProposalConfig(ConfigBaseclass):
    name: Literal["langevin"] = "langevin"                            # configuration.py:956-963
    langevin_scale: float = 1.0
    r_min: float = 0.2
    r_max: float = 2.0


class LocalStepsizeProposalConfig(ConfigBaseclass):
    name: Literal["local", "local_one_el"] = "local"                  # configuration.py:966-975
    r_min: float = 0.1
    r_max: float = 1


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
    proposal: Union[MCMCSimpleProposalConfig, MCMCLangevinProposalConfig, LocalStepsizeProposalConfig] = MCMCSimpleProposalConfig()
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
    """configuration.py:1262-1330.  `optimizer`, `checkpoints`, `shared_optimization` belong to the control plane around the hot
    path: they are carried through YAML round trips untouched and never read here."""
    mcmc: MCMCConfigOptimization = MCMCConfigOptimization()
    clipping: ClippingConfig = ClippingConfig()
    n_epochs: int = 60_000
    forward_lap: bool = True
    max_batch_size: int = 64
    stop_on_nan: bool = True
    optimizer: Optional[Dict[str, Any]] = None
    n_epochs_prev: int = 0
    use_batch_reweighting: bool = False
    checkpoints: Optional[Dict[str, Any]] = None
    shared_optimization: Optional[Dict[str, Any]] = None
    params_ema_factor: Optional[float] = None


class EvaluationConfig(ConfigBaseclass):
    mcmc: MCMCConfigEvaluation = MCMCConfigEvaluation()
    n_epochs: int = 0
    forward_lap: bool = True
    max_batch_size: int = 64
    opt_epochs: List[int] = []
    evaluate_final: bool = True
    calculate_energies: bool = True
    forces: Optional[Dict[str, Any]] = None
    localization_metric: Optional[str] = None
    structure_factor_grid: Optional[Any] = None
    density: Optional[Any] = None


class ComputationConfig(ConfigBaseclass):
    float_precision: Literal["float32"] = "float32"
    rng_seed: Optional[int] = None
    n_local_devices: Optional[int] = None
    n_nodes: int = 1
    disable_jit: bool = False
    use_gpu: bool = True
    require_gpu: bool = False
    force_device_count: bool = False
    disable_tensor_cores: bool = True
    """The reference forces true-FP32 matmuls (process_molecule.py:20-29); here the dense layers run FP32-accurate 3xTF32 on the
    tensor cores (three tf32 products per FP32 product), which honours that contract; `dpe_set_gemm_path(0)` selects FP32 SIMT."""
    use_profiler: bool = False
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


def generate_el_ion_mapping(R, Z, n_el, n_up):
    """The reference's default electron -> ion assignment (PhysicalConfig._generate_el_ion_mapping, configuration.py:1571-1615):
    electrons are placed one at a time; every (spin, ion) candidate is scored by the largest |local spin| it would leave, where
    the local spin of an ion is the exp(-distance)-weighted sum of (n_up - n_dn) over all ions; the best-scoring candidate whose
    ion is not full and whose spin is not exhausted wins.  Output: spin-up electrons ion by ion, then spin-down.  Setup-time
    host code; the same numpy calls on the same arrays, so ties (symmetric molecules) break as in the reference."""
    R = np.array(R)
    n_ions = len(Z)
    weights = np.exp(-np.linalg.norm(R[:, None, :] - R[None, :, :], axis=-1))
    occupation = np.zeros([n_ions, 2], int)          # [ion, spin]
    left = [n_up, n_el - n_up]
    sign = np.array([1.0, -1.0])
    for _ in range(n_el):
        local_spin = weights @ (occupation[:, 0] - occupation[:, 1])
        # candidate c = spin * n_ions + ion adds sign[spin] * weights[:, ion] to the local spins
        outcome = [local_spin + weights @ (sign[c // n_ions] * np.eye(n_ions)[c % n_ions]) for c in range(2 * n_ions)]
        for c in np.argsort(np.max(np.abs(outcome), axis=-1)):
            spin, ion = divmod(int(c), n_ions)
            if occupation[ion].sum() == Z[ion] or left[spin] == 0:
                continue
            occupation[ion, spin] += 1
            left[spin] -= 1
            break
    return [ion for spin in range(2) for ion in range(n_ions) for _ in range(occupation[ion, spin])]


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
    changes: Optional[Any] = None
    weight_for_shared: Optional[float] = None

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
            elif self.Z is not None and self.R is not None and self.n_electrons is not None:
                self.el_ion_mapping = generate_el_ion_mapping(self.R, self.Z, self.n_electrons, self.n_up)
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
    # sections of the reference's root config (configuration.py:1934-1986) that configure the control plane around the hot path
    # (pre-training, PySCF baseline, loggers, SLURM dispatch, checkpoint reuse): accepted, kept verbatim, never read here
    pre_training: Optional[Dict[str, Any]] = None
    baseline: Optional[Dict[str, Any]] = None
    logging: Optional[Dict[str, Any]] = None
    dispatch: Optional[Dict[str, Any]] = None
    reuse: Optional[Dict[str, Any]] = None

    @classmethod
    def load_configuration_file(cls, config_file):
        """configuration.py:1988-1992."""
        with open(config_file) as f:
            return cls.model_validate(yaml.safe_load(f) or {})

    @classmethod
    def load(cls, config_file):
        return cls.load_configuration_file(config_file)
