"""Host-side data model of the cell_population likelihood for the slice built in round 1:
independent, non-dividing cells with per-cell quasi-random variability and ONE time_course_population_average data set.

Mirrors what CellPopulationLikelihood / Experiment::Load leave in memory after parsing likelihood.xml, the SBML model and
the NetCDF data (reference: src/cellpop/Experiment.cpp:404-633, CellPopulationLikelihood.cpp:27-80). The SBML reader and
code generator stay on the reference side: the model enters as the TEXT its generator emits
(SBMLModel::GenerateCode, src/sbml/SBMLModel.cpp:291-389).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

FLT_EPSILON = float(np.finfo(np.float32).eps)

APPLY_TYPES = {  # VariabilityDescriptionVariable.cpp:139-157
    "additive": 0, "additive_log": 1, "additive_log2": 2, "multiplicative": 3, "multiplicative_log": 4,
    "multiplicative_log2": 5, "replace": 6,
}


@dataclass
class Variability:
    """One <variable> of a diagonal_gaussian <cell_variability> block (VariabilityDescriptionVariable.cpp:112-162)."""

    apply: str
    model_parameter: int | None = None          # index of the sampled variable it perturbs ...
    initial_condition_species: int | None = None  # ... or index of the ODE species whose initial value it perturbs
    scale_ix: int | None = None                 # scale = transformed variable[scale_ix] ...
    scale_fixed: float = 0.0                    # ... or a constant; the quasi-random normal is multiplied by exp(scale)
    negate: bool = False
    # <variable entry_time=...>: the reference reads it, gives it a quasi-random dimension and never applies it
    # (VariabilityDescription::ApplyVariabilityEntryTime has no caller); kind 2 of the ABI's variability rows
    entry_time: bool = False
    # <variable only_initial_cells="true">: applied to cells whose "initial cell" flag is set only -- not to daughters, and (a
    # quirk of Experiment.cpp:662-670) not to the single initial cell of a num_cells="1" experiment; + 4 on the kind column
    only_initial_cells: bool = False

    def row(self):
        is_ic = self.initial_condition_species is not None
        target = 0 if self.entry_time else (self.initial_condition_species if is_ic else self.model_parameter)
        return [(2.0 if self.entry_time else float(is_ic)) + (4.0 if self.only_initial_cells else 0.0), float(target), float(APPLY_TYPES[self.apply]), float(-1 if self.scale_ix is None else self.scale_ix),
                float(self.scale_fixed), float(self.negate)]


@dataclass
class Marker:
    """A further marker of a per-cell data set -- species_name="a+b;c", stdev="s1;s2" ...: the entries after a ';'
    (DataLikelihoodTimeCourseBase.cpp:79-87, DataLikelihoodBase.cpp:130-233). Error model, weight and the switches are the data set's."""

    obs_species: list
    observed: np.ndarray                  # [observed cells][T], NaN = missing
    stdev_ix: int | None = None
    stdev: float = 1.0
    proportional_stdev_ix: int | None = None
    proportional_stdev: float = 1.0
    offset_ix: int | None = None
    offset: float = 0.0
    scale_ix: int | None = None
    scale: float = 1.0
    log_ratio_denominator: int | None = None  # use_log_ratio: species b of "a/b" (obs_species = [a])


@dataclass
class CellPopProblem:
    derivative_code: str                 # generated_derivative text in the reference generator's ABI
    num_species: int
    initial_conditions: np.ndarray       # [N]
    transforms: np.ndarray               # [nvar] VariableSet transform codes
    num_cells: int
    timepoints: np.ndarray               # [T] data timepoints (sorted)
    observed: np.ndarray                 # [R][T], NaN = missing
    obs_species: list[int]               # species_name="a+b": indices of the summed ODE species
    constant_species: np.ndarray = field(default_factory=lambda: np.zeros(0))
    non_sampled_parameters: np.ndarray = field(default_factory=lambda: np.zeros(0))
    sobol: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))  # [num_cells][D] uniforms in (0, 1)
    variability: list[Variability] = field(default_factory=list)
    variability_distribution: str = "diagonal_gaussian"   # or "full_gaussian" (VariabilityDescription.cpp:50-139)
    covariance: list = field(default_factory=list)        # full_gaussian: D (D - 1) / 2 entries, each a variable index (int) or a fixed float
    # one <treatment_trajectory type="pulses" species_name= times="t1,t2,..."/> (TreatmentTrajectoryPulses.cpp): index of the
    # CONSTANT species it drives and the pulse times; None = no treatment
    treatment_species: int | None = None
    treatment_times: np.ndarray = field(default_factory=lambda: np.zeros(0))
    entry_time_ix: int | None = None
    entry_time: float = 0.0
    error_model: str = "normal"
    # <data type=>: "time_course_population_average" (observed [replicates][T]) or "time_course" -- per-cell trajectories:
    # observed [observed cells][T], as many observed as simulated cells, every observed cell matched to one simulated cell by
    # minimum-cost perfect matching (DataLikelihoodTimeCourse.cpp:230-365); synchronize="none", no parent information, one marker
    data_kind: str = "time_course_population_average"
    # "time_points" (DataLikelihoodTimePoints.cpp): observed [observed cell slots][T], NaN = no such cell at that timepoint -- at
    # every timepoint the observed cells present are matched to simulated cells (snapshots of different cells per time);
    # value_relative_to_timepoint_ix (DataLikelihoodBase.cpp:49): the simulated value relative to its own value at that timepoint
    value_relative_to_timepoint_ix: int | None = None
    # time_course: <data optimize_offset_scale="true" ...> (DataLikelihoodTimeCourseBase.cpp:43-57, 317-322): every observed
    # trajectory is regressed on every simulated one, offset and scale clamped to these ranges (the reference's defaults)
    optimize_offset_scale: bool = False
    # time_course: <data saturation_scale="variable"> (DataLikelihoodTimeCourse.cpp:243-254): index of the variable s of the signal
    # saturation s / (1 + exp(-x)) - s / 2 applied to the scaled and shifted trajectories
    saturation_scale_ix: int | None = None
    # time_course: <data use_log_ratio="true" species_name="a/b"> (DataLikelihoodTimeCourse.cpp:380-397): the cell's value is
    # log10(a / b), b replaced by 1e-16 when smaller; obs_species = [a], this = b. With markers, every marker is a ratio.
    log_ratio_denominator: int | None = None
    # population average over the cells that entered mitosis only: the model's "nuclear_envelope" species (by index) fell below 0.5
    # after some accepted step (Cell.cpp:487-492; DataLikelihoodTimeCoursePopulationAverage.cpp:171-176)
    include_only_cells_that_went_through_mitosis: bool = False
    nuclear_envelope_species: int | None = None
    use_only_nondivided: bool = False  # time_points, dividing population: daughters are left out (DataLikelihoodTimePoints.cpp:349-351)
    # per-cell data kinds: the markers after the first one (which is obs_species / observed / stdev ... of this problem)
    extra_markers: list = field(default_factory=list)
    optimize_offset_range: tuple = (-1.0, 1.0)
    optimize_scale_range: tuple = (0.1, 10.0)
    relative_to_time_average: bool = False   # <data relative_to_time_average="true">: log of the value over its time average
    # the time the experiment integrates its cells to when it has further data sets that end later (Experiment.cpp:655-656);
    # None: the last of `timepoints`
    simulation_end_time: float | None = None
    stdev_relative_to_scale: bool = False    # <data stdev_relative_to_scale="true">: stdev *= data scale (DataLikelihoodBase.cpp:151-153)
    weight: float = 1.0
    stdev_ix: int | None = None
    stdev: float = 1.0
    proportional_stdev_ix: int | None = None   # error_model proportional_normal / additive_proportional_normal
    proportional_stdev: float = 1.0
    offset_ix: int | None = None
    offset: float = 0.0
    scale_ix: int | None = None
    scale: float = 1.0
    missing_simulation_time_stdev: float = 300.0
    solver_relative_tolerance: float = 4 * FLT_EPSILON   # Experiment.cpp:415-416
    solver_absolute_tolerance: float = 4 * FLT_EPSILON
    solver_min_timestep: float = 1e-8                     # Experiment.cpp:412
    solver_max_timestep: float = float("inf")             # Experiment.cpp:413 -> CVodeSetMaxStep (0 / inf: no ceiling)
    solver_max_steps: int = 10000                         # Experiment.cpp:414
    # <experiment divide_cells="true" max_cells=>: a cell whose ODE species "cytokinesis" exceeds 1 after a step ends there and two
    # daughters start from its state with seven named species reset (Experiment.cpp:726-782, Cell.cpp:119-148, 463-538);
    # a cell whose "apoptosis" species exceeds 1 ends. Species are given by index; `sobol` then has more rows than num_cells
    # (the reference makes 100 * num_cells): daughters of the cell with row r take rows num_cells + 2 r + child.
    divide_cells: bool = False
    max_cells: int = 0
    cytokinesis_species: int | None = None
    apoptosis_species: int | None = None
    # indices of cytokinesis, nuclear_envelope, G1S_break, G2_break, spindle_components, assembled_spindle, chromatid_separation
    division_reset_species: tuple = ()

    @property
    def capacity(self) -> int:
        """Number of cell columns of every per-cell output: max_cells for a dividing population, num_cells otherwise."""
        return int(self.max_cells) if self.divide_cells else int(self.num_cells)

    @property
    def num_variables(self) -> int:
        return int(self.transforms.shape[0])

    @property
    def num_timepoints(self) -> int:
        return int(self.timepoints.shape[0])

    @property
    def num_replicates(self) -> int:
        return int(self.observed.shape[0])

    @property
    def variability_dim(self) -> int:
        return len(self.variability)

    def covariance_rows(self) -> np.ndarray:
        """[D (D - 1) / 2][2]: variable index (or -1), fixed value -- entry (i - 1) i / 2 + k is the angle (in units of pi) of
        row i, column k of the spherical Cholesky parametrisation."""
        D = len(self.variability)
        n = D * (D - 1) // 2
        if self.variability_distribution != "full_gaussian":
            return np.zeros((0, 2))
        if len(self.covariance) != n:
            raise ValueError(f"full_gaussian with {D} variables needs {n} covariance entries")
        return np.array([[float(c), 0.0] if isinstance(c, (int, np.integer)) else [-1.0, float(c)] for c in self.covariance], dtype=np.float64).reshape(n, 2)

    def variability_rows(self) -> np.ndarray:
        return np.array([v.row() for v in self.variability], dtype=np.float64).reshape(len(self.variability), 6)
