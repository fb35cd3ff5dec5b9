// GPU-backed drop-in for CellPopulationLikelihood (src/cellpop/CellPopulationLikelihood.{h,cpp}) on top of the C ABI
// (include/bcm3b200.h, model kind "cell_population"). likelihood.xml surface as in the reference
// (CellPopulationLikelihood.cpp:27-35, Experiment.cpp:404-633):
//
//   <bcm_likelihood type="cell_population">
//     <experiment name= model_file= entry_time=<variable|number> num_cells= max_cells= divide_cells="false"
//                 [solver_min_timestep=] [solver_max_steps=] [solver_absolute_tolerance=] [solver_relative_tolerance=]
//                 [trailing_simulation_time=]>
//       <cell_variability distribution="diagonal_gaussian|full_gaussian" [covar_base_name=]>   (several diagonal_gaussian blocks allowed)
//         <variable (initial_condition_species=|model_parameter=|entry_time=) apply= scale=<variable|number> [negate=]
//                   [only_initial_cells=]/> ...
//       </cell_variability>
//       <data type="time_course_population_average" species_name="a[+b]" stdev=<variable|number> [proportional_stdev=]
//             [offset=] [scale=] [error_model=] [weight=] [missing_simulation_time_stdev=] [relative_to_time_average=]
//             [stdev_relative_to_scale=]/> ...
//       [<treatment_trajectory type="pulses" species_name=<constant species> times="t1,t2,..."/>]
//       [<experiment_specific_parameter parameter_name= replacement_parameter_name=/>] [<set_parameter parameter_name= value=/>]
//       [<set_species species_name= value=/>]
//     </experiment> ...
//   </bcm_likelihood>
//
// Several <experiment> elements and several <data> elements per experiment are accepted: the log-likelihood is the sum over
// experiments of the sum over their data sets (CellPopulationLikelihood.cpp:82-101, Experiment.cpp:346-355). Every data set is
// one handle of the C ABI; the data sets of one experiment all integrate their cells to the experiment's common end time
// (descriptor key simulation_end_time = the last time any of them requests, Experiment.cpp:190-214,655-656), which makes
// each handle's trajectories the ones the reference's single simulation of that experiment produces. The cells of an
// experiment are therefore integrated once per data set -- sharing one integration between the data sets of an experiment
// is a device-side optimisation that is not built yet.
//
// The SBML reader/code generator and the NetCDF reader stay on the reference side: the generated model
// (SetModel) and the data sets (SetData) are supplied before PostInitialize(), which is where the reference compiles its
// generated code too (Experiment::PostInitialize -> SolverCodeGenerator). Anything the device path does not implement
// (cell division, treatment trajectories from data, per-cell likelihood types) is refused here.
//
// <experiment_specific_parameter> (Experiment.cpp:515-527, :640-642): the cells of the experiment see the replacement
// variable's transformed value in place of the parameter's; done here by handing the experiment's handles a copy of the
// chain values with that column replaced (both variables must share a transform, and no data set of the experiment may
// read the replaced variable, because the reference's data likelihoods see the unreplaced values, Experiment.cpp:350).
// <set_parameter> (:508-514) fixes a parameter in the cell model BEFORE code generation: it is recorded
// (GetFixedParameters) for whoever supplies the generated model. <set_species> is read and then ignored, as in the
// reference, whose use of it is compiled out (Cell.cpp:88-93 under #if 0).
#pragma once

#include <utility>

#include "Likelihood.h"

class CellPopulationLikelihoodB200 : public bcm3::Likelihood {
public:
	struct Model { // what SBMLModel leaves behind: SBMLModel.cpp:92-125 (species order), :291-389 (generated text)
		std::string derivative_code;                 // SBMLModel::GenerateCode()
		std::vector<std::string> species_names;      // simulated species, in the generator's order
		std::vector<double> initial_conditions;      // [N]
		std::vector<double> constant_species;        // [Nc]
		std::vector<std::string> constant_species_names; // needed only when a treatment trajectory drives one of them
		std::vector<double> non_sampled_parameters;  // [Nn]
	};
	struct Data { // one time_course_population_average data set (DataLikelihoodTimeCourse.cpp:66-160)
		std::vector<double> timepoints;              // [T]
		std::vector<double> observed;                // [R][T], NaN = missing
		size_t num_replicates = 1;
	};

	CellPopulationLikelihoodB200(size_t sampling_threads, size_t evaluation_threads);
	~CellPopulationLikelihoodB200() override;

	bool Initialize(std::shared_ptr<const bcm3::VariableSet> varset, const bcm3::XmlNode& likelihood_node) override;
	size_t GetNumExperiments() const { return experiments.size(); }
	size_t GetNumDataSets(size_t experiment) const { return experiments[experiment].data.size(); }
	const std::string& GetExperimentName(size_t experiment) const { return experiments[experiment].name; }
	const std::string& GetModelFile(size_t experiment) const { return experiments[experiment].model_file; }
	// the generated model of one experiment / of every experiment (the usual case: all experiments share one model_file)
	void SetModel(size_t experiment, const Model& m) { experiments[experiment].model = m; }
	void SetModel(const Model& m)
	{
		for (auto& e : experiments) e.model = m;
	}
	void SetData(size_t experiment, size_t data_set, const Data& d) { experiments[experiment].data[data_set].data = d; }
	void SetData(const Data& d) { SetData(0, 0, d); }
	// the quasi-random table of VariabilityPseudoRandomIterator.cpp:14-26, [num_cells][D] uniforms in (0, 1); one per experiment
	void SetSobolTable(size_t experiment, const std::vector<double>& table) { experiments[experiment].sobol = table; }
	void SetSobolTable(const std::vector<double>& table) { SetSobolTable(0, table); }
	void SetDevice(int dev, bool compile_only_ = false) { device = dev; compile_only = compile_only_; }
	// false: one handle (and one integration of the experiment's cells) per <data> element instead of one per experiment
	void SetShareIntegration(bool share) { share_integration = share; }
	// CellPopulationLikelihood.cpp:46-61 -> Experiment.cpp:132-143: the names become non_sampled_parameters[i] of the generated
	// code (the generator is on the reference side: the Model handed to SetModel must have been generated with the same names),
	// the values start as NaN and are replaced between evaluations (bcmopt/main.cpp:231)
	bool AddNonSampledParameters(const std::vector<std::string>& variable_names) override;
	void SetNonSampledParameters(const bcm3::VectorReal& values) override;
	const std::vector<std::string>& GetNonSampledParameterNames() const { return non_sampled_names; }
	// CellPopulationLikelihood.cpp:73-80 (the reference's body is compiled out): evaluation and launch counts per data set
	void OutputEvaluationStatistics(const std::string& path) const override;
	bool PostInitialize() override;
	bool IsReentrant() override { return true; } // the reference's is not (CellPopulationLikelihood.h:22): one object per sampling thread
	bool EvaluateLogProbability(size_t threadix, const bcm3::VectorReal& values, bcm3::Real& logp) override;
	bool EvaluateLogProbabilityBatch(const bcm3::MatrixReal& values, bcm3::VectorReal& logp) override;

	// <set_parameter> elements of the experiment: to be applied to the cell model before its code is generated
	const std::vector<std::pair<std::string, double>>& GetFixedParameters(size_t experiment) const { return experiments[experiment].fixed_parameters; }
	size_t GetNumCells(size_t experiment = 0) const { return experiments[experiment].num_cells; }
	size_t GetVariabilityDimension(size_t experiment = 0) const { return experiments[experiment].variables.size(); }
	const std::string& GetDescriptor(size_t experiment = 0, size_t data_set = 0) const { return experiments[experiment].data[data_set].descriptor; }
	const std::string& LastError() const { return last_error; }

private:
	struct ValueRef { // "<variable name | number>" attributes (VariabilityDescriptionVariable.cpp:112-162, DataLikelihoodBase.cpp:38-73)
		long ix = -1;
		double fixed = 0.0;
	};
	struct VarEntry {
		bool is_ic = false;
		bool entry_time = false; // read, given a quasi-random dimension, never applied -- as in the reference
		bool only_initial_cells = false; // not applied to daughters (nor to the single cell of a num_cells="1" experiment)
		std::string target;
		int apply = 0;
		ValueRef scale;
		bool negate = false;
	};
	struct DataSet { // one <data> element = one handle of the C ABI
		std::string species_name, error_model = "normal";
		std::string type = "time_course_population_average"; // or "time_course" / "time_points": per-cell data + matching (DataLikelihoodTimeCourse.cpp, DataLikelihoodTimePoints.cpp)
		long value_relative_to_timepoint_ix = -1;             // time_points, DataLikelihoodBase.cpp:49
		bool use_only_nondivided = false;                     // time_points, DataLikelihoodTimePoints.cpp:27
		bool include_only_mitotic = false;                    // population average, include_only_cells_that_went_through_mitosis
		long saturation_scale_ix = -1;                        // time_course, <data saturation_scale="variable">
		std::string denominator_name;                         // use_log_ratio: species b of species_name="a/b" (species_name then holds a)
		long marker_of = -1;                                  // >= 0: a further marker (species_name="a;b") of the data set with that index in Experiment::data
		bool optimize_offset_scale = false;                   // time_course, DataLikelihoodTimeCourseBase.cpp:43-57
		double optimize_offset_min = -1.0, optimize_offset_max = 1.0, optimize_scale_min = 0.1, optimize_scale_max = 10.0;
		ValueRef stdev, proportional_stdev, offset, scale;
		bool have_proportional_stdev = false, relative_to_time_average = false, stdev_relative_to_scale = false;
		double weight = 1.0, missing_stdev = 300.0;
		Data data;
		std::string descriptor;
		void* handle = nullptr;
	};
	struct Experiment {
		std::string name, model_file, distribution = "diagonal_gaussian", covar_base_name;
		size_t num_cells = 1, max_cells = 20;
		bool divide_cells = true; // Experiment.cpp:488
		ValueRef entry_time;
		double solver_max_timestep = std::numeric_limits<double>::infinity();
		double solver_min_timestep = 1e-8, solver_abs_tol = 4.0 * 1.1920928955078125e-07, solver_rel_tol = 4.0 * 1.1920928955078125e-07;
		long solver_max_steps = 10000;
		double trailing_simulation_time = 0.0; // the cells are integrated this much past the last requested time
		std::vector<VarEntry> variables;
		std::string treatment_species_name; // <treatment_trajectory type="pulses">
		std::vector<double> treatment_times;
		Model model;
		std::vector<double> sobol;
		std::vector<DataSet> data;
		std::vector<std::pair<size_t, size_t>> specific_parameters; // (parameter, replacement) variable indices
		std::vector<std::pair<std::string, double>> fixed_parameters; // <set_parameter>
		std::vector<std::string> set_species;                         // <set_species>: names only (validated, otherwise unused)
	};
	bool Resolve(const std::string& text, ValueRef& out, const char* what);
	bool InitializeExperiment(const bcm3::XmlNode& node, Experiment& e);
	bool CreateHandle(Experiment& e, DataSet& ds, double simulation_end_time, const std::vector<DataSet*>& followers);
	// the [C][nvar] block the handles of `e` are given: `values` itself, or the copy with the experiment-specific columns replaced
	static const double* ExperimentValues(const Experiment& e, const double* values, size_t C, size_t nvar, std::vector<double>& scratch);
	bool Fail(const std::string& m)
	{
		last_error = m;
		return false;
	}

	std::shared_ptr<const bcm3::VariableSet> varset;
	std::vector<std::string> non_sampled_names;
	bool have_non_sampled_names = false;
	std::vector<Experiment> experiments;
	int device = 0;
	bool compile_only = false, share_integration = true;
	std::vector<int> status;
	std::vector<double> part; // one handle's per-chain results (batched entry: one caller at a time, like the reference's sampler)
	std::vector<double> replaced; // ExperimentValues scratch of the batched entry
	std::string last_error;
};
