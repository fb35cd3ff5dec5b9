// GPU-backed drop-in for CellPopulationLikelihood (src/cellpop/CellPopulationLikelihood.{h,cpp}) on top of the C ABI
// (include/bcm3b200.h, model kind "cell_population"). likelihood.xml surface as in the reference
// (CellPopulationLikelihood.cpp:27-35, Experiment.cpp:404-633):
//
//   <bcm_likelihood type="cell_population">
//     <experiment name= model_file= entry_time=<variable|number> num_cells= max_cells= divide_cells="false"
//                 [solver_min_timestep=] [solver_max_steps=] [solver_absolute_tolerance=] [solver_relative_tolerance=]>
//       <cell_variability distribution="diagonal_gaussian|full_gaussian" [covar_base_name=]>
//         <variable (initial_condition_species=|model_parameter=) apply= scale=<variable|number> [negate=]/> ...
//       </cell_variability>
//       <data type="time_course_population_average" species_name="a[+b]" stdev=<variable|number> [proportional_stdev=]
//             [offset=] [scale=] [error_model=] [weight=] [missing_simulation_time_stdev=]/>
//       [<treatment_trajectory type="pulses" species_name=<constant species> times="t1,t2,..."/>]
//     </experiment>
//   </bcm_likelihood>
//
// The SBML reader/code generator and the NetCDF reader stay on the reference side: the generated model
// (SetModel) and the data set (SetData) are supplied before PostInitialize(), which is where the reference compiles its
// generated code too (Experiment::PostInitialize -> SolverCodeGenerator). Anything the device path does not implement
// (cell division, several experiments / data sets, treatment trajectories from data, per-cell likelihood types) is refused here.
#pragma once

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
	void SetModel(const Model& m) { model = m; }
	void SetData(const Data& d) { data = d; }
	// the quasi-random table of VariabilityPseudoRandomIterator.cpp:14-26, [num_cells][D] uniforms in (0, 1)
	void SetSobolTable(const std::vector<double>& table) { sobol = table; }
	void SetDevice(int dev, bool compile_only_ = false) { device = dev; compile_only = compile_only_; }
	bool PostInitialize() override;
	bool IsReentrant() override { return true; } // the reference's is not (CellPopulationLikelihood.h:22): one object per sampling thread
	bool EvaluateLogProbability(size_t threadix, const bcm3::VectorReal& values, bcm3::Real& logp) override;
	bool EvaluateLogProbabilityBatch(const bcm3::MatrixReal& values, bcm3::VectorReal& logp) override;

	size_t GetNumCells() const { return num_cells; }
	size_t GetVariabilityDimension() const { return variables.size(); }
	const std::string& GetDescriptor() const { return descriptor; }
	const std::string& LastError() const { return last_error; }

private:
	struct ValueRef { // "<variable name | number>" attributes (VariabilityDescriptionVariable.cpp:112-162, DataLikelihoodBase.cpp:38-73)
		long ix = -1;
		double fixed = 0.0;
	};
	struct VarEntry {
		bool is_ic = false;
		std::string target;
		int apply = 0;
		ValueRef scale;
		bool negate = false;
	};
	bool Resolve(const std::string& text, ValueRef& out, const char* what);
	bool Fail(const std::string& m)
	{
		last_error = m;
		return false;
	}

	std::shared_ptr<const bcm3::VariableSet> varset;
	std::string experiment_name, model_file, distribution = "diagonal_gaussian", covar_base_name, species_name, error_model = "normal";
	size_t num_cells = 1;
	ValueRef entry_time, stdev, proportional_stdev, offset, scale;
	bool have_proportional_stdev = false, relative_to_time_average = false;
	double weight = 1.0, missing_stdev = 300.0;
	double solver_min_timestep = 1e-8, solver_abs_tol = 4.0 * 1.1920928955078125e-07, solver_rel_tol = 4.0 * 1.1920928955078125e-07;
	long solver_max_steps = 10000;
	std::vector<VarEntry> variables;
	std::string treatment_species_name; // <treatment_trajectory type="pulses">
	std::vector<double> treatment_times;
	Model model;
	Data data;
	std::vector<double> sobol;
	std::string descriptor;
	void* handle = nullptr;
	int device = 0;
	bool compile_only = false;
	std::vector<int> status;
	std::string last_error;
};
