#include "CellPopulationLikelihoodB200.h"

#include <algorithm>
#include <cstdlib>
#include <sstream>

extern "C" {
#include "bcm3b200.h"
}

using bcm3::Real;

namespace {

// VariabilityDescriptionVariable.cpp:139-157
int apply_code(const std::string& s)
{
	static const char* names[] = { "additive", "additive_log", "additive_log2", "multiplicative", "multiplicative_log", "multiplicative_log2", "replace" };
	for (int i = 0; i < 7; i++)
		if (s == names[i]) return i;
	return -1;
}

bool parse_number(const std::string& s, double& v)
{
	if (s.empty()) return false;
	char* end = nullptr;
	v = strtod(s.c_str(), &end);
	return end && *end == 0;
}

} // namespace

CellPopulationLikelihoodB200::CellPopulationLikelihoodB200(size_t, size_t) {}

CellPopulationLikelihoodB200::~CellPopulationLikelihoodB200()
{
	if (handle) bcm3b200_destroy(handle);
}

// a variable of the variable set by name, else a number
bool CellPopulationLikelihoodB200::Resolve(const std::string& text, ValueRef& out, const char* what)
{
	const size_t ix = varset->GetVariableIndex(text);
	if (ix != std::numeric_limits<size_t>::max()) {
		out.ix = (long)ix;
		return true;
	}
	if (parse_number(text, out.fixed)) {
		out.ix = -1;
		return true;
	}
	return Fail(std::string("Could not find variable for ") + what + " \"" + text + "\", and could also not cast it to a constant real value");
}

bool CellPopulationLikelihoodB200::Initialize(std::shared_ptr<const bcm3::VariableSet> vs, const bcm3::XmlNode& node)
{
	varset = vs;
	const bcm3::XmlNode* exp = nullptr;
	size_t num_experiments = 0;
	for (const auto& c : node.children)
		if (c.name == "experiment") {
			if (!exp) exp = &c;
			num_experiments++;
		}
	if (!exp) return Fail("Error parsing likelihood file: no experiment");
	if (num_experiments > 1) return Fail("the GPU path evaluates one experiment per likelihood");
	experiment_name = exp->get("name");
	model_file = exp->get("model_file");
	if (exp->get_bool("divide_cells", true)) return Fail("divide_cells=\"true\" (the reference's default, Experiment.cpp:488) is not supported by the GPU path: set divide_cells=\"false\"");
	num_cells = (size_t)exp->get_int("num_cells", 1);
	const size_t max_cells = (size_t)exp->get_int("max_cells", 20);
	if (num_cells > max_cells) return Fail("num_cells exceeds max_cells");
	if (exp->get("solver_type", "CVODE") != "CVODE") return Fail("only solver_type=\"CVODE\" is supported");
	solver_min_timestep = exp->get_real("solver_min_timestep", solver_min_timestep);
	solver_max_steps = exp->get_int("solver_max_steps", solver_max_steps);
	solver_abs_tol = exp->get_real("solver_absolute_tolerance", solver_abs_tol);
	solver_rel_tol = exp->get_real("solver_relative_tolerance", solver_rel_tol);
	if (!Resolve(exp->get("entry_time", "0"), entry_time, "entry_time")) return false;

	size_t num_data = 0, num_variability = 0;
	for (const auto& c : exp->children) {
		if (c.name == "cell_variability") {
			if (++num_variability > 1) return Fail("one cell_variability block per experiment is supported");
			distribution = c.get("distribution", "diagonal_gaussian");
			if (distribution != "diagonal_gaussian" && distribution != "full_gaussian") return Fail("Unknown cell_variability distribution \"" + distribution + "\"");
			covar_base_name = c.get("covar_base_name");
			for (const auto& v : c.children) {
				if (v.name != "variable") continue;
				VarEntry e;
				if (v.has("initial_condition_species")) {
					e.is_ic = true;
					e.target = v.get("initial_condition_species");
				} else if (v.has("model_parameter")) {
					e.target = v.get("model_parameter");
				} else {
					return Fail("cell_variability variable needs initial_condition_species or model_parameter (entry_time variability is not supported)");
				}
				e.apply = apply_code(v.get("apply"));
				if (e.apply < 0) return Fail("Unknown cell variability apply type \"" + v.get("apply") + "\"");
				if (!Resolve(v.get("scale", "0"), e.scale, "scale")) return false;
				e.negate = v.get_bool("negate", false);
				if (v.get_bool("only_initial_cells", false)) return Fail("only_initial_cells is not supported");
				variables.push_back(e);
			}
		} else if (c.name == "data") {
			if (++num_data > 1) return Fail("one data set per experiment is supported");
			if (c.get("type") != "time_course_population_average") return Fail("data type \"" + c.get("type") + "\" is not supported by the GPU path (time_course_population_average only)");
			species_name = c.get("species_name");
			if (species_name.find(';') != std::string::npos) return Fail("one observed quantity per data set is supported");
			error_model = c.get("error_model", "normal");
			if (!Resolve(c.get("stdev", "1"), stdev, "stdev")) return false;
			if (c.has("proportional_stdev")) {
				have_proportional_stdev = true;
				if (!Resolve(c.get("proportional_stdev"), proportional_stdev, "proportional_stdev")) return false;
			}
			if (!Resolve(c.get("offset", "0"), offset, "offset")) return false;
			if (!Resolve(c.get("scale", "1"), scale, "scale")) return false;
			relative_to_time_average = c.get_bool("relative_to_time_average", false);
			weight = c.get_real("weight", 1.0);
			missing_stdev = c.get_real("missing_simulation_time_stdev", 300.0);
		} else if (c.name == "treatment_trajectory") {
			// Experiment.cpp:566-590 + TreatmentTrajectoryPulses::Load
			if (c.get("type") != "pulses") return Fail("treatment_trajectory type \"" + c.get("type") + "\" is not supported by the GPU path (pulses only)");
			if (!treatment_species_name.empty()) return Fail("one treatment_trajectory per experiment is supported");
			treatment_species_name = c.get("species_name");
			std::stringstream ss(c.get("times"));
			std::string tok;
			while (std::getline(ss, tok, ',')) {
				double v;
				if (!parse_number(tok, v)) return Fail("treatment_trajectory times: cannot parse \"" + tok + "\"");
				treatment_times.push_back(v);
			}
		} else if (c.name == "set_species" || c.name == "experiment_specific_parameter" || c.name == "set_parameter") {
			return Fail("<" + c.name + "> is not supported by the GPU path");
		}
	}
	if (num_data == 0) return Fail("experiment has no data");
	return true;
}

bool CellPopulationLikelihoodB200::PostInitialize()
{
	const size_t N = model.species_names.size(), nvar = varset->GetNumVariables(), T = data.timepoints.size(), D = variables.size();
	if (N == 0 || model.initial_conditions.size() != N || model.derivative_code.empty()) return Fail("SetModel() has not supplied the generated model");
	if (T == 0 || data.observed.size() != data.num_replicates * T) return Fail("SetData() has not supplied a consistent data set");
	if (D > 0 && sobol.size() != num_cells * D) return Fail("SetSobolTable(): expected num_cells x variability dimension entries");

	// species_name="a+b": the summed simulated species
	std::vector<size_t> obs;
	{
		std::stringstream ss(species_name);
		std::string part;
		while (std::getline(ss, part, '+')) {
			auto it = std::find(model.species_names.begin(), model.species_names.end(), part);
			if (it == model.species_names.end()) return Fail("Species \"" + part + "\" of the data set is not a simulated species of the model");
			obs.push_back((size_t)(it - model.species_names.begin()));
		}
	}
	std::ostringstream d;
	d.precision(17);
	d << "num_species=" << N << ";num_constant_species=" << model.constant_species.size() << ";num_variables=" << nvar << ";num_non_sampled="
	  << model.non_sampled_parameters.size() << ";num_cells=" << num_cells << ";num_timepoints=" << T << ";num_replicates=" << data.num_replicates
	  << ";variability_dim=" << D << ";variability_distribution=" << distribution << ";solver_relative_tolerance=" << solver_rel_tol
	  << ";solver_absolute_tolerance=" << solver_abs_tol << ";solver_min_timestep=" << solver_min_timestep << ";solver_max_steps=" << solver_max_steps
	  << ";relative_to_time_average=" << (relative_to_time_average ? 1 : 0) << ";error_model=" << error_model << ";weight=" << weight << ";missing_simulation_time_stdev=" << missing_stdev << ";device=" << device
	  << ";compile_only=" << (compile_only ? 1 : 0);
	auto ref = [&](const char* name, const ValueRef& r) {
		if (r.ix >= 0) d << ";" << name << "_ix=" << r.ix;
		else d << ";" << name << "=" << r.fixed;
	};
	ref("entry_time", entry_time);
	ref("stdev", stdev);
	if (have_proportional_stdev) ref("proportional_stdev", proportional_stdev);
	ref("offset", offset);
	ref("scale", scale);
	d << ";obs_species=";
	for (size_t k = 0; k < obs.size(); k++) d << (k ? "+" : "") << obs[k];
	if (!treatment_species_name.empty()) {
		auto it = std::find(model.constant_species_names.begin(), model.constant_species_names.end(), treatment_species_name);
		if (it == model.constant_species_names.end()) return Fail("Treatment species \"" + treatment_species_name + "\" is not a constant species of the model");
		d << ";treatment_species=" << (it - model.constant_species_names.begin());
	}
	descriptor = d.str();
	if (bcm3b200_create("cell_population", descriptor.data(), descriptor.size(), 1, &handle) != BCM3B200_OK) return Fail(bcm3b200_last_error());

	auto set = [&](const char* name, const std::vector<double>& v, std::vector<size_t> shape) {
		static const double zero = 0.0;
		const double* p = v.empty() ? &zero : v.data();
		if (bcm3b200_set_data(handle, name, p, shape.data(), (int)shape.size()) != BCM3B200_OK) return Fail(bcm3b200_last_error());
		return true;
	};
	std::vector<double> transforms(nvar);
	for (size_t i = 0; i < nvar; i++) transforms[i] = (double)varset->GetTransform(i);
	bool ok = set("initial_conditions", model.initial_conditions, { N }) && set("constant_species", model.constant_species, { model.constant_species.size() }) &&
	          set("non_sampled_parameters", model.non_sampled_parameters, { model.non_sampled_parameters.size() }) && set("timepoints", data.timepoints, { T }) &&
	          set("observed", data.observed, { data.num_replicates, T }) && set("transforms", transforms, { nvar });
	if (!ok) return false;
	if (!treatment_species_name.empty() && !treatment_times.empty() && !set("treatment_times", treatment_times, { treatment_times.size() })) return false;
	if (D > 0) {
		std::vector<double> rows(D * 6);
		for (size_t i = 0; i < D; i++) {
			const VarEntry& e = variables[i];
			size_t target;
			if (e.is_ic) {
				auto it = std::find(model.species_names.begin(), model.species_names.end(), e.target);
				if (it == model.species_names.end()) return Fail("Variability initial_condition_species \"" + e.target + "\" is not a simulated species");
				target = (size_t)(it - model.species_names.begin());
			} else {
				target = varset->GetVariableIndex(e.target);
				if (target == std::numeric_limits<size_t>::max()) return Fail("Variability model_parameter \"" + e.target + "\" is not a sampled variable");
			}
			double* r = rows.data() + i * 6;
			r[0] = e.is_ic ? 1.0 : 0.0;
			r[1] = (double)target;
			r[2] = (double)e.apply;
			r[3] = (double)e.scale.ix;
			r[4] = e.scale.fixed;
			r[5] = e.negate ? 1.0 : 0.0;
		}
		if (!set("sobol", sobol, { num_cells, D }) || !set("variability", rows, { D, 6 })) return false;
		if (distribution == "full_gaussian" && D > 1) {
			// covariance values are variables named <covar_base_name><k+1>_<i+1>, k < i (VariabilityDescription.cpp:203-216)
			std::vector<double> cov(D * (D - 1));
			for (size_t i = 1; i < D; i++)
				for (size_t k = 0; k < i; k++) {
					const std::string name = covar_base_name + std::to_string(k + 1) + "_" + std::to_string(i + 1);
					ValueRef r;
					if (!Resolve(name, r, "covariance")) return false;
					const size_t e = (i - 1) * i / 2 + k;
					cov[2 * e] = (double)r.ix;
					cov[2 * e + 1] = r.fixed;
				}
			if (!set("variability_covariance", cov, { D * (D - 1) / 2, 2 })) return false;
		}
	}
	if (bcm3b200_set_text(handle, "derivative_code", model.derivative_code.data(), model.derivative_code.size()) != BCM3B200_OK) return Fail(bcm3b200_last_error());
	if (bcm3b200_finalize(handle) != BCM3B200_OK) return Fail(bcm3b200_last_error());
	return true;
}

bool CellPopulationLikelihoodB200::EvaluateLogProbability(size_t, const bcm3::VectorReal& values, Real& logp)
{
	int st = 0;
	if (bcm3b200_evaluate_batch(handle, 1, values.size(), values.data(), &logp, &st) != BCM3B200_OK) return Fail(bcm3b200_last_error());
	return true; // -inf (a failed cell, Experiment.cpp:356-358) is a legal value; NaN is turned into an error by the sampler
}

bool CellPopulationLikelihoodB200::EvaluateLogProbabilityBatch(const bcm3::MatrixReal& values, bcm3::VectorReal& logp)
{
	logp.assign(values.cols(), -bcm3::kInf);
	status.assign(values.cols(), 0);
	if (values.cols() == 0) return true;
	if (bcm3b200_evaluate_batch(handle, values.cols(), values.rows(), values.data.data(), logp.data(), status.data()) != BCM3B200_OK) return Fail(bcm3b200_last_error());
	return true;
}
