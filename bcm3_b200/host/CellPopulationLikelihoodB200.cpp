#include "CellPopulationLikelihoodB200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
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
	for (auto& e : experiments)
		for (auto& ds : e.data)
			if (ds.handle) bcm3b200_destroy(ds.handle);
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

// CellPopulationLikelihood::Initialize (CellPopulationLikelihood.cpp:27-35): one Experiment per <experiment> element
bool CellPopulationLikelihoodB200::Initialize(std::shared_ptr<const bcm3::VariableSet> vs, const bcm3::XmlNode& node)
{
	varset = vs;
	experiments.clear();
	for (const auto& c : node.children)
		if (c.name == "experiment") {
			experiments.emplace_back();
			if (!InitializeExperiment(c, experiments.back())) return false;
		}
	if (experiments.empty()) return Fail("Error parsing likelihood file: no experiment");
	return true;
}

// Experiment::Create + Experiment::Load (Experiment.cpp:404-614)
bool CellPopulationLikelihoodB200::InitializeExperiment(const bcm3::XmlNode& xml, Experiment& e)
{
	const bcm3::XmlNode* exp = &xml;
	e.name = exp->get("name");
	e.model_file = exp->get("model_file");
	e.divide_cells = exp->get_bool("divide_cells", true); // Experiment.cpp:488: true unless said otherwise
	e.num_cells = (size_t)exp->get_int("num_cells", 1);
	e.max_cells = (size_t)exp->get_int("max_cells", 20);
	if (e.num_cells > e.max_cells) return Fail("num_cells exceeds max_cells");
	// Cell.cpp:489-497: a non-zero value extends a cell's integration past the end of the experiment from a threshold-crossing
	// time that only exists with stored integration points (per-cell data types)
	if (exp->get_real("simulate_past_chromatid_separation_time", 0.0) != 0.0) return Fail("simulate_past_chromatid_separation_time is not supported by the GPU path");
	if (exp->get("solver_type", "CVODE") != "CVODE") return Fail("only solver_type=\"CVODE\" is supported");
	e.solver_min_timestep = exp->get_real("solver_min_timestep", e.solver_min_timestep);
	e.solver_max_steps = exp->get_int("solver_max_steps", e.solver_max_steps);
	e.solver_abs_tol = exp->get_real("solver_absolute_tolerance", e.solver_abs_tol);
	e.solver_rel_tol = exp->get_real("solver_relative_tolerance", e.solver_rel_tol);
	if (!Resolve(exp->get("entry_time", "0"), e.entry_time, "entry_time")) return false;
	e.solver_max_timestep = exp->get_real("solver_max_timestep", bcm3::kInf); // Experiment.cpp:413 -> CVodeSetMaxStep
	// Experiment.cpp:172-180: only meaningful for the synchronised per-cell data types (and it overwrites the fixed entry time there)
	if (!exp->get("synchronization_time_offset").empty()) return Fail("synchronization_time_offset is not supported by the GPU path");
	e.trailing_simulation_time = exp->get_real("trailing_simulation_time", 0.0); // Experiment.cpp:489, 655-656
	if (!(e.trailing_simulation_time >= 0.0)) return Fail("trailing_simulation_time must not be negative");

	size_t num_variability = 0;
	for (const auto& c : exp->children) {
		if (c.name == "cell_variability") {
			// Several blocks take successive quasi-random dimensions and are applied in order (Cell.cpp:163-175,
			// VariabilityDescription.cpp:54-64): for diagonal_gaussian blocks that is one block with all their variables.
			const std::string distribution = c.get("distribution", "diagonal_gaussian");
			if (distribution != "diagonal_gaussian" && distribution != "full_gaussian") return Fail("Unknown cell_variability distribution \"" + distribution + "\"");
			if (++num_variability > 1 && (distribution != "diagonal_gaussian" || e.distribution != "diagonal_gaussian"))
				return Fail("several cell_variability blocks per experiment are supported only when all of them are diagonal_gaussian");
			e.distribution = distribution;
			e.covar_base_name = c.get("covar_base_name");
			for (const auto& v : c.children) {
				if (v.name != "variable") continue;
				VarEntry ve;
				if (v.has("initial_condition_species")) {
					ve.is_ic = true;
					ve.target = v.get("initial_condition_species");
				} else if (v.has("model_parameter")) {
					ve.target = v.get("model_parameter");
				} else if (!v.get("entry_time").empty()) {
					// VariabilityDescriptionVariable.cpp:116-117: the reference has no caller of ApplyVariabilityEntryTime, so
					// the variable only occupies a quasi-random dimension
					ve.entry_time = true;
				} else {
					return Fail("Cell variability description has neither a initial_condition_species name nor a model_parameter name, nor entry_time");
				}
				ve.apply = apply_code(v.get("apply"));
				if (ve.apply < 0) return Fail("Unknown cell variability apply type \"" + v.get("apply") + "\"");
				if (!Resolve(v.get("scale", "0"), ve.scale, "scale")) return false;
				ve.negate = v.get_bool("negate", false);
				// VariabilityDescriptionVariable.cpp:132-136: defaults to true for an entry-time variable, false otherwise
				ve.only_initial_cells = v.get_bool("only_initial_cells", ve.entry_time);
				e.variables.push_back(ve);
			}
		} else if (c.name == "data") {
			// DataLikelihoodBase::Create (DataLikelihoodBase.cpp:17-36) + DataLikelihoodTimeCoursePopulationAverage::Load
			// DataLikelihoodTimeCourse::Load (.cpp:20-215) for the per-cell type: what it can do beyond the slice built here is refused
			const std::string type = c.get("type");
			if (type != "time_course_population_average" && type != "time_course" && type != "time_points")
				return Fail("data type \"" + type + "\" is not supported by the GPU path (time_course_population_average, time_course, time_points)");
			DataSet ds;
			ds.type = type;
			if (type == "time_points") {
				// DataLikelihoodTimePoints::Load (DataLikelihoodTimePoints.cpp:20-43) + DataLikelihoodBase.cpp:49
				const std::string sync = c.get("synchronize", "");
				if (!sync.empty() && sync != "none") return Fail("time_points data with synchronize=\"" + sync + "\" needs stored integration points: not supported by the GPU path");
				if (c.get_bool("relative_to_time_average", false)) return Fail("relative_to_time_average belongs to time_course_population_average data");
				const std::string em = c.get("error_model", "normal");
				if (em != "normal" && em != "student_t4") return Fail("time_points data knows the normal and student_t4 error models (DataLikelihoodTimePoints.cpp:280-287)");
				ds.value_relative_to_timepoint_ix = (long)c.get_real("value_relative_to_timepoint_ix", -1.0);
				ds.use_only_nondivided = c.get_bool("use_only_nondivided", false); // .cpp:27: daughters of a dividing population are left out
			}
			if (type == "time_course") {
				const std::string sync = c.get("synchronize", "");
				if (!sync.empty() && sync != "none") return Fail("time_course data with synchronize=\"" + sync + "\" needs stored integration points: not supported by the GPU path");
				// DataLikelihoodTimeCourseBase::Load, .cpp:43-57
				ds.optimize_offset_scale = c.get_bool("optimize_offset_scale", false);
				ds.optimize_offset_min = c.get_real("optimize_offset_min", -1.0);
				ds.optimize_offset_max = c.get_real("optimize_offset_max", 1.0);
				ds.optimize_scale_min = c.get_real("optimize_scale_min", 0.1);
				ds.optimize_scale_max = c.get_real("optimize_scale_max", 10.0);
				const std::string em = c.get("error_model", "normal");
				if (ds.optimize_offset_scale && (em == "proportional_normal" || em == "additive_proportional_normal"))
					return Fail("optimize_offset_scale with a proportional error model is undefined in the reference (its per-cell sigma tables are not filled): not supported");
				if (!c.get("saturation_scale", "").empty()) {
					// DataLikelihoodTimeCourseBase::PostInitialize (.cpp:120-133) accepts a variable name or a number, but a number is
					// overwritten with DBL_MAX at every evaluation (PrepateEvaluation, .cpp:243-246): only the variable form works
					const size_t ix = varset->GetVariableIndex(c.get("saturation_scale"));
					if (ix == std::numeric_limits<size_t>::max())
						return Fail("saturation_scale must name a variable (the reference overwrites a numeric saturation scale with the largest double at every evaluation)");
					ds.saturation_scale_ix = (long)ix;
				}
				if (c.get_bool("relative_to_time_average", false)) return Fail("relative_to_time_average belongs to time_course_population_average data");
			}
			// species_name="a+b;c": several MARKERS per cell (DataLikelihoodTimeCourseBase.cpp:79-87), each with the entry of the
			// ';'-separated stdev / proportional_stdev / offset / scale lists that has its index -- or the only entry when a list
			// has one (DataLikelihoodBase.cpp:75-119, 130-233). Per-cell data types only; every marker after the first becomes a
			// data set of its own here (SetData supplies its observed block) that names the first as its owner.
			auto split_list = [](const std::string& text) {
				std::vector<std::string> out;
				std::stringstream ss(text);
				std::string part;
				while (std::getline(ss, part, ';')) {
					const size_t b = part.find_first_not_of(" \t"), en = part.find_last_not_of(" \t");
					out.push_back(b == std::string::npos ? std::string() : part.substr(b, en - b + 1));
				}
				if (out.empty()) out.push_back(std::string());
				return out;
			};
			const std::vector<std::string> marker_species = split_list(c.get("species_name"));
			if (marker_species.size() > 1 && type == "time_course_population_average") return Fail("one observed quantity per population-average data set is supported");
			if (marker_species.size() > 4) return Fail("more than four markers per data set are not supported by the GPU path");
			auto list_entry = [&](const char* attr, const char* def, size_t l, std::string& out) {
				const std::vector<std::string> tokens = split_list(c.get(attr, def));
				if (tokens.size() == 1) out = tokens[0];
				else if (l < tokens.size()) out = tokens[l];
				else return Fail(std::string("the ") + attr + " list is shorter than the list of markers (DataLikelihoodBase.cpp: Out of bounds)");
				return true;
			};
			ds.species_name = marker_species[0];
			ds.error_model = c.get("error_model", "normal");
			std::string entry;
			if (!list_entry("stdev", "1", 0, entry) || !Resolve(entry, ds.stdev, "stdev")) return false;
			if (c.has("proportional_stdev")) {
				ds.have_proportional_stdev = true;
				if (!list_entry("proportional_stdev", "1", 0, entry) || !Resolve(entry, ds.proportional_stdev, "proportional_stdev")) return false;
			}
			// DataLikelihoodBase.cpp:64-69
			if ((ds.error_model == "proportional_normal" || ds.error_model == "additive_proportional_normal") && !ds.have_proportional_stdev)
				return Fail("Proportional error model is selected, but proportional stdev has not been specified.");
			if (!list_entry("offset", "0", 0, entry) || !Resolve(entry, ds.offset, "offset")) return false;
			if (!list_entry("scale", "1", 0, entry) || !Resolve(entry, ds.scale, "scale")) return false;
			ds.relative_to_time_average = c.get_bool("relative_to_time_average", false);
			ds.stdev_relative_to_scale = c.get_bool("stdev_relative_to_scale", false); // DataLikelihoodBase.cpp:48
			// attributes of DataLikelihoodTimeCourseBase::Load (.cpp:41-57) that change the population average and are not built
			// use_log_ratio (DataLikelihoodTimeCourseBase.cpp:142-147, 171-201): per-cell time_course data only, every marker "a/b"
			const bool use_log_ratio = c.get_bool("use_log_ratio", false);
			if (use_log_ratio && type != "time_course") return Fail("use_log_ratio is supported for time_course data only (DataLikelihoodTimeCourse.cpp:380-397)");
			auto split_ratio = [&](DataSet& target) {
				const size_t slash = target.species_name.find('/');
				if (use_log_ratio) {
					if (slash == std::string::npos) return Fail("use_log_ratio is specified as true, but the species_name does not contain a division");
					target.denominator_name = target.species_name.substr(slash + 1);
					target.species_name = target.species_name.substr(0, slash);
					auto trim = [](std::string& t) {
						const size_t b = t.find_first_not_of(" \t"), en = t.find_last_not_of(" \t");
						t = (b == std::string::npos) ? std::string() : t.substr(b, en - b + 1);
					};
					trim(target.denominator_name);
					trim(target.species_name);
					if (target.denominator_name.find('/') != std::string::npos || target.species_name.find('+') != std::string::npos || target.denominator_name.find('+') != std::string::npos)
						return Fail("only division of exactly two species is supported");
				} else if (slash != std::string::npos) {
					return Fail("simulated species reference has a division, but use_log_ratio has not been specified; only log ratios are supported for now");
				}
				return true;
			};
			if (!split_ratio(ds)) return false;
			// DataLikelihoodTimeCourseBase.cpp:44; it acts in DataLikelihoodTimeCoursePopulationAverage::NotifySimulatedValue only (.cpp:171-176)
			ds.include_only_mitotic = c.get_bool("include_only_cells_that_went_through_mitosis", false) && type == "time_course_population_average";
			// optimize_offset_scale, saturation_scale and value_relative_to_timepoint_ix act on the per-cell data types only
			ds.weight = c.get_real("weight", 1.0);
			ds.missing_stdev = c.get_real("missing_simulation_time_stdev", 300.0);
			e.data.push_back(ds);
			const long owner = (long)e.data.size() - 1;
			for (size_t l = 1; l < marker_species.size(); l++) {
				DataSet mk = ds;
				mk.marker_of = owner;
				mk.species_name = marker_species[l];
				mk.denominator_name.clear();
				if (!split_ratio(mk)) return false;
				if (!list_entry("stdev", "1", l, entry) || !Resolve(entry, mk.stdev, "stdev")) return false;
				if (mk.have_proportional_stdev && (!list_entry("proportional_stdev", "1", l, entry) || !Resolve(entry, mk.proportional_stdev, "proportional_stdev"))) return false;
				if (!list_entry("offset", "0", l, entry) || !Resolve(entry, mk.offset, "offset")) return false;
				if (!list_entry("scale", "1", l, entry) || !Resolve(entry, mk.scale, "scale")) return false;
				e.data.push_back(mk);
			}
		} else if (c.name == "treatment_trajectory") {
			// Experiment.cpp:566-590 + TreatmentTrajectoryPulses::Load
			if (c.get("type") != "pulses") return Fail("treatment_trajectory type \"" + c.get("type") + "\" is not supported by the GPU path (pulses only)");
			if (!e.treatment_species_name.empty()) return Fail("one treatment_trajectory per experiment is supported");
			e.treatment_species_name = c.get("species_name");
			std::stringstream ss(c.get("times"));
			std::string tok;
			while (std::getline(ss, tok, ',')) {
				double v;
				if (!parse_number(tok, v)) return Fail("treatment_trajectory times: cannot parse \"" + tok + "\"");
				e.treatment_times.push_back(v);
			}
		} else if (c.name == "experiment_specific_parameter") {
			// Experiment.cpp:515-527
			const size_t p = varset->GetVariableIndex(c.get("parameter_name")), r = varset->GetVariableIndex(c.get("replacement_parameter_name"));
			const size_t none = std::numeric_limits<size_t>::max();
			if (p == none) return Fail("experiment_specific_parameter: \"" + c.get("parameter_name") + "\" is not a variable");
			if (r == none) return Fail("experiment_specific_parameter: replacement \"" + c.get("replacement_parameter_name") + "\" is not a variable");
			if (varset->GetTransform(p) != varset->GetTransform(r))
				return Fail("experiment_specific_parameter: \"" + c.get("parameter_name") + "\" and its replacement must share a variable transform on the GPU path");
			e.specific_parameters.emplace_back(p, r);
		} else if (c.name == "set_parameter") {
			double v;
			if (!c.has("parameter_name") || !parse_number(c.get("value"), v)) return Fail("set_parameter needs parameter_name and a numeric value");
			e.fixed_parameters.emplace_back(c.get("parameter_name"), v);
		} else if (c.name == "set_species") {
			double v;
			if (!c.has("species_name") || !parse_number(c.get("value"), v)) return Fail("set_species needs species_name and a numeric value");
			e.set_species.push_back(c.get("species_name"));
		}
	}
	if (e.data.empty()) return Fail("experiment \"" + e.name + "\" has no data");
	// the reference evaluates its data likelihoods on the unreplaced values (Experiment.cpp:350)
	for (const auto& sp : e.specific_parameters)
		for (const auto& ds : e.data)
			for (const ValueRef* r : { &ds.stdev, &ds.proportional_stdev, &ds.offset, &ds.scale })
				if (r->ix == (long)sp.first) return Fail("experiment_specific_parameter replaces a variable that a data set of the experiment reads");
	return true;
}

// Experiment::PostInitialize (Experiment.cpp:120-232): every data set becomes one handle; the experiment's simulation end is
// the last time any of its data sets requests (:190-214, :655-656)
bool CellPopulationLikelihoodB200::AddNonSampledParameters(const std::vector<std::string>& variable_names)
{
	non_sampled_names = variable_names;
	have_non_sampled_names = true;
	// Experiment.cpp:134-135: values unknown until SetNonSampledParameters
	for (auto& e : experiments) e.model.non_sampled_parameters.assign(variable_names.size(), std::numeric_limits<double>::quiet_NaN());
	return true;
}

void CellPopulationLikelihoodB200::SetNonSampledParameters(const bcm3::VectorReal& values)
{
	for (auto& e : experiments) {
		e.model.non_sampled_parameters.assign(values.begin(), values.end());
		for (auto& ds : e.data) {
			if (!ds.handle || values.empty()) continue;
			const size_t shape[1] = { values.size() };
			if (bcm3b200_set_data(ds.handle, "non_sampled_parameters", values.data(), shape, 1) != BCM3B200_OK) last_error = bcm3b200_last_error();
		}
	}
}

void CellPopulationLikelihoodB200::OutputEvaluationStatistics(const std::string& path) const
{
	FILE* f = fopen((path + "/cellpop_evaluation_statistics.txt").c_str(), "w");
	if (!f) return;
	fprintf(f, "experiment\tdata_set\tevaluations\tkernel_launches\n");
	for (const auto& e : experiments) {
		for (size_t k = 0; k < e.data.size(); k++) {
			int64_t evals = 0, launches = 0;
			if (e.data[k].handle) {
				bcm3b200_get_stat(e.data[k].handle, "num_evaluations", &evals);
				bcm3b200_get_stat(e.data[k].handle, "total_kernel_launches", &launches);
			}
			fprintf(f, "%s\t%zu\t%lld\t%lld\n", e.name.c_str(), k, (long long)evals, (long long)launches);
		}
	}
	fclose(f);
}

bool CellPopulationLikelihoodB200::PostInitialize()
{
	for (auto& e : experiments) {
		if (have_non_sampled_names && e.model.non_sampled_parameters.size() != non_sampled_names.size())
			return Fail("the model of experiment \"" + e.name + "\" was generated for " + std::to_string(e.model.non_sampled_parameters.size()) +
			            " non-sampled parameters, AddNonSampledParameters named " + std::to_string(non_sampled_names.size()));
		double end_time = -std::numeric_limits<double>::infinity();
		for (const auto& ds : e.data) {
			const size_t T = ds.data.timepoints.size();
			if (T == 0 || ds.data.observed.size() != ds.data.num_replicates * T) return Fail("SetData() has not supplied a consistent data set for experiment \"" + e.name + "\"");
			if (!std::is_sorted(ds.data.timepoints.begin(), ds.data.timepoints.end())) return Fail("data set timepoints must be sorted");
			end_time = std::max(end_time, ds.data.timepoints.back());
		}
		for (const auto& name : e.set_species) // Experiment.cpp:497-500
			if (std::find(e.model.species_names.begin(), e.model.species_names.end(), name) == e.model.species_names.end())
				return Fail("set_species: \"" + name + "\" is not a simulated species of the model");
		// Cell::integration_step_cb (Cell.cpp:463-540): with these species in the model a cell's integration ends at a
		// threshold crossing (death) or is extended past anaphase -- events the device path does not have. The other
		// special species only record times for the per-cell data types.
		// Built: division at "cytokinesis" > 1 (divide_cells), death at "apoptosis" > 1, both without stored integration points.
		// "chromatid_separation" only moves the end of the integration when simulate_past_chromatid_separation_time is set
		// (refused above); the remaining special species only record times for the per-cell data types.
		end_time += e.trailing_simulation_time;
		// One integration per experiment, shared by its data sets (Experiment.cpp:190-214, 298-312): up to four data sets per
		// handle -- the first of a group owns it, the others ride along. Needs strictly increasing timepoints and the lane-group
		// kernel (N <= 96); otherwise, and with share_integration off, one handle per data set as before.
		bool can_share = share_integration && e.model.species_names.size() <= 96;
		for (const auto& ds : e.data)
			for (size_t i = 1; i < ds.data.timepoints.size(); i++) can_share = can_share && ds.data.timepoints[i] > ds.data.timepoints[i - 1];
		// a data set and its further markers (marker_of) always share a handle: they are ONE likelihood
		auto with_markers = [&](size_t k) { // number of entries of e.data that data set k and its markers occupy
			size_t n = 1;
			while (k + n < e.data.size() && e.data[k + n].marker_of == (long)k) n++;
			return n;
		};
		auto slots = [&](size_t k, size_t n) { // value-row blocks of the handle that entries k .. k + n - 1 need: one each + one per log-ratio denominator
			size_t v = n;
			for (size_t j = 0; j < n; j++) v += e.data[k + j].denominator_name.empty() ? 0 : 1;
			return v;
		};
		for (size_t k = 0; k < e.data.size();) {
			std::vector<DataSet*> followers;
			size_t group = with_markers(k);
			if (slots(k, group) > 4) return Fail("a data set with its further markers and log-ratio denominators needs more than the four value blocks of a handle");
			while (can_share && k + group < e.data.size() && slots(k, group + with_markers(k + group)) <= 4) group += with_markers(k + group);
			for (size_t j = 1; j < group; j++) {
				followers.push_back(&e.data[k + j]);
				if (e.data[k + j].marker_of >= 0) {
					const DataSet& owner = e.data[(size_t)e.data[k + j].marker_of];
					if (e.data[k + j].data.timepoints != owner.data.timepoints || e.data[k + j].data.num_replicates != owner.data.num_replicates)
						return Fail("SetData(): a marker shares the timepoints and the observed cells of the data set it belongs to");
				}
			}
			if (!CreateHandle(e, e.data[k], end_time, followers)) return false;
			k += group;
		}
	}
	return true;
}

bool CellPopulationLikelihoodB200::CreateHandle(Experiment& e, DataSet& ds, double simulation_end_time, const std::vector<DataSet*>& followers)
{
	const Model& model = e.model;
	const Data& data = ds.data;
	const size_t N = model.species_names.size(), nvar = varset->GetNumVariables(), T = data.timepoints.size(), D = e.variables.size();
	const size_t num_cells = e.num_cells;
	if (N == 0 || model.initial_conditions.size() != N || model.derivative_code.empty()) return Fail("SetModel() has not supplied the generated model");
	auto species_index = [&](const char* name) -> long {
		auto it = std::find(model.species_names.begin(), model.species_names.end(), name);
		return it == model.species_names.end() ? -1 : (long)(it - model.species_names.begin());
	};
	// Cell::Cell looks these up by name (Cell.cpp:40-55); a model without "cytokinesis" never divides
	const long cytokinesis = species_index("cytokinesis"), apoptosis = species_index("apoptosis");
	const bool divides = e.divide_cells && cytokinesis >= 0;
	if (D > 0 && (divides ? (e.sobol.size() < num_cells * D || e.sobol.size() % D != 0) : e.sobol.size() != num_cells * D))
		return Fail(divides ? "SetSobolTable(): a dividing population needs at least num_cells rows (the reference makes 100 x num_cells)"
		                    : "SetSobolTable(): expected num_cells x variability dimension entries");

	// species_name="a+b": the summed simulated species
	std::vector<size_t> obs;
	auto observed_species = [&](const std::string& names, std::vector<size_t>& out) {
		std::stringstream ss(names);
		std::string part;
		while (std::getline(ss, part, '+')) {
			auto it = std::find(model.species_names.begin(), model.species_names.end(), part);
			if (it == model.species_names.end()) return Fail("Species \"" + part + "\" of the data set is not a simulated species of the model");
			out.push_back((size_t)(it - model.species_names.begin()));
		}
		return true;
	};
	if (!observed_species(ds.species_name, obs)) return false;
	std::ostringstream d;
	d.precision(17);
	d << "num_species=" << N << ";num_constant_species=" << model.constant_species.size() << ";num_variables=" << nvar << ";num_non_sampled="
	  << model.non_sampled_parameters.size() << ";num_cells=" << num_cells << ";num_timepoints=" << T << ";num_replicates=" << data.num_replicates
	  << ";variability_dim=" << D << ";variability_distribution=" << e.distribution << ";solver_relative_tolerance=" << e.solver_rel_tol
	  << ";solver_absolute_tolerance=" << e.solver_abs_tol << ";solver_min_timestep=" << e.solver_min_timestep << ";solver_max_steps=" << e.solver_max_steps
	  << ";relative_to_time_average=" << (ds.relative_to_time_average ? 1 : 0) << ";stdev_relative_to_scale=" << (ds.stdev_relative_to_scale ? 1 : 0) << ";error_model=" << ds.error_model << ";weight=" << ds.weight
	  << ";missing_simulation_time_stdev=" << ds.missing_stdev << ";device=" << device << ";compile_only=" << (compile_only ? 1 : 0);
	if (ds.type != "time_course_population_average") d << ";data_kind=" << ds.type;
	if (ds.value_relative_to_timepoint_ix >= 0) d << ";value_relative_to_timepoint_ix=" << ds.value_relative_to_timepoint_ix;
	{
		bool wants_mitosis = ds.include_only_mitotic;
		for (const DataSet* f : followers) wants_mitosis = wants_mitosis || f->include_only_mitotic;
		if (wants_mitosis) { // Cell::Cell finds the species by name (Cell.cpp:40-55); without it no cell ever "enters mitosis"
			const long ne = species_index("nuclear_envelope");
			if (ne < 0) return Fail("include_only_cells_that_went_through_mitosis: the model has no species \"nuclear_envelope\" (Cell.cpp:487-492)");
			d << ";nuclear_envelope_species=" << ne;
		}
	}
	if (ds.include_only_mitotic) d << ";include_only_cells_that_went_through_mitosis=1";
	if (ds.use_only_nondivided) d << ";use_only_nondivided=1";
	if (ds.saturation_scale_ix >= 0) d << ";saturation_scale_ix=" << ds.saturation_scale_ix;
	if (ds.optimize_offset_scale)
		d << ";optimize_offset_scale=1;optimize_offset_min=" << ds.optimize_offset_min << ";optimize_offset_max=" << ds.optimize_offset_max
		  << ";optimize_scale_min=" << ds.optimize_scale_min << ";optimize_scale_max=" << ds.optimize_scale_max;
	if (simulation_end_time > data.timepoints.back()) d << ";simulation_end_time=" << simulation_end_time;
	if (std::isfinite(e.solver_max_timestep)) d << ";solver_max_timestep=" << e.solver_max_timestep;
	if (divides) {
		// Cell::SetInitialConditionsFromOtherCell (Cell.cpp:127-133) resets these seven species by name and does not check that
		// they exist: a dividing model without one of them is an error here
		d << ";divide_cells=1;max_cells=" << e.max_cells << ";cytokinesis_species=" << cytokinesis << ";division_reset_species=";
		const char* reset_names[7] = { "cytokinesis", "nuclear_envelope", "G1S_break", "G2_break", "spindle_components", "assembled_spindle", "chromatid_separation" };
		for (int k = 0; k < 7; k++) {
			const long ix = species_index(reset_names[k]);
			if (ix < 0) return Fail(std::string("divide_cells: the model has no species \"") + reset_names[k] + "\" (Cell.cpp:127-133 resets it in every daughter)");
			d << (k ? "+" : "") << ix;
		}
	}
	if (apoptosis >= 0) d << ";apoptosis_species=" << apoptosis;
	auto ref = [&](const char* name, const ValueRef& r) {
		if (r.ix >= 0) d << ";" << name << "_ix=" << r.ix;
		else d << ";" << name << "=" << r.fixed;
	};
	ref("entry_time", e.entry_time);
	ref("stdev", ds.stdev);
	if (ds.have_proportional_stdev) ref("proportional_stdev", ds.proportional_stdev);
	ref("offset", ds.offset);
	ref("scale", ds.scale);
	d << ";obs_species=";
	for (size_t k = 0; k < obs.size(); k++) d << (k ? "+" : "") << obs[k];
	// the experiment's further data sets that share this handle's integration: the data-set keys again, suffixed @1, @2, ...
	// log-ratio denominators: further value-row blocks of the handle after the followers, each flagged with the entry it divides
	std::vector<std::pair<size_t, size_t>> denominators; // (entry of the handle, species index)
	for (size_t j = 0; j <= followers.size(); j++) {
		const DataSet& owner = (j == 0) ? ds : *followers[j - 1];
		if (owner.denominator_name.empty()) continue;
		std::vector<size_t> den;
		if (!observed_species(owner.denominator_name, den)) return false;
		denominators.emplace_back(j, den[0]);
	}
	if (!followers.empty() || !denominators.empty()) d << ";num_data_sets=" << (1 + followers.size() + denominators.size());
	for (size_t q = 0; q < denominators.size(); q++) {
		const std::string sfx = "@" + std::to_string(1 + followers.size() + q);
		d << ";denominator_of" << sfx << "=" << denominators[q].first << ";num_timepoints" << sfx << "=" << T << ";num_replicates" << sfx << "=1;obs_species" << sfx << "="
		  << denominators[q].second;
	}
	for (size_t j = 0; j < followers.size(); j++) {
		const DataSet& f = *followers[j];
		const std::string sfx = "@" + std::to_string(j + 1);
		std::vector<size_t> fobs;
		if (!observed_species(f.species_name, fobs)) return false;
		d << ";num_timepoints" << sfx << "=" << f.data.timepoints.size() << ";num_replicates" << sfx << "=" << f.data.num_replicates << ";relative_to_time_average" << sfx
		  << "=" << (f.relative_to_time_average ? 1 : 0) << ";stdev_relative_to_scale" << sfx << "=" << (f.stdev_relative_to_scale ? 1 : 0) << ";error_model" << sfx << "="
		  << f.error_model << ";weight" << sfx << "=" << f.weight << ";missing_simulation_time_stdev" << sfx << "=" << f.missing_stdev;
		if (f.type != "time_course_population_average") d << ";data_kind" << sfx << "=" << f.type;
		if (f.value_relative_to_timepoint_ix >= 0) d << ";value_relative_to_timepoint_ix" << sfx << "=" << f.value_relative_to_timepoint_ix;
		if (f.include_only_mitotic) d << ";include_only_cells_that_went_through_mitosis" << sfx << "=1";
		if (f.use_only_nondivided) d << ";use_only_nondivided" << sfx << "=1";
		if (f.saturation_scale_ix >= 0) d << ";saturation_scale_ix" << sfx << "=" << f.saturation_scale_ix;
		if (f.marker_of >= 0) { // position of the owner inside this handle: 0 = ds, q + 1 = followers[q]
			long at = (&e.data[(size_t)f.marker_of] == &ds) ? 0 : -1;
			for (size_t q = 0; q < followers.size() && at < 0; q++)
				if (followers[q] == &e.data[(size_t)f.marker_of]) at = (long)q + 1;
			if (at < 0) return Fail("internal: a marker and its data set ended up in different handles");
			d << ";marker_of" << sfx << "=" << at;
		}
		if (f.optimize_offset_scale)
			d << ";optimize_offset_scale" << sfx << "=1;optimize_offset_min" << sfx << "=" << f.optimize_offset_min << ";optimize_offset_max" << sfx << "=" << f.optimize_offset_max
			  << ";optimize_scale_min" << sfx << "=" << f.optimize_scale_min << ";optimize_scale_max" << sfx << "=" << f.optimize_scale_max;
		auto fref = [&](const char* name, const ValueRef& r) {
			if (r.ix >= 0) d << ";" << name << "_ix" << sfx << "=" << r.ix;
			else d << ";" << name << sfx << "=" << r.fixed;
		};
		fref("stdev", f.stdev);
		if (f.have_proportional_stdev) fref("proportional_stdev", f.proportional_stdev);
		fref("offset", f.offset);
		fref("scale", f.scale);
		d << ";obs_species" << sfx << "=";
		for (size_t k = 0; k < fobs.size(); k++) d << (k ? "+" : "") << fobs[k];
	}
	if (!e.treatment_species_name.empty()) {
		auto it = std::find(model.constant_species_names.begin(), model.constant_species_names.end(), e.treatment_species_name);
		if (it == model.constant_species_names.end()) return Fail("Treatment species \"" + e.treatment_species_name + "\" is not a constant species of the model");
		d << ";treatment_species=" << (it - model.constant_species_names.begin());
	}
	ds.descriptor = d.str();
	void*& handle = ds.handle;
	if (bcm3b200_create("cell_population", ds.descriptor.data(), ds.descriptor.size(), 1, &handle) != BCM3B200_OK) return Fail(bcm3b200_last_error());

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
	for (size_t j = 0; j < followers.size(); j++) {
		const Data& fd = followers[j]->data;
		const std::string sfx = "@" + std::to_string(j + 1);
		if (!set(("timepoints" + sfx).c_str(), fd.timepoints, { fd.timepoints.size() }) ||
		    !set(("observed" + sfx).c_str(), fd.observed, { fd.num_replicates, fd.timepoints.size() }))
			return false;
	}
	for (size_t q = 0; q < denominators.size(); q++) { // value rows only: the timepoints of its numerator, no observations
		const Data& od = (denominators[q].first == 0) ? data : followers[denominators[q].first - 1]->data;
		const std::string sfx = "@" + std::to_string(1 + followers.size() + q);
		const std::vector<double> none(od.timepoints.size(), 0.0);
		if (!set(("timepoints" + sfx).c_str(), od.timepoints, { od.timepoints.size() }) || !set(("observed" + sfx).c_str(), none, { 1, od.timepoints.size() })) return false;
	}
	if (!e.treatment_species_name.empty() && !e.treatment_times.empty() && !set("treatment_times", e.treatment_times, { e.treatment_times.size() })) return false;
	if (D > 0) {
		std::vector<double> rows(D * 6);
		for (size_t i = 0; i < D; i++) {
			const VarEntry& ve = e.variables[i];
			size_t target = 0;
			if (ve.entry_time) {
			} else if (ve.is_ic) {
				auto it = std::find(model.species_names.begin(), model.species_names.end(), ve.target);
				if (it == model.species_names.end()) return Fail("Variability initial_condition_species \"" + ve.target + "\" is not a simulated species");
				target = (size_t)(it - model.species_names.begin());
			} else {
				target = varset->GetVariableIndex(ve.target);
				if (target == std::numeric_limits<size_t>::max()) return Fail("Variability model_parameter \"" + ve.target + "\" is not a sampled variable");
			}
			double* r = rows.data() + i * 6;
			r[0] = (ve.entry_time ? 2.0 : ve.is_ic ? 1.0 : 0.0) + (ve.only_initial_cells ? 4.0 : 0.0);
			r[1] = (double)target;
			r[2] = (double)ve.apply;
			r[3] = (double)ve.scale.ix;
			r[4] = ve.scale.fixed;
			r[5] = ve.negate ? 1.0 : 0.0;
		}
		if (!set("sobol", e.sobol, { divides ? e.sobol.size() / D : num_cells, D }) || !set("variability", rows, { D, 6 })) return false;
		if (e.distribution == "full_gaussian" && D > 1) {
			// covariance values are variables named <covar_base_name><k+1>_<i+1>, k < i (VariabilityDescription.cpp:203-216)
			std::vector<double> cov(D * (D - 1));
			for (size_t i = 1; i < D; i++)
				for (size_t k = 0; k < i; k++) {
					const std::string name = e.covar_base_name + std::to_string(k + 1) + "_" + std::to_string(i + 1);
					ValueRef r;
					if (!Resolve(name, r, "covariance")) return false;
					const size_t ix = (i - 1) * i / 2 + k;
					cov[2 * ix] = (double)r.ix;
					cov[2 * ix + 1] = r.fixed;
				}
			if (!set("variability_covariance", cov, { D * (D - 1) / 2, 2 })) return false;
		}
	}
	if (bcm3b200_set_text(handle, "derivative_code", model.derivative_code.data(), model.derivative_code.size()) != BCM3B200_OK) return Fail(bcm3b200_last_error());
	if (bcm3b200_finalize(handle) != BCM3B200_OK) return Fail(bcm3b200_last_error());
	return true;
}

// Experiment::Simulate (Experiment.cpp:635-642): transformed[parameter] = transformed[replacement] for the cells of this
// experiment; with equal transforms that is the same as replacing the untransformed column
const double* CellPopulationLikelihoodB200::ExperimentValues(const Experiment& e, const double* values, size_t C, size_t nvar, std::vector<double>& replaced)
{
	if (e.specific_parameters.empty()) return values;
	replaced.assign(values, values + C * nvar);
	for (size_t c = 0; c < C; c++)
		for (const auto& sp : e.specific_parameters) replaced[c * nvar + sp.first] = replaced[c * nvar + sp.second];
	return replaced.data();
}

bool CellPopulationLikelihoodB200::EvaluateLogProbability(size_t, const bcm3::VectorReal& values, Real& logp)
{
	// CellPopulationLikelihood.cpp:84-98 / Experiment.cpp:346-355: sums start from 0.0, data sets inside experiments
	logp = 0.0;
	for (auto& e : experiments) {
		Real experiment_logp = 0.0;
		std::vector<double> scratch; // local: this entry may be called from several sampling threads (IsReentrant)
		const double* v = ExperimentValues(e, values.data(), 1, values.size(), scratch);
		for (auto& ds : e.data) {
			if (!ds.handle) continue; // rides along with the first data set of its group
			int st = 0;
			Real dl_logp = 0.0;
			if (bcm3b200_evaluate_batch(ds.handle, 1, values.size(), v, &dl_logp, &st) != BCM3B200_OK) return Fail(bcm3b200_last_error());
			experiment_logp += dl_logp;
		}
		logp += experiment_logp;
	}
	return true; // -inf (a failed cell, Experiment.cpp:356-358) is a legal value; NaN is turned into an error by the sampler
}

bool CellPopulationLikelihoodB200::EvaluateLogProbabilityBatch(const bcm3::MatrixReal& values, bcm3::VectorReal& logp)
{
	const size_t C = values.cols();
	logp.assign(C, -bcm3::kInf);
	status.assign(C, 0);
	if (C == 0) return true;
	std::fill(logp.begin(), logp.end(), 0.0);
	part.resize(C);
	std::vector<double> experiment_logp(C);
	for (auto& e : experiments) {
		std::fill(experiment_logp.begin(), experiment_logp.end(), 0.0);
		const double* v = ExperimentValues(e, values.data.data(), C, values.rows(), replaced);
		for (auto& ds : e.data) {
			if (!ds.handle) continue; // rides along with the first data set of its group
			if (bcm3b200_evaluate_batch(ds.handle, C, values.rows(), v, part.data(), status.data()) != BCM3B200_OK) return Fail(bcm3b200_last_error());
			for (size_t c = 0; c < C; c++) experiment_logp[c] += part[c];
		}
		for (size_t c = 0; c < C; c++) logp[c] += experiment_logp[c];
	}
	return true;
}
