// C entry points of the host mirror (libbcm3host.so) so that the Python test-suite can drive the C++ sampler,
// factory and likelihood classes exactly as a C++ program would.
#include <cstring>

#include "CellPopulationLikelihoodB200.h"
#include "LikelihoodFactory.h"
#include "LikelihoodPopPKTrajectoryB200.h"
#include "PharmacoLikelihoodPopulationB200.h"
#include "SamplerPT.h"

using namespace bcm3;

namespace {

void set_err(char* err, size_t errlen, const std::string& m)
{
	if (err && errlen) {
		strncpy(err, m.c_str(), errlen - 1);
		err[errlen - 1] = 0;
	}
}

struct Setup {
	std::shared_ptr<VariableSet> varset;
	std::shared_ptr<Prior> prior;
	std::shared_ptr<Likelihood> likelihood;
};

bool make_setup(const char* prior_xml, const char* likelihood_xml, Setup& s, std::string& error)
{
	XmlNode root;
	if (!ParseXml(prior_xml, root, error)) return false;
	const XmlNode* node = root.child("prior");
	if (!node) node = root.child("variableset");
	if (!node) {
		error = "Incorrect prior XML format";
		return false;
	}
	s.varset = std::make_shared<VariableSet>();
	if (!s.varset->LoadFromNode(*node)) {
		error = "Error parsing variable file";
		return false;
	}
	s.prior = Prior::CreateFromNode(*node, s.varset);
	if (!s.prior) {
		error = "Error creating prior";
		return false;
	}
	s.likelihood = LikelihoodFactory::CreateLikelihoodFromText(likelihood_xml, s.varset, 1, 1, &error);
	return s.likelihood != nullptr;
}

// optional sinks of a run: a SampleHandlerTSV file and a SampleHandlerStoreMaxAPosteriori whose result goes to map_out
// [lposterior, llikelihood, values...]
struct RunSinks {
	const char* tsv_file = nullptr;
	double* map_out = nullptr;
};

int run(Setup& st, const char* config_text, int batched, unsigned long long seed, double* out, size_t max_rows, size_t* num_rows,
        size_t* stats, char* err, size_t errlen, const RunSinks* sinks = nullptr)
{
	SamplerPTSettings settings;
	std::string error;
	if (!settings.LoadFromConfigText(config_text ? config_text : "", &error)) {
		set_err(err, errlen, error);
		return -1;
	}
	settings.batched = batched != 0;
	settings.rngseed = seed;
	SamplerPT sampler(settings);
	sampler.SetVariableSet(st.varset);
	sampler.SetPrior(st.prior);
	sampler.SetLikelihood(st.likelihood);
	std::shared_ptr<SampleHandlerTSV> tsv;
	std::shared_ptr<SampleHandlerStoreMaxAPosteriori> map;
	if (sinks && sinks->tsv_file && sinks->tsv_file[0]) { // as bcminf installs its NetCDF handler (bcminf/main.cpp:96-99)
		tsv = std::make_shared<SampleHandlerTSV>();
		tsv->SetFile(sinks->tsv_file);
		std::vector<std::string> names;
		for (size_t i = 0; i < st.varset->GetNumVariables(); i++) names.push_back(st.varset->GetVariableName(i));
		if (!tsv->Initialize(settings.num_samples, names, VectorReal())) {
			set_err(err, errlen, std::string("Failed to open output file \"") + sinks->tsv_file + "\"");
			return -4;
		}
		sampler.AddSampleHandler(tsv);
	}
	if (sinks && sinks->map_out) {
		map = std::make_shared<SampleHandlerStoreMaxAPosteriori>();
		sampler.AddSampleHandler(map);
	}
	if (!sampler.Initialize() || !sampler.Run()) {
		set_err(err, errlen, sampler.LastError());
		return -2;
	}
	const size_t nvar = st.varset->GetNumVariables();
	const auto& samples = sampler.GetSamples();
	size_t n = std::min(max_rows, samples.size());
	for (size_t r = 0; r < n; r++) {
		double* row = out + r * (nvar + 3);
		row[0] = samples[r].temperature;
		row[1] = samples[r].lprior;
		row[2] = samples[r].llh;
		for (size_t i = 0; i < nvar; i++) row[3 + i] = samples[r].values[i];
	}
	if (num_rows) *num_rows = samples.size();
	if (map) {
		sinks->map_out[0] = map->GetMAPlposterior();
		sinks->map_out[1] = map->GetMAPllikelihood();
		for (size_t i = 0; i < nvar && i < (size_t)map->GetMAP().size(); i++) sinks->map_out[2 + i] = map->GetMAP()[i];
	}
	if (stats) {
		stats[0] = sampler.GetNumLikelihoodEvaluations();
		stats[1] = sampler.GetNumBatchedCalls();
		stats[2] = sampler.GetTemperatures().size();
		stats[3] = sampler.GetBlocks(sampler.GetTemperatures().size() - 1).size(); // variable blocks of the posterior chain at the end
	}
	return 0;
}

} // namespace

extern "C" {

// Parallel-tempered run with a host likelihood (banana). out rows: [temperature, lprior, llh, values...].
int bcm3host_run_pt(const char* prior_xml, const char* likelihood_xml, const char* config_text, int batched, unsigned long long seed,
                    double* out, size_t max_rows, size_t* num_rows, size_t* stats, char* err, size_t errlen)
{
	Setup st;
	std::string error;
	if (!make_setup(prior_xml, likelihood_xml, st, error)) {
		set_err(err, errlen, error);
		return -1;
	}
	if (!st.likelihood->PostInitialize()) {
		set_err(err, errlen, "PostInitialize failed");
		return -1;
	}
	return run(st, config_text, batched, seed, out, max_rows, num_rows, stats, err, errlen);
}

// The same run with the reference's sample sinks attached: SampleHandlerTSV writing to tsv_file (may be NULL) and
// SampleHandlerStoreMaxAPosteriori reporting into map_out [2 + nvar] = lposterior, llikelihood, values (may be NULL).
int bcm3host_run_pt_with_handlers(const char* prior_xml, const char* likelihood_xml, const char* config_text, int batched, unsigned long long seed,
                                  const char* tsv_file, double* map_out, double* out, size_t max_rows, size_t* num_rows, size_t* stats, char* err,
                                  size_t errlen)
{
	Setup st;
	std::string error;
	if (!make_setup(prior_xml, likelihood_xml, st, error)) {
		set_err(err, errlen, error);
		return -1;
	}
	if (!st.likelihood->PostInitialize()) {
		set_err(err, errlen, "PostInitialize failed");
		return -1;
	}
	RunSinks sinks;
	sinks.tsv_file = tsv_file;
	sinks.map_out = map_out;
	return run(st, config_text, batched, seed, out, max_rows, num_rows, stats, err, errlen, &sinks);
}

// Same with the GPU-backed pop_pk_trajectory likelihood; the trial arrays stand in for the NetCDF file.
int bcm3host_run_pt_poppk(const char* prior_xml, const char* likelihood_xml, const char* config_text, int batched, unsigned long long seed,
                          size_t P, size_t T, const double* time, const double* obs, const double* dose, const double* dosing_interval,
                          const double* dose_after, const double* dose_change_time, const double* intermittent, const double* interruptions,
                          int device, int device_count, double* out, size_t max_rows, size_t* num_rows, size_t* stats, char* err, size_t errlen)
{
	Setup st;
	std::string error;
	if (!make_setup(prior_xml, likelihood_xml, st, error)) {
		set_err(err, errlen, error);
		return -1;
	}
	auto* ll = dynamic_cast<LikelihoodPopPKTrajectoryB200*>(st.likelihood.get());
	if (!ll) {
		set_err(err, errlen, "likelihood.xml is not of type pop_pk_trajectory");
		return -1;
	}
	LikelihoodPopPKTrajectoryB200::TrialData td;
	td.time.assign(time, time + T);
	td.observed_concentration.assign(obs, obs + P * T);
	td.dose.assign(dose, dose + P);
	td.dosing_interval.assign(dosing_interval, dosing_interval + P);
	td.dose_after_dose_change.assign(dose_after, dose_after + P);
	td.dose_change_time.assign(dose_change_time, dose_change_time + P);
	td.intermittent.assign(intermittent, intermittent + P);
	td.treatment_interruptions.assign(interruptions, interruptions + P * 29);
	ll->SetTrialData(td);
	ll->SetDevices(device, device_count);
	if (!ll->PostInitialize()) {
		set_err(err, errlen, ll->LastError());
		return -3;
	}
	return run(st, config_text, batched, seed, out, max_rows, num_rows, stats, err, errlen);
}

// likelihood.xml type="pharmaco_population" -> LikelihoodFactory -> PharmacoLikelihoodPopulationB200 with the trial arrays the NetCDF
// reader would supply; evaluates C variable vectors batched or one by one
int bcm3host_pharmaco_evaluate(const char* prior_xml, const char* likelihood_xml, size_t P, size_t T, const double* time, const double* obs,
                               const double* dose, const double* dosing_interval, const double* dose_after, const double* dose_change_time,
                               const double* intermittent, const double* interruptions, int device, const double* values, size_t C, int batched,
                               double* logp, char* err, size_t errlen)
{
	Setup st;
	std::string error;
	if (!make_setup(prior_xml, likelihood_xml, st, error)) {
		set_err(err, errlen, error);
		return -1;
	}
	auto* ll = dynamic_cast<PharmacoLikelihoodPopulationB200*>(st.likelihood.get());
	if (!ll) {
		set_err(err, errlen, "likelihood.xml is not of type pharmaco_population");
		return -1;
	}
	PharmacoLikelihoodPopulationB200::TrialData td;
	td.time.assign(time, time + T);
	td.observed_concentration.assign(obs, obs + P * T);
	td.dose.assign(dose, dose + P);
	td.dosing_interval.assign(dosing_interval, dosing_interval + P);
	td.dose_after_dose_change.assign(dose_after, dose_after + P);
	td.dose_change_time.assign(dose_change_time, dose_change_time + P);
	td.intermittent.assign(intermittent, intermittent + P);
	td.treatment_interruptions.assign(interruptions, interruptions + P * 29);
	ll->SetTrialData(td);
	ll->SetDevice(device);
	if (!ll->PostInitialize()) {
		set_err(err, errlen, ll->LastError());
		return -3;
	}
	const size_t nvar = st.varset->GetNumVariables();
	if (batched) {
		MatrixReal m(nvar, C);
		std::copy(values, values + nvar * C, m.data.begin());
		VectorReal lp;
		if (!st.likelihood->EvaluateLogProbabilityBatch(m, lp)) {
			set_err(err, errlen, ll->LastError());
			return -2;
		}
		std::copy(lp.begin(), lp.end(), logp);
	} else {
		for (size_t c = 0; c < C; c++) {
			VectorReal v(values + c * nvar, values + (c + 1) * nvar);
			if (!st.likelihood->EvaluateLogProbability(0, v, logp[c])) {
				set_err(err, errlen, ll->LastError());
				return -2;
			}
		}
	}
	return 0;
}

// Factory + plugin surface check: evaluate C variable vectors (values[C][nvar]) through EvaluateLogProbability (batched = 0)
// or EvaluateLogProbabilityBatch (batched = 1) of the likelihood named by likelihood.xml (host likelihoods only).
int bcm3host_evaluate(const char* prior_xml, const char* likelihood_xml, const double* values, size_t C, int batched, double* logp, char* err,
                      size_t errlen)
{
	Setup st;
	std::string error;
	if (!make_setup(prior_xml, likelihood_xml, st, error)) {
		set_err(err, errlen, error);
		return -1;
	}
	const size_t nvar = st.varset->GetNumVariables();
	if (batched) {
		MatrixReal m(nvar, C);
		std::copy(values, values + nvar * C, m.data.begin());
		VectorReal lp;
		if (!st.likelihood->EvaluateLogProbabilityBatch(m, lp)) return -2;
		std::copy(lp.begin(), lp.end(), logp);
	} else {
		for (size_t c = 0; c < C; c++) {
			VectorReal v(values + c * nvar, values + (c + 1) * nvar);
			if (!st.likelihood->EvaluateLogProbability(0, v, logp[c])) return -2;
		}
	}
	return 0;
}

// likelihood.xml -> LikelihoodFactory -> CellPopulationLikelihoodB200, with the generated model, data set and quasi-random
// table that the SBML / NetCDF readers would supply; PostInitialize included
static bool make_cellpop(const char* prior_xml, const char* likelihood_xml, const char* derivative_code, size_t N, const char* const* species_names,
                         const double* initial_conditions, size_t Nc, const double* constant_species, size_t T, size_t R, const double* timepoints,
                         const double* observed, size_t sobol_count, const double* sobol, int device, int compile_only, Setup& st,
                         CellPopulationLikelihoodB200*& ll, char* err, size_t errlen)
{
	std::string error;
	if (!make_setup(prior_xml, likelihood_xml, st, error)) {
		set_err(err, errlen, error);
		return false;
	}
	ll = dynamic_cast<CellPopulationLikelihoodB200*>(st.likelihood.get());
	if (!ll) {
		set_err(err, errlen, "likelihood.xml is not of type cell_population");
		return false;
	}
	CellPopulationLikelihoodB200::Model m;
	m.derivative_code = derivative_code;
	for (size_t i = 0; i < N; i++) m.species_names.push_back(species_names[i]);
	m.initial_conditions.assign(initial_conditions, initial_conditions + N);
	m.constant_species.assign(constant_species, constant_species + Nc);
	for (size_t i = 0; i < Nc; i++) m.constant_species_names.push_back("c" + std::to_string(i));
	ll->SetModel(m);
	CellPopulationLikelihoodB200::Data d;
	d.timepoints.assign(timepoints, timepoints + T);
	d.observed.assign(observed, observed + R * T);
	d.num_replicates = R;
	ll->SetData(d);
	ll->SetSobolTable(std::vector<double>(sobol, sobol + sobol_count));
	ll->SetDevice(device, compile_only != 0);
	if (!ll->PostInitialize()) {
		set_err(err, errlen, ll->LastError());
		return false;
	}
	return true;
}

// cell_population through the plugin surface: likelihood.xml + the generated model, data set and quasi-random table that
// the SBML / NetCDF readers would supply; evaluates values[C][nvar] batched (1) or chain by chain (0). compile_only = 1
// stops after PostInitialize (CPU container: the model library is compiled, nothing runs); the descriptor handed to the
// C ABI is returned in desc_out.
int bcm3host_cellpop_evaluate(const char* prior_xml, const char* likelihood_xml, const char* derivative_code, size_t N,
                              const char* const* species_names, const double* initial_conditions, size_t Nc, const double* constant_species,
                              size_t T, size_t R, const double* timepoints, const double* observed, size_t sobol_count, const double* sobol,
                              int device, int compile_only, const double* values, size_t C, int batched, double* logp, char* desc_out,
                              size_t desc_len, char* err, size_t errlen)
{
	Setup st;
	CellPopulationLikelihoodB200* ll = nullptr;
	if (!make_cellpop(prior_xml, likelihood_xml, derivative_code, N, species_names, initial_conditions, Nc, constant_species, T, R, timepoints, observed,
	                  sobol_count, sobol, device, compile_only, st, ll, err, errlen))
		return -3;
	set_err(desc_out, desc_len, ll->GetDescriptor());
	if (compile_only) return 0;
	const size_t nvar = st.varset->GetNumVariables();
	if (batched) {
		MatrixReal mat(nvar, C);
		std::copy(values, values + nvar * C, mat.data.begin());
		VectorReal lp;
		if (!st.likelihood->EvaluateLogProbabilityBatch(mat, lp)) {
			set_err(err, errlen, ll->LastError());
			return -2;
		}
		std::copy(lp.begin(), lp.end(), logp);
	} else {
		for (size_t c = 0; c < C; c++) {
			VectorReal v(values + c * nvar, values + (c + 1) * nvar);
			if (!st.likelihood->EvaluateLogProbability(0, v, logp[c])) {
				set_err(err, errlen, ll->LastError());
				return -2;
			}
		}
	}
	return 0;
}

// Parallel-tempered run on the GPU-backed cell_population likelihood (batched = one call per mutate round).
int bcm3host_run_pt_cellpop(const char* prior_xml, const char* likelihood_xml, const char* config_text, int batched, unsigned long long seed,
                            const char* derivative_code, size_t N, const char* const* species_names, const double* initial_conditions, size_t Nc,
                            const double* constant_species, size_t T, size_t R, const double* timepoints, const double* observed, size_t sobol_count,
                            const double* sobol, int device, double* out, size_t max_rows, size_t* num_rows, size_t* stats, char* err, size_t errlen)
{
	Setup st;
	CellPopulationLikelihoodB200* ll = nullptr;
	if (!make_cellpop(prior_xml, likelihood_xml, derivative_code, N, species_names, initial_conditions, Nc, constant_species, T, R, timepoints, observed,
	                  sobol_count, sobol, device, 0, st, ll, err, errlen))
		return -3;
	return run(st, config_text, batched, seed, out, max_rows, num_rows, stats, err, errlen);
}

// ---- cell_population with several experiments / data sets: a session object, filled the way the reference's readers would ----
struct CellpopSession {
	Setup st;
	CellPopulationLikelihoodB200* ll = nullptr;
};

void* bcm3host_cellpop_open(const char* prior_xml, const char* likelihood_xml, char* err, size_t errlen)
{
	auto* s = new CellpopSession;
	std::string error;
	if (!make_setup(prior_xml, likelihood_xml, s->st, error)) {
		set_err(err, errlen, error);
		delete s;
		return nullptr;
	}
	s->ll = dynamic_cast<CellPopulationLikelihoodB200*>(s->st.likelihood.get());
	if (!s->ll) {
		set_err(err, errlen, "likelihood.xml is not of type cell_population");
		delete s;
		return nullptr;
	}
	return s;
}

void bcm3host_cellpop_close(void* session) { delete static_cast<CellpopSession*>(session); }

// number of experiments; num_data[i] = data sets of experiment i (up to max_experiments entries)
size_t bcm3host_cellpop_layout(void* session, size_t* num_data, size_t max_experiments)
{
	auto* s = static_cast<CellpopSession*>(session);
	const size_t n = s->ll->GetNumExperiments();
	for (size_t i = 0; i < n && i < max_experiments; i++) num_data[i] = s->ll->GetNumDataSets(i);
	return n;
}

// experiment < 0: every experiment gets this model
int bcm3host_cellpop_set_model_ns(void* session, long experiment, const char* derivative_code, size_t N, const char* const* species_names,
                                  const double* initial_conditions, size_t Nc, const double* constant_species, size_t Nn, const double* non_sampled)
{
	auto* s = static_cast<CellpopSession*>(session);
	if (experiment >= (long)s->ll->GetNumExperiments()) return -1;
	CellPopulationLikelihoodB200::Model m;
	m.derivative_code = derivative_code;
	for (size_t i = 0; i < N; i++) m.species_names.push_back(species_names[i]);
	m.initial_conditions.assign(initial_conditions, initial_conditions + N);
	m.constant_species.assign(constant_species, constant_species + Nc);
	for (size_t i = 0; i < Nc; i++) m.constant_species_names.push_back("c" + std::to_string(i));
	m.non_sampled_parameters.assign(non_sampled, non_sampled + Nn);
	if (experiment < 0) s->ll->SetModel(m);
	else s->ll->SetModel((size_t)experiment, m);
	return 0;
}

int bcm3host_cellpop_set_model(void* session, long experiment, const char* derivative_code, size_t N, const char* const* species_names,
                               const double* initial_conditions, size_t Nc, const double* constant_species)
{
	auto* s = static_cast<CellpopSession*>(session);
	if (experiment >= (long)s->ll->GetNumExperiments()) return -1;
	CellPopulationLikelihoodB200::Model m;
	m.derivative_code = derivative_code;
	for (size_t i = 0; i < N; i++) m.species_names.push_back(species_names[i]);
	m.initial_conditions.assign(initial_conditions, initial_conditions + N);
	m.constant_species.assign(constant_species, constant_species + Nc);
	for (size_t i = 0; i < Nc; i++) m.constant_species_names.push_back("c" + std::to_string(i));
	if (experiment < 0) s->ll->SetModel(m);
	else s->ll->SetModel((size_t)experiment, m);
	return 0;
}

// Likelihood::AddNonSampledParameters / SetNonSampledParameters / OutputEvaluationStatistics through the session
int bcm3host_cellpop_add_non_sampled(void* session, size_t count, const char* const* names)
{
	auto* s = static_cast<CellpopSession*>(session);
	std::vector<std::string> v;
	for (size_t i = 0; i < count; i++) v.push_back(names[i]);
	return s->ll->AddNonSampledParameters(v) ? 0 : -1;
}

int bcm3host_cellpop_set_non_sampled(void* session, size_t count, const double* values)
{
	auto* s = static_cast<CellpopSession*>(session);
	s->ll->SetNonSampledParameters(bcm3::VectorReal(values, values + count));
	return 0;
}

int bcm3host_cellpop_output_statistics(void* session, const char* path)
{
	static_cast<CellpopSession*>(session)->ll->OutputEvaluationStatistics(path);
	return 0;
}

int bcm3host_cellpop_set_data(void* session, size_t experiment, size_t data_set, size_t T, size_t R, const double* timepoints, const double* observed)
{
	auto* s = static_cast<CellpopSession*>(session);
	if (experiment >= s->ll->GetNumExperiments() || data_set >= s->ll->GetNumDataSets(experiment)) return -1;
	CellPopulationLikelihoodB200::Data d;
	d.timepoints.assign(timepoints, timepoints + T);
	d.observed.assign(observed, observed + R * T);
	d.num_replicates = R;
	s->ll->SetData(experiment, data_set, d);
	return 0;
}

int bcm3host_cellpop_set_sobol(void* session, size_t experiment, size_t count, const double* sobol)
{
	auto* s = static_cast<CellpopSession*>(session);
	if (experiment >= s->ll->GetNumExperiments()) return -1;
	s->ll->SetSobolTable(experiment, std::vector<double>(sobol, sobol + count));
	return 0;
}

int bcm3host_cellpop_post_initialize(void* session, int device, int compile_only, char* err, size_t errlen)
{
	auto* s = static_cast<CellpopSession*>(session);
	s->ll->SetDevice(device, compile_only != 0);
	if (!s->ll->PostInitialize()) {
		set_err(err, errlen, s->ll->LastError());
		return -3;
	}
	return 0;
}

// 0: one handle (one integration of the cells) per <data> element; 1 (default): one per experiment, shared by its data sets
void bcm3host_cellpop_share_integration(void* session, int share) { static_cast<CellpopSession*>(session)->ll->SetShareIntegration(share != 0); }

int bcm3host_cellpop_descriptor(void* session, size_t experiment, size_t data_set, char* out, size_t len)
{
	auto* s = static_cast<CellpopSession*>(session);
	if (experiment >= s->ll->GetNumExperiments() || data_set >= s->ll->GetNumDataSets(experiment)) return -1;
	set_err(out, len, s->ll->GetDescriptor(experiment, data_set));
	return 0;
}

// k-th <set_parameter> of an experiment; returns the number of them
size_t bcm3host_cellpop_fixed_parameter(void* session, size_t experiment, size_t k, char* name_out, size_t len, double* value)
{
	auto* s = static_cast<CellpopSession*>(session);
	if (experiment >= s->ll->GetNumExperiments()) return 0;
	const auto& fp = s->ll->GetFixedParameters(experiment);
	if (k < fp.size()) {
		set_err(name_out, len, fp[k].first);
		*value = fp[k].second;
	}
	return fp.size();
}

int bcm3host_cellpop_session_evaluate(void* session, const double* values, size_t C, int batched, double* logp, char* err, size_t errlen)
{
	auto* s = static_cast<CellpopSession*>(session);
	const size_t nvar = s->st.varset->GetNumVariables();
	if (batched) {
		MatrixReal mat(nvar, C);
		std::copy(values, values + nvar * C, mat.data.begin());
		VectorReal lp;
		if (!s->st.likelihood->EvaluateLogProbabilityBatch(mat, lp)) {
			set_err(err, errlen, s->ll->LastError());
			return -2;
		}
		std::copy(lp.begin(), lp.end(), logp);
	} else {
		for (size_t c = 0; c < C; c++) {
			VectorReal v(values + c * nvar, values + (c + 1) * nvar);
			if (!s->st.likelihood->EvaluateLogProbability(0, v, logp[c])) {
				set_err(err, errlen, s->ll->LastError());
				return -2;
			}
		}
	}
	return 0;
}

// ---- GaussianMixture (the fit behind proposal_type=gaussian_mixture) for tests ----
// samples[n][D] row-major. Outputs: weights[K], means[K][D], covariances[K][D][D], stats = { AIC, log-likelihood }.
// TreeClusterCompleteLinkage for tests: cluster[i] = index of item i's block (blocks in the order the sampler would use them)
int bcm3host_tree_cluster(const double* distance, size_t n, double cut_height, int* cluster)
{
	MatrixReal d(n, n);
	for (size_t j = 0; j < n; j++)
		for (size_t i = 0; i < n; i++) d(i, j) = distance[i * n + j];
	const auto blocks = TreeClusterCompleteLinkage(d, cut_height);
	for (size_t i = 0; i < n; i++) cluster[i] = -1;
	for (size_t b = 0; b < blocks.size(); b++)
		for (size_t item : blocks[b]) cluster[item] = (int)b;
	return (int)blocks.size();
}

int bcm3host_gmm_fit(const double* samples, size_t n, size_t D, size_t K, unsigned long long seed, double ess_factor, double* weights,
                     double* means, double* covariances, double* stats)
{
	std::vector<VectorReal> rows(n, VectorReal(D));
	for (size_t r = 0; r < n; r++) rows[r].assign(samples + r * D, samples + (r + 1) * D);
	RNG rng(seed, 0);
	GaussianMixture g;
	if (!g.Fit(rows, K, rng, ess_factor)) return -1;
	for (size_t k = 0; k < K; k++) {
		weights[k] = g.GetWeights()[k];
		const auto& c = g.GetComponent(k);
		std::copy(c.mean.begin(), c.mean.end(), means + k * D);
		std::copy(c.covariance.begin(), c.covariance.end(), covariances + k * D * D);
	}
	stats[0] = g.GetAIC();
	stats[1] = g.GetLogLikelihood();
	return 0;
}

// mixture given explicitly: log pdf and responsibilities at x[m][D]
int bcm3host_gmm_evaluate(size_t K, size_t D, const double* weights, const double* means, const double* covariances, const double* x, size_t m,
                          double* logpdf, double* responsibilities)
{
	std::vector<VectorReal> mu(K, VectorReal(D));
	std::vector<std::vector<Real>> cov(K, std::vector<Real>(D * D));
	for (size_t k = 0; k < K; k++) {
		mu[k].assign(means + k * D, means + (k + 1) * D);
		cov[k].assign(covariances + k * D * D, covariances + (k + 1) * D * D);
	}
	GaussianMixture g;
	if (!g.Set(mu, cov, VectorReal(weights, weights + K))) return -1;
	VectorReal r;
	for (size_t i = 0; i < m; i++) {
		const VectorReal xi(x + i * D, x + (i + 1) * D);
		logpdf[i] = g.LogPdf(xi);
		g.CalculateResponsibilities(xi, r);
		std::copy(r.begin(), r.end(), responsibilities + i * K);
	}
	return 0;
}

// symmetric eigen-decomposition of a[n][n]: values ascending, vectors column-major
void bcm3host_symmetric_eigen(const double* a, size_t n, double* values, double* vectors)
{
	VectorReal v;
	std::vector<Real> vec;
	GaussianMixture::SymmetricEigen(std::vector<Real>(a, a + n * n), n, v, vec);
	std::copy(v.begin(), v.end(), values);
	std::copy(vec.begin(), vec.end(), vectors);
}

// VariableSet / Prior surface for tests: number of variables, transform codes, index lookup
int bcm3host_varset_info(const char* prior_xml, const char* lookup_name, size_t* num_variables, int* transforms, size_t max_n, size_t* index)
{
	XmlNode root;
	std::string error;
	if (!ParseXml(prior_xml, root, error)) return -1;
	const XmlNode* node = root.child("prior");
	if (!node) node = root.child("variableset");
	if (!node) return -1;
	VariableSet vs;
	if (!vs.LoadFromNode(*node)) return -1;
	*num_variables = vs.GetNumVariables();
	for (size_t i = 0; i < vs.GetNumVariables() && i < max_n; i++) transforms[i] = (int)vs.GetTransform(i);
	if (lookup_name && index) *index = vs.GetVariableIndex(lookup_name);
	return 0;
}

} // extern "C"
