// pharmaco_host.cuh -- host side of the "pharmaco_population" evaluator inside libbcm3b200.so: mirrors
// PharmacoLikelihoodPopulation::Initialize / PostInitialize (src/pharmaco/PharmacoLikelihoodPopulation.cpp:43-188) and
// Patient::Load (src/pharmaco/PharmacoPatient.cpp:8-116) for the derived per-patient arrays, and turns one batched call into
// pharmaco_kernel + the per-chain reduction shared with the PopPK path.
#pragma once

#include <cmath>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "common_host.cuh"
#include "pharmaco_kernel.cuh"

namespace bcm3b200 {

// one translation unit per family of matrix sizes (pharmaco_inst.cu): N = 2 + peripheral + transit compartments
int launch_pharmaco(int N, dim3 grid, int block, cudaStream_t stream, const PhArgs& a);

struct PharmacoState {
	// description (bcm3b200.h lists the keys)
	std::string drug;
	int P = 0, T = 0, nvar = 0;
	int use_peripheral = 0, num_transit = 0, use_bioavailability = 0;
	bool single = false; // model kind "pharmaco_single"
	int use_biphasic = 0, use_metabolite = 0;
	std::map<std::string, int> ix; // variable indices by role, -1 = absent
	int shard_rank = 0, shard_count = 1, device = 0;
	std::map<std::string, std::vector<double>> data;
	bool diagnostics = false;
	// derived
	bool finalized = false;
	double mol_weight = 0.0;
	int offset = 0, P_local = 0, N = 2;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	DevBuf<double> d_values, d_treat_time, d_treat_dose, d_obs_time, d_obs_value, d_patient_ll, d_partial, d_conc;
	DevBuf<int32_t> d_transforms, d_treat_begin, d_obs_begin, d_obs_grid, d_pix[6];
	double* h_partial = nullptr;
	size_t h_partial_n = 0;
	int last_C = 0;
	int64_t total_launches = 0, last_launches = 0, num_evaluations = 0;
	double last_kernel_ms = 0.0;
	PhArgs args;

	int var(const char* role) const
	{
		auto it = ix.find(role);
		return it == ix.end() ? -1 : it->second;
	}
	~PharmacoState()
	{
		if (h_partial) cudaFreeHost(h_partial);
		if (ev0) cudaEventDestroy(ev0);
		if (ev1) cudaEventDestroy(ev1);
		if (stream) cudaStreamDestroy(stream);
	}
};

static const char* const kPharmacoPatientArrays[6] = { "patient_absorption_ix", "patient_excretion_ix", "patient_clearance_ix",
	                                                   "patient_volume_of_distribution_ix", "patient_transit_time_ix", "patient_bioavailability_ix" };

inline int pharmaco_finalize(PharmacoState& ph, double mol_weight)
{
	if (ph.finalized) return BCM3B200_OK;
	const int P = ph.P, T = ph.T;
	static const char* required[] = { "time", "observed_concentration", "dose", "dosing_interval", "dose_after_dose_change", "dose_change_time",
		                              "intermittent", "treatment_interruptions", "transforms" };
	for (const char* name : required)
		if (!ph.data.count(name)) return fail(BCM3B200_ERR_STATE, "missing data \"%s\"", name);
	if (std::isnan(mol_weight)) return fail(BCM3B200_ERR_ARG, "Unknown drug \"%s\"", ph.drug.c_str());
	ph.mol_weight = mol_weight;
	// PostInitialize, cpp:102-188
	if (ph.var("additive_sd") < 0 && ph.var("proportional_sd") < 0)
		return fail(BCM3B200_ERR_ARG, "Neither \"additive_error_standard_deviation\" nor \"proportional_error_standard_deviation\" has been specified in the prior");
	for (const char* role : { "mean_absorption", "mean_clearance", "mean_volume_of_distribution" })
		if (ph.var(role) < 0) return fail(BCM3B200_ERR_ARG, "Could not find variable \"%s\"", ph.single ? role + 5 : role);
	if (ph.single) {
		// PharmacoLikelihoodSingle::PostInitialize, PharmacoLikelihoodSingle.cpp:75-150
		if (P != 1) return fail(BCM3B200_ERR_ARG, "pharmaco_single is the likelihood of ONE patient: num_patients=1");
		if (ph.use_biphasic && ph.var("direct_absorption") < 0)
			return fail(BCM3B200_ERR_ARG, "Biphasic absorption was specified, but direct absorption rate has not been specified in prior.");
		if (ph.use_metabolite && ph.var("metabolite_conversion_rate") < 0)
			return fail(BCM3B200_ERR_ARG, "Use of metabolite was specified, but metabolite conversion rate has not been specified in prior.");
	}
	if (ph.use_peripheral && (ph.var("peripheral_forward_rate") < 0 || ph.var("peripheral_backward_rate") < 0))
		return fail(BCM3B200_ERR_ARG, "Peripheral compartment was specified, but forward or backward rates have not both been specified in prior");
	if (ph.num_transit > 0 && ph.var("mean_transit_time") < 0)
		return fail(BCM3B200_ERR_ARG, "Transit compartments were specified, but mean transit time has not been specified in prior");
	ph.N = 2 + (ph.use_peripheral ? 1 : 0) + (ph.use_metabolite ? 1 : 0) + ph.num_transit;
	if (ph.N > 8) return fail(BCM3B200_ERR_UNSUPPORTED, "more than 8 compartments (2 + peripheral + metabolite + transit)");
	struct Marginal {
		const char* sigma_role;
		int array;
		bool needed;
	};
	const Marginal marginals[6] = { { "sigma_absorption", 0, ph.var("sigma_absorption") >= 0 },
		                            { "sigma_excretion", 1, ph.var("sigma_excretion") >= 0 && ph.var("mean_excretion") >= 0 },
		                            { "sigma_clearance", 2, ph.var("sigma_clearance") >= 0 },
		                            { "sigma_volume_of_distribution", 3, ph.var("sigma_volume_of_distribution") >= 0 },
		                            { "sigma_transit_time", 4, ph.var("sigma_transit_time") >= 0 && ph.num_transit > 0 },
		                            { "bioavailability", 5, ph.use_bioavailability != 0 } };
	for (const Marginal& m : marginals)
		if (m.needed && !ph.data.count(kPharmacoPatientArrays[m.array]))
			return fail(BCM3B200_ERR_STATE, "\"%s\" is in the prior but the per-patient variable indices \"%s\" were not supplied", m.sigma_role, kPharmacoPatientArrays[m.array]);

	// Patient::Load, PharmacoPatient.cpp:48-112: dose times up to 696 h on the patient's schedule, observations with a value only
	const std::vector<double>& time = ph.data["time"];
	const std::vector<double>& obs = ph.data["observed_concentration"];
	const std::vector<double>& dose = ph.data["dose"];
	const std::vector<double>& interval = ph.data["dosing_interval"];
	const std::vector<double>& dac = ph.data["dose_after_dose_change"];
	const std::vector<double>& dct = ph.data["dose_change_time"];
	const std::vector<double>& inter = ph.data["intermittent"];
	const std::vector<double>& interruptions = ph.data["treatment_interruptions"];
	std::vector<int32_t> treat_begin(P + 1, 0), obs_begin(P + 1, 0), obs_grid;
	std::vector<double> treat_time, treat_dose, obs_time, obs_value;
	for (int j = 0; j < P; j++) {
		if (!(interval[j] > 0.0)) return fail(BCM3B200_ERR_ARG, "patient %d: dosing_interval must be positive", j);
		const int intermittent = (int)inter[j];
		const double last_time = 696.0;
		double t = 0.0;
		while (t < last_time) {
			bool give = true;
			const int day = (int)floor(t / 24.0);
			if (day >= 0 && day < 29 && interruptions[(size_t)j * 29 + day] != 0.0) give = false;
			if (intermittent == 1) {
				if (t - 7.0 * 24.0 * floor(t / (7.0 * 24.0)) >= 5.0 * 24.0) give = false;
			} else if (intermittent == 2) {
				if (t - 28.0 * 24.0 * floor(t / (28.0 * 24.0)) >= 21.0 * 24.0) give = false;
			} else if (intermittent == 3) {
				if (t - 7.0 * 24.0 * floor(t / (7.0 * 24.0)) >= 4.0 * 24.0) give = false;
			}
			if (give) {
				treat_time.push_back(t);
				treat_dose.push_back((!std::isnan(dct[j]) && t >= dct[j]) ? dac[j] : dose[j]);
			}
			t += interval[j];
		}
		treat_begin[j + 1] = (int32_t)treat_time.size();
		double prev = -std::numeric_limits<double>::infinity();
		for (int i = 0; i < T; i++) {
			if (time[i] < prev) return fail(BCM3B200_ERR_ARG, "Observation timepoints need to be sorted");
			prev = time[i];
			const double y = obs[(size_t)j * T + i];
			if (!std::isnan(y)) {
				obs_time.push_back(time[i]);
				obs_value.push_back(y);
				obs_grid.push_back(i);
			}
		}
		obs_begin[j + 1] = (int32_t)obs_time.size();
	}

	const long long lo = (long long)P * ph.shard_rank / ph.shard_count, hi = (long long)P * (ph.shard_rank + 1) / ph.shard_count;
	ph.offset = (int)lo;
	ph.P_local = (int)(hi - lo);
	CUDA_TRY(cudaSetDevice(ph.device));
	if (!ph.stream) CUDA_TRY(cudaStreamCreateWithFlags(&ph.stream, cudaStreamNonBlocking));
	if (!ph.ev0) CUDA_TRY(cudaEventCreate(&ph.ev0));
	if (!ph.ev1) CUDA_TRY(cudaEventCreate(&ph.ev1));
	auto upd = [&](DevBuf<double>& b, const std::vector<double>& v) -> cudaError_t {
		cudaError_t e = b.ensure(v.size() ? v.size() : 1);
		if (e != cudaSuccess || v.empty()) return e;
		return cudaMemcpy(b.p, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice);
	};
	auto upi = [&](DevBuf<int32_t>& b, const std::vector<int32_t>& v) -> cudaError_t {
		cudaError_t e = b.ensure(v.size() ? v.size() : 1);
		if (e != cudaSuccess || v.empty()) return e;
		return cudaMemcpy(b.p, v.data(), sizeof(int32_t) * v.size(), cudaMemcpyHostToDevice);
	};
	CUDA_TRY(upd(ph.d_treat_time, treat_time));
	CUDA_TRY(upd(ph.d_treat_dose, treat_dose));
	CUDA_TRY(upd(ph.d_obs_time, obs_time));
	CUDA_TRY(upd(ph.d_obs_value, obs_value));
	CUDA_TRY(upi(ph.d_treat_begin, treat_begin));
	CUDA_TRY(upi(ph.d_obs_begin, obs_begin));
	CUDA_TRY(upi(ph.d_obs_grid, obs_grid));
	std::vector<int32_t> tr(ph.nvar);
	for (int i = 0; i < ph.nvar; i++) tr[i] = (int32_t)ph.data["transforms"][i];
	CUDA_TRY(upi(ph.d_transforms, tr));
	PhArgs& a = ph.args;
	memset(&a, 0, sizeof(a));
	const int32_t** parr[6] = { &a.p_absorption_ix, &a.p_excretion_ix, &a.p_clearance_ix, &a.p_vod_ix, &a.p_transit_ix, &a.p_bioavailability_ix };
	for (const Marginal& m : marginals) {
		*parr[m.array] = nullptr;
		if (!m.needed) continue;
		const std::vector<double>& src = ph.data[kPharmacoPatientArrays[m.array]];
		std::vector<int32_t> v(P);
		for (int j = 0; j < P; j++) {
			v[j] = (int32_t)src[j];
			if (v[j] < 0 || v[j] >= ph.nvar) return fail(BCM3B200_ERR_ARG, "%s[%d] is not a variable index", kPharmacoPatientArrays[m.array], j);
		}
		CUDA_TRY(upi(ph.d_pix[m.array], v));
		*parr[m.array] = ph.d_pix[m.array].p;
	}
	a.P_local = ph.P_local;
	a.patient_offset = ph.offset;
	a.nvar = ph.nvar;
	a.transforms = ph.d_transforms.p;
	a.additive_sd_ix = ph.var("additive_sd");
	a.proportional_sd_ix = ph.var("proportional_sd");
	a.mean_absorption_ix = ph.var("mean_absorption");
	a.mean_excretion_ix = ph.var("mean_excretion");
	a.mean_clearance_ix = ph.var("mean_clearance");
	a.mean_vod_ix = ph.var("mean_volume_of_distribution");
	a.sigma_absorption_ix = marginals[0].needed ? ph.var("sigma_absorption") : -1;
	a.sigma_excretion_ix = marginals[1].needed ? ph.var("sigma_excretion") : -1;
	a.sigma_clearance_ix = marginals[2].needed ? ph.var("sigma_clearance") : -1;
	a.sigma_vod_ix = marginals[3].needed ? ph.var("sigma_volume_of_distribution") : -1;
	a.sigma_transit_ix = marginals[4].needed ? ph.var("sigma_transit_time") : -1;
	a.periph_fwd_ix = ph.var("peripheral_forward_rate");
	a.periph_bwd_ix = ph.var("peripheral_backward_rate");
	a.mean_transit_time_ix = ph.var("mean_transit_time");
	a.use_peripheral = ph.use_peripheral;
	a.num_transit = ph.num_transit;
	a.use_bioavailability = ph.use_bioavailability;
	a.single = ph.single ? 1 : 0;
	a.use_biphasic = ph.use_biphasic;
	a.use_metabolite = ph.use_metabolite;
	a.direct_absorption_ix = ph.var("direct_absorption");
	a.metabolite_conversion_ix = ph.var("metabolite_conversion_rate");
	a.conv_base = 1e6 / ph.mol_weight;
	a.treat_begin = ph.d_treat_begin.p;
	a.treat_time = ph.d_treat_time.p;
	a.treat_dose = ph.d_treat_dose.p;
	a.obs_begin = ph.d_obs_begin.p;
	a.obs_time = ph.d_obs_time.p;
	a.obs_value = ph.d_obs_value.p;
	a.obs_grid = ph.d_obs_grid.p;
	a.T = T;
	ph.finalized = true;
	return BCM3B200_OK;
}

// upload the batch, run K1 + the chain reduction: partial [3][C] (sum of the finite terms, first -inf patient, first NaN patient) on the device
inline int pharmaco_enqueue(PharmacoState& ph, size_t C, size_t nvar, const double* values, double* d_partial, cudaStream_t st)
{
	if ((int)nvar != ph.nvar) return fail(BCM3B200_ERR_ARG, "num_variables %zu != %d", nvar, ph.nvar);
	if (!ph.finalized) return fail(BCM3B200_ERR_STATE, "not finalized");
	if (C > 65535) return fail(BCM3B200_ERR_UNSUPPORTED, "more than 65535 chains in one batch");
	CUDA_TRY(cudaSetDevice(ph.device));
	CUDA_TRY(ph.d_values.ensure(C * nvar));
	CUDA_TRY(ph.d_patient_ll.ensure(C * (size_t)(ph.P_local ? ph.P_local : 1)));
	CUDA_TRY(cudaMemcpyAsync(ph.d_values.p, values, sizeof(double) * C * nvar, cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaEventRecord(ph.ev0, st));
	PhArgs a = ph.args;
	a.num_chains = (int)C;
	a.values = ph.d_values.p;
	a.patient_ll = ph.d_patient_ll.p;
	a.diag_conc = nullptr;
	if (ph.diagnostics) {
		CUDA_TRY(ph.d_conc.ensure(C * (size_t)(ph.P_local ? ph.P_local : 1) * ph.T));
		a.diag_conc = ph.d_conc.p;
	}
	ph.last_launches = 0;
	if (ph.P_local > 0) {
		const int block = 128;
		int rc = launch_pharmaco(ph.N, dim3((ph.P_local + block - 1) / block, (unsigned)C), block, st, a);
		if (rc != 0) return fail(BCM3B200_ERR_CUDA, "pharmaco_kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
		ph.last_launches++;
	}
	poppk_chain_reduce<<<(unsigned)C, 256, 0, st>>>(ph.d_patient_ll.p, ph.P_local, ph.offset, (int)C, d_partial);
	CUDA_TRY(cudaGetLastError());
	ph.last_launches++;
	ph.total_launches += ph.last_launches;
	CUDA_TRY(cudaEventRecord(ph.ev1, st));
	ph.last_C = (int)C;
	return BCM3B200_OK;
}

// the reference adds every patient's term (cpp:243): a NaN anywhere poisons the sum, otherwise a -inf anywhere makes it -inf
inline void pharmaco_combine(size_t C, const double* partial, double* logp, int* status)
{
	for (size_t c = 0; c < C; c++) {
		double v = partial[c];
		if (partial[2 * C + c] < std::numeric_limits<double>::infinity()) v = std::numeric_limits<double>::quiet_NaN();
		else if (partial[C + c] < std::numeric_limits<double>::infinity()) v = -std::numeric_limits<double>::infinity();
		logp[c] = v;
		if (status) status[c] = std::isnan(v) ? BCM3B200_STATUS_NAN : BCM3B200_STATUS_OK;
	}
}

} // namespace bcm3b200
