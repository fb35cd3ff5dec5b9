// bcm3b200.cu -- C ABI (include/bcm3b200.h) of the B200-native batched likelihood evaluator.
//
// Host side of the drop-in boundary: owns device buffers, streams and the static trial data, mirrors
// LikelihoodPopPKTrajectory::Initialize (src/likelihoods/LikelihoodPopPKTrajectory.cpp:50-252) for the
// derived quantities, and turns one batched call into kernel launches. No CPU fallback exists: without a
// usable CUDA device every entry point that computes returns BCM3B200_ERR_CUDA.
#include "../../include/bcm3b200.h"

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common_host.cuh"
#define BCM3_POPPK_AUX_KERNELS
#include "poppk_kernel.cuh"
#include "cellpop_host.cuh"
#include "comm_host.cuh"
#include "pharmaco_host.cuh"

using namespace bcm3b200;

namespace {

// the float literal of LikelihoodPopPKTrajectory.cpp:238, widened to double
const double kTol = (double)1e-6f;

double molecular_weight(const std::string& drug)
{
	// LikelihoodPopPKTrajectory.cpp:377-393
	if (drug == "lapatinib") return 581.06;
	if (drug == "dacomitinib") return 469.95;
	if (drug == "afatinib") return 485.94;
	if (drug == "trametinib") return 615.404;
	if (drug == "mirdametinib") return 482.19;
	if (drug == "selumetinib") return 457.68;
	return std::numeric_limits<double>::quiet_NaN();
}

struct Shard {
	int device = 0;
	int offset = 0; // global index of the first patient
	int P = 0;      // patients in this shard
	int P_pad = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	DevBuf<double> time, obs_t, dose, interval, dac, dct, values, patient_ll, partial, diag_conc;
	DevBuf<unsigned long long> rank_keys, rank_keys_sorted; // [C][P] (chain, ka) sort keys
	DevBuf<int> rank_patients, order;                        // [C][P] patient indices before / after the sort
	DevBuf<unsigned char> sort_temp;
	DevBuf<int32_t> intermittent, simulate_until, diag_counters;
	DevBuf<uint32_t> skipped;
	double* h_partial = nullptr; // pinned [3][C]
	size_t h_partial_n = 0;
	int last_C = 0;
	~Shard()
	{
		if (h_partial) cudaFreeHost(h_partial);
		if (ev0) cudaEventDestroy(ev0);
		if (ev1) cudaEventDestroy(ev1);
		if (stream) cudaStreamDestroy(stream);
	}
};

struct Handle {
	std::unique_ptr<PharmacoState> ph; // set for model kind "pharmaco_population" (matrix-exponential PK, no ODE solver)
	std::unique_ptr<CellPopState> cp; // set for model kind "cell_population"; the fields below are the PopPK evaluator
	// cell_population with device_count > 1 in ONE process: the states of the further devices (cp is the first device's) and
	// one NCCL end per device (ncclCommInitAll); the first device combines the gathered partials and finishes
	std::vector<std::unique_ptr<CellPopState>> cp_more;
	std::vector<std::unique_ptr<CommEnd>> dev_comm;
	// one process per GPU: this rank's end of the communicator attached with bcm3b200_comm_init (both model kinds)
	CommEnd comm;
	DevBuf<double> xchg; // partial block of a host-buffer evaluate on a handle with a communicator
	// description
	int pk_type = PK_ONE;
	std::string drug;
	int P = 0, T = 0, nvar = 0, sd_ix = -1, max_steps = 2000;
	int shard_rank = 0, shard_count = 1, device0 = 0, device_count = 1;
	// <pk_model volume_of_distribution= k_periphery_fwd= k_periphery_bwd=>, NaN = sampled (cpp:64-67)
	double fixed_vod = std::numeric_limits<double>::quiet_NaN(), fixed_periphery_fwd = std::numeric_limits<double>::quiet_NaN(),
	       fixed_periphery_bwd = std::numeric_limits<double>::quiet_NaN();
	// host copies of the static data
	std::map<std::string, std::vector<double>> data;
	// derived
	bool finalized = false;
	double rtol = 0, atol = 0, mol_weight = 0;
	int npk = 4;
	int named_a_ix = -1, named_b_ix = -1; // n_transit / biphasic_uptake_time, mean_transit_time / mean_absorption2
	int tr[SV_COUNT];
	int ix[SV_COUNT];
	std::vector<std::unique_ptr<Shard>> shards;
	// options / stats
	bool diagnostics = false;
	int block_size = 0;
	bool sort_patients = true;                    // option "sort_patients"
	bool single = false;                          // model kind "pharmacokinetic_trajectory": one patient, the chain's variables are its rates
	bool chain_fastest_grid = true;               // option "chain_fastest_grid": with ranked patients, launch all chains' expensive blocks first
	long long sort_min_systems = 60000;           // option "sort_min_systems": rank patients when P * C reaches this (measured:
	                                              // 80 k systems 5.59 -> 4.75 ms, 40 k and below no gain; tools/rank_check.py)
	int64_t total_launches = 0, last_launches = 0, num_evaluations = 0;
	double last_kernel_ms = 0;
};

bool parse_desc(const char* desc, size_t n, std::map<std::string, std::string>& kv)
{
	std::string s(desc, n);
	size_t pos = 0;
	while (pos < s.size()) {
		size_t end = s.find(';', pos);
		if (end == std::string::npos) end = s.size();
		std::string item = s.substr(pos, end - pos);
		pos = end + 1;
		size_t a = item.find_first_not_of(" \t\n\r");
		if (a == std::string::npos) continue;
		size_t b = item.find_last_not_of(" \t\n\r\0", std::string::npos, 5);
		item = item.substr(a, b - a + 1);
		size_t eq = item.find('=');
		if (eq == std::string::npos) return false;
		kv[item.substr(0, eq)] = item.substr(eq + 1);
	}
	return true;
}

int get_int(const std::map<std::string, std::string>& kv, const char* key, int def, bool* present = nullptr)
{
	auto it = kv.find(key);
	if (present) *present = it != kv.end();
	if (it == kv.end()) return def;
	return atoi(it->second.c_str());
}

std::vector<CellPopState*> cellpop_states(Handle* h)
{
	std::vector<CellPopState*> v;
	if (h->cp) v.push_back(h->cp.get());
	for (auto& m : h->cp_more) v.push_back(m.get());
	return v;
}

// "key=value" description -> one cell_population state (bcm3b200.h lists the keys)
int make_cellpop_state(std::map<std::string, std::string>& kv, std::unique_ptr<CellPopState>& out)
{
	std::unique_ptr<CellPopState> cp(new CellPopState);
	bool ok = true, present;
	auto need = [&](const char* key) {
		int v = get_int(kv, key, 0, &present);
		ok = ok && present;
		return v;
	};
	auto real = [&](const char* key, double def) { return kv.count(key) ? strtod(kv[key].c_str(), nullptr) : def; };
	cp->N = need("num_species");
	cp->nvar = need("num_variables");
	cp->num_cells = need("num_cells");
	cp->T = need("num_timepoints");
	// 96 species = the lane-group kernel at 32 lanes x 3 components per lane with the Newton matrix (N x (N | 1) doubles,
	// 74 KB at N = 96) in the cell's shared-memory block; larger models do not fit one SM's shared memory per cell
	if (!ok || cp->N < 1 || cp->N > 96 || cp->num_cells < 0 || cp->T < 1)
		return fail(BCM3B200_ERR_ARG, "num_species (1..96), num_variables, num_cells and num_timepoints are required");
	cp->Nc = get_int(kv, "num_constant_species", 0);
	cp->Nn = get_int(kv, "num_non_sampled", 0);
	cp->R = get_int(kv, "num_replicates", 1);
	cp->D = get_int(kv, "variability_dim", 0);
	cp->entry_time_ix = get_int(kv, "entry_time_ix", -1);
	cp->entry_time_fixed = real("entry_time", 0.0);
	cp->rel_tol = real("solver_relative_tolerance", cp->rel_tol);
	cp->abs_tol = real("solver_absolute_tolerance", cp->abs_tol);
	cp->min_dt = real("solver_min_timestep", cp->min_dt);
	cp->max_dt = real("solver_max_timestep", cp->max_dt);
	cp->max_steps = get_int(kv, "solver_max_steps", cp->max_steps);
	const std::string em = kv.count("error_model") ? kv["error_model"] : "normal";
	if (em == "normal" || em == "additive_normal") cp->error_model = CP_ERR_NORMAL;
	else if (em == "student_t4" || em == "t4") cp->error_model = CP_ERR_STUDENT_T4;
	else if (em == "proportional_normal") cp->error_model = CP_ERR_PROPORTIONAL_NORMAL;
	else if (em == "additive_proportional_normal") cp->error_model = CP_ERR_ADDITIVE_PROPORTIONAL_NORMAL;
	else return fail(BCM3B200_ERR_UNSUPPORTED, "error_model \"%s\" is not supported (normal, student_t4, proportional_normal, additive_proportional_normal)", em.c_str());
	{
		const std::string dk = kv.count("data_kind") ? kv["data_kind"] : "time_course_population_average";
		if (dk == "time_course") cp->data_kind = 1;
		else if (dk == "time_points") cp->data_kind = 2;
		else if (dk != "time_course_population_average") return fail(BCM3B200_ERR_UNSUPPORTED, "data_kind \"%s\" is not supported (time_course_population_average, time_course, time_points)", dk.c_str());
		cp->value_relative_to_timepoint_ix = get_int(kv, "value_relative_to_timepoint_ix", -1);
		cp->saturation_scale_ix = get_int(kv, "saturation_scale_ix", -1);
		cp->use_only_nondivided = get_int(kv, "use_only_nondivided", 0) != 0;
		cp->include_only_mitotic = get_int(kv, "include_only_cells_that_went_through_mitosis", 0) != 0;
		cp->nuclear_envelope_ix = get_int(kv, "nuclear_envelope_species", -1);
		cp->optimize_offset_scale = get_int(kv, "optimize_offset_scale", 0) != 0;
		cp->optimize_offset_min = real("optimize_offset_min", -1.0);
		cp->optimize_offset_max = real("optimize_offset_max", 1.0);
		cp->optimize_scale_min = real("optimize_scale_min", 0.1);
		cp->optimize_scale_max = real("optimize_scale_max", 10.0);
	}
	cp->treatment_species = get_int(kv, "treatment_species", -1);
	cp->relative_to_time_average = get_int(kv, "relative_to_time_average", 0) != 0;
	cp->stdev_relative_to_scale = get_int(kv, "stdev_relative_to_scale", 0) != 0;
	cp->prop_stdev_ix = get_int(kv, "proportional_stdev_ix", -1);
	cp->prop_stdev_fixed = real("proportional_stdev", 1.0);
	const std::string vd = kv.count("variability_distribution") ? kv["variability_distribution"] : "diagonal_gaussian";
	if (vd == "full_gaussian") cp->full_gaussian = true;
	else if (vd != "diagonal_gaussian") return fail(BCM3B200_ERR_UNSUPPORTED, "variability_distribution \"%s\" is not supported (diagonal_gaussian, full_gaussian)", vd.c_str());
	cp->weight = real("weight", 1.0);
	cp->stdev_ix = get_int(kv, "stdev_ix", -1);
	cp->stdev_fixed = real("stdev", 1.0);
	cp->offset_ix = get_int(kv, "offset_ix", -1);
	cp->offset_fixed = real("offset", 0.0);
	cp->scale_ix = get_int(kv, "scale_ix", -1);
	cp->scale_fixed = real("scale", 1.0);
	cp->missing_simulation_time_stdev = real("missing_simulation_time_stdev", 300.0);
	cp->have_sim_end_time = kv.count("simulation_end_time") != 0;
	cp->sim_end_time = real("simulation_end_time", 0.0);
	if (kv.count("obs_species")) {
		std::string v = kv["obs_species"];
		size_t pos = 0;
		while (pos <= v.size()) {
			size_t e = v.find('+', pos);
			if (e == std::string::npos) e = v.size();
			if (e > pos) cp->obs_species.push_back(atoi(v.substr(pos, e - pos).c_str()));
			pos = e + 1;
		}
	}
	cp->shard_rank = get_int(kv, "shard_rank", 0);
	cp->shard_count = get_int(kv, "shard_count", 1);
	cp->device = get_int(kv, "device", 0);
	// further data sets of the same experiment: the data-set keys again with the suffix @1, @2, @3
	const int num_data_sets = get_int(kv, "num_data_sets", 1);
	if (num_data_sets < 1 || num_data_sets > 4) return fail(BCM3B200_ERR_UNSUPPORTED, "num_data_sets must be 1..4");
	for (int k = 1; k < num_data_sets; k++) {
		std::unique_ptr<CellPopState::MoreData> m(new CellPopState::MoreData);
		const std::string sfx = "@" + std::to_string(k);
		auto key = [&](const char* base) { return std::string(base) + sfx; };
		auto realk = [&](const char* base, double def) { return kv.count(key(base)) ? strtod(kv[key(base)].c_str(), nullptr) : def; };
		m->T = get_int(kv, key("num_timepoints").c_str(), 0);
		m->R = get_int(kv, key("num_replicates").c_str(), 1);
		const std::string emk = kv.count(key("error_model")) ? kv[key("error_model")] : "normal";
		if (emk == "normal" || emk == "additive_normal") m->error_model = CP_ERR_NORMAL;
		else if (emk == "student_t4" || emk == "t4") m->error_model = CP_ERR_STUDENT_T4;
		else if (emk == "proportional_normal") m->error_model = CP_ERR_PROPORTIONAL_NORMAL;
		else if (emk == "additive_proportional_normal") m->error_model = CP_ERR_ADDITIVE_PROPORTIONAL_NORMAL;
		else return fail(BCM3B200_ERR_UNSUPPORTED, "error_model \"%s\" is not supported", emk.c_str());
		{
			const std::string dk = kv.count(key("data_kind")) ? kv[key("data_kind")] : "time_course_population_average";
			if (dk == "time_course") m->data_kind = 1;
			else if (dk == "time_points") m->data_kind = 2;
			else if (dk != "time_course_population_average") return fail(BCM3B200_ERR_UNSUPPORTED, "data_kind \"%s\" is not supported", dk.c_str());
			m->value_relative_to_timepoint_ix = get_int(kv, key("value_relative_to_timepoint_ix").c_str(), -1);
			m->saturation_scale_ix = get_int(kv, key("saturation_scale_ix").c_str(), -1);
			m->use_only_nondivided = get_int(kv, key("use_only_nondivided").c_str(), 0) != 0;
			m->include_only_mitotic = get_int(kv, key("include_only_cells_that_went_through_mitosis").c_str(), 0) != 0;
			m->marker_of = get_int(kv, key("marker_of").c_str(), -1);
			m->denominator_of = get_int(kv, key("denominator_of").c_str(), -1);
			m->optimize_offset_scale = get_int(kv, key("optimize_offset_scale").c_str(), 0) != 0;
			m->optimize_offset_min = realk("optimize_offset_min", -1.0);
			m->optimize_offset_max = realk("optimize_offset_max", 1.0);
			m->optimize_scale_min = realk("optimize_scale_min", 0.1);
			m->optimize_scale_max = realk("optimize_scale_max", 10.0);
		}
		m->stdev_ix = get_int(kv, key("stdev_ix").c_str(), -1);
		m->stdev_fixed = realk("stdev", 1.0);
		m->offset_ix = get_int(kv, key("offset_ix").c_str(), -1);
		m->offset_fixed = realk("offset", 0.0);
		m->scale_ix = get_int(kv, key("scale_ix").c_str(), -1);
		m->scale_fixed = realk("scale", 1.0);
		m->prop_stdev_ix = get_int(kv, key("proportional_stdev_ix").c_str(), -1);
		m->prop_stdev_fixed = realk("proportional_stdev", 1.0);
		m->weight = realk("weight", 1.0);
		m->missing_stdev = realk("missing_simulation_time_stdev", 300.0);
		m->relative_to_time_average = get_int(kv, key("relative_to_time_average").c_str(), 0) != 0;
		m->stdev_relative_to_scale = get_int(kv, key("stdev_relative_to_scale").c_str(), 0) != 0;
		if (kv.count(key("obs_species"))) {
			std::string v = kv[key("obs_species")];
			size_t pos = 0;
			while (pos <= v.size()) {
				size_t e = v.find('+', pos);
				if (e == std::string::npos) e = v.size();
				if (e > pos) m->obs_species.push_back(atoi(v.substr(pos, e - pos).c_str()));
				pos = e + 1;
			}
		}
		cp->more.push_back(std::move(m));
	}
	// dividing / dying cells (Experiment.cpp:488-489 divide_cells, max_cells; Cell.cpp:40-55 the species found by name)
	cp->divide_cells = get_int(kv, "divide_cells", 0) != 0;
	cp->max_cells = get_int(kv, "max_cells", cp->num_cells);
	cp->cytokinesis_ix = get_int(kv, "cytokinesis_species", -1);
	cp->apoptosis_ix = get_int(kv, "apoptosis_species", -1);
	cp->sobol_rows = cp->num_cells;
	if (kv.count("division_reset_species")) {
		std::string v = kv["division_reset_species"];
		size_t pos = 0;
		int k = 0;
		while (pos <= v.size() && k < 7) {
			size_t e = v.find('+', pos);
			if (e == std::string::npos) e = v.size();
			if (e > pos) cp->reset_ix[k++] = atoi(v.substr(pos, e - pos).c_str());
			pos = e + 1;
		}
	}
	out = std::move(cp);
	return BCM3B200_OK;
}

int pick_block_size(const Handle& h, const Shard& s, size_t C)
{
	if (h.block_size) return h.block_size;
	// Large batches: 12 warps per SM at 168 registers per thread -- the cold part of the integrator state lives in shared
	// memory (bdf_thread.cuh) -- with the warps of a block running the integrator in lock-step (BCM3_BLOCK_LOCKSTEP) so that
	// they share instruction fetches (the hot loop is several times larger than the 32 KB instruction cache). With the
	// patients ranked by absorption rate the blocks are homogeneous in step count, and three blocks of 128 threads per SM
	// (264.0 ms at config 5) schedule slightly better than one block of 384 (268.9 ms); unranked it was the other way round
	// (355.7 vs 345.7 ms). Small batches are latency-bound: spread them over as many SMs as possible with small blocks.
	int dev_sms = 148;
	cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, s.device);
	size_t threads = (size_t)s.P * C;
	if (threads >= (size_t)dev_sms * 2 * 128) return 128;
	if (threads >= (size_t)dev_sms * 2 * 64) return 64;
	return 32;
}

int finalize(Handle* h)
{
	if (h->finalized) return BCM3B200_OK;
	const int P = h->P, T = h->T;
	static const char* required[] = { "time", "observed_concentration", "dose", "dosing_interval", "dose_after_dose_change",
		                              "dose_change_time", "intermittent", "treatment_interruptions", "transforms" };
	for (const char* name : required) {
		if (!h->data.count(name)) return fail(BCM3B200_ERR_STATE, "missing data \"%s\"", name);
	}
	h->mol_weight = molecular_weight(h->drug);
	if (std::isnan(h->mol_weight)) return fail(BCM3B200_ERR_ARG, "Unknown drug \"%s\"", h->drug.c_str());
	static const int npk_of_type[6] = { 4, 6, 7, 7, 6, 8 }; // cpp:99-120 (7 for both biphasic types, as in the reference)
	h->npk = npk_of_type[h->pk_type];
	if (h->pk_type >= PK_ONE_BIPHASIC && (h->named_a_ix < 0 || h->named_b_ix < 0 || h->named_a_ix >= h->nvar || h->named_b_ix >= h->nvar))
		return fail(BCM3B200_ERR_ARG, h->pk_type >= PK_ONE_TRANSIT ? "transit models need n_transit_ix and mean_transit_time_ix"
		                                                           : "biphasic models need biphasic_uptake_time_ix and mean_absorption2_ix");
	// cpp:122-130: every fixed attribute takes one variable out of the prior -- and nothing else changes: the reference keeps
	// reading the vector at the all-sampled positions (cpp:267-272, 283-286), which is reproduced here as it is
	const int fixed_var_count = (std::isnan(h->fixed_vod) ? 0 : 1) + (std::isnan(h->fixed_periphery_fwd) ? 0 : 1) + (std::isnan(h->fixed_periphery_bwd) ? 0 : 1);
	const bool two_cmt_model = (h->pk_type == PK_TWO || h->pk_type == PK_TWO_BIPHASIC || h->pk_type == PK_TWO_TRANSIT);
	if (h->single) {
		// the single-patient likelihood checks no variable count (its check is compiled out, LikelihoodPharmacokineticTrajectory.cpp:128-151);
		// every position it reads (its cpp:264-305) has to exist
		int last = 3;
		if (two_cmt_model && std::isnan(h->fixed_periphery_fwd)) last = 5;
		if (h->pk_type == PK_ONE_BIPHASIC || h->pk_type == PK_TWO_BIPHASIC) last = 7;
		if (h->nvar <= last) return fail(BCM3B200_ERR_ARG, "the model reads variable %d, the prior has %d", last, h->nvar);
	} else {
		if (h->nvar != h->npk - fixed_var_count + 2 * (P + 1) + 2) return fail(BCM3B200_ERR_ARG, "Incorrect number of variables in prior");
		if (P > 0 && h->npk + 2 * P + 1 >= h->nvar)
			return fail(BCM3B200_ERR_ARG, "with these fixed pk_model attributes the reference reads past the end of the variable vector (cpp:283-286)");
	}
	if (h->sd_ix < 0 || h->sd_ix + 1 >= h->nvar) return fail(BCM3B200_ERR_ARG, "sd_ix out of range");

	const std::vector<double>& time = h->data["time"];
	const std::vector<double>& obs = h->data["observed_concentration"];
	const std::vector<double>& dose = h->data["dose"];
	const std::vector<double>& dac = h->data["dose_after_dose_change"];
	const std::vector<double>& dct = h->data["dose_change_time"];
	const std::vector<double>& inter = h->data["treatment_interruptions"];
	const std::vector<double>& transforms = h->data["transforms"];

	// chain-level variable indices (positional, cpp:267-272,283-286) and their transforms
	const int named_a = h->named_a_ix >= 0 ? h->named_a_ix : 0, named_b = h->named_b_ix >= 0 ? h->named_b_ix : 0;
	const int sigma0 = h->single ? 0 : h->npk; // no population spreads in the single-patient likelihood: slots unused
	const int ix[SV_COUNT] = { 0, 1, 2, 3, h->single && h->nvar <= 4 ? 0 : 4, h->single && h->nvar <= 5 ? 0 : 5, sigma0, h->single ? 0 : h->npk + 1, h->sd_ix, h->sd_ix + 1, named_a, named_b };
	for (int k = 0; k < SV_COUNT; k++) {
		h->ix[k] = ix[k];
		h->tr[k] = (int)transforms[ix[k]];
	}

	// simulate_until (cpp:163-184), skipped-day masks (cpp:154-161), minimum dose (cpp:197-203)
	std::vector<int32_t> simulate_until(P);
	std::vector<uint32_t> skipped(P);
	double minimum_dose = std::numeric_limits<double>::max();
	for (int j = 0; j < P; j++) {
		uint32_t mask = 0;
		for (int d = 0; d < 29; d++)
			if (inter[(size_t)j * 29 + d] != 0.0) mask |= (1u << d);
		skipped[j] = mask;
		int su = 0;
		if (mask & 2u) {
			for (int i = 0; i < T; i++) {
				if (time[i] >= 24.0) {
					su = i;
					break;
				}
			}
		} else {
			su = T;
		}
		for (int i = 0; i < T; i++) {
			if (!std::isnan(obs[(size_t)j * T + i])) {
				if (time[i] > 15 * 24) su = 0;
				break;
			}
		}
		if (h->single) su = T; // the whole time vector (LikelihoodPharmacokineticTrajectory.cpp:341)
		simulate_until[j] = su;
		if (!std::isnan(dac[j]) && !h->single) {
			if (std::isnan(dct[j]))
				return fail(BCM3B200_ERR_ARG, "Patient %d has dose change, but time of dose change is not specified.", j);
		}
		if (dose[j] < minimum_dose) minimum_dose = dose[j];
		if (!std::isnan(dac[j]) && dac[j] < minimum_dose) minimum_dose = dac[j];
	}
	h->rtol = kTol;
	h->atol = minimum_dose * kTol; // SetTolerance(1e-6f, minimum_dose * 1e-6f), cpp:238
	if (h->single) h->atol = dose[0] * kTol; // SetTolerance(1e-6f, dose * 1e-6f), LikelihoodPharmacokineticTrajectory.cpp:226

	// partition this handle's contiguous slice of patients over its devices
	const long long lo = (long long)P * h->shard_rank / h->shard_count;
	const long long hi = (long long)P * (h->shard_rank + 1) / h->shard_count;
	const long long Pl = hi - lo;
	h->shards.clear();
	for (int d = 0; d < h->device_count; d++) {
		std::unique_ptr<Shard> s(new Shard);
		s->device = h->device0 + d;
		s->offset = (int)(lo + Pl * d / h->device_count);
		s->P = (int)(lo + Pl * (d + 1) / h->device_count) - s->offset;
		s->P_pad = (s->P + 31) / 32 * 32;
		CUDA_TRY(cudaSetDevice(s->device));
		CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
		CUDA_TRY(cudaEventCreate(&s->ev0));
		CUDA_TRY(cudaEventCreate(&s->ev1));
		const int Ps = s->P, off = s->offset;
		const size_t Pa = (size_t)(Ps > 0 ? Ps : 1);
		CUDA_TRY(s->time.ensure(T > 0 ? T : 1));
		CUDA_TRY(cudaMemcpy(s->time.p, time.data(), sizeof(double) * T, cudaMemcpyHostToDevice));
		// observations, transposed to time-major so that a warp reads 32 consecutive patients
		std::vector<double> obs_t((size_t)T * s->P_pad, std::numeric_limits<double>::quiet_NaN());
		for (int j = 0; j < Ps; j++)
			for (int i = 0; i < T; i++) obs_t[(size_t)i * s->P_pad + j] = obs[(size_t)(off + j) * T + i];
		CUDA_TRY(s->obs_t.ensure(obs_t.size() ? obs_t.size() : 1));
		CUDA_TRY(cudaMemcpy(s->obs_t.p, obs_t.data(), sizeof(double) * obs_t.size(), cudaMemcpyHostToDevice));
		auto up = [&](DevBuf<double>& b, const std::vector<double>& v) -> cudaError_t {
			cudaError_t e = b.ensure(Pa);
			if (e != cudaSuccess) return e;
			return cudaMemcpy(b.p, v.data() + off, sizeof(double) * Ps, cudaMemcpyHostToDevice);
		};
		CUDA_TRY(up(s->dose, dose));
		CUDA_TRY(up(s->interval, h->data["dosing_interval"]));
		CUDA_TRY(up(s->dac, dac));
		CUDA_TRY(up(s->dct, dct));
		std::vector<int32_t> im(Pa);
		const std::vector<double>& imd = h->data["intermittent"];
		for (int j = 0; j < Ps; j++) im[j] = (int32_t)imd[off + j];
		if (h->single) // the single-patient likelihood keeps the schedule in a bool (its cpp:184-186): any schedule is schedule 1
			for (int j = 0; j < Ps; j++) im[j] = im[j] != 0 ? 1 : 0;
		CUDA_TRY(s->intermittent.ensure(Pa));
		CUDA_TRY(cudaMemcpy(s->intermittent.p, im.data(), sizeof(int32_t) * Ps, cudaMemcpyHostToDevice));
		CUDA_TRY(s->skipped.ensure(Pa));
		CUDA_TRY(cudaMemcpy(s->skipped.p, skipped.data() + off, sizeof(uint32_t) * Ps, cudaMemcpyHostToDevice));
		CUDA_TRY(s->simulate_until.ensure(Pa));
		CUDA_TRY(cudaMemcpy(s->simulate_until.p, simulate_until.data() + off, sizeof(int32_t) * Ps, cudaMemcpyHostToDevice));
		h->shards.push_back(std::move(s));
	}
	h->finalized = true;
	return BCM3B200_OK;
}

// enqueue K0+K1 and K3 for one shard; values already on the device in the layout (row_stride, col_patient0, ix)
int launch_shard(Handle* h, Shard* s, size_t C, const double* d_values, long long row_stride, long long col_patient0,
                 const int* ix, double* d_partial, cudaStream_t stream)
{
	const int block = pick_block_size(*h, *s, C);
	const int nblk = (s->P + block - 1) / block;
	if (nblk == 0) {
		// empty shard: neutral partial
		std::vector<double> neutral(3 * C, std::numeric_limits<double>::infinity());
		for (size_t c = 0; c < C; c++) neutral[c] = 0.0;
		CUDA_TRY(cudaMemcpyAsync(d_partial, neutral.data(), sizeof(double) * 3 * C, cudaMemcpyHostToDevice, stream));
		CUDA_TRY(cudaStreamSynchronize(stream));
		h->last_launches = 0;
		return BCM3B200_OK;
	}
	CUDA_TRY(s->patient_ll.ensure((size_t)s->P * C));

	PkArgs a;
	a.time = s->time.p;
	a.obs_t = s->obs_t.p;
	a.dose = s->dose.p;
	a.dosing_interval = s->interval.p;
	a.dose_after_dose_change = s->dac.p;
	a.dose_change_time = s->dct.p;
	a.intermittent = s->intermittent.p;
	a.skipped_days = s->skipped.p;
	a.simulate_until = s->simulate_until.p;
	a.P_local = s->P;
	a.P_pad = s->P_pad;
	a.patient_offset = s->offset;
	a.T = h->T;
	a.max_steps = h->max_steps;
	a.rtol = h->rtol;
	a.atol = h->atol;
	a.conv_base = 1e6 / h->mol_weight;
	a.fixed_vod = h->fixed_vod;
	a.fixed_periphery_fwd = h->fixed_periphery_fwd;
	a.fixed_periphery_bwd = h->fixed_periphery_bwd;
	a.values = d_values;
	a.row_stride = row_stride;
	a.col_patient0 = col_patient0;
	for (int k = 0; k < SV_COUNT; k++) {
		a.ix[k] = ix[k];
		a.tr[k] = h->tr[k];
	}
	a.patient_ll = s->patient_ll.p;
	a.order = nullptr;
	a.chain_fastest = 0;
	a.single = h->single ? 1 : 0;
	// large batches: rank every chain's patients by absorption rate first (see poppk_kernel)
	if (h->sort_patients && (long long)s->P * (long long)C >= h->sort_min_systems && s->P > 0 && (long long)s->P * (long long)C < (1ll << 31)) {
		const size_t n = (size_t)s->P * C;
		CUDA_TRY(s->rank_keys.ensure(n));
		CUDA_TRY(s->rank_keys_sorted.ensure(n));
		CUDA_TRY(s->rank_patients.ensure(n));
		CUDA_TRY(s->order.ensure(n));
		poppk_rank_kernel<<<dim3((s->P + 255) / 256, (unsigned)C), 256, 0, stream>>>(a, (int)C, s->rank_keys.p, s->rank_patients.p);
		CUDA_TRY(cudaGetLastError());
		int chain_bits = 1;
		while (((size_t)1 << chain_bits) < C) chain_bits++;
		size_t temp_bytes = 0;
		CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, s->rank_keys.p, s->rank_keys_sorted.p, s->rank_patients.p, s->order.p, (int)n, 0,
		                                         32 + chain_bits, stream));
		CUDA_TRY(s->sort_temp.ensure(temp_bytes ? temp_bytes : 1));
		CUDA_TRY(cub::DeviceRadixSort::SortPairs(s->sort_temp.p, temp_bytes, s->rank_keys.p, s->rank_keys_sorted.p, s->rank_patients.p, s->order.p, (int)n, 0,
		                                         32 + chain_bits, stream));
		a.order = s->order.p;
		a.chain_fastest = (h->chain_fastest_grid && nblk <= 65535) ? 1 : 0;
		h->last_launches += 1; // + the library's sort passes
		h->total_launches += 1;
	}
	a.diag_conc = nullptr;
	a.diag_counters = nullptr;
	if (h->diagnostics) {
		CUDA_TRY(s->diag_conc.ensure(C * s->P * h->T));
		CUDA_TRY(s->diag_counters.ensure(C * s->P * 8));
		a.diag_conc = s->diag_conc.p;
		a.diag_counters = s->diag_counters.p;
	}
	s->last_C = (int)C;

	// the integrator's thread-private state columns [slots][stride] first, then s_time [T] and s_sim [T][block]
	if (block > BCM3_POPPK_STRIDE_BIG) return fail(BCM3B200_ERR_ARG, "block_size above %d is not supported", BCM3_POPPK_STRIDE_BIG);
	const int stride = block <= BCM3_POPPK_STRIDE_SMALL ? BCM3_POPPK_STRIDE_SMALL : BCM3_POPPK_STRIDE_BIG; // the two instantiations of the state-column stride
	const bool two_cmt = (h->pk_type == PK_TWO || h->pk_type == PK_TWO_BIPHASIC || h->pk_type == PK_TWO_TRANSIT);
	const int slots = two_cmt ? (int)BdfSlots<3>::COUNT : (int)BdfSlots<2>::COUNT;
	const size_t smem_bytes = sizeof(double) * ((size_t)h->T + (size_t)h->T * block + (size_t)slots * stride);
	if (smem_bytes > 226 * 1024) return fail(BCM3B200_ERR_UNSUPPORTED, "too many timepoints (%d) for block size %d", h->T, block);
	if (C > 65535) return fail(BCM3B200_ERR_UNSUPPORTED, "more than 65535 chains in one batch");
	const dim3 grid = a.chain_fastest ? dim3((unsigned)C, nblk) : dim3(nblk, (unsigned)C);

	int lrc;
	if (h->pk_type == PK_ONE || h->pk_type == PK_TWO) lrc = launch_poppk_plain(two_cmt, h->diagnostics, stride, grid, block, smem_bytes, stream, a);
	else if (h->pk_type == PK_ONE_BIPHASIC || h->pk_type == PK_TWO_BIPHASIC) lrc = launch_poppk_biphasic(two_cmt, h->diagnostics, stride, grid, block, smem_bytes, stream, a);
	else lrc = launch_poppk_transit(two_cmt, h->diagnostics, stride, grid, block, smem_bytes, stream, a);
	if (lrc != 0) return fail(BCM3B200_ERR_CUDA, "poppk_kernel launch failed: %s", cudaGetErrorString((cudaError_t)lrc));
	CUDA_TRY(cudaGetLastError());
	poppk_chain_reduce<<<(unsigned)C, 256, 0, stream>>>(s->patient_ll.p, s->P, s->offset, (int)C, d_partial);
	CUDA_TRY(cudaGetLastError());
	h->last_launches += 2;
	h->total_launches += 2;
	return BCM3B200_OK;
}

void combine(size_t C, const double* partial, double* logp, int* status)
{
	// serial semantics of cpp:427-440: the sum stops at the first -inf patient; a NaN before it poisons the sum
	for (size_t c = 0; c < C; c++) {
		const double s = partial[c], first_inf = partial[C + c], first_nan = partial[2 * C + c];
		double v;
		if (first_nan < first_inf) v = std::numeric_limits<double>::quiet_NaN();
		else if (first_inf < std::numeric_limits<double>::infinity()) v = -std::numeric_limits<double>::infinity();
		else v = s;
		logp[c] = v;
		if (status) status[c] = std::isnan(v) ? BCM3B200_STATUS_NAN : BCM3B200_STATUS_OK;
	}
}

// FP64 FMA-chain microbenchmark: the roofline denominator for the BDF kernels (MEASURED_PEAKS.json has no FP64 entry).
__global__ void fp64_peak_kernel(double* out, double a, double b, int iters)
{
	double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 16; u++) {
			x0 = fma(x0, a, b);
			x1 = fma(x1, a, b);
			x2 = fma(x2, a, b);
			x3 = fma(x3, a, b);
			x4 = fma(x4, a, b);
			x5 = fma(x5, a, b);
			x6 = fma(x6, a, b);
			x7 = fma(x7, a, b);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

} // namespace

extern "C" {

int bcm3b200_measure_fp64_peak(int device, double* tflops)
{
	if (!tflops) return fail(BCM3B200_ERR_ARG, "null argument");
	CUDA_TRY(cudaSetDevice(device));
	int sms = 0;
	CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
	const int block = 256, grid = sms * 8, iters = 2048;
	double* d = nullptr;
	CUDA_TRY(cudaMalloc((void**)&d, sizeof(double) * block * grid));
	cudaEvent_t e0, e1;
	CUDA_TRY(cudaEventCreate(&e0));
	CUDA_TRY(cudaEventCreate(&e1));
	double best = 0.0;
	for (int rep = 0; rep < 6; rep++) {
		CUDA_TRY(cudaEventRecord(e0));
		fp64_peak_kernel<<<grid, block>>>(d, 0.999999, 1e-9, iters);
		CUDA_TRY(cudaEventRecord(e1));
		CUDA_TRY(cudaEventSynchronize(e1));
		float ms = 0.f;
		CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
		const double flop = 2.0 * 8 * 16 * (double)iters * block * grid;
		const double tf = flop / (ms * 1e-3) / 1e12;
		if (rep > 0 && tf > best) best = tf;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(d);
	*tflops = best;
	return BCM3B200_OK;
}

const char* bcm3b200_last_error(void) { return last_error_ref().c_str(); }

int bcm3b200_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

int bcm3b200_create(const char* model_kind, const void* model_desc, size_t desc_bytes, int device_count, void** handle)
{
	if (!model_kind || !handle) return fail(BCM3B200_ERR_ARG, "null argument");
	*handle = nullptr;
	const bool is_pharmaco_single = strcmp(model_kind, "pharmaco_single") == 0;
	const bool is_cellpop = strcmp(model_kind, "cell_population") == 0, is_pharmaco = strcmp(model_kind, "pharmaco_population") == 0 || is_pharmaco_single;
	const bool is_single = strcmp(model_kind, "pharmacokinetic_trajectory") == 0;
	if (!is_cellpop && !is_pharmaco && !is_single && strcmp(model_kind, "pop_pk_trajectory") != 0)
		return fail(BCM3B200_ERR_UNSUPPORTED, "unknown model kind \"%s\"", model_kind);
	std::map<std::string, std::string> kv;
	if (model_desc && !parse_desc((const char*)model_desc, desc_bytes, kv)) return fail(BCM3B200_ERR_ARG, "malformed model description");
	std::unique_ptr<Handle> h(new Handle);
	if (is_pharmaco) {
		std::unique_ptr<PharmacoState> ph(new PharmacoState);
		bool hasP, hasT, hasN;
		ph->P = get_int(kv, "num_patients", 0, &hasP);
		ph->T = get_int(kv, "num_timepoints", 0, &hasT);
		ph->nvar = get_int(kv, "num_variables", 0, &hasN);
		if (!hasP || !hasT || !hasN || ph->P < 0 || ph->T < 0) return fail(BCM3B200_ERR_ARG, "num_patients, num_timepoints and num_variables are required");
		ph->drug = kv.count("drug") ? kv["drug"] : "";
		ph->use_peripheral = get_int(kv, "peripheral_compartment", 0) != 0;
		ph->num_transit = get_int(kv, "num_transit_compartments", 0);
		ph->use_bioavailability = get_int(kv, "bioavailability", 0) != 0;
		ph->single = is_pharmaco_single;
		if (is_pharmaco_single) {
			// PharmacoLikelihoodSingle: <pk_model biphasic_absorption= metabolite=> (cpp:44-45), variables by the names of cpp:75-146
			ph->use_biphasic = get_int(kv, "biphasic_absorption", 0) != 0;
			ph->use_metabolite = get_int(kv, "metabolite", 0) != 0;
			ph->use_bioavailability = 0;
			static const char* keys[][2] = { { "additive_sd", "additive_sd" }, { "proportional_sd", "proportional_sd" }, { "absorption", "mean_absorption" },
				                             { "excretion", "mean_excretion" }, { "clearance", "mean_clearance" }, { "volume_of_distribution", "mean_volume_of_distribution" },
				                             { "peripheral_forward_rate", "peripheral_forward_rate" }, { "peripheral_backward_rate", "peripheral_backward_rate" },
				                             { "mean_transit_time", "mean_transit_time" }, { "direct_absorption", "direct_absorption" },
				                             { "metabolite_conversion_rate", "metabolite_conversion_rate" } };
			for (auto& k : keys) {
				const int ix = get_int(kv, (std::string(k[0]) + "_ix").c_str(), -1);
				if (ix >= ph->nvar) return fail(BCM3B200_ERR_ARG, "%s_ix out of range", k[0]);
				ph->ix[k[1]] = ix;
			}
		} else {
			static const char* roles[] = { "additive_sd", "proportional_sd", "mean_absorption", "mean_excretion", "mean_clearance", "mean_volume_of_distribution",
				                           "sigma_absorption", "sigma_excretion", "sigma_clearance", "sigma_volume_of_distribution", "sigma_transit_time",
				                           "peripheral_forward_rate", "peripheral_backward_rate", "mean_transit_time" };
			for (const char* role : roles) {
				const int ix = get_int(kv, (std::string(role) + "_ix").c_str(), -1);
				if (ix >= ph->nvar) return fail(BCM3B200_ERR_ARG, "%s_ix out of range", role);
				ph->ix[role] = ix;
			}
		}
		ph->shard_rank = get_int(kv, "shard_rank", 0);
		ph->shard_count = get_int(kv, "shard_count", 1);
		ph->device = get_int(kv, "device", 0);
		if (ph->shard_count < 1 || ph->shard_rank < 0 || ph->shard_rank >= ph->shard_count) return fail(BCM3B200_ERR_ARG, "bad shard_rank / shard_count");
		if (device_count != 1) return fail(BCM3B200_ERR_UNSUPPORTED, "pharmaco_population needs device_count == 1 (one process per GPU + bcm3b200_comm_init to use several)");
		const int ndev = bcm3b200_device_count();
		if (ndev == 0) return fail(BCM3B200_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
		if (ph->device < 0 || ph->device >= ndev) return fail(BCM3B200_ERR_CUDA, "device %d requested but only %d visible", ph->device, ndev);
		h->shard_rank = ph->shard_rank;
		h->shard_count = ph->shard_count;
		h->device0 = ph->device;
		h->device_count = 1;
		h->ph = std::move(ph);
		*handle = h.release();
		return BCM3B200_OK;
	}
	if (is_cellpop) {
		std::unique_ptr<CellPopState> cp;
		int mrc = make_cellpop_state(kv, cp);
		if (mrc != BCM3B200_OK) return mrc;
		if (device_count < 1) return fail(BCM3B200_ERR_ARG, "device_count must be >= 1");
		if (cp->shard_count < 1 || cp->shard_rank < 0 || cp->shard_rank >= cp->shard_count) return fail(BCM3B200_ERR_ARG, "bad shard_rank / shard_count");
		const bool compile_only = get_int(kv, "compile_only", 0) != 0;
		if (!compile_only) {
			const int ndev = bcm3b200_device_count();
			if (ndev == 0) return fail(BCM3B200_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
			if (cp->device < 0 || cp->device + device_count > ndev)
				return fail(BCM3B200_ERR_CUDA, "devices %d..%d requested but only %d visible", cp->device, cp->device + device_count - 1, ndev);
		}
		if (device_count > 1) {
			// the handle's slice of the cells is split once more over its devices: device d is shard rank * D + d of count * D
			const int r0 = cp->shard_rank, cnt = cp->shard_count, dev0 = cp->device;
			for (int d = 0; d < device_count; d++) {
				std::unique_ptr<CellPopState> more;
				if (d > 0) {
					mrc = make_cellpop_state(kv, more);
					if (mrc != BCM3B200_OK) return mrc;
				}
				CellPopState& st = d == 0 ? *cp : *more;
				st.shard_rank = r0 * device_count + d;
				st.shard_count = cnt * device_count;
				st.device = dev0 + d;
				if (d > 0) h->cp_more.push_back(std::move(more));
			}
		}
		h->device0 = cp->device;
		h->device_count = device_count;
		h->shard_rank = get_int(kv, "shard_rank", 0);
		h->shard_count = get_int(kv, "shard_count", 1);
		h->cp = std::move(cp);
		*handle = h.release();
		return BCM3B200_OK;
	}
	const std::string type = kv.count("type") ? kv["type"] : "";
	if (type == "one") h->pk_type = PK_ONE;
	else if (type == "two") h->pk_type = PK_TWO;
	// cpp:73-76: BOTH biphasic strings select the TWO-compartment biphasic model in the reference -- the same
	// likelihood.xml must give the same log-likelihood here. The one-compartment biphasic right-hand side the reference
	// also carries (cpp:496-530) is unreachable from its XML; it is kept under a name of its own.
	else if (type == "one_biphasic_uptake" || type == "two_biphasic_uptake") h->pk_type = PK_TWO_BIPHASIC;
	else if (type == "one_compartment_biphasic_uptake") h->pk_type = PK_ONE_BIPHASIC;
	else if (type == "one_transit") h->pk_type = PK_ONE_TRANSIT;
	else if (type == "two_transit") h->pk_type = PK_TWO_TRANSIT;
	else return fail(BCM3B200_ERR_UNSUPPORTED, "Unknown PK model type \"%s\"", type.c_str()); // cpp:84-87
	if (h->pk_type == PK_ONE_TRANSIT || h->pk_type == PK_TWO_TRANSIT) {
		h->named_a_ix = get_int(kv, "n_transit_ix", -1);
		h->named_b_ix = get_int(kv, "mean_transit_time_ix", -1);
	} else if (h->pk_type == PK_ONE_BIPHASIC || h->pk_type == PK_TWO_BIPHASIC) {
		h->named_a_ix = get_int(kv, "biphasic_uptake_time_ix", -1);
		h->named_b_ix = get_int(kv, "mean_absorption2_ix", -1);
	}
	h->single = is_single;
	if (is_single && (h->pk_type == PK_ONE_BIPHASIC || h->pk_type == PK_TWO_BIPHASIC)) {
		// LikelihoodPharmacokineticTrajectory.cpp:302-303: the single-patient likelihood reads the switching time and the second
		// absorption rate at positions 6 and 7, not by name
		h->named_a_ix = 6;
		h->named_b_ix = 7;
	}
	h->drug = kv.count("drug") ? kv["drug"] : "";
	if (kv.count("volume_of_distribution")) h->fixed_vod = strtod(kv["volume_of_distribution"].c_str(), nullptr);
	if (kv.count("k_periphery_fwd")) h->fixed_periphery_fwd = strtod(kv["k_periphery_fwd"].c_str(), nullptr);
	if (kv.count("k_periphery_bwd")) h->fixed_periphery_bwd = strtod(kv["k_periphery_bwd"].c_str(), nullptr);
	bool hasP, hasT, hasN, hasS;
	h->P = get_int(kv, "num_patients", 0, &hasP);
	h->T = get_int(kv, "num_timepoints", 0, &hasT);
	h->nvar = get_int(kv, "num_variables", 0, &hasN);
	h->sd_ix = get_int(kv, "sd_ix", -1, &hasS);
	if (!hasP || !hasT || !hasN || !hasS || h->P < 0 || h->T < 0)
		return fail(BCM3B200_ERR_ARG, "num_patients, num_timepoints, num_variables and sd_ix are required");
	if (is_single && (h->P != 1 || device_count != 1 || get_int(kv, "shard_count", 1) != 1))
		return fail(BCM3B200_ERR_ARG, "pharmacokinetic_trajectory is the likelihood of ONE patient: num_patients=1, one device, no shards");
	h->max_steps = get_int(kv, "max_steps", 2000);
	h->shard_rank = get_int(kv, "shard_rank", 0);
	h->shard_count = get_int(kv, "shard_count", 1);
	h->device0 = get_int(kv, "device", 0);
	h->device_count = device_count;
	if (h->shard_count < 1 || h->shard_rank < 0 || h->shard_rank >= h->shard_count) return fail(BCM3B200_ERR_ARG, "bad shard_rank / shard_count");
	if (device_count < 1) return fail(BCM3B200_ERR_ARG, "device_count must be >= 1");
	const int ndev = bcm3b200_device_count();
	if (ndev == 0) return fail(BCM3B200_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
	if (h->device0 < 0 || h->device0 + device_count > ndev)
		return fail(BCM3B200_ERR_CUDA, "devices %d..%d requested but only %d visible", h->device0, h->device0 + device_count - 1, ndev);
	*handle = h.release();
	return BCM3B200_OK;
}

int bcm3b200_set_data(void* handle, const char* name, const double* data, const size_t* shape, int ndim)
{
	Handle* h = (Handle*)handle;
	if (!h || !name || !data || !shape || ndim < 1 || ndim > 2) return fail(BCM3B200_ERR_ARG, "bad argument");
	if (h->ph) {
		PharmacoState& ph = *h->ph;
		const size_t P = (size_t)ph.P, T = (size_t)ph.T;
		const std::string n(name);
		size_t want0 = 0, want1 = 0;
		if (n == "time") want0 = T;
		else if (n == "observed_concentration") { want0 = P; want1 = T; }
		else if (n == "dose" || n == "dosing_interval" || n == "dose_after_dose_change" || n == "dose_change_time" || n == "intermittent") want0 = P;
		else if (n == "treatment_interruptions") { want0 = P; want1 = 29; }
		else if (n == "transforms") want0 = (size_t)ph.nvar;
		else {
			bool known = false;
			for (const char* arr : kPharmacoPatientArrays) known = known || n == arr;
			if (!known) return fail(BCM3B200_ERR_ARG, "unknown data name \"%s\"", name);
			want0 = P;
		}
		if (ndim != (want1 ? 2 : 1) || shape[0] != want0 || (want1 && shape[1] != want1)) return fail(BCM3B200_ERR_ARG, "shape mismatch for \"%s\"", name);
		ph.data[n].assign(data, data + want0 * (want1 ? want1 : 1));
		ph.finalized = false;
		return BCM3B200_OK;
	}
	if (h->cp) {
		CellPopState& cp = *h->cp;
		const std::string n(name);
		size_t w0 = 0, w1 = 0;
		if (n == "initial_conditions") w0 = cp.N;
		else if (n == "constant_species") w0 = cp.Nc;
		else if (n == "non_sampled_parameters") w0 = cp.Nn;
		else if (n == "sobol") {
			// a dividing population takes rows beyond its initial cells (daughters of row r: num_cells + 2 r + child,
			// CellPopulation.cpp:75; the reference makes 100 * num_cells rows, VariabilityPseudoRandomIterator.cpp:17)
			w0 = (cp.divide_cells && shape[0] >= (size_t)cp.num_cells) ? shape[0] : (size_t)cp.num_cells;
			w1 = cp.D;
		}
		else if (n == "timepoints") w0 = cp.T;
		else if (n == "observed") { w0 = cp.R; w1 = cp.T; }
		else if (n == "transforms") w0 = cp.nvar;
		else if (n == "variability") { w0 = cp.D; w1 = 6; }
		else if (n == "variability_covariance") { w0 = (size_t)cp.D * (cp.D - 1) / 2; w1 = 2; }
		else if (n == "treatment_times") w0 = shape[0]; // any number of pulses
		else if ((n.rfind("timepoints@", 0) == 0 || n.rfind("observed@", 0) == 0) && n.size() > 1) {
			// a further data set of the experiment (num_data_sets > 1)
			const int k = atoi(n.c_str() + n.find('@') + 1);
			if (k < 1 || k > (int)cp.more.size()) return fail(BCM3B200_ERR_ARG, "\"%s\": the handle was created with %zu data set(s)", name, cp.more.size() + 1);
			const bool is_time = n[0] == 't';
			const size_t Tk = (size_t)cp.more[k - 1]->T, Rk = (size_t)cp.more[k - 1]->R;
			if (is_time ? (ndim != 1 || shape[0] != Tk) : (ndim != 2 || shape[0] != Rk || shape[1] != Tk)) return fail(BCM3B200_ERR_ARG, "shape mismatch for \"%s\"", name);
			for (CellPopState* st : cellpop_states(h)) {
				std::vector<double>& dst = is_time ? st->more[k - 1]->timepoints : st->more[k - 1]->observed;
				dst.assign(data, data + (is_time ? Tk : Rk * Tk));
				st->finalized = false;
			}
			return BCM3B200_OK;
		}
		else return fail(BCM3B200_ERR_ARG, "unknown data name \"%s\"", name);
		const int want_ndim = w1 ? 2 : 1;
		if (ndim != want_ndim || shape[0] != w0 || (w1 && shape[1] != w1)) return fail(BCM3B200_ERR_ARG, "shape mismatch for \"%s\"", name);
		for (CellPopState* st : cellpop_states(h)) {
			st->data[n].assign(data, data + w0 * (w1 ? w1 : 1));
			if (n == "sobol") st->sobol_rows = (int)w0;
			st->finalized = false;
		}
		return BCM3B200_OK;
	}
	const size_t P = (size_t)h->P, T = (size_t)h->T;
	const std::string n(name);
	size_t want0 = 0, want1 = 0;
	if (n == "time") want0 = T;
	else if (n == "observed_concentration") { want0 = P; want1 = T; }
	else if (n == "dose" || n == "dosing_interval" || n == "dose_after_dose_change" || n == "dose_change_time" || n == "intermittent") want0 = P;
	else if (n == "treatment_interruptions") { want0 = P; want1 = 29; }
	else if (n == "transforms") want0 = (size_t)h->nvar;
	else return fail(BCM3B200_ERR_ARG, "unknown data name \"%s\"", name);
	const int want_ndim = want1 ? 2 : 1;
	if (ndim != want_ndim || shape[0] != want0 || (want1 && shape[1] != want1))
		return fail(BCM3B200_ERR_ARG, "shape mismatch for \"%s\"", name);
	const size_t count = want0 * (want1 ? want1 : 1);
	h->data[n].assign(data, data + count);
	h->finalized = false;
	return BCM3B200_OK;
}

int bcm3b200_set_text(void* handle, const char* name, const char* text, size_t text_bytes)
{
	Handle* h = (Handle*)handle;
	if (!h || !name || !text) return fail(BCM3B200_ERR_ARG, "null argument");
	if (!h->cp) return fail(BCM3B200_ERR_ARG, "this model kind takes no text inputs");
	if (strcmp(name, "derivative_code") != 0) return fail(BCM3B200_ERR_ARG, "unknown text name \"%s\"", name);
	for (CellPopState* st : cellpop_states(h)) {
		st->derivative_code.assign(text, text_bytes);
		st->finalized = false;
	}
	return BCM3B200_OK;
}

int bcm3b200_get_cell_diagnostics(void* handle, double* cell_values, int32_t* cell_status, int32_t* cell_steps, double* population_average)
{
	Handle* h = (Handle*)handle;
	if (!h || !h->cp) return fail(BCM3B200_ERR_ARG, "not a cell_population handle");
	CellPopState& cp = *h->cp;
	if (!cp.finalized || cp.last_C == 0) return fail(BCM3B200_ERR_STATE, "no evaluation yet");
	CUDA_TRY(cudaSetDevice(cp.device));
	CUDA_TRY(cudaDeviceSynchronize());
	const size_t C = (size_t)cp.last_C, nc = (size_t)cp.capacity(), T = (size_t)cp.rows(); // max_cells columns for a dividing population; every data set's timepoints
	if (cell_values) CUDA_TRY(cudaMemcpy(cell_values, cp.d_cellvals.p, sizeof(double) * C * T * nc, cudaMemcpyDeviceToHost));
	if (cell_status) CUDA_TRY(cudaMemcpy(cell_status, cp.d_status.p, sizeof(int32_t) * C * nc, cudaMemcpyDeviceToHost));
	if (cell_steps) CUDA_TRY(cudaMemcpy(cell_steps, cp.d_steps.p, sizeof(int32_t) * C * nc, cudaMemcpyDeviceToHost));
	if (population_average) CUDA_TRY(cudaMemcpy(population_average, cp.d_avg.p, sizeof(double) * C * T, cudaMemcpyDeviceToHost));
	return BCM3B200_OK;
}

int bcm3b200_finalize(void* handle)
{
	Handle* h = (Handle*)handle;
	if (!h) return fail(BCM3B200_ERR_ARG, "null handle");
	if (h->ph) return pharmaco_finalize(*h->ph, molecular_weight(h->ph->drug));
	if (h->cp) {
		for (CellPopState* st : cellpop_states(h)) {
			int rc = cellpop_finalize(*st, bcm3b200_device_count() > 0);
			if (rc != BCM3B200_OK) return rc;
		}
		return BCM3B200_OK;
	}
	return finalize(h);
}

// pharmaco_population, host buffers in, log-likelihoods out: this handle's slice of the patients, then (with a communicator)
// the all-gather + rank-order combination of the [3][C] blocks, exactly as for pop_pk_trajectory
static int pharmaco_evaluate_handle(Handle* h, size_t C, size_t nvar, const double* values, double* logp, int* status)
{
	PharmacoState& ph = *h->ph;
	int rc = pharmaco_finalize(ph, molecular_weight(ph.drug));
	if (rc != BCM3B200_OK) return rc;
	if (C == 0) return BCM3B200_OK;
	CUDA_TRY(cudaSetDevice(ph.device));
	CUDA_TRY(ph.d_partial.ensure(3 * C));
	if (ph.h_partial_n < 3 * C) {
		if (ph.h_partial) cudaFreeHost(ph.h_partial);
		ph.h_partial = nullptr;
		CUDA_TRY(cudaMallocHost((void**)&ph.h_partial, sizeof(double) * 3 * C));
		ph.h_partial_n = 3 * C;
	}
	rc = pharmaco_enqueue(ph, C, nvar, values, ph.d_partial.p, ph.stream);
	if (rc != BCM3B200_OK) return rc;
	if (h->comm.active()) {
		rc = h->comm.gather_combine(ph.d_partial.p, 3 * C, C, ph.d_partial.p, ph.stream);
		if (rc != BCM3B200_OK) return rc;
		ph.total_launches += 2;
	}
	CUDA_TRY(cudaMemcpyAsync(ph.h_partial, ph.d_partial.p, sizeof(double) * 3 * C, cudaMemcpyDeviceToHost, ph.stream));
	CUDA_TRY(cudaStreamSynchronize(ph.stream));
	float ms = 0.f;
	if (cudaEventElapsedTime(&ms, ph.ev0, ph.ev1) == cudaSuccess) ph.last_kernel_ms = ms;
	pharmaco_combine(C, ph.h_partial, logp, status);
	ph.num_evaluations += (int64_t)C;
	return BCM3B200_OK;
}

// cell_population, host buffers in, complete log-likelihoods out, for every way a handle can be spread:
//   one device                          -> cellpop_evaluate
//   one process per GPU + communicator  -> this rank's partial, all-gather + rank-order sum, finish (every rank gets logp)
//   several devices in this process     -> every device's partial on its own stream, grouped all-gather over the
//                                          ncclCommInitAll ends, the first device sums in device order and finishes
static int cellpop_evaluate_handle(Handle* h, size_t C, size_t nvar, const double* values, double* logp, int* status)
{
	CellPopState& cp = *h->cp;
	if (h->cp_more.empty() && !h->comm.active()) return cellpop_evaluate(cp, C, nvar, values, logp, status);
	if (C == 0) return BCM3B200_OK;
	const size_t width = 2 * (size_t)cp.rows() + 1, n = C * width;
	if (h->cp_more.empty()) {
		int rc = cellpop_finalize(cp, true);
		if (rc != BCM3B200_OK) return rc;
		CUDA_TRY(cudaSetDevice(cp.device));
		CUDA_TRY(h->xchg.ensure(n));
		rc = cellpop_enqueue_partial(cp, C, nvar, values, h->xchg.p, cp.stream);
		if (rc != BCM3B200_OK) return rc;
		rc = h->comm.gather_combine(h->xchg.p, n, n, h->xchg.p, cp.stream);
		if (rc != BCM3B200_OK) return rc;
		cp.total_launches += 2;
		return cellpop_finish(cp, C, h->xchg.p, logp, status, cp.stream);
	}
	std::vector<CellPopState*> states = cellpop_states(h);
	for (CellPopState* st : states) {
		int rc = cellpop_finalize(*st, true);
		if (rc != BCM3B200_OK) return rc;
	}
	if (h->dev_comm.empty()) {
		NcclApi& api = NcclApi::get();
		if (!api.ok()) return fail(BCM3B200_ERR_CUDA, "NCCL is needed to combine the devices of a cell_population handle and could not be loaded: %s", api.why.c_str());
		std::vector<int> devs;
		for (CellPopState* st : states) devs.push_back(st->device);
		std::vector<ncclComm_t> comms(devs.size());
		NCCL_TRY(api.CommInitAll(comms.data(), (int)devs.size(), devs.data()));
		for (size_t d = 0; d < devs.size(); d++) {
			std::unique_ptr<CommEnd> e(new CommEnd);
			e->comm = comms[d];
			e->world = (int)devs.size();
			e->rank = (int)d;
			h->dev_comm.push_back(std::move(e));
		}
	}
	for (size_t d = 0; d < states.size(); d++) {
		CellPopState& st = *states[d];
		CUDA_TRY(cudaSetDevice(st.device));
		CUDA_TRY(st.d_partial.ensure(n));
		CUDA_TRY(h->dev_comm[d]->gathered.ensure(states.size() * n));
		int rc = cellpop_enqueue_partial(st, C, nvar, values, st.d_partial.p, st.stream);
		if (rc != BCM3B200_OK) return rc;
	}
	NCCL_TRY(NcclApi::get().GroupStart());
	for (size_t d = 0; d < states.size(); d++) {
		CUDA_TRY(cudaSetDevice(states[d]->device));
		int rc = h->dev_comm[d]->gather(states[d]->d_partial.p, n, states[d]->stream);
		if (rc != BCM3B200_OK) {
			NcclApi::get().GroupEnd();
			return rc;
		}
	}
	NCCL_TRY(NcclApi::get().GroupEnd());
	CUDA_TRY(cudaSetDevice(cp.device));
	int rc = h->dev_comm[0]->combine(n, n, cp.d_partial.p, cp.stream);
	if (rc != BCM3B200_OK) return rc;
	cp.total_launches += 1;
	rc = cellpop_finish(cp, C, cp.d_partial.p, logp, status, cp.stream);
	if (rc != BCM3B200_OK) return rc;
	double max_ms = cp.last_kernel_ms;
	for (size_t d = 1; d < states.size(); d++) { // the other devices have nothing left but their end of the gather
		CUDA_TRY(cudaSetDevice(states[d]->device));
		CUDA_TRY(cudaEventRecord(states[d]->ev1, states[d]->stream));
		CUDA_TRY(cudaStreamSynchronize(states[d]->stream));
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, states[d]->ev0, states[d]->ev1) == cudaSuccess && ms > max_ms) max_ms = ms;
		cp.last_launches += states[d]->last_launches;
	}
	cp.last_kernel_ms = max_ms;
	CUDA_TRY(cudaSetDevice(cp.device));
	return BCM3B200_OK;
}

// Upload this shard's slice of the host batch in the compact layout [C][16 + 2 * P_shard] and enqueue the kernels on
// `stream`. Slots of a row: [0, npk + 2) = the head of the variable vector (the positional chain-level entries and the two
// population spreads, cpp:267-272, 283-286), 10-11 = standard_deviation and its successor, 12-13 = the two variables the
// variants find by name, 16.. = the shard's per-patient probabilities. Every piece is a strided copy straight from the
// caller's buffer (asynchronous when that is page-locked): nothing is staged in the handle, so enqueueing a second batch
// while the first is still in flight is safe as long as the caller leaves the first `values` alone until the stream has
// passed its copies (stated in bcm3b200.h).
static int upload_and_launch(Handle* h, Shard* s, size_t C, size_t num_variables, const double* values, double* d_partial,
                             cudaStream_t stream)
{
	const int SH = 16;
	const int head = h->single ? std::min(h->nvar, 10) : h->npk + 2;
	const int cix[SV_COUNT] = { h->ix[0], h->ix[1], h->ix[2], h->ix[3], h->ix[4], h->ix[5], h->ix[6], h->ix[7], 10, 11, 12, 13 };
	const size_t stride = SH + 2 * (size_t)s->P;
	CUDA_TRY(s->values.ensure(C * stride));
	auto piece = [&](int slot, size_t column, size_t count) -> cudaError_t {
		return cudaMemcpy2DAsync(s->values.p + slot, stride * sizeof(double), values + column, num_variables * sizeof(double), count * sizeof(double), C,
		                         cudaMemcpyHostToDevice, stream);
	};
	CUDA_TRY(piece(0, 0, (size_t)head));
	CUDA_TRY(piece(10, (size_t)h->sd_ix, 2));
	if (h->named_a_ix >= 0) CUDA_TRY(piece(12, (size_t)h->named_a_ix, 1));
	if (h->named_b_ix >= 0) CUDA_TRY(piece(13, (size_t)h->named_b_ix, 1));
	if (s->P > 0 && !h->single) CUDA_TRY(piece(SH, (size_t)h->npk + 2 + 2 * (size_t)s->offset, 2 * (size_t)s->P));
	CUDA_TRY(cudaEventRecord(s->ev0, stream));
	int rc = launch_shard(h, s, C, s->values.p, (long long)stride, SH, cix, d_partial, stream);
	if (rc != BCM3B200_OK) return rc;
	CUDA_TRY(cudaEventRecord(s->ev1, stream));
	return BCM3B200_OK;
}

int bcm3b200_evaluate_batch(void* handle, size_t num_chains, size_t num_variables, const double* values, double* logp, int* status)
{
	Handle* h = (Handle*)handle;
	if (!h || !values || !logp) return fail(BCM3B200_ERR_ARG, "null argument");
	if (h->ph) return pharmaco_evaluate_handle(h, num_chains, num_variables, values, logp, status);
	if (h->cp) return cellpop_evaluate_handle(h, num_chains, num_variables, values, logp, status);
	if ((int)num_variables != h->nvar) return fail(BCM3B200_ERR_ARG, "num_variables %zu != %d", num_variables, h->nvar);
	int rc = finalize(h);
	if (rc != BCM3B200_OK) return rc;
	const size_t C = num_chains;
	if (C == 0) return BCM3B200_OK;
	h->last_launches = 0;
	if (h->comm.active()) {
		// one process per GPU with the communicator inside the library: this rank's slice, then ONE all-gather of the
		// [3][C] blocks and their combination in rank order on the device, then 3 C doubles back -- the complete result on
		// every rank, bit-identical between the ranks
		Shard* s = h->shards[0].get();
		CUDA_TRY(cudaSetDevice(s->device));
		CUDA_TRY(s->partial.ensure(3 * C));
		if (s->h_partial_n < 3 * C) {
			if (s->h_partial) cudaFreeHost(s->h_partial);
			s->h_partial = nullptr;
			CUDA_TRY(cudaMallocHost((void**)&s->h_partial, sizeof(double) * 3 * C));
			s->h_partial_n = 3 * C;
		}
		rc = upload_and_launch(h, s, C, num_variables, values, s->partial.p, s->stream);
		if (rc != BCM3B200_OK) return rc;
		rc = h->comm.gather_combine(s->partial.p, 3 * C, C, s->partial.p, s->stream);
		if (rc != BCM3B200_OK) return rc;
		h->last_launches += 2;
		h->total_launches += 2;
		CUDA_TRY(cudaMemcpyAsync(s->h_partial, s->partial.p, sizeof(double) * 3 * C, cudaMemcpyDeviceToHost, s->stream));
		CUDA_TRY(cudaStreamSynchronize(s->stream));
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess) h->last_kernel_ms = ms;
		combine(C, s->h_partial, logp, status);
		h->num_evaluations += (int64_t)C;
		return BCM3B200_OK;
	}

	for (auto& sp : h->shards) {
		Shard* s = sp.get();
		CUDA_TRY(cudaSetDevice(s->device));
		CUDA_TRY(s->partial.ensure(3 * C));
		if (s->h_partial_n < 3 * C) {
			if (s->h_partial) cudaFreeHost(s->h_partial);
			s->h_partial = nullptr;
			CUDA_TRY(cudaMallocHost((void**)&s->h_partial, sizeof(double) * 3 * C));
			s->h_partial_n = 3 * C;
		}
		rc = upload_and_launch(h, s, C, num_variables, values, s->partial.p, s->stream);
		if (rc != BCM3B200_OK) return rc;
		CUDA_TRY(cudaMemcpyAsync(s->h_partial, s->partial.p, sizeof(double) * 3 * C, cudaMemcpyDeviceToHost, s->stream));
	}
	// combine the shards in patient order
	std::vector<double> total(3 * C, std::numeric_limits<double>::infinity());
	for (size_t c = 0; c < C; c++) total[c] = 0.0;
	double max_ms = 0.0;
	for (auto& sp : h->shards) {
		Shard* s = sp.get();
		CUDA_TRY(cudaSetDevice(s->device));
		CUDA_TRY(cudaStreamSynchronize(s->stream));
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess && ms > max_ms) max_ms = ms;
		for (size_t c = 0; c < C; c++) {
			total[c] += s->h_partial[c];
			total[C + c] = std::fmin(total[C + c], s->h_partial[C + c]);
			total[2 * C + c] = std::fmin(total[2 * C + c], s->h_partial[2 * C + c]);
		}
	}
	h->last_kernel_ms = max_ms;
	combine(C, total.data(), logp, status);
	h->num_evaluations += (int64_t)C;
	return BCM3B200_OK;
}

int bcm3b200_enqueue_batch(void* handle, size_t num_chains, size_t num_variables, const double* values, double* d_partial, void* stream)
{
	Handle* h = (Handle*)handle;
	if (!h || !values || !d_partial) return fail(BCM3B200_ERR_ARG, "null argument");
	if (h->ph) {
		int prc = pharmaco_finalize(*h->ph, molecular_weight(h->ph->drug));
		if (prc != BCM3B200_OK) return prc;
		if (num_chains == 0) return BCM3B200_OK;
		prc = pharmaco_enqueue(*h->ph, num_chains, num_variables, values, d_partial, (cudaStream_t)stream);
		if (prc == BCM3B200_OK) h->ph->num_evaluations += (int64_t)num_chains;
		return prc;
	}
	if (h->cp) {
		if (!h->cp_more.empty()) return fail(BCM3B200_ERR_UNSUPPORTED, "device-buffer entries need device_count == 1");
		return cellpop_enqueue_partial(*h->cp, num_chains, num_variables, values, d_partial, (cudaStream_t)stream);
	}
	if ((int)num_variables != h->nvar) return fail(BCM3B200_ERR_ARG, "num_variables %zu != %d", num_variables, h->nvar);
	int rc = finalize(h);
	if (rc != BCM3B200_OK) return rc;
	if (h->shards.size() != 1) return fail(BCM3B200_ERR_UNSUPPORTED, "device-buffer entries need device_count == 1");
	if (num_chains == 0) return BCM3B200_OK;
	Shard* s = h->shards[0].get();
	CUDA_TRY(cudaSetDevice(s->device));
	h->last_launches = 0;
	rc = upload_and_launch(h, s, num_chains, num_variables, values, d_partial, (cudaStream_t)stream);
	if (rc != BCM3B200_OK) return rc;
	h->num_evaluations += (int64_t)num_chains;
	return BCM3B200_OK;
}

int bcm3b200_evaluate_batch_device(void* handle, size_t num_chains, size_t num_variables, const double* d_values, double* d_partial,
                                   void* stream)
{
	Handle* h = (Handle*)handle;
	if (!h || !d_values || !d_partial) return fail(BCM3B200_ERR_ARG, "null argument");
	if (h->cp || h->ph) return fail(BCM3B200_ERR_UNSUPPORTED, "the device-values entry is pop_pk_trajectory only");
	if ((int)num_variables != h->nvar) return fail(BCM3B200_ERR_ARG, "num_variables %zu != %d", num_variables, h->nvar);
	int rc = finalize(h);
	if (rc != BCM3B200_OK) return rc;
	if (h->shards.size() != 1) return fail(BCM3B200_ERR_UNSUPPORTED, "device-buffer entry needs device_count == 1");
	if (num_chains == 0) return BCM3B200_OK;
	Shard* s = h->shards[0].get();
	CUDA_TRY(cudaSetDevice(s->device));
	h->last_launches = 0;
	const long long col0 = (long long)h->npk + 2 + 2ll * s->offset;
	rc = launch_shard(h, s, num_chains, d_values, (long long)num_variables, col0, h->ix, d_partial, (cudaStream_t)stream);
	if (rc != BCM3B200_OK) return rc;
	h->num_evaluations += (int64_t)num_chains;
	return BCM3B200_OK;
}

int bcm3b200_cellpop_finish(void* handle, size_t num_chains, const double* d_partial, double* logp, int* status, void* stream)
{
	Handle* h = (Handle*)handle;
	if (!h || !d_partial || !logp) return fail(BCM3B200_ERR_ARG, "null argument");
	if (!h->cp) return fail(BCM3B200_ERR_ARG, "not a cell_population handle");
	return cellpop_finish(*h->cp, num_chains, d_partial, logp, status, (cudaStream_t)stream);
}

int bcm3b200_comm_unique_id(void* id, size_t id_bytes)
{
	if (!id || id_bytes < sizeof(ncclUniqueId)) return fail(BCM3B200_ERR_ARG, "the id buffer needs %zu bytes (BCM3B200_COMM_ID_BYTES)", sizeof(ncclUniqueId));
	NcclApi& api = NcclApi::get();
	if (!api.ok()) return fail(BCM3B200_ERR_CUDA, "NCCL could not be loaded: %s", api.why.c_str());
	ncclUniqueId uid;
	NCCL_TRY(api.GetUniqueId(&uid));
	memset(id, 0, id_bytes);
	memcpy(id, &uid, sizeof(uid));
	return BCM3B200_OK;
}

int bcm3b200_comm_init(void* handle, const void* id, size_t id_bytes)
{
	Handle* h = (Handle*)handle;
	if (!h || !id || id_bytes < sizeof(ncclUniqueId)) return fail(BCM3B200_ERR_ARG, "bad argument");
	if (h->device_count != 1) return fail(BCM3B200_ERR_UNSUPPORTED, "a communicator joins handles with one device each (device_count == 1)");
	if (h->comm.comm) return fail(BCM3B200_ERR_STATE, "this handle already has a communicator");
	const int world = h->cp ? h->cp->shard_count : h->shard_count, rank = h->cp ? h->cp->shard_rank : h->shard_rank; // pharmaco: copied into the handle
	if (world == 1) return BCM3B200_OK; // nothing to exchange
	NcclApi& api = NcclApi::get();
	if (!api.ok()) return fail(BCM3B200_ERR_CUDA, "NCCL could not be loaded: %s", api.why.c_str());
	CUDA_TRY(cudaSetDevice(h->cp ? h->cp->device : h->device0));
	ncclUniqueId uid;
	memcpy(&uid, id, sizeof(uid));
	NCCL_TRY(api.CommInitRank(&h->comm.comm, world, uid, rank));
	h->comm.world = world;
	h->comm.rank = rank;
	return BCM3B200_OK;
}

int bcm3b200_exchange_partials(void* handle, size_t num_chains, double* d_partial, void* stream)
{
	Handle* h = (Handle*)handle;
	if (!h || !d_partial) return fail(BCM3B200_ERR_ARG, "null argument");
	if (num_chains == 0 || !h->comm.active()) return BCM3B200_OK;
	CUDA_TRY(cudaSetDevice(h->cp ? h->cp->device : h->device0));
	const size_t C = num_chains;
	int rc;
	if (h->cp) rc = h->comm.gather_combine(d_partial, C * (2 * (size_t)h->cp->rows() + 1), C * (2 * (size_t)h->cp->rows() + 1), d_partial, (cudaStream_t)stream);
	else rc = h->comm.gather_combine(d_partial, 3 * C, C, d_partial, (cudaStream_t)stream);
	if (rc != BCM3B200_OK) return rc;
	(h->cp ? h->cp->total_launches : h->total_launches) += 2;
	return BCM3B200_OK;
}

int bcm3b200_combine_partials(size_t num_chains, const double* partial, double* logp, int* status)
{
	if (!partial || !logp) return fail(BCM3B200_ERR_ARG, "null argument");
	combine(num_chains, partial, logp, status);
	return BCM3B200_OK;
}

int bcm3b200_get_diagnostics(void* handle, double* conc, double* patient_ll, int32_t* counters)
{
	Handle* h = (Handle*)handle;
	if (!h) return fail(BCM3B200_ERR_ARG, "null handle");
	if (h->cp) return fail(BCM3B200_ERR_UNSUPPORTED, "use bcm3b200_get_cell_diagnostics for cell_population");
	if (h->ph) {
		PharmacoState& ph = *h->ph;
		if (!ph.finalized || ph.last_C == 0) return fail(BCM3B200_ERR_STATE, "no evaluation yet");
		if (counters) return fail(BCM3B200_ERR_UNSUPPORTED, "pharmaco_population has no solver counters (matrix exponential, no ODE solver)");
		CUDA_TRY(cudaSetDevice(ph.device));
		CUDA_TRY(cudaDeviceSynchronize());
		const size_t C = (size_t)ph.last_C, Pl = (size_t)ph.P_local;
		if (patient_ll && Pl) CUDA_TRY(cudaMemcpy(patient_ll, ph.d_patient_ll.p, sizeof(double) * C * Pl, cudaMemcpyDeviceToHost));
		if (conc) {
			if (!ph.diagnostics) return fail(BCM3B200_ERR_STATE, "diagnostics were not enabled before the last evaluate");
			if (Pl) CUDA_TRY(cudaMemcpy(conc, ph.d_conc.p, sizeof(double) * C * Pl * ph.T, cudaMemcpyDeviceToHost));
		}
		return BCM3B200_OK;
	}
	if (!h->diagnostics || !h->finalized) return fail(BCM3B200_ERR_STATE, "diagnostics were not enabled before the last evaluate");
	// layout over the handle's patients: [C][P_handle][...], shards are contiguous slices
	size_t Ph = 0;
	for (auto& sp : h->shards) Ph += (size_t)sp->P;
	size_t base = 0;
	for (auto& sp : h->shards) {
		Shard* s = sp.get();
		const size_t C = (size_t)s->last_C, Ps = (size_t)s->P, T = (size_t)h->T;
		if (C == 0 || Ps == 0) {
			base += Ps;
			continue;
		}
		CUDA_TRY(cudaSetDevice(s->device));
		CUDA_TRY(cudaDeviceSynchronize());
		if (conc)
			CUDA_TRY(cudaMemcpy2D(conc + base * T, Ph * T * sizeof(double), s->diag_conc.p, Ps * T * sizeof(double), Ps * T * sizeof(double), C,
			                      cudaMemcpyDeviceToHost));
		if (patient_ll)
			CUDA_TRY(cudaMemcpy2D(patient_ll + base, Ph * sizeof(double), s->patient_ll.p, Ps * sizeof(double), Ps * sizeof(double), C,
			                      cudaMemcpyDeviceToHost));
		if (counters)
			CUDA_TRY(cudaMemcpy2D(counters + base * 8, Ph * 8 * sizeof(int32_t), s->diag_counters.p, Ps * 8 * sizeof(int32_t),
			                      Ps * 8 * sizeof(int32_t), C, cudaMemcpyDeviceToHost));
		base += Ps;
	}
	return BCM3B200_OK;
}

int bcm3b200_set_option(void* handle, const char* name, int64_t value)
{
	Handle* h = (Handle*)handle;
	if (!h || !name) return fail(BCM3B200_ERR_ARG, "null argument");
	if (h->ph) {
		if (!strcmp(name, "diagnostics")) h->ph->diagnostics = value != 0;
		else return fail(BCM3B200_ERR_ARG, "unknown option \"%s\"", name);
		return BCM3B200_OK;
	}
	if (h->cp) {
		for (CellPopState* st : cellpop_states(h)) {
			if (!strcmp(name, "diagnostics")) st->diagnostics = value != 0;
			else if (!strcmp(name, "cellpop_kernel")) st->kernel_choice = (int)value; // 0 auto, 1 one cell per warp, 2 one cell per thread, 3 one cell per lane group
			else if (!strcmp(name, "cellpop_steps_report")) st->steps_report = (int)value;
			else if (!strcmp(name, "cellpop_rhs_lanes")) { st->rhs_lanes = (int)value; st->finalized = false; } // before finalize: lane-parallel right-hand side, 0 never / 1 where it pays (default) / 2 always
			else if (!strcmp(name, "cellpop_group_lanes")) st->group_lanes = (int)value; // before finalize: lanes per cell of the group kernel (0 auto)
			else return fail(BCM3B200_ERR_ARG, "unknown option \"%s\"", name);
		}
		return BCM3B200_OK;
	}
	if (!strcmp(name, "diagnostics")) h->diagnostics = value != 0;
	else if (!strcmp(name, "sort_patients")) h->sort_patients = value != 0;
	else if (!strcmp(name, "sort_min_systems")) h->sort_min_systems = value;
	else if (!strcmp(name, "chain_fastest_grid")) h->chain_fastest_grid = value != 0;
	else if (!strcmp(name, "block_size")) {
		if (value != 0 && (value < 32 || value > BCM3_POPPK_STRIDE_BIG || value % 32 != 0)) return fail(BCM3B200_ERR_ARG, "block_size must be 0 or a multiple of 32 up to %d", BCM3_POPPK_STRIDE_BIG);
		h->block_size = (int)value;
	} else return fail(BCM3B200_ERR_ARG, "unknown option \"%s\"", name);
	return BCM3B200_OK;
}

int bcm3b200_get_stat(void* handle, const char* name, int64_t* value)
{
	Handle* h = (Handle*)handle;
	if (!h || !name || !value) return fail(BCM3B200_ERR_ARG, "null argument");
	if (h->ph) {
		PharmacoState& ph = *h->ph;
		if (!strcmp(name, "last_kernel_launches")) *value = ph.last_launches;
		else if (!strcmp(name, "total_kernel_launches")) *value = ph.total_launches;
		else if (!strcmp(name, "num_evaluations")) *value = ph.num_evaluations;
		else if (!strcmp(name, "last_kernel_us")) *value = (int64_t)(ph.last_kernel_ms * 1000.0);
		else if (!strcmp(name, "num_patients_local") || !strcmp(name, "patient_offset")) {
			if (!ph.finalized) return fail(BCM3B200_ERR_STATE, "not finalized");
			*value = !strcmp(name, "num_patients_local") ? ph.P_local : ph.offset;
		} else if (!strcmp(name, "num_compartments")) *value = ph.N;
		else return fail(BCM3B200_ERR_ARG, "unknown stat \"%s\"", name);
		return BCM3B200_OK;
	}
	if (h->cp) {
		CellPopState& cp = *h->cp;
		if (!strcmp(name, "last_kernel_launches")) *value = cp.last_launches;
		else if (!strcmp(name, "total_kernel_launches")) *value = cp.total_launches;
		else if (!strcmp(name, "num_evaluations")) *value = cp.num_evaluations;
		else if (!strcmp(name, "last_kernel_us")) *value = (int64_t)(cp.last_kernel_ms * 1000.0);
		else if (!strcmp(name, "num_cells_local")) *value = cp.cells_local;
		else if (!strcmp(name, "cell_columns")) *value = cp.capacity();
		else if (!strcmp(name, "partial_doubles_per_chain")) *value = 2 * cp.rows() + 1;
		else if (!strcmp(name, "value_rows")) *value = cp.rows();
		else return fail(BCM3B200_ERR_ARG, "unknown stat \"%s\"", name);
		return BCM3B200_OK;
	}
	if (!strcmp(name, "last_kernel_launches")) *value = h->last_launches;
	else if (!strcmp(name, "total_kernel_launches")) *value = h->total_launches;
	else if (!strcmp(name, "num_evaluations")) *value = h->num_evaluations;
	else if (!strcmp(name, "last_kernel_us")) *value = (int64_t)(h->last_kernel_ms * 1000.0);
	else if (!strcmp(name, "num_patients_local") || !strcmp(name, "patient_offset")) {
		if (!h->finalized) return fail(BCM3B200_ERR_STATE, "not finalized");
		int64_t n = 0;
		for (auto& sp : h->shards) n += sp->P;
		*value = !strcmp(name, "num_patients_local") ? n : (h->shards.empty() ? 0 : h->shards[0]->offset);
	} else return fail(BCM3B200_ERR_ARG, "unknown stat \"%s\"", name);
	return BCM3B200_OK;
}

void* bcm3b200_host_alloc(size_t bytes)
{
	void* p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
		last_error_ref() = "cudaMallocHost failed";
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

void bcm3b200_host_free(void* p)
{
	if (p) cudaFreeHost(p);
}

void bcm3b200_destroy(void* handle)
{
	Handle* h = (Handle*)handle;
	if (!h) return;
	for (auto& sp : h->shards) {
		cudaSetDevice(sp->device);
		cudaDeviceSynchronize();
	}
	for (CellPopState* st : cellpop_states(h)) {
		if (st->stream) {
			cudaSetDevice(st->device);
			cudaStreamSynchronize(st->stream);
		}
	}
	if (h->ph && h->ph->stream) {
		cudaSetDevice(h->ph->device);
		cudaStreamSynchronize(h->ph->stream);
	}
	delete h;
}

} // extern "C"

int bcm3b200_match_cells(int n, const double* cost, int32_t* match)
{
	if (n < 0 || (n > 0 && (!cost || !match))) return fail(BCM3B200_ERR_ARG, "bad argument");
	const std::vector<int> m = payor_matching_complete(n, cost);
	if ((int)m.size() != n) {
		for (int i = 0; i < n; i++) match[i] = -1;
		return n == 0 ? BCM3B200_OK : fail(BCM3B200_ERR_STATE, "no perfect matching found");
	}
	for (int i = 0; i < n; i++) match[i] = m[i];
	return BCM3B200_OK;
}
