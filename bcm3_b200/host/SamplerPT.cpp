#include "SamplerPT.h"

#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>

namespace bcm3 {

// ---------------------------------------------------------------------------------------------- settings
bool SamplerPTSettings::LoadFromConfigFile(const std::string& filename, std::string* error)
{
	std::ifstream f(filename);
	if (!f) {
		if (error) *error = "cannot open " + filename;
		return false;
	}
	std::stringstream ss;
	ss << f.rdbuf();
	return LoadFromConfigText(ss.str(), error);
}

bool SamplerPTSettings::LoadFromConfigText(const std::string& text, std::string* error)
{
	// boost::program_options INI style as in examples/banana/config.txt
	std::map<std::string, std::string> kv;
	std::istringstream in(text);
	std::string line, section;
	auto trim = [](std::string v) {
		size_t a = v.find_first_not_of(" \t\r\n"), b = v.find_last_not_of(" \t\r\n");
		return a == std::string::npos ? std::string() : v.substr(a, b - a + 1);
	};
	while (std::getline(in, line)) {
		line = trim(line.substr(0, line.find('#')));
		if (line.empty()) continue;
		if (line.front() == '[' && line.back() == ']') {
			section = line.substr(1, line.size() - 2);
			continue;
		}
		size_t eq = line.find('=');
		if (eq == std::string::npos) {
			if (error) *error = "bad config line: " + line;
			return false;
		}
		kv[section + "." + trim(line.substr(0, eq))] = trim(line.substr(eq + 1));
	}
	auto U = [&](const char* k, size_t& v) { if (kv.count(k)) v = (size_t)strtoull(kv[k].c_str(), nullptr, 10); };
	auto R = [&](const char* k, Real& v) { if (kv.count(k)) v = strtod(kv[k].c_str(), nullptr); };
	auto S = [&](const char* k, std::string& v) { if (kv.count(k)) v = kv[k]; };
	U("sampler.num_samples", num_samples);
	U("sampler.use_every_nth", use_every_nth);
	if (kv.count("sampler.rngseed")) rngseed = strtoull(kv["sampler.rngseed"].c_str(), nullptr, 10);
	U("ptmhsampler.num_chains", num_chains);
	S("ptmhsampler.proposal_type", proposal_type);
	S("ptmhsampler.blocking_strategy", blocking_strategy);
	S("ptmhsampler.swapping_scheme", swapping_scheme);
	U("ptmhsampler.num_exploration_steps", num_exploration_steps);
	U("ptmhsampler.max_history_size", max_history_size);
	U("ptmhsampler.adapt_proposal_samples", adapt_proposal_samples);
	U("ptmhsampler.adapt_proposal_times", adapt_proposal_times);
	U("ptmhsampler.adapt_proposal_max_history_samples", adapt_proposal_max_history_samples);
	U("ptmhsampler.stop_proposal_scaling", stop_proposal_scaling);
	R("ptmhsampler.exchange_probability", exchange_probability);
	R("ptmhsampler.temperature_schedule_power", temperature_schedule_power);
	R("ptmhsampler.temperature_schedule_max", temperature_schedule_max);
	R("ptmhsampler.proposal_t_dof", proposal_t_dof);
	U("ptmhsampler.initial_position_tries", initial_position_tries);
	return true;
}

// ---------------------------------------------------------------------------------------------- history
void SampleHistory::Initialize(size_t num_variables, size_t history_size, size_t sub)
{
	nvar = num_variables;
	capacity = history_size;
	subsampling = sub ? sub : 1;
	sample_n = sample_n_s = 0;
	samples.assign(nvar * capacity, 0.f);
}

void SampleHistory::AddSample(const VectorReal& sample)
{
	// SampleHistory.cpp:32-45
	if (capacity == 0) return;
	sample_n_s++;
	if (sample_n_s == subsampling) {
		size_t ix = sample_n % capacity;
		for (size_t i = 0; i < nvar; i++) samples[ix * nvar + i] = (float)sample[i];
		sample_n++;
		sample_n_s = 0;
	}
}

void SampleHistory::GetHistory(std::vector<VectorReal>& rows) const
{
	const size_t n = GetSampleCount();
	rows.assign(n, VectorReal(nvar));
	for (size_t r = 0; r < n; r++)
		for (size_t i = 0; i < nvar; i++) rows[r][i] = (Real)samples[r * nvar + i];
}

void SampleHistory::GetHistory(const std::vector<size_t>& variable_indices, std::vector<VectorReal>& rows) const
{
	const size_t n = GetSampleCount();
	rows.assign(n, VectorReal(variable_indices.size()));
	for (size_t r = 0; r < n; r++)
		for (size_t i = 0; i < variable_indices.size(); i++) rows[r][i] = (Real)samples[r * nvar + variable_indices[i]];
}

MatrixReal SampleHistory::GetEmpiricalCorrelation() const
{
	// cor() of SummaryStats.h on the float history cast to Real: centred columns, covariance / (sd_i sd_j); a variable
	// that never moved has no correlation (NaN), as there
	const size_t n = GetSampleCount();
	MatrixReal c(nvar, nvar, kNaN);
	if (n < 2) return c;
	std::vector<Real> mean(nvar, 0.0);
	for (size_t r = 0; r < n; r++)
		for (size_t i = 0; i < nvar; i++) mean[i] += (Real)samples[r * nvar + i];
	for (size_t i = 0; i < nvar; i++) mean[i] /= (Real)n;
	MatrixReal cov(nvar, nvar, 0.0);
	for (size_t r = 0; r < n; r++)
		for (size_t j = 0; j < nvar; j++) {
			const Real dj = (Real)samples[r * nvar + j] - mean[j];
			for (size_t i = j; i < nvar; i++) cov(i, j) += ((Real)samples[r * nvar + i] - mean[i]) * dj;
		}
	for (size_t j = 0; j < nvar; j++)
		for (size_t i = j; i < nvar; i++) {
			const Real v = (cov(i, j) / (Real)(n - 1)) / (sqrt(cov(i, i) / (Real)(n - 1)) * sqrt(cov(j, j) / (Real)(n - 1)));
			c(i, j) = v;
			c(j, i) = v;
		}
	return c;
}

// ---------------------------------------------------------------------------------------------- variable blocks
std::vector<std::vector<size_t>> TreeClusterCompleteLinkage(const MatrixReal& distance, Real cut_height)
{
	// The merge list: n - 1 nodes in the order in which complete linkage joins the closest pair of clusters. The tie-breaking
	// has to be the C Clustering Library's (cluster.c: find_closest_pair scans the lower triangle row by row and keeps the
	// FIRST minimum; pmlcluster keeps the merged cluster in the smaller slot and moves the last slot into the freed one),
	// or equal distances would give other blocks than the reference's.
	const int n = (int)distance.cols();
	std::vector<std::vector<size_t>> clusters;
	if (n < 1) return clusters;
	struct Node { int left, right; Real distance; };
	std::vector<Node> nodes;
	{
		std::vector<std::vector<Real>> d(n);
		for (int i = 0; i < n; i++) {
			d[i].resize(i);
			for (int j = 0; j < i; j++) d[i][j] = distance(j, i);
		}
		std::vector<int> id(n);
		for (int j = 0; j < n; j++) id[j] = j;
		for (int m = n; m > 1; m--) {
			int is = 1, js = 0;
			Real best = d[1][0];
			for (int i = 1; i < m; i++)
				for (int j = 0; j < i; j++)
					if (d[i][j] < best) {
						best = d[i][j];
						is = i;
						js = j;
					}
			// the joined cluster lives in slot js: its distance to every other slot is the larger of the two
			for (int j = 0; j < js; j++) d[js][j] = std::max(d[is][j], d[js][j]);
			for (int j = js + 1; j < is; j++) d[j][js] = std::max(d[is][j], d[j][js]);
			for (int j = is + 1; j < m; j++) d[j][js] = std::max(d[j][is], d[j][js]);
			// slot is takes over the last slot
			for (int j = 0; j < is; j++) d[is][j] = d[m - 1][j];
			for (int j = is + 1; j < m - 1; j++) d[j][is] = d[m - 1][j];
			nodes.push_back(Node{ id[is], id[js], best });
			id[js] = m - n - 1; // node k is referred to as -(k + 1)
			id[is] = id[m - 1];
		}
	}
	// The cut (Clustering.cpp:61-96): walking the merge list, a node below the cut height joins its two sides, one above
	// it leaves each side that is a single item as a cluster of its own. cluster_of[item] = position in `clusters`.
	std::vector<int> cluster_of(n, -1);
	auto any_item = [&](int el) { // an item below a node (all of them are in one cluster when the node was joined)
		while (el < 0) {
			const Node& nd = nodes[(size_t)(-el) - 1];
			el = nd.left >= 0 ? nd.left : nd.right;
		}
		return el;
	};
	auto add_item = [&](int item, int ci) {
		clusters[ci].push_back((size_t)item);
		cluster_of[item] = ci;
	};
	for (const Node& nd : nodes) {
		if (nd.distance < cut_height) {
			if (nd.left >= 0 && nd.right >= 0) {
				clusters.emplace_back();
				add_item(nd.left, (int)clusters.size() - 1);
				add_item(nd.right, (int)clusters.size() - 1);
			} else if (nd.left >= 0) {
				add_item(nd.left, cluster_of[any_item(nd.right)]);
			} else if (nd.right >= 0) {
				add_item(nd.right, cluster_of[any_item(nd.left)]);
			} else {
				const int keep = cluster_of[any_item(nd.left)], drop = cluster_of[any_item(nd.right)];
				for (size_t item : clusters[drop]) add_item((int)item, keep);
				clusters.erase(clusters.begin() + drop);
				for (int& c : cluster_of)
					if (c > drop) c--;
			}
		} else {
			if (nd.left >= 0) {
				clusters.emplace_back();
				add_item(nd.left, (int)clusters.size() - 1);
			}
			if (nd.right >= 0) {
				clusters.emplace_back();
				add_item(nd.right, (int)clusters.size() - 1);
			}
		}
	}
	for (auto& c : clusters) std::sort(c.begin(), c.end()); // std::set order
	return clusters;
}

bool GetVariableBlocks(const std::string& strategy, const SampleHistory& history, size_t num_variables, std::vector<std::vector<size_t>>& blocks)
{
	blocks.clear();
	auto singletons = [&]() {
		blocks.resize(num_variables);
		for (size_t i = 0; i < num_variables; i++) blocks[i].assign(1, i);
	};
	if (strategy == "one_block") {
		blocks.resize(1);
		for (size_t i = 0; i < num_variables; i++) blocks[0].push_back(i);
	} else if (strategy == "no_blocking") {
		singletons();
	} else if (strategy == "Turek") {
		// Turek et al. automated blocking: variables whose |correlation| in the chain's history exceeds 0.5 share a block
		// (distance 1 - |r|, complete linkage, cut at 0.5). One variable: the reference's merge list is empty and it ends up
		// with no block at all; here it is one block.
		if (history.GetSampleCount() > 2 && num_variables > 1) {
			const MatrixReal r = history.GetEmpiricalCorrelation();
			MatrixReal dist(num_variables, num_variables, 0.0);
			for (size_t j = 0; j < num_variables; j++)
				for (size_t i = 0; i < num_variables; i++) dist(i, j) = 1.0 - fabs(r(i, j));
			blocks = TreeClusterCompleteLinkage(dist, 0.5);
		} else {
			singletons();
		}
	} else {
		return false;
	}
	return true;
}

// ---------------------------------------------------------------------------------------------- proposal
bool Proposal::Initialize(const SampleHistory& history, size_t max_history_samples, const Prior& prior, const std::vector<size_t>& variable_indices, RNG& rng)
{
	// Proposal::Initialize (Proposal.cpp:39-140)
	var_ix = variable_indices;
	n = var_ix.size();
	target_acceptance_rate = (n == 1) ? 0.44 : (n == 2) ? 0.35 : (n == 3) ? 0.3 : 0.234;
	lower.resize(n);
	upper.resize(n);
	for (size_t i = 0; i < n; i++) {
		lower[i] = prior.GetLowerBound(var_ix[i]);
		upper[i] = prior.GetUpperBound(var_ix[i]);
	}
	std::vector<VectorReal> rows;
	history.GetHistory(var_ix, rows);
	if (rows.size() > max_history_samples && max_history_samples > 0) {
		std::vector<size_t> use;
		size_t subsample = rows.size() / max_history_samples;
		if (subsample > 1) {
			use.resize(rows.size() / subsample);
			for (size_t i = 0; i < use.size(); i++) use[i] = i * subsample;
		} else {
			use.resize(rows.size());
			for (size_t i = 0; i < use.size(); i++) use[i] = i;
		}
		while (use.size() > max_history_samples) use.erase(use.begin() + rng.GetUnsignedInt((unsigned)use.size() - 1));
		std::vector<VectorReal> sel;
		for (size_t ix : use) sel.push_back(rows[ix]);
		rows.swap(sel);
	}
	return InitializeImpl(rows, prior, rng);
}

bool ProposalGlobalCovariance::InitializeImpl(const std::vector<VectorReal>& rows, const Prior& prior, RNG&)
{
	// ProposalGlobalCovariance::InitializeImpl (ProposalGlobalCovariance.cpp:66-112)
	covariance.assign(n * n, 0.0);
	if (rows.size() < 2) {
		for (size_t j = 0; j < n; j++) {
			Real var;
			prior.EvaluateMarginalVariance(var_ix[j], var);
			covariance[j + j * n] = var;
		}
	} else {
		VectorReal mean(n, 0.0);
		for (auto& r : rows)
			for (size_t i = 0; i < n; i++) mean[i] += r[i];
		for (size_t i = 0; i < n; i++) mean[i] /= (Real)rows.size();
		for (auto& r : rows)
			for (size_t i = 0; i < n; i++)
				for (size_t j = 0; j < n; j++) covariance[i + j * n] += (r[i] - mean[i]) * (r[j] - mean[j]);
		for (Real& v : covariance) v /= (Real)(rows.size() - 1);
		for (size_t j = 0; j < n; j++) {
			Real var;
			prior.EvaluateMarginalVariance(var_ix[j], var);
			covariance[j + j * n] = std::max(covariance[j + j * n], 1e-6 * var);
		}
	}
	// Cholesky (covariance.llt()). With fewer history samples than variables the empirical covariance is singular; Eigen's
	// LLT would silently produce NaNs there, here the proposal falls back to the diagonal of the covariance.
	auto cholesky = [&]() {
		chol.assign(n * n, 0.0);
		for (size_t j = 0; j < n; j++) {
			Real d = covariance[j + j * n];
			for (size_t k = 0; k < j; k++) d -= chol[j + k * n] * chol[j + k * n];
			if (!(d > 1e-12 * covariance[j + j * n])) return false;
			chol[j + j * n] = sqrt(d);
			for (size_t i = j + 1; i < n; i++) {
				Real v = covariance[i + j * n];
				for (size_t k = 0; k < j; k++) v -= chol[i + k * n] * chol[j + k * n];
				chol[i + j * n] = v / chol[j + j * n];
			}
		}
		return true;
	};
	if (!cholesky()) {
		for (size_t i = 0; i < n; i++)
			for (size_t j = 0; j < n; j++)
				if (i != j) covariance[i + j * n] = 0.0;
		if (!cholesky()) return false;
	}
	return true;
}

void ProposalGlobalCovariance::Update(RNG& rng, bool scaling_frozen)
{
	// Proposal::Update, Proposal.cpp:196-211 (the draw happens regardless so that the stream does not depend on the flag)
	Real learn_rate = 1.0 + rng.GetReal() * scaling_learning_rate;
	if (scaling_frozen) return;
	if (current_acceptance_rate_ema < 0.952381 * target_acceptance_rate) {
		adaptive_scale /= learn_rate;
		adaptive_scale = std::max(adaptive_scale, (Real)1e-4);
	} else if (current_acceptance_rate_ema > 1.05 * target_acceptance_rate) {
		adaptive_scale *= learn_rate;
		adaptive_scale = std::min(adaptive_scale, (Real)10.0);
	}
}

void ProposalGlobalCovariance::NotifyAccepted(bool accepted)
{
	// Proposal.cpp:213-222
	const Real ema_alpha = 2.0 / (scaling_ema_period + 1);
	current_acceptance_rate_ema += ((accepted ? 1.0 : 0.0) - current_acceptance_rate_ema) * ema_alpha;
}

Real Proposal::ReflectOnBounds(Real x, Real lb, Real ub)
{
	// Proposal.cpp:385-397
	for (;;) {
		if (x < lb) x = lb + (lb - x);
		else if (x > ub) x = ub - (x - ub);
		else break;
	}
	return x;
}

void ProposalGlobalCovariance::GetNewSample(const VectorReal& current, VectorReal& proposed, RNG& rng)
{
	// ProposalGlobalCovariance.cpp:19-41
	const Real t_scale = TScale(rng);
	VectorReal z(n);
	for (size_t i = 0; i < n; i++) z[i] = rng.GetNormal();
	proposed.assign(n, 0.0);
	for (size_t i = 0; i < n; i++) {
		Real v = 0.0;
		for (size_t k = 0; k <= i; k++) v += chol[i + k * n] * z[k];
		proposed[i] = ReflectOnBounds(current[i] + v * (t_scale * adaptive_scale), lower[i], upper[i]);
	}
}

// ---------------------------------------------------------------------------------------------- mixture proposal
namespace {

// lag-`lag` autocorrelation of a history column (SummaryStats.h acf)
Real autocorrelation(const std::vector<VectorReal>& rows, size_t col, size_t lag, Real mu, Real var)
{
	const size_t n = rows.size();
	Real c = 0.0;
	for (size_t i = 0; i + lag < n; i++) c += (rows[i][col] - mu) * (rows[i + lag][col] - mu);
	return c / ((Real)(n - lag) * var);
}

} // namespace

bool ProposalGaussianMixture::InitializeImpl(const std::vector<VectorReal>& rows, const Prior& prior, RNG& rng)
{
	// ProposalGaussianMixture::InitializeImpl (ProposalGaussianMixture.cpp:117-251)
	bool have = false;
	const size_t ns = rows.size();
	if (ns >= 2) {
		// the smallest per-variable effective sample size scales the log-likelihood in the adjusted AIC and the
		// regularisation of the covariance estimates
		Real min_ess = kInf;
		const int lag_max = std::max(5, (int)(10 * log10((Real)ns)));
		for (size_t i = 0; i < n; i++) {
			Real mu = 0.0, var = 0.0;
			for (auto& r : rows) mu += r[i];
			mu /= (Real)ns;
			for (auto& r : rows) var += (r[i] - mu) * (r[i] - mu);
			var /= (Real)(ns - 1);
			Real rho = 0.0;
			if (var > 0.0)
				for (int lag = 1; lag < lag_max && (size_t)lag < ns; lag++) rho += autocorrelation(rows, i, (size_t)lag, mu, var);
			min_ess = std::min(min_ess, (Real)ns / (1.0 + 2.0 * rho));
		}
		if (!(min_ess > 0.0)) min_ess = 1.0; // strongly anti-correlated or constant history
		min_ess = std::min(min_ess, (Real)ns);
		const Real aic_adjust = min_ess / (Real)ns;
		Real best = kInf;
		static const size_t num_components[7] = { 1, 2, 3, 4, 5, 8, 13 };
		for (size_t k : num_components) {
			if (min_ess < (Real)(k * (1 + std::min(n / 2, (size_t)10)))) continue; // not enough effective samples
			GaussianMixture candidate;
			if (!candidate.Fit(rows, k, rng, (Real)ns / min_ess)) continue;
			const Real nparam = 0.5 * candidate.GetAIC() + candidate.GetLogLikelihood();
			const Real score = adjusted_aic ? 2.0 * nparam - 2.0 * aic_adjust * candidate.GetLogLikelihood() : candidate.GetAIC();
			if (score < best) {
				gmm = candidate;
				best = candidate.GetAIC(); // as the reference does, also when the adjusted score selects (:163-166)
				have = true;
			}
		}
	}
	if (!have) {
		// first initialisation, or no mixture could be fitted: one component with the prior's marginal variances
		VectorReal mean(n, 0.0);
		std::vector<Real> cov(n * n, 0.0);
		for (size_t i = 0; i < n; i++) {
			Real m = 0.0, v = 1.0;
			if (!prior.EvaluateMarginalMean(var_ix[i], m)) m = 0.0;
			if (!prior.EvaluateMarginalVariance(var_ix[i], v)) v = 1.0;
			mean[i] = m;
			cov[i + i * n] = v;
		}
		if (!gmm.Set({ mean }, { cov }, VectorReal(1, 1.0))) return false;
	}
	scales.assign(gmm.GetNumComponents(), 2.38 / sqrt((Real)n));
	acceptance_rate_emas.assign(gmm.GetNumComponents(), target_acceptance_rate);
	selected_component = -1;
	return true;
}

void ProposalGaussianMixture::GetNewSample(const VectorReal& current, VectorReal& proposed, RNG& rng)
{
	// ProposalGaussianMixture.cpp:21-45 (the Metropolis-Hastings ratio below ignores the t scale, as the reference's does)
	VectorReal resp;
	gmm.CalculateResponsibilities(current, resp);
	selected_component = (long)SampleIndex(rng, resp);
	const Real t_scale = TScale(rng);
	const std::vector<Real>& L = gmm.GetComponent((size_t)selected_component).chol;
	VectorReal z(n);
	for (size_t i = 0; i < n; i++) z[i] = rng.GetNormal();
	proposed.assign(n, 0.0);
	for (size_t i = 0; i < n; i++) {
		Real v = 0.0;
		for (size_t k = 0; k <= i; k++) v += L[i + k * n] * z[k];
		proposed[i] = ReflectOnBounds(current[i] + v * (t_scale * scales[(size_t)selected_component]), lower[i], upper[i]);
	}
}

Real ProposalGaussianMixture::CalculateMHRatio(const VectorReal& current, const VectorReal& proposed) const
{
	// ProposalGaussianMixture.cpp:47-67: q(new | cur) = sum_k resp_k(cur) N(new - cur; 0, scale_k^2 Sigma_k), and the reverse
	VectorReal fwd, rev, v(n);
	gmm.CalculateResponsibilities(current, fwd);
	gmm.CalculateResponsibilities(proposed, rev);
	Real fwd_logp = -kInf, rev_logp = -kInf;
	for (size_t k = 0; k < gmm.GetNumComponents(); k++) {
		const GaussianMixture::Component& c = gmm.GetComponent(k);
		for (size_t i = 0; i < n; i++) v[i] = (proposed[i] - current[i]) / scales[k];
		GaussianMixture::SolveLower(c.chol, n, v);
		Real q = 0.0;
		for (size_t i = 0; i < n; i++) q += v[i] * v[i];
		// the step density is even in the step, so forward and reverse share the quadratic form
		const Real base = -log(scales[k] * scales[k]) + c.logC - 0.5 * q;
		fwd_logp = GaussianMixture::LogSum(fwd_logp, base + log(fwd[k]));
		rev_logp = GaussianMixture::LogSum(rev_logp, base + log(rev[k]));
	}
	return rev_logp - fwd_logp;
}

void ProposalGaussianMixture::Update(RNG& rng, bool scaling_frozen)
{
	// ProposalGaussianMixture.cpp:69-90: only the scale of the component that made the previous proposal moves
	const Real learn_rate = 1.0 + rng.GetReal() * scaling_learning_rate * (Real)gmm.GetNumComponents();
	if (scaling_frozen || selected_component < 0) return;
	Real& scale = scales[(size_t)selected_component];
	const Real ema = acceptance_rate_emas[(size_t)selected_component];
	if (ema < target_acceptance_rate / (1.0 - scaling_learning_rate)) scale = std::max(scale / learn_rate, (Real)1e-4);
	else if (ema > (1.0 + scaling_learning_rate) * target_acceptance_rate) scale = std::min(scale * learn_rate, (Real)10.0);
}

void ProposalGaussianMixture::NotifyAccepted(bool accepted)
{
	// ProposalGaussianMixture.cpp:92-104
	if (selected_component < 0) return;
	const Real ema_alpha = 2.0 / (scaling_ema_period + 1);
	Real& ema = acceptance_rate_emas[(size_t)selected_component];
	ema += ((accepted ? 1.0 : 0.0) - ema) * ema_alpha;
}

// ---------------------------------------------------------------------------------------------- sampler
bool SamplerPT::Initialize()
{
	if (!varset || !prior || !likelihood) {
		last_error = "variable set, prior and likelihood must be set";
		return false;
	}
	if (!MakeProposal()) {
		last_error = "proposal_type \"" + s.proposal_type + "\" is not supported (global_covariance, gaussian_mixture, gaussian_mixture_adjustedAIC)";
		return false;
	}
	if (s.blocking_strategy != "one_block" && s.blocking_strategy != "no_blocking" && s.blocking_strategy != "Turek") {
		last_error = s.blocking_strategy == "clustered_autoblock" ? "blocking_strategy \"clustered_autoblock\" (clustered sample history) is not supported: one_block, no_blocking, Turek"
		                                                          : "Unknown blocking strategy \"" + s.blocking_strategy + "\"";
		return false;
	}
	if (s.swapping_scheme != "deterministic_even_odd" && s.swapping_scheme != "stochastic_even_odd" && s.swapping_scheme != "stochastic_random") {
		last_error = "Unknown swapping scheme \"" + s.swapping_scheme + "\"";
		return false;
	}
	num_variables = varset->GetNumVariables();
	if (s.num_chains < 1) return false;

	// temperatures, SamplerPT.cpp:83-93: power law below temperature_max, chain 0 at temperature 0
	temperatures.assign(s.num_chains, 0.0);
	for (size_t i = 1; i + 1 < s.num_chains; i++) {
		Real alpha = i / (Real)(s.num_chains - 1);
		temperatures[i] = s.temperature_schedule_max * pow(alpha, s.temperature_schedule_power);
	}
	temperatures[s.num_chains - 1] = s.temperature_schedule_max;

	if (s.adapt_proposal_samples > 0) {
		// SamplerPT.cpp:75-78
		proposal_scaling_ema_period = (size_t)ceil(s.adapt_proposal_samples * s.use_every_nth * (1.0 - s.exchange_probability) / 10);
		proposal_scaling_learning_rate = pow(100.0, 1.0 / proposal_scaling_ema_period) - 1.0;
	}

	// history sizing, SamplerPT.cpp:113-126
	size_t expected = s.adapt_proposal_samples * s.use_every_nth;
	if (temperatures.size() > 1 && s.swapping_scheme == "deterministic_even_odd") expected *= (s.num_exploration_steps + 1);
	size_t subsampling = 1, sample_history = expected;
	if (sample_history > s.max_history_size) {
		subsampling = (expected + s.max_history_size - 1) / s.max_history_size;
		sample_history = expected / subsampling;
	}

	rng.Seed(s.rngseed, 0);
	chains.assign(temperatures.size(), Chain());
	for (size_t i = 0; i < chains.size(); i++) {
		Chain& c = chains[i];
		c.temperature = temperatures[i];
		c.rng.Seed(s.rngseed, i + 1);
		c.current_var_values.assign(num_variables, 0.0);
		c.history.Initialize(num_variables, sample_history, subsampling);
		if (!AdaptChainProposal(c)) { // SamplerPTChain::Initialize ends with AdaptProposal on the empty history (SamplerPTChain.cpp:95)
			if (last_error.empty()) last_error = "Proposal initialization failed.";
			return false;
		}
	}
	previous_swap_even = false;
	proposal_adaptations_done = 0;
	proposal_scaling_adaptations_done = false;
	samples.clear();
	return true;
}

// SamplerPTChain.cpp:431-437
std::shared_ptr<Proposal> SamplerPT::MakeProposal() const
{
	if (s.proposal_type == "global_covariance") return std::make_shared<ProposalGlobalCovariance>();
	if (s.proposal_type == "gaussian_mixture") return std::make_shared<ProposalGaussianMixture>(false);
	if (s.proposal_type == "gaussian_mixture_adjustedAIC") return std::make_shared<ProposalGaussianMixture>(true);
	return nullptr;
}

bool SamplerPT::EvaluateAll(const std::vector<size_t>& which, const MatrixReal& proposals, VectorReal& lpriors, VectorReal& llhs)
{
	// Sampler::EvaluatePriorLikelihood (Sampler.cpp:150-188) for the listed chains: priors on the host, likelihoods
	// either in one batched call or one by one -- both give the same numbers for a likelihood whose batched entry is exact.
	const size_t n = which.size();
	lpriors.assign(n, -kInf);
	llhs.assign(n, -kInf);
	MatrixReal sub(num_variables, n);
	for (size_t k = 0; k < n; k++) {
		std::copy(proposals.col(which[k]), proposals.col(which[k]) + num_variables, sub.col(k));
		if (!prior->EvaluateLogPDF(0, sub.col(k), lpriors[k])) {
			last_error = "Prior evaluation failed";
			return false;
		}
		if (lpriors[k] != lpriors[k]) {
			last_error = "NAN in prior calculation";
			return false;
		}
	}
	if (s.batched) {
		if (!likelihood->EvaluateLogProbabilityBatch(sub, llhs)) {
			last_error = "Likelihood evaluation failed";
			return false;
		}
		num_batched_calls++;
	} else {
		VectorReal v(num_variables);
		for (size_t k = 0; k < n; k++) {
			v.assign(sub.col(k), sub.col(k) + num_variables);
			if (!likelihood->EvaluateLogProbability(0, v, llhs[k])) {
				last_error = "Likelihood evaluation failed";
				return false;
			}
		}
	}
	for (size_t k = 0; k < n; k++) {
		llhs[k] *= likelihood->GetLearningRate(); // Sampler.cpp:168
		num_likelihood_evaluations++;             // :169
		if (llhs[k] != llhs[k]) {
			last_error = "NAN in likelihood calculation";
			return false;
		}
	}
	return true;
}

bool SamplerPT::FindStartingPositions()
{
	// SamplerPTChain::FindStartingPosition (SamplerPTChain.cpp:188-215), all chains per try in one batch
	std::vector<size_t> pending(chains.size());
	for (size_t i = 0; i < chains.size(); i++) pending[i] = i;
	MatrixReal proposals(num_variables, chains.size());
	for (size_t t = 0; t < s.initial_position_tries && !pending.empty(); t++) {
		for (size_t ci : pending) prior->Sample(proposals.col(ci), &chains[ci].rng);
		VectorReal lp, ll;
		if (!EvaluateAll(pending, proposals, lp, ll)) return false;
		std::vector<size_t> still;
		for (size_t k = 0; k < pending.size(); k++) {
			Chain& c = chains[pending[k]];
			c.current_var_values.assign(proposals.col(pending[k]), proposals.col(pending[k]) + num_variables);
			c.lprior = lp[k];
			c.llh = ll[k];
			c.lpowerposterior = c.lprior + c.temperature * c.llh;
			if (!(c.lpowerposterior > -kInf)) still.push_back(pending[k]);
		}
		pending.swap(still);
	}
	if (!pending.empty()) {
		last_error = "Could not find starting position with power posterior != inf after " + std::to_string(s.initial_position_tries) + " tries";
		return false;
	}
	return true;
}

bool SamplerPT::ExchangeMove(Chain& chain1, Chain& chain2)
{
	// SamplerPTChain::ExchangeMove, SamplerPTChain.cpp:328-381: uses cached lprior / llh only, no likelihood call
	chain1.attempted_exchange++;
	Real p1 = (chain1.temperature == 0.0) ? chain2.lprior : chain1.temperature * chain2.llh + chain2.lprior;
	Real p2 = (chain2.temperature == 0.0) ? chain1.lprior : chain2.temperature * chain1.llh + chain1.lprior;
	Real tp = exp((p1 + p2) - (chain1.lpowerposterior + chain2.lpowerposterior));
	tp = std::min((Real)1.0, tp);
	const bool swap = rng.GetReal() < tp;
	if (swap) {
		chain1.accepted_exchange++;
		std::swap(chain1.current_var_values, chain2.current_var_values);
		std::swap(chain1.llh, chain2.llh);
		std::swap(chain1.lprior, chain2.lprior);
		chain1.lpowerposterior = p1;
		chain2.lpowerposterior = p2;
	}
	if (chain1.temperature != 0.0) chain1.history.AddSample(chain1.current_var_values);
	if (chain2.temperature != 0.0) chain2.history.AddSample(chain2.current_var_values);
	return swap;
}

void SamplerPT::DoExchangeMove()
{
	// SamplerPT.cpp:277-306
	if (s.swapping_scheme == "stochastic_random") {
		size_t ci = rng.GetUnsignedInt((unsigned)chains.size() - 2);
		ExchangeMove(chains[ci], chains[ci + 1]);
		return;
	}
	size_t start_ix;
	if (previous_swap_even) {
		start_ix = 1;
		previous_swap_even = false;
	} else {
		start_ix = 0;
		previous_swap_even = true;
	}
	for (size_t ci = start_ix; ci < chains.size(); ci += 2) {
		size_t ix2 = ci + 1;
		if (ix2 == chains.size()) ix2 = 0;
		ExchangeMove(chains[ci], chains[ix2]);
	}
}

bool SamplerPT::DoMutateMove()
{
	// SamplerPT::DoMutateMove (SamplerPT.cpp:308-319) + SamplerPTChain::MutateMove (SamplerPTChain.cpp:217-313). A chain updates
	// its variable blocks one after the other, each with a likelihood evaluation at the vector that differs from the chain's
	// current one in that block only; the chains are independent of each other, so round b evaluates block b of EVERY chain
	// that has one in ONE batched call (the T = 0 chain draws from the prior in round 0 and has no further rounds).
	const size_t C = chains.size();
	size_t rounds = 1;
	for (const Chain& c : chains) rounds = std::max(rounds, c.blocks.size());
	std::vector<VectorReal> block_current(C), block_new(C);
	for (size_t round = 0; round < rounds; round++) {
		// 1. every participating chain proposes, host side, per-chain streams
		std::vector<size_t> which;
		for (size_t ci = 0; ci < C; ci++)
			if (chains[ci].temperature == 0.0 ? round == 0 : round < chains[ci].blocks.size()) which.push_back(ci);
		MatrixReal proposals(num_variables, which.size());
		for (size_t k = 0; k < which.size(); k++) {
			const size_t ci = which[k];
			Chain& c = chains[ci];
			if (c.temperature == 0.0) {
				prior->Sample(proposals.col(k), &c.rng);
				continue;
			}
			Chain::Block& b = c.blocks[round];
			b.proposal->Update(c.rng, proposal_scaling_adaptations_done);
			block_current[ci].resize(b.variable_indices.size());
			for (size_t i = 0; i < b.variable_indices.size(); i++) block_current[ci][i] = c.current_var_values[b.variable_indices[i]];
			b.proposal->GetNewSample(block_current[ci], block_new[ci], c.rng);
			std::copy(c.current_var_values.begin(), c.current_var_values.end(), proposals.col(k));
			for (size_t i = 0; i < b.variable_indices.size(); i++) proposals(b.variable_indices[i], k) = block_new[ci][i];
		}
		// 2. ONE batched evaluation of all proposals of this round
		VectorReal lp, ll;
		std::vector<size_t> slots(which.size());
		for (size_t k = 0; k < which.size(); k++) slots[k] = k; // columns of `proposals`
		if (!EvaluateAll(slots, proposals, lp, ll)) return false;
		// 3. every chain tests and accepts / rejects
		for (size_t k = 0; k < which.size(); k++) {
			Chain& c = chains[which[k]];
			if (c.temperature == 0.0) {
				c.current_var_values.assign(proposals.col(k), proposals.col(k) + num_variables);
				c.lprior = lp[k];
				c.llh = ll[k];
				c.lpowerposterior = (c.llh == -kInf) ? c.lprior : c.lprior + c.temperature * c.llh; // :230-236
				c.attempted_mutate++;
				c.accepted_mutate++;
				continue;
			}
			Chain::Block& b = c.blocks[round];
			const Real new_lpp = lp[k] + c.temperature * ll[k];
			// TestSample, SamplerPTChain.cpp:465-482
			c.attempted_mutate++;
			bool accept = false;
			if (new_lpp > -kInf) {
				Real tp = exp(new_lpp - c.lpowerposterior + b.proposal->CalculateMHRatio(block_current[which[k]], block_new[which[k]]));
				tp = std::min((Real)1.0, tp);
				accept = c.rng.GetReal() < tp;
			}
			if (accept) {
				c.accepted_mutate++;
				c.current_var_values.assign(proposals.col(k), proposals.col(k) + num_variables);
				c.lprior = lp[k];
				c.llh = ll[k];
				c.lpowerposterior = new_lpp;
			}
			b.proposal->NotifyAccepted(accept);
		}
	}
	for (Chain& c : chains)
		if (c.temperature != 0.0) c.history.AddSample(c.current_var_values);
	return true;
}

void SamplerPT::EmitSample()
{
	// SamplerPT::EmitSample (SamplerPT.cpp:321-330): every fixed-temperature chain reports its state
	for (const Chain& c : chains) samples.push_back(EmittedSample{ c.current_var_values, c.lprior, c.llh, c.temperature });
	// handler by handler, chain by chain, weight 1 (SamplerPT.cpp:323-329)
	for (const auto& handler : sample_handlers)
		for (const Chain& c : chains) handler->ReceiveSample(c.current_var_values, c.lprior, c.llh, c.temperature, 1.0);
}

bool SamplerPT::AdaptChainProposal(Chain& c)
{
	// SamplerPTChain::AdaptProposal (SamplerPTChain.cpp:109-170): nothing for the chain that samples the prior; the blocking
	// strategy's blocks from the history, a fresh proposal per block from the history of its variables, history discarded
	if (c.temperature == 0.0) return true;
	std::vector<std::vector<size_t>> blocks;
	if (!GetVariableBlocks(s.blocking_strategy, c.history, num_variables, blocks)) {
		last_error = "Unknown blocking strategy \"" + s.blocking_strategy + "\"";
		return false;
	}
	c.blocks.clear();
	c.blocks.resize(blocks.size());
	for (size_t b = 0; b < blocks.size(); b++) {
		c.blocks[b].variable_indices = blocks[b];
		c.blocks[b].proposal = MakeProposal();
		if (!c.blocks[b].proposal->Initialize(c.history, s.adapt_proposal_max_history_samples, *prior, blocks[b], c.rng)) {
			last_error = "Proposal adaptation failed";
			return false;
		}
		c.blocks[b].proposal->SetScalingSchedule(proposal_scaling_ema_period, proposal_scaling_learning_rate);
		c.blocks[b].proposal->SetTDof(s.proposal_t_dof);
	}
	c.history.Reset();
	return true;
}

bool SamplerPT::AdaptProposals()
{
	for (Chain& c : chains)
		if (!AdaptChainProposal(c)) return false;
	return true;
}

bool SamplerPT::Run()
{
	// SamplerPT::RunImpl, SamplerPT.cpp:174-260
	if (!FindStartingPositions()) return false;
	const size_t total_samples = s.num_samples * s.use_every_nth;
	for (size_t si = 0; si < total_samples; si++) {
		const size_t sample_ix = si / s.use_every_nth;
		bool result = true;
		if (chains.size() > 1) {
			if (s.swapping_scheme == "deterministic_even_odd") {
				DoExchangeMove();
				for (size_t ei = 0; ei < s.num_exploration_steps && result; ei++) result = DoMutateMove();
			} else {
				if (rng.GetReal() < s.exchange_probability) DoExchangeMove();
				else result = DoMutateMove();
			}
		} else {
			result = DoMutateMove();
		}
		if (!result) {
			if (last_error.empty()) last_error = "Sample step failed";
			return false;
		}
		if ((si + 1) % s.use_every_nth == 0) {
			EmitSample();
			if (s.adapt_proposal_samples > 0 && ((sample_ix + 1) % s.adapt_proposal_samples == 0) && si + 1 != total_samples &&
			    proposal_adaptations_done < s.adapt_proposal_times) {
				if (!AdaptProposals()) return false;
				proposal_adaptations_done++;
			}
			if (s.stop_proposal_scaling > 0 && sample_ix > s.stop_proposal_scaling) proposal_scaling_adaptations_done = true;
		}
	}
	return true;
}

} // namespace bcm3
