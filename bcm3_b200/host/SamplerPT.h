// Host PT driver: mirror of the reference's parallel-tempered Metropolis-Hastings sampler for the part that drives the
// likelihood (src/sampler/SamplerPT.cpp:174-330, SamplerPTChain.cpp:188-381,465-482, Proposal.cpp:196-225,
// ProposalGlobalCovariance.cpp:19-114, SampleHistory.cpp:32-67), restructured around ONE batched likelihood call per
// mutate round:   propose for all chains -> EvaluateLogProbabilityBatch -> accept/reject for all chains.
// Proposal generation, MH tests and temperature swaps stay on the host and are deterministic (RNG.h).
// Proposals: global_covariance and gaussian_mixture (the one examples/banana/config.txt asks for; ProposalGaussianMixture.cpp,
// GMM.cpp). Blocking strategies one_block, no_blocking and Turek (BlockingStrategy*.cpp): a chain's mutate move updates its
// variable blocks one after the other, each with its own proposal and its own likelihood evaluation
// (SamplerPTChain.cpp:249-307); batched, block index b of ALL chains is one EvaluateLogProbabilityBatch. The clustered
// strategy and proposal (clustered_autoblock, SampleHistoryClustering) are refused.
#pragma once

#include <functional>

#include "GaussianMixture.h"
#include "Likelihood.h"
#include "Prior.h"
#include "RNG.h"

#include "SampleHandler.h"

namespace bcm3 {

struct SamplerPTSettings {
	// [sampler]
	size_t num_samples = 1000;
	size_t use_every_nth = 1;
	uint64_t rngseed = 0;
	// [ptmhsampler] (defaults of SamplerPT::AddOptionsDescription, SamplerPT.cpp:147-172)
	size_t num_chains = 6;
	std::string proposal_type = "gaussian_mixture"; // the reference default (SamplerPT.cpp:11,152); also global_covariance, gaussian_mixture_adjustedAIC (SamplerPTChain.cpp:431-437)
	std::string blocking_strategy = "one_block";     // one_block | no_blocking | Turek (SamplerPT.cpp:10,151; SamplerPTChain.cpp:66-77)
	std::string swapping_scheme = "deterministic_even_odd";
	size_t num_exploration_steps = 1;
	size_t max_history_size = 2000;
	size_t adapt_proposal_samples = 2000;
	size_t adapt_proposal_times = 2;
	size_t adapt_proposal_max_history_samples = 2000;
	size_t stop_proposal_scaling = 6000;
	Real exchange_probability = 0.5;
	Real temperature_schedule_power = 3.0;
	Real temperature_schedule_max = 1.0;
	Real proposal_t_dof = 0.0; // ptmhsampler.proposal_t_dof: 0 = normal proposals (SamplerPT.cpp:169)
	size_t initial_position_tries = 100;
	// evaluation mode: true = one EvaluateLogProbabilityBatch per mutate round, false = one EvaluateLogProbability per chain
	bool batched = true;

	bool LoadFromConfigFile(const std::string& filename, std::string* error = nullptr); // INI: [section] key=value
	bool LoadFromConfigText(const std::string& text, std::string* error = nullptr);
};

class SampleHistory {
public:
	void Initialize(size_t num_variables, size_t history_size, size_t subsampling);
	void AddSample(const VectorReal& sample);
	void Reset() { sample_n = 0; sample_n_s = 0; } // SampleHistory.cpp:26-30: an adaptation discards the history it used
	size_t GetSampleCount() const { return sample_n < capacity ? sample_n : capacity; }
	// rows = samples, columns = variables (stored as float like the reference, SampleHistory.cpp:41)
	void GetHistory(std::vector<VectorReal>& rows) const;
	void GetHistory(const std::vector<size_t>& variable_indices, std::vector<VectorReal>& rows) const; // SampleHistory.cpp:58-68
	// Pearson correlation of the variables over the stored samples, num_variables x num_variables column-major
	// (SampleHistory.cpp:76-81 -> cor(), SummaryStats.h)
	MatrixReal GetEmpiricalCorrelation() const;
	size_t GetNumVariables() const { return nvar; }

private:
	size_t nvar = 0, capacity = 0, subsampling = 1, sample_n = 0, sample_n_s = 0;
	std::vector<float> samples;
};

// Proposal (src/sampler/Proposal.{h,cpp}): what a chain asks of its proposal distribution for the one variable block
// Complete-linkage hierarchical clustering of `n` items from their n x n distance matrix, cut at `cut_height`
// (bcm3::TreeCluster, utils/Clustering.cpp:35-99, around the C Clustering Library's pairwise-maximum-linkage treecluster):
// the clusters in the order in which the reference's walk over the merge list creates them, members ascending.
std::vector<std::vector<size_t>> TreeClusterCompleteLinkage(const MatrixReal& distance, Real cut_height);

// BlockingStrategy::GetBlocks (BlockingStrategyOneBlock.cpp, BlockingStrategyNoBlocking.cpp, BlockingStrategyTurek.cpp:8-40)
bool GetVariableBlocks(const std::string& strategy, const SampleHistory& history, size_t num_variables, std::vector<std::vector<size_t>>& blocks);

class Proposal {
public:
	virtual ~Proposal() {}
	// Proposal::Initialize (Proposal.cpp:39-140): the block's variables, their bounds, target acceptance rate, the (sub)sampled
	// history of those variables -> InitializeImpl
	bool Initialize(const SampleHistory& history, size_t max_history_samples, const Prior& prior, const std::vector<size_t>& variable_indices, RNG& rng);
	void SetScalingSchedule(size_t ema_period, Real learning_rate) { scaling_ema_period = ema_period; scaling_learning_rate = learning_rate; }
	// ptmhsampler.proposal_t_dof (SamplerPT.cpp:63,169; Proposal.cpp:45): > 0 scales every step by 1 / sqrt(w), w drawn as
	// rng.GetGamma(t_dof / 2, t_dof / 2) -- the second argument is the SCALE of RNG::GetGamma, as the reference passes it
	void SetTDof(Real dof) { t_dof = dof; }
	virtual void Update(RNG& rng, bool scaling_frozen) = 0;
	virtual void GetNewSample(const VectorReal& current, VectorReal& proposed, RNG& rng) = 0;
	virtual Real CalculateMHRatio(const VectorReal& current, const VectorReal& proposed) const = 0; // log q(cur|new) - log q(new|cur)
	virtual void NotifyAccepted(bool accepted) = 0;

protected:
	virtual bool InitializeImpl(const std::vector<VectorReal>& history_rows, const Prior& prior, RNG& rng) = 0;
	static Real ReflectOnBounds(Real x, Real lb, Real ub);
	size_t n = 0;
	std::vector<size_t> var_ix; // the block's variables (indices into the prior / variable set)
	std::vector<Real> lower, upper;
	size_t scaling_ema_period = 1000;
	Real scaling_learning_rate = 0.05, target_acceptance_rate = 0.234;
	Real t_dof = 0.0;
	Real TScale(RNG& rng) const { return t_dof > 0.0 ? 1.0 / std::sqrt(rng.GetGamma(0.5 * t_dof, 0.5 * t_dof)) : 1.0; }
};

class ProposalGlobalCovariance : public Proposal {
public:
	void Update(RNG& rng, bool scaling_frozen) override;
	void GetNewSample(const VectorReal& current, VectorReal& proposed, RNG& rng) override;
	Real CalculateMHRatio(const VectorReal&, const VectorReal&) const override { return 0.0; }
	void NotifyAccepted(bool accepted) override;
	Real GetScale() const { return adaptive_scale; }
	const std::vector<Real>& GetCovariance() const { return covariance; }

protected:
	bool InitializeImpl(const std::vector<VectorReal>& history_rows, const Prior& prior, RNG& rng) override;

private:
	std::vector<Real> covariance, chol; // n x n, column-major; chol lower-triangular
	Real adaptive_scale = 1.0, current_acceptance_rate_ema = 0.23;
};

// ProposalGaussianMixture (src/sampler/ProposalGaussianMixture.cpp): a mixture fitted to the chain's history; a proposal
// picks the component by its responsibility for the current position, steps with that component's covariance and scale,
// and the Metropolis-Hastings ratio accounts for the position-dependent choice
class ProposalGaussianMixture : public Proposal {
public:
	explicit ProposalGaussianMixture(bool select_with_adjusted_aic) : adjusted_aic(select_with_adjusted_aic) {}
	void Update(RNG& rng, bool scaling_frozen) override;
	void GetNewSample(const VectorReal& current, VectorReal& proposed, RNG& rng) override;
	Real CalculateMHRatio(const VectorReal& current, const VectorReal& proposed) const override;
	void NotifyAccepted(bool accepted) override;
	const GaussianMixture& GetMixture() const { return gmm; }
	const VectorReal& GetScales() const { return scales; }

protected:
	bool InitializeImpl(const std::vector<VectorReal>& history_rows, const Prior& prior, RNG& rng) override;

private:
	GaussianMixture gmm;
	VectorReal scales, acceptance_rate_emas;
	long selected_component = -1;
	bool adjusted_aic;
};

struct EmittedSample {
	VectorReal values;
	Real lprior, llh, temperature;
};

class SamplerPT {
public:
	SamplerPT(const SamplerPTSettings& settings) : s(settings) {}
	void SetVariableSet(std::shared_ptr<const VariableSet> v) { varset = v; }
	void SetPrior(std::shared_ptr<Prior> p) { prior = p; }
	void SetLikelihood(std::shared_ptr<Likelihood> l) { likelihood = l; }
	void AddSampleHandler(std::shared_ptr<SampleHandler> handler) { sample_handlers.push_back(handler); } // Sampler.cpp:49-52
	bool Initialize();
	bool Run();

	const std::vector<EmittedSample>& GetSamples() const { return samples; }
	const VectorReal& GetTemperatures() const { return temperatures; }
	size_t GetNumLikelihoodEvaluations() const { return num_likelihood_evaluations; }
	size_t GetNumBatchedCalls() const { return num_batched_calls; }
	std::vector<std::vector<size_t>> GetBlocks(size_t chain) const
	{
		std::vector<std::vector<size_t>> out;
		for (auto& b : chains[chain].blocks) out.push_back(b.variable_indices);
		return out;
	}
	Real GetMutateAcceptance(size_t chain) const { return chains[chain].attempted_mutate ? chains[chain].accepted_mutate / (Real)chains[chain].attempted_mutate : 0.0; }
	Real GetExchangeAcceptance(size_t chain) const { return chains[chain].attempted_exchange ? chains[chain].accepted_exchange / (Real)chains[chain].attempted_exchange : 0.0; }
	const std::string& LastError() const { return last_error; }

private:
	struct Chain {
		Real temperature = 1.0;
		VectorReal current_var_values;
		Real lprior = -kInf, llh = -kInf, lpowerposterior = -kInf;
		size_t attempted_mutate = 0, accepted_mutate = 0, attempted_exchange = 0, accepted_exchange = 0;
		SampleHistory history;
		struct Block { // SamplerPTChain::Block, SamplerPTChain.h
			std::vector<size_t> variable_indices;
			std::shared_ptr<Proposal> proposal;
		};
		std::vector<Block> blocks;
		RNG rng;
	};
	bool AdaptChainProposal(Chain& c);

	bool EvaluateAll(const std::vector<size_t>& which, const MatrixReal& proposals, VectorReal& lpriors, VectorReal& llhs);
	bool FindStartingPositions();
	void DoExchangeMove();
	bool ExchangeMove(Chain& chain1, Chain& chain2);
	bool DoMutateMove();
	void EmitSample();
	bool AdaptProposals();
	std::shared_ptr<Proposal> MakeProposal() const;

	SamplerPTSettings s;
	std::shared_ptr<const VariableSet> varset;
	std::shared_ptr<Prior> prior;
	std::shared_ptr<Likelihood> likelihood;
	size_t num_variables = 0;
	VectorReal temperatures;
	std::vector<Chain> chains;
	RNG rng; // exchange moves, prior draws of the start-up
	bool previous_swap_even = false;
	size_t proposal_adaptations_done = 0;
	bool proposal_scaling_adaptations_done = false;
	size_t proposal_scaling_ema_period = 1000;
	Real proposal_scaling_learning_rate = 0.05;
	size_t num_likelihood_evaluations = 0, num_batched_calls = 0;
	std::vector<std::shared_ptr<SampleHandler>> sample_handlers;
	std::vector<EmittedSample> samples;
	std::string last_error;
};

} // namespace bcm3
