// GPU-backed drop-in for LikelihoodPopPKTrajectory (src/likelihoods/LikelihoodPopPKTrajectory.{h,cpp}) on top of the
// C ABI (include/bcm3b200.h). likelihood.xml surface: <bcm_likelihood type="pop_pk_trajectory"><pk_model drug=
// type= trial= pkdata_file= .../> as in the reference (cpp:58-87). The NetCDF reader is out of scope (SURVEY 8f row
// 3): the trial arrays are supplied with SetTrialData() before PostInitialize().
#pragma once

#include "Likelihood.h"

class LikelihoodPopPKTrajectoryB200 : public bcm3::Likelihood {
public:
	struct TrialData { // the NetCDF variables read at cpp:94-161
		std::vector<double> time;                    // [T]
		std::vector<double> observed_concentration;  // [P][T]
		std::vector<double> dose, dosing_interval, dose_after_dose_change, dose_change_time; // [P]
		std::vector<double> intermittent;            // [P]
		std::vector<double> treatment_interruptions; // [P][29]
	};

	LikelihoodPopPKTrajectoryB200(size_t sampling_threads, size_t evaluation_threads);
	~LikelihoodPopPKTrajectoryB200() override;

	bool Initialize(std::shared_ptr<const bcm3::VariableSet> varset, const bcm3::XmlNode& likelihood_node) override;
	void SetTrialData(const TrialData& data) { trial = data; }
	void SetDevices(int first_device, int device_count) { device0 = first_device; num_devices = device_count; }
	bool PostInitialize() override;
	bool IsReentrant() override { return true; }
	bool EvaluateLogProbability(size_t threadix, const bcm3::VectorReal& values, bcm3::Real& logp) override;
	bool EvaluateLogProbabilityBatch(const bcm3::MatrixReal& values, bcm3::VectorReal& logp) override;

	size_t GetNumPatients() const { return trial.dose.size(); }
	const std::string& LastError() const { return last_error; }

private:
	std::shared_ptr<const bcm3::VariableSet> varset;
	std::string drug, pk_type_str, trial_name, pkdata_file, fixed_attributes;
	TrialData trial;
	void* handle = nullptr;
	int device0 = 0, num_devices = 1;
	std::vector<int> status;
	std::string last_error;
};
