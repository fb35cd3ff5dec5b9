// GPU-backed drop-in for LikelihoodPopPKTrajectory (src/likelihoods/LikelihoodPopPKTrajectory.{h,cpp}) on top of the
// C ABI (include/bcm3b200.h). likelihood.xml surface: <bcm_likelihood type="pop_pk_trajectory"><pk_model drug=
// type= trial= pkdata_file= .../> as in the reference (cpp:58-87). The NetCDF reader is out of scope (SURVEY 8f row
// 3): the trial arrays are supplied with SetTrialData() before PostInitialize().
// The same class, constructed with single_patient = true, is the drop-in for LikelihoodPharmacokineticTrajectory
// (src/likelihoods/LikelihoodPharmacokineticTrajectory.{h,cpp}; type="pharmacokinetic_trajectory"): <pk_model ... patient=>
// names ONE patient of the trial (its cpp:96, 161-166: looked up in the "patients" dimension -- SetPatientIDs() supplies that
// dimension; without it the attribute is the patient's index), whose rates the chain's variables are.
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

	LikelihoodPopPKTrajectoryB200(size_t sampling_threads, size_t evaluation_threads, bool single_patient = false);
	void SetPatientIDs(const std::vector<std::string>& ids) { patient_ids = ids; }
	void SetPatientID(const std::string& patient) { patient_id = patient; } // LikelihoodPharmacokineticTrajectory::SetPatientID / the pk.patient option
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
	std::string drug, pk_type_str, trial_name, pkdata_file, fixed_attributes, patient_id;
	std::vector<std::string> patient_ids;
	bool single = false;
	TrialData trial;
	void* handle = nullptr;
	int device0 = 0, num_devices = 1;
	std::vector<int> status;
	std::string last_error;
};
