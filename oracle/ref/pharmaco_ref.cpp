// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- pharmaco_population checker around the reference's REAL compartment model:
// src/pharmaco/PharmacokineticModel.cpp is compiled from where it lies (oracle/ref/Makefile) together with Eigen's matrix
// exponential (unsupported/Eigen/MatrixFunctions); the glue is oracle/pharmaco_glue.hpp.
#include "Utils.h"
#include "PharmacokineticModel.h"

#include "../pharmaco_glue.hpp"

namespace refglue {
double ndtri(double p); // poppk_ref.cpp
}
namespace pharmaco_glue {
double ndtri(double p) { return refglue::ndtri(p); }
}

namespace {

struct RefModel {
	PharmacokineticModel model;
	bool peripheral = false;
	int num_transit = 0;
	void configure(bool use_peripheral, int transit)
	{
		// PostInitialize, PharmacoLikelihoodPopulation.cpp:110-118
		peripheral = use_peripheral;
		num_transit = transit;
		model.SetUsePeripheralCompartment(use_peripheral);
		model.SetNumTransitCompartments((size_t)transit);
	}
	bool biphasic = false, metabolite = false;
	void configure_single(bool use_biphasic, bool use_metabolite)
	{
		// PharmacoLikelihoodSingle::PostInitialize, PharmacoLikelihoodSingle.cpp:127-149
		biphasic = use_biphasic;
		metabolite = use_metabolite;
		model.SetUseBiphasicAbsorption(use_biphasic);
		if (use_metabolite) model.SetMetaboliteElimination(1.0);
		model.SetUseMetabolite(use_metabolite);
	}
	void set_single(double direct_absorption_rate, double metabolite_conversion_rate)
	{
		if (biphasic) model.SetDirectAbsorptionRate(direct_absorption_rate);
		if (metabolite) model.SetMetaboliteConversionRate(metabolite_conversion_rate);
	}
	bool solve(double absorption, double excretion, double elimination, double kf, double kb, double transit_rate, double bioavailability,
	           const std::vector<double>& tt, const std::vector<double>& td, const std::vector<double>& ot, std::vector<double>& out)
	{
		// SetupSimulation, cpp:309-338
		if (peripheral) {
			model.SetPeripheralForwardRate(kf);
			model.SetPeripheralBackwardRate(kb);
		}
		if (num_transit > 0) model.SetTransitRate(transit_rate);
		model.SetBioavailability(bioavailability);
		model.SetAbsorption(absorption);
		model.SetExcretion(excretion);
		model.SetElimination(elimination);
		VectorReal treatment_times = Eigen::Map<const VectorReal>(tt.data(), (Eigen::Index)tt.size());
		VectorReal treatment_doses = Eigen::Map<const VectorReal>(td.data(), (Eigen::Index)td.size());
		VectorReal observation_timepoints = Eigen::Map<const VectorReal>(ot.data(), (Eigen::Index)ot.size());
		VectorReal central = VectorReal::Constant((Eigen::Index)ot.size(), std::numeric_limits<Real>::quiet_NaN());
		const bool ok = model.Solve(treatment_times, treatment_doses, observation_timepoints, central, nullptr);
		for (size_t i = 0; i < ot.size(); i++) out[i] = central((Eigen::Index)i);
		return ok;
	}
};

} // namespace

extern "C" int oracle_pharmaco_evaluate(const oracle_pharmaco_problem* prob, size_t num_chains, const double* values, double* logp, double* conc,
                                        double* patient_ll, int num_threads)
{
	return pharmaco_glue::evaluate<RefModel>(prob, num_chains, values, logp, conc, patient_ll, num_threads);
}
