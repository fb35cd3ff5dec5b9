// The reference's sample sinks (src/sampler/SampleHandler.h:5-9): the sampler hands every emitted sample of every
// fixed-temperature chain to every handler added with Sampler::AddSampleHandler (Sampler.cpp:49-52, SamplerPT.cpp:321-330).
//   SampleHandlerTSV                  src/sampler/SampleHandlerTSV.cpp:16-52  -- tab-separated text, the samples at temperature 1 only
//   SampleHandlerStoreMaxAPosteriori  src/sampler/SampleHandlerStoreMaxAPosteriori.cpp:5-32 -- the best log-posterior seen, at any temperature
// The NetCDF-4 writer (SampleHandlerNetCDF, what bcminf installs) is not built: no HDF5 in this image.
#pragma once

#include <cstdio>
#include <limits>
#include <string>
#include <vector>

#include "Types.h"

namespace bcm3 {

class SampleHandler {
public:
	virtual ~SampleHandler() {}
	virtual void ReceiveSample(const VectorReal& sample, Real lprior, Real llh, Real temperature, Real weight) = 0;
};

class SampleHandlerTSV : public SampleHandler {
public:
	void SetFile(const std::string& fn) { filename = fn; }
	// header: the variable names, then "log prior", "log likelihood", "weight" (sample_count and the temperatures are not used)
	bool Initialize(size_t /*sample_count*/, const std::vector<std::string>& variables, const VectorReal& /*output_temperatures*/)
	{
		FILE* file = fopen(filename.c_str(), "w");
		if (!file) return false;
		for (const std::string& v : variables) fprintf(file, "%s\t", v.c_str());
		fprintf(file, "log prior\tlog likelihood\tweight\n");
		fclose(file);
		return true;
	}
	// The file is opened and closed per sample and every number is "%.6g", as in the reference -- including its line structure:
	// the log likelihood ends the line and the weight stands on a line of its own (SampleHandlerTSV.cpp:45-47).
	void ReceiveSample(const VectorReal& values, Real lprior, Real llh, Real temperature, Real weight) override
	{
		if (temperature != 1.0) return;
		FILE* file = fopen(filename.c_str(), "a");
		if (!file) return;
		for (size_t i = 0; i < (size_t)values.size(); i++) fprintf(file, "%.6g\t", values[i]);
		fprintf(file, "%.6g\t", lprior);
		fprintf(file, "%.6g\n", llh);
		fprintf(file, "%.6g\n", weight);
		fclose(file);
	}

private:
	std::string filename;
};

class SampleHandlerStoreMaxAPosteriori : public SampleHandler {
public:
	void Reset() { MAP_lposterior = MAP_llikelihood = -std::numeric_limits<Real>::infinity(); }
	void ReceiveSample(const VectorReal& values, Real lprior, Real llh, Real /*temperature*/, Real /*weight*/) override
	{
		const Real lposterior = lprior + llh; // any temperature
		if (lposterior > MAP_lposterior) {
			MAP_lposterior = lposterior;
			MAP_llikelihood = llh;
			MAP_values = values;
		}
	}
	Real GetMAPlposterior() const { return MAP_lposterior; }
	Real GetMAPllikelihood() const { return MAP_llikelihood; }
	const VectorReal& GetMAP() const { return MAP_values; }

private:
	Real MAP_lposterior = -std::numeric_limits<Real>::infinity(), MAP_llikelihood = -std::numeric_limits<Real>::infinity();
	VectorReal MAP_values;
};

} // namespace bcm3
