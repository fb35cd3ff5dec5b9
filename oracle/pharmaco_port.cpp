// TEST INFRASTRUCTURE ONLY -- pharmaco_population checker in plain C++: PharmacokineticModel::ConstructMatrix / Solve
// (src/pharmaco/PharmacokineticModel.cpp:111-247) restated with a dense matrix exponential that follows Eigen's
// (unsupported/Eigen/src/MatrixFunctions/MatrixExponential.h:64-345: Pade 3/5/7/9 by the 1-norm, Pade 13 with scaling and
// squaring, (V - U) X = (V + U) by partially pivoted LU). Pinned against the compiled reference in tests/test_oracle.py.
#include <cmath>
#include <cstring>

#include "pharmaco_glue.hpp"

extern "C" {
double oracle_ndtri(double p);
}
namespace pharmaco_glue {
double ndtri(double p) { return oracle_ndtri(p); }
}

namespace {

typedef std::vector<double> Mat; // n x n row-major

Mat mul(const Mat& a, const Mat& b, int n)
{
	Mat c((size_t)n * n, 0.0);
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++) {
			double s = 0.0;
			for (int k = 0; k < n; k++) s += a[i * n + k] * b[k * n + j];
			c[i * n + j] = s;
		}
	return c;
}

// sum_k coef[k] * mats[k] + c0 * I
Mat poly(const std::vector<const Mat*>& mats, const std::vector<double>& coef, double c0, int n)
{
	Mat out((size_t)n * n, 0.0);
	for (int e = 0; e < n * n; e++) {
		double v = 0.0;
		for (size_t k = 0; k < mats.size(); k++) v += coef[k] * (*mats[k])[e];
		out[e] = v;
	}
	for (int i = 0; i < n; i++) out[i * n + i] += c0;
	return out;
}

Mat solve_lu(Mat D, Mat X, int n) // D X' = X with partial pivoting
{
	for (int k = 0; k < n; k++) {
		int piv = k;
		for (int i = k + 1; i < n; i++)
			if (fabs(D[i * n + k]) > fabs(D[piv * n + k])) piv = i;
		if (piv != k)
			for (int j = 0; j < n; j++) {
				std::swap(D[k * n + j], D[piv * n + j]);
				std::swap(X[k * n + j], X[piv * n + j]);
			}
		for (int i = k + 1; i < n; i++) {
			const double l = D[i * n + k] / D[k * n + k];
			for (int j = k; j < n; j++) D[i * n + j] -= l * D[k * n + j];
			for (int j = 0; j < n; j++) X[i * n + j] -= l * X[k * n + j];
		}
	}
	for (int k = n - 1; k >= 0; k--)
		for (int j = 0; j < n; j++) {
			double s = X[k * n + j];
			for (int i = k + 1; i < n; i++) s -= D[k * n + i] * X[i * n + j];
			X[k * n + j] = s / D[k * n + k];
		}
	return X;
}

Mat expm(const Mat& M, int n)
{
	double l1 = 0.0;
	for (int j = 0; j < n; j++) {
		double s = 0.0;
		for (int i = 0; i < n; i++) s += fabs(M[i * n + j]);
		l1 = std::max(l1, s);
	}
	Mat A = M, U, V;
	int squarings = 0;
	if (l1 < 1.495585217958292e-002) {
		const Mat A2 = mul(A, A, n);
		U = mul(A, poly({ &A2 }, { 1.0 }, 60.0, n), n);
		V = poly({ &A2 }, { 12.0 }, 120.0, n);
	} else if (l1 < 2.539398330063230e-001) {
		const Mat A2 = mul(A, A, n), A4 = mul(A2, A2, n);
		U = mul(A, poly({ &A4, &A2 }, { 1.0, 420.0 }, 15120.0, n), n);
		V = poly({ &A4, &A2 }, { 30.0, 3360.0 }, 30240.0, n);
	} else if (l1 < 9.504178996162932e-001) {
		const Mat A2 = mul(A, A, n), A4 = mul(A2, A2, n), A6 = mul(A4, A2, n);
		U = mul(A, poly({ &A6, &A4, &A2 }, { 1.0, 1512.0, 277200.0 }, 8648640.0, n), n);
		V = poly({ &A6, &A4, &A2 }, { 56.0, 25200.0, 1995840.0 }, 17297280.0, n);
	} else if (l1 < 2.097847961257068e+000) {
		const Mat A2 = mul(A, A, n), A4 = mul(A2, A2, n), A6 = mul(A4, A2, n), A8 = mul(A6, A2, n);
		U = mul(A, poly({ &A8, &A6, &A4, &A2 }, { 1.0, 3960.0, 2162160.0, 302702400.0 }, 8821612800.0, n), n);
		V = poly({ &A8, &A6, &A4, &A2 }, { 90.0, 110880.0, 30270240.0, 2075673600.0 }, 17643225600.0, n);
	} else {
		const double maxnorm = 5.371920351148152;
		frexp(l1 / maxnorm, &squarings);
		if (squarings < 0) squarings = 0;
		for (double& x : A) x = ldexp(x, -squarings);
		const Mat A2 = mul(A, A, n), A4 = mul(A2, A2, n), A6 = mul(A4, A2, n);
		Mat tmp = mul(A6, poly({ &A6, &A4, &A2 }, { 1.0, 16380.0, 40840800.0 }, 0.0, n), n);
		const Mat w = poly({ &A6, &A4, &A2 }, { 33522128640.0, 10559470521600.0, 1187353796428800.0 }, 32382376266240000.0, n);
		for (int e = 0; e < n * n; e++) tmp[e] += w[e];
		U = mul(A, tmp, n);
		V = mul(A6, poly({ &A6, &A4, &A2 }, { 182.0, 960960.0, 1323241920.0 }, 0.0, n), n);
		const Mat w2 = poly({ &A6, &A4, &A2 }, { 670442572800.0, 129060195264000.0, 7771770303897600.0 }, 64764752532480000.0, n);
		for (int e = 0; e < n * n; e++) V[e] += w2[e];
	}
	Mat numer((size_t)n * n), denom((size_t)n * n);
	for (int e = 0; e < n * n; e++) {
		numer[e] = U[e] + V[e];
		denom[e] = -U[e] + V[e];
	}
	Mat R = solve_lu(denom, numer, n);
	for (int s = 0; s < squarings; s++) R = mul(R, R, n);
	return R;
}

struct PortModel {
	bool peripheral = false;
	int num_transit = 0;
	void configure(bool use_peripheral, int transit)
	{
		peripheral = use_peripheral;
		num_transit = transit;
	}
	bool biphasic = false, metabolite = false;
	double direct_absorption = 0.0, metabolite_conversion = 0.0;
	void configure_single(bool use_biphasic, bool use_metabolite)
	{
		biphasic = use_biphasic;
		metabolite = use_metabolite;
	}
	void set_single(double direct_absorption_rate, double metabolite_conversion_rate)
	{
		direct_absorption = direct_absorption_rate;
		metabolite_conversion = metabolite_conversion_rate;
	}
	bool solve(double absorption, double excretion, double elimination, double kf, double kb, double transit_rate, double bioavailability,
	           const std::vector<double>& tt, const std::vector<double>& td, const std::vector<double>& ot, std::vector<double>& out)
	{
		// ConstructMatrix, PharmacokineticModel.cpp:188-247
		int n = 2;
		if (peripheral) n++;
		const int metabolite_ix = n;
		if (metabolite) n++;
		const int first_transit = n;
		n += num_transit;
		Mat A((size_t)n * n, 0.0);
		auto a = [&](int r, int c) -> double& { return A[(size_t)r * n + c]; };
		a(0, 0) -= excretion;
		a(0, 0) -= absorption;
		if (num_transit > 0) {
			a(first_transit, 0) += absorption;
			if (num_transit > 2) {
				for (int i = 0; i < num_transit - 1; i++) {
					a(first_transit + i, first_transit + i) -= transit_rate;
					a(first_transit + i + 1, first_transit + i) += transit_rate;
				}
			}
			a(first_transit + num_transit - 1, first_transit + num_transit - 1) = -transit_rate;
			a(1, first_transit + num_transit - 1) += transit_rate;
		} else {
			a(1, 0) += absorption;
		}
		if (peripheral) {
			a(1, 1) -= kf;
			a(2, 1) += kf;
			a(1, 2) += kb;
			a(2, 2) -= kb;
		}
		if (biphasic) {
			a(0, 0) -= direct_absorption;
			a(1, 0) += direct_absorption;
		}
		if (metabolite) {
			a(1, 1) -= metabolite_conversion;
			a(metabolite_ix, 1) += metabolite_conversion;
			a(metabolite_ix, metabolite_ix) -= 1.0; // metabolite elimination, fixed (PharmacoLikelihoodSingle.cpp:143)
		}
		a(1, 1) -= elimination;
		// Solve, :111-177
		std::vector<double> y(n, 0.0), yn(n);
		const double simulate_until = ot.back();
		size_t tti = 0, oti = 0;
		double current_t = 0;
		while (tti < tt.size() && current_t < simulate_until) {
			const double target_t = (tti + 1 < tt.size()) ? tt[tti + 1] : simulate_until;
			y[0] += td[tti] * bioavailability;
			while (oti < ot.size() && ot[oti] <= target_t) {
				const double offset_t = ot[oti] - current_t;
				Mat Mt = A;
				for (double& x : Mt) x *= offset_t;
				const Mat E = expm(Mt, n);
				double central = 0.0;
				for (int k = 0; k < n; k++) central += E[(size_t)1 * n + k] * y[k];
				out[oti] = central;
				oti++;
			}
			const double dt = target_t - current_t;
			Mat Mt = A;
			for (double& x : Mt) x *= dt;
			const Mat E = expm(Mt, n);
			for (int i = 0; i < n; i++) {
				double s = 0.0;
				for (int k = 0; k < n; k++) s += E[(size_t)i * n + k] * y[k];
				yn[i] = s;
			}
			for (int i = 0; i < n; i++)
				if (std::isnan(yn[i])) return false;
			current_t = target_t;
			y = yn;
			tti++;
		}
		return true;
	}
};

} // namespace

extern "C" int oracle_pharmaco_evaluate(const oracle_pharmaco_problem* prob, size_t num_chains, const double* values, double* logp, double* conc,
                                        double* patient_ll, int num_threads)
{
	return pharmaco_glue::evaluate<PortModel>(prob, num_chains, values, logp, conc, patient_ll, num_threads);
}
