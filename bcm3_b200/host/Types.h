// Host-side mirror of the reference's plugin / sampler interface for the batched-evaluation path.
// Boost- and Eigen-free (neither is available to this build); names follow the reference (src/utils/Typedefs.h).
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <memory>
#include <string>
#include <vector>

namespace bcm3 {

typedef double Real;
typedef std::vector<Real> VectorReal;

// Dense column-major matrix, nvar x C in the batched call: column c = chain c's variable vector (contiguous),
// which is exactly row c of the C ABI's values[C][nvar].
struct MatrixReal {
	size_t rows_ = 0, cols_ = 0;
	std::vector<Real> data;
	MatrixReal() {}
	MatrixReal(size_t r, size_t c, Real v = 0.0) : rows_(r), cols_(c), data(r * c, v) {}
	void resize(size_t r, size_t c) { rows_ = r; cols_ = c; data.assign(r * c, 0.0); }
	size_t rows() const { return rows_; }
	size_t cols() const { return cols_; }
	Real& operator()(size_t i, size_t j) { return data[i + j * rows_]; }
	Real operator()(size_t i, size_t j) const { return data[i + j * rows_]; }
	Real* col(size_t j) { return data.data() + j * rows_; }
	const Real* col(size_t j) const { return data.data() + j * rows_; }
};

const Real kInf = std::numeric_limits<Real>::infinity();
const Real kNaN = std::numeric_limits<Real>::quiet_NaN();

} // namespace bcm3
