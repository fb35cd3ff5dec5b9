// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- not part of the product.
//
// Boost-free stand-in for the reference's src/utils/Utils.h, used ONLY when
// compiling the reference's own, unmodified src/odecommon/*.cpp (and vendored
// CVODE 5.3.0 + Eigen) from /root/reference into oracle/_ref/libbcm3ref.so.
// The reference's Utils.h pulls in Boost (property_tree, foreach) and the
// Logger, neither of which the ODE layer needs for its arithmetic.
// The include directory holding this file is passed first on the -I line so
// `#include "Utils.h"` in the reference sources resolves here.
#pragma once

#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <ctype.h>

#include <vector>
#include <list>
#include <set>
#include <queue>
#include <map>
#include <algorithm>
#include <limits>
#include <string>
#include <functional>
#include <thread>
#include <mutex>
#include <memory>

#include "Eigen/Dense"
#include "Typedefs.h"  // the reference's own src/utils/Typedefs.h (Real = double)

#define ASSERT(expr) do {} while (0)
#define LOG(message, ...) do {} while (0)
#define LOGWARNING(message, ...) do {} while (0)
#define LOGERROR(message, ...) do { fprintf(stderr, "[bcm3ref] " message "\n", ## __VA_ARGS__); } while (0)

namespace bcm3 {
// Same definitions as the reference's src/utils/MathFunctions.h:7-13 (that
// header cannot be included: it drags in boost/math/special_functions/log1p.hpp).
inline Real square(Real x) { return x * x; }
inline double log10(double x) { return 0.4342944819032518276511289189166 * log(x); }
inline double fastpow10(double x) { return exp(x * 2.3025850929940459); }
}
