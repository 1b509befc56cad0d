// Host harness for gscf_b200/csrc/smallsolve.hpp (the header is compiled for host and device): the bordered DIIS system
// + LU of PulayDiisScf.hh:326-393 and the line search of TruncatedOptDampScf.hh:370-394, callable from the CPU tests.
#include <cmath>
using std::fabs;      // nvcc sees the global overloads; a plain host compiler needs them named
using std::isfinite;
#include "../../gscf_b200/csrc/smallsolve.hpp"

extern "C" int smallsolve_diis(const double* gram, int gram_ld, const int* order, int k, double* coeff) {
  return gscfb::diis_build_and_solve(gram, gram_ld, order, k, coeff) ? 1 : 0;
}
extern "C" double smallsolve_damping(double s, double c, double min_pref) {
  return gscfb::compute_damping_coeff(s, c, min_pref);
}
