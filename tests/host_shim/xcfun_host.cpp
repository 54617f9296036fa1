// Host build of the PRODUCT's pointwise functional header (csrc/xc_functionals.cuh) so that the
// CPU test-suite can compare it with the oracle without a GPU.  Test scaffolding only.
#include "xc_functionals.cuh"

template <int T, bool E>
static void run(long n, const double* rho, const double* gx, const double* gy, const double* gz,
                const double* w, double* out) {
    for (long i = 0; i < n; ++i) {
        xcfun::PointCoef c = xcfun::evaluate_point<T, E>(rho[i], gx[i], gy[i], gz[i], w[i]);
        out[5 * i + 0] = c.exc; out[5 * i + 1] = c.a; out[5 * i + 2] = c.bx;
        out[5 * i + 3] = c.by;  out[5 * i + 4] = c.bz;
    }
}

extern "C" void xcfun_host_eval(int type, int exact, long n, const double* rho, const double* gx,
                                const double* gy, const double* gz, const double* w, double* out) {
    if (type == 0) { exact ? run<0, true>(n, rho, gx, gy, gz, w, out) : run<0, false>(n, rho, gx, gy, gz, w, out); }
    else if (type == 1) { exact ? run<1, true>(n, rho, gx, gy, gz, w, out) : run<1, false>(n, rho, gx, gy, gz, w, out); }
    else { run<2, false>(n, rho, gx, gy, gz, w, out); }
}
