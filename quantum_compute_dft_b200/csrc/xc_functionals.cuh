// xc_functionals.cuh -- pointwise exchange-correlation functionals (FP64, device).
//
// Subsystem (c) of BASELINE.json's north_star.  Evaluated ONCE per grid point inside the
// epilogue of the density kernel (the reference evaluates every functional twice per point:
// lda/gga/b3lyp_fused_kernel with compute_B=false and then true, dft_solver.cu:569/578,
// 605/614, 642/656).
//
// Reference routines restated here (math only; file:line into /root/reference/src/dft_solver.cu):
//   Slater :61-76   B88 gradient part :78-104   VWN-RPA :106-138   LYP :140-178
//   VWN5 :180-205   PW92 :207-220   PBE-x :222-242   PBE-c :244-283
//   thresholds RHO_EPS / MIN_GRAD :12-13, mixing coefficients :33-36
//
// template<bool EXACT>: false reproduces the reference's potentials bug for bug
// (SURVEY.md deviations D1: VWN5 derivative without the atan terms, D2: sign of dx/drho in
// PBE-c, D3: beta = 0.066725); true gives potentials that are the exact derivatives of the
// energies (libxc / PySCF numint).  Energies differ only through D3.
#pragma once
#include <math.h>

// The header also compiles as plain C++ (tests/test_functionals_host.py builds it with g++ to
// check the product's formulas against the oracle on the CPU, before any GPU time is spent).
#if defined(__CUDACC__)
#define XCFUN_HD __host__ __device__ __forceinline__
#else
#define XCFUN_HD inline
#endif

namespace xcfun {

constexpr double kRhoFloor = 1e-12;    // dft_solver.cu:12
constexpr double kSigmaFloor = 1e-20;  // dft_solver.cu:13
constexpr double kPi = 3.14159265358979323846;
constexpr double kCx = 0.7385587663820224;  // (3/4)(3/pi)^(1/3)

struct Point {
    double eps;     // energy per particle
    double vrho;    // d(rho eps)/d rho      (as the reference device routines return it)
    double vsigma;  // d(rho eps)/d sigma
};

// rs = (3/(4 pi rho))^(1/3) from rho^(1/3)
XCFUN_HD double wigner_seitz(double rho13) {
    return 0.62035049089940001667 / rho13;  // (3/(4 pi))^(1/3)
}

// VWN interpolation: eps(x) and d eps/dx at x = sqrt(rs).  FULL=false drops the derivative of
// the two atan terms exactly as dft_solver.cu:192-193 does.
template <bool FULL>
XCFUN_HD void vwn_interp(double x, double A, double b, double c, double x0,
                                           double& eps, double& deps_dx) {
    const double X = fma(x, x + b, c);
    const double Q = sqrt(4.0 * c - b * b);
    const double X0 = fma(x0, x0 + b, c);
    const double tx = 2.0 * x + b;
    const double at = atan(Q / tx);
    const double pref = b * x0 / X0;
    const double xm = x - x0;
    eps = A * (log(x * x / X) + (2.0 * b / Q) * at
               - pref * (log(xm * xm / X) + (2.0 * (2.0 * x0 + b) / Q) * at));
    double d = 2.0 / x - tx / X - pref * (2.0 / xm - tx / X);
    if (FULL) d += (pref * (2.0 * x0 + b) - b) / X;
    deps_dx = A * d;
}

template <bool EXACT>
XCFUN_HD Point lda_slater_vwn5(double rho) {
    const double r13 = cbrt(rho);
    const double ex = -kCx * r13;
    const double rs = wigner_seitz(r13);
    const double x = sqrt(rs);
    double ec, dec;
    vwn_interp<EXACT>(x, 0.0310907, 3.72744, 12.9352, -0.10498, ec, dec);
    Point p;
    p.eps = ex + ec;
    p.vrho = (4.0 / 3.0) * ex + (ec - (x / 6.0) * dec);  // (rs/3) dec/(2x) = x dec / 6
    p.vsigma = 0.0;
    return p;
}

// PW92 (modified constant A), closed shell: eps and v = eps - (rs/3) d eps/d rs
XCFUN_HD void pw92(double rs, double& ec, double& vc) {
    const double A = 0.03109069086965489503, a1 = 0.21370;
    const double b1 = 7.5957, b2 = 3.5876, b3 = 1.6382, b4 = 0.49294;
    const double sr = sqrt(rs);
    const double q = 2.0 * A * (sr * (b1 + b3 * rs) + rs * (b2 + b4 * rs));
    const double dq = 2.0 * A * (0.5 * b1 / sr + b2 + 1.5 * b3 * sr + 2.0 * b4 * rs);
    const double lg = log(1.0 + 1.0 / q);
    const double f = -2.0 * A * (1.0 + a1 * rs);
    // d/drs [f ln(1+1/q)] = f' ln(1+1/q) - f q' / (q (q+1))
    const double de = -2.0 * A * a1 * lg - f * dq / (q * (q + 1.0));
    ec = f * lg;
    vc = ec - (rs / 3.0) * de;
}

template <bool EXACT>
XCFUN_HD Point gga_pbe(double rho, double sigma) {
    const double r13 = cbrt(rho);
    const double r43 = rho * r13;
    const double kF = 3.09366772628013593097 * r13;  // (3 pi^2)^(1/3) rho^(1/3)
    const double rho2 = rho * rho;
    Point p;
    // ---- exchange (:222-242)
    {
        const double kappa = 0.804, mu = 0.2195149727645171;
        const double den = 4.0 * kF * kF * rho2;
        double s2 = (sigma > kSigmaFloor && den > 1e-50) ? sigma / den : 0.0;
        s2 = fmin(s2, 1e12);
        const double u = 1.0 + mu * s2 / kappa;
        const double F = 1.0 + kappa * (1.0 - 1.0 / u);
        const double dF = mu / (u * u);
        const double ex = -kCx * r13 * F;
        p.eps = ex;
        p.vsigma = (-kCx * r43) * dF / den;
        p.vrho = (4.0 / 3.0) * ex - (8.0 / 3.0) * (-kCx * r43) * s2 * dF / rho;
    }
    // ---- correlation (:244-283)
    {
        const double beta = EXACT ? 0.06672455060314922 : 0.066725;
        const double gamma = 0.03109069086965489503;
        const double bg = beta / gamma;
        double el, vl;
        pw92(wigner_seitz(r13), el, vl);
        const double den = 16.0 * kF * rho2;
        const bool ok = den > 1e-50;
        double t2 = (sigma > kSigmaFloor && ok) ? (sigma * kPi) / den : 0.0;
        t2 = fmin(t2, 1.0e20);
        const double x = -el / gamma;
        const double em1 = expm1(x);
        const double A = (fabs(em1) < 1e-20) ? 1.0e20 : bg / em1;
        const double At2 = A * t2;
        const double num = 1.0 + At2;
        const double dnm = 1.0 + At2 + At2 * At2;
        const double Qf = num / dnm;
        const double arg = 1.0 + bg * t2 * Qf;
        const double H = gamma * log(arg);
        const double dQ = (dnm - num * (1.0 + 2.0 * At2)) / (dnm * dnm);
        const double pre = beta / arg;
        const double dH_dt2 = pre * (Qf + At2 * dQ);
        const double dH_dA = pre * t2 * t2 * dQ;
        const double dt2_dsig = ok ? kPi / den : 0.0;
        double dx_drho = (vl - el) / (rho * gamma);  // sign as coded at :277 (deviation D2)
        if (EXACT) dx_drho = -dx_drho;
        const double dA_dx = -A * (em1 + 1.0) / em1;  // exp(x) = expm1(x) + 1
        const double dt2_drho = t2 * (-7.0 / 3.0) / rho;
        p.eps += el + H;
        p.vsigma += rho * dH_dt2 * dt2_dsig;
        p.vrho += vl + H + rho * (dH_dA * dA_dx * dx_drho + dH_dt2 * dt2_drho);
    }
    return p;
}

// B3LYP local part: 0.80 Slater + 0.72 dB88 + 0.19 VWN-RPA + 0.81 LYP (:434-513).  The 0.20 exact
// exchange stays in the driver (dft.py:197,217-221,234).
XCFUN_HD Point hyb_b3lyp(double rho, double sigma) {
    const double r13 = cbrt(rho);
    Point p;
    // Slater (:69-76)
    const double exs = -kCx * r13;
    double eps = 0.80 * exs;
    double vrho = 0.80 * (4.0 / 3.0) * exs;
    double vsig = 0.0;
    // B88 gradient correction for one spin channel (:78-104), rho_s = rho/2, sigma_s = sigma/4
    {
        const double rs_ = 0.5 * rho, ss_ = 0.25 * sigma;
        if (rs_ >= kRhoFloor && ss_ >= kSigmaFloor) {
            const double beta = 0.0042;
            const double q13 = cbrt(rs_);
            const double q43 = rs_ * q13;
            const double g = sqrt(ss_);
            const double x = g / q43;
            const double as = asinh(x);
            const double dn = 1.0 + 6.0 * beta * x * as;
            const double term = beta * x * x / dn;
            const double ddn = 6.0 * beta * (as + x / sqrt(1.0 + x * x));
            const double dF = beta * (2.0 * x * dn - x * x * ddn) / (dn * dn);
            const double dE_dx = -q43 * dF;
            eps += 0.72 * (-term * q13);
            vsig += 0.72 * 0.5 * (dE_dx / (2.0 * q43 * g));  // :468 halves the spin-channel vsigma
            vrho += 0.72 * ((4.0 / 3.0) * (-(q43 * term) / rs_) - (4.0 / 3.0) * dE_dx * (x / rs_));
        }
    }
    // VWN-RPA (:106-138), complete derivative
    {
        const double rs = wigner_seitz(r13);
        const double x = sqrt(rs);
        double ec, dec;
        vwn_interp<true>(x, 0.0310907, 13.0720, 42.7198, -0.409286, ec, dec);
        eps += 0.19 * ec;
        vrho += 0.19 * (ec - (x / 6.0) * dec);
    }
    // LYP closed shell (:140-178)
    {
        const double a = 0.04918, b = 0.132, c = 0.2533, d = 0.349;
        const double CF = 2.87123400018819108;
        const double rm13 = 1.0 / r13;
        const double rm53 = rm13 * rm13 * rm13 * rm13 * rm13;
        const double ex = exp(-c * rm13);
        const double dn = 1.0 + d * rm13;
        const double idn = 1.0 / dn;
        const double G = ex * idn;
        const double td = d * rm13 * idn;
        const double delta = c * rm13 + td;
        const double br = 3.0 + 7.0 * delta;
        const double k72 = a * b / 72.0;
        const double H = -a * rho * idn - a * b * CF * rho * G + k72 * sigma * rm53 * G * br;
        const double d_rm13 = -(1.0 / 3.0) * rm13 / rho;
        const double d_dn = d * d_rm13;
        const double d_G = G * delta / (3.0 * rho);
        const double d_td = d * (d_rm13 * idn - rm13 * idn * idn * d_dn);
        const double d_delta = c * d_rm13 + d_td;
        const double d_H1 = -a * (dn - rho * d_dn) * (idn * idn);
        const double d_H2a = -a * b * CF * (G + rho * d_G);
        const double tder = ((delta - 5.0) / (3.0 * rho)) * br + 7.0 * d_delta;
        const double d_H2b = k72 * sigma * (rm53 * G) * tder;
        eps += 0.81 * (H / rho);
        vrho += 0.81 * (d_H1 + d_H2a + d_H2b);
        vsig += 0.81 * (k72 * rm53 * G * br);
    }
    p.eps = eps;
    p.vrho = vrho;
    p.vsigma = vsig;
    return p;
}

// Per-point result in the engine's unified convention.  With
//     B[g,:] = a_g Phi[g,:] + b_g . grad Phi[g,:]      and     V_out = B^T Phi + Phi^T B
// the symmetric output equals 1/2 (V_ref + V_ref^T) for every solver type:
//   LDA   (:336-341)  B_ref = w v Phi                       -> a = w v / 2,     b = 0
//   GGA   (:429)      B_ref = w (vrho Phi + 4 vsig g.dPhi)  -> a = w vrho / 2,  b = 2 w vsig g
//   B3LYP (:492,:510) B_ref = w (vrho/2 Phi + 2 vsig g.dPhi), V_ref = M + M^T -> same a, b
// exc = w rho eps is the point's contribution to E_xc.  Row gate rho < 1e-12 -> all zero
// (:318-324, :394-400, :447-453).
struct PointCoef {
    double exc, a, bx, by, bz;
};

template <int XC_TYPE, bool EXACT>
XCFUN_HD PointCoef evaluate_point(double rho, double gx, double gy, double gz, double w) {
    PointCoef o = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (rho < kRhoFloor) return o;
    Point p;
    if (XC_TYPE == 0) {
        p = lda_slater_vwn5<EXACT>(rho);
    } else {
        const double sigma = gx * gx + gy * gy + gz * gz;
        p = (XC_TYPE == 1) ? gga_pbe<EXACT>(rho, sigma) : hyb_b3lyp(rho, sigma);
    }
    o.exc = w * rho * p.eps;
    o.a = 0.5 * w * p.vrho;
    if (XC_TYPE != 0) {
        const double s = 2.0 * w * p.vsigma;
        o.bx = s * gx;
        o.by = s * gy;
        o.bz = s * gz;
    }
    return o;
}

}  // namespace xcfun
