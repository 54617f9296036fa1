/*
 * xc_oracle.c -- CPU restatement of the reference's XC numerical-integration path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under quantum_compute_dft_b200/ may link,
 * import or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * What it restates (all citations are into /root/reference/src/dft_solver.cu):
 *   thresholds             :12-13   (RHO_EPS 1e-12, MIN_GRAD 1e-20)
 *   functional parameters  :17-49
 *   Slater                 :61-76
 *   B88 (gradient part)    :78-104
 *   VWN-RPA (B3LYP)        :106-138
 *   LYP                    :140-178
 *   VWN5                   :180-205   (compat mode keeps deviation D1)
 *   PW92                   :207-220
 *   PBE exchange           :222-242
 *   PBE correlation        :244-283   (compat mode keeps D2 and D3)
 *   density rho / grad rho :294-307, :346-380
 *   exc / B conventions    :309-344 (LDA) :382-432 (GGA) :434-513 (B3LYP)
 *   V = B^T Phi            :580, :616, :663 (cublasDgemm N,T in column-major)
 *   B3LYP M + M^T          :515-527, :665-667
 *
 * mode 0 ("compat") reproduces the reference bug for bug; mode 1 ("exact")
 * repairs D1-D3 so that every potential is the derivative of its own energy
 * (= libxc / PySCF numint conventions).  B3LYP is identical in both modes.
 *
 * Pinning status: the reference ships no golden vectors for this path
 * (SURVEY.md section 4).  The oracle is pinned by (i) the pointwise
 * known-answer tables and the H2/h2_grid.txt fixture of SURVEY.md 8(c)
 * (tests/test_oracle_kat.py) and (ii) outputs of the reference CUDA source
 * compiled unmodified for sm_100a (oracle/Makefile -> oracle/_ref/) and run on
 * a B200, committed under tests/golden/ref_*.npz by tools/make_reference_golden.py.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_RHO_FLOOR 1e-12   /* dft_solver.cu:12 */
#define ORACLE_SIGMA_FLOOR 1e-20 /* dft_solver.cu:13 */

static const double kPi = 3.14159265358979323846;

typedef struct { double eps, vrho, vsigma; } xc_out;

/* ---------------------------------------------------------------- LDA pieces */

/* dft_solver.cu:61-76: eps_x = -(3/4)(3/pi)^(1/3) rho^(1/3), v_x = 4/3 eps_x */
static xc_out slater_x(double rho)
{
    xc_out o = {0.0, 0.0, 0.0};
    if (rho < ORACLE_RHO_FLOOR) return o;
    o.eps = -0.7385587663820224 * pow(rho, 1.0 / 3.0);
    o.vrho = (4.0 / 3.0) * o.eps;
    return o;
}

/* Generic VWN interpolation formula with parameters (A,b,c,x0); returns eps and
 * d eps / dx at x = sqrt(rs).  full_derivative = 0 reproduces dft_solver.cu:192-193
 * (the two atan terms are not differentiated: deviation D1); = 1 is the complete
 * derivative as coded for the RPA parameter set at :129-135. */
static void vwn_form(double x, double A, double b, double c, double x0,
                     int full_derivative, double *eps, double *deps_dx)
{
    double X = x * x + b * x + c;
    double Q = sqrt(4.0 * c - b * b);
    double X0 = x0 * x0 + b * x0 + c;
    double at = atan(Q / (2.0 * x + b));
    double pref = b * x0 / X0;
    double e = log(x * x / X) + (2.0 * b / Q) * at
             - pref * (log((x - x0) * (x - x0) / X) + (2.0 * (2.0 * x0 + b) / Q) * at);
    double d = 2.0 / x - (2.0 * x + b) / X
             - pref * (2.0 / (x - x0) - (2.0 * x + b) / X);
    if (full_derivative)
        d += -b / X + pref * (2.0 * x0 + b) / X;
    *eps = A * e;
    *deps_dx = A * d;
}

/* dft_solver.cu:196-205 (VWN5 paramagnetic set at :22) */
static xc_out vwn5_c(double rho, int mode)
{
    xc_out o = {0.0, 0.0, 0.0};
    if (rho < ORACLE_RHO_FLOOR) return o;
    double rs = pow(3.0 / (4.0 * kPi * rho), 1.0 / 3.0);
    double x = sqrt(rs);
    double e, de;
    vwn_form(x, 0.0310907, 3.72744, 12.9352, -0.10498, mode == 1, &e, &de);
    o.eps = e;
    o.vrho = e - (rs / 3.0) * (de / (2.0 * x));
    return o;
}

/* dft_solver.cu:106-138 (VWN-RPA set at :38-41), derivative complete */
static xc_out vwn_rpa_c(double rho)
{
    xc_out o = {0.0, 0.0, 0.0};
    if (rho < ORACLE_RHO_FLOOR) return o;
    double rs = pow(3.0 / (4.0 * kPi * rho), 1.0 / 3.0);
    double x = sqrt(rs);
    double e, de;
    vwn_form(x, 0.0310907, 13.0720, 42.7198, -0.409286, 1, &e, &de);
    o.eps = e;
    o.vrho = e - (rs / 3.0) * (de / (2.0 * x));
    return o;
}

/* dft_solver.cu:207-220 */
static xc_out pw92_c(double rho)
{
    xc_out o = {0.0, 0.0, 0.0};
    if (rho < ORACLE_RHO_FLOOR) return o;
    const double A = 0.03109069086965489503, a1 = 0.21370;
    const double b1 = 7.5957, b2 = 3.5876, b3 = 1.6382, b4 = 0.49294;
    double rs = pow(3.0 / (4.0 * kPi * rho), 1.0 / 3.0);
    double sr = sqrt(rs);
    double q = 2.0 * A * (b1 * sr + b2 * rs + b3 * rs * sr + b4 * rs * rs);
    double dq = 2.0 * A * (0.5 * b1 / sr + b2 + 1.5 * b3 * sr + 2.0 * b4 * rs);
    double lg = log(1.0 + 1.0 / q);
    double f = -2.0 * A * (1.0 + a1 * rs);
    double de = -2.0 * A * a1 * lg + f * (1.0 / (1.0 + 1.0 / q)) * (-1.0 / (q * q)) * dq;
    o.eps = f * lg;
    o.vrho = o.eps - (rs / 3.0) * de;
    return o;
}

/* ---------------------------------------------------------------- GGA pieces */

/* dft_solver.cu:222-242 */
static xc_out pbe_x(double rho, double sigma)
{
    xc_out o = {0.0, 0.0, 0.0};
    if (rho < ORACLE_RHO_FLOOR) return o;
    const double Cx = -0.7385587663820224, kappa = 0.804, mu = 0.2195149727645171;
    double r13 = pow(rho, 1.0 / 3.0);
    double r43 = rho * r13;
    double kF = pow(3.0 * kPi * kPi * rho, 1.0 / 3.0);
    double den = 4.0 * kF * kF * rho * rho;
    double s2 = 0.0;
    if (sigma > ORACLE_SIGMA_FLOOR && den > 1e-50) s2 = sigma / den;
    if (s2 > 1e12) s2 = 1e12;
    double u = 1.0 + mu * s2 / kappa;
    double F = 1.0 + kappa * (1.0 - 1.0 / u);
    double dF = mu / (u * u);
    o.eps = Cx * r13 * F;
    o.vsigma = (Cx * r43) * dF * (1.0 / den);
    o.vrho = (4.0 / 3.0) * o.eps - (8.0 / 3.0) * (Cx * r43) * s2 * dF / rho;
    return o;
}

/* dft_solver.cu:244-283.  compat: beta = 0.066725 (D3) and the sign of dx/drho as
 * coded at :277 (D2).  exact: beta = 0.06672455060314922 and dx/drho = -(v-e)/(rho*gamma). */
static xc_out pbe_c(double rho, double sigma, int mode)
{
    xc_out o = {0.0, 0.0, 0.0};
    if (rho < ORACLE_RHO_FLOOR) return o;
    xc_out l = pw92_c(rho);
    const double beta = (mode == 1) ? 0.06672455060314922 : 0.066725;
    const double gamma = 0.03109069086965489503;
    double kF = pow(3.0 * kPi * kPi * rho, 1.0 / 3.0);
    double den = 16.0 * kF * rho * rho;
    double t2 = 0.0;
    if (sigma > ORACLE_SIGMA_FLOOR && den > 1e-50) t2 = (sigma * kPi) / den;
    if (t2 > 1.0e20) t2 = 1.0e20;
    double x = -l.eps / gamma;
    double em1 = expm1(x);
    double A = (fabs(em1) < 1e-20) ? 1.0e20 : (beta / gamma) / em1;
    double At2 = A * t2;
    double num = 1.0 + At2;
    double dnm = 1.0 + At2 + At2 * At2;
    double Qf = num / dnm;
    double arg = 1.0 + (beta / gamma) * t2 * Qf;
    double H = gamma * log(arg);
    double dQ = (dnm - num * (1.0 + 2.0 * At2)) / (dnm * dnm);
    double pre = gamma / arg * (beta / gamma);
    double dH_dt2 = pre * (Qf + At2 * dQ);
    double dH_dA = pre * t2 * t2 * dQ;
    double dt2_dsig = (den > 1e-50) ? kPi / den : 0.0;
    double dx_drho = (l.vrho - l.eps) / (rho * gamma);
    if (mode == 1) dx_drho = -dx_drho;
    double dA_dx = -A * exp(x) / em1;
    double dt2_drho = t2 * (-7.0 / 3.0) / rho;
    o.eps = l.eps + H;
    o.vsigma = rho * dH_dt2 * dt2_dsig;
    o.vrho = l.vrho + H + rho * (dH_dA * dA_dx * dx_drho + dH_dt2 * dt2_drho);
    return o;
}

/* dft_solver.cu:78-104: gradient-correction part of B88 for ONE spin channel
 * (called with rho/2, sigma/4 at :458-468).  eps is per spin-particle. */
static xc_out b88_dx(double rho_s, double sigma_s)
{
    xc_out o = {0.0, 0.0, 0.0};
    if (rho_s < ORACLE_RHO_FLOOR) return o;
    if (sigma_s < ORACLE_SIGMA_FLOOR) return o;
    const double beta = 0.0042;
    double r13 = pow(rho_s, 1.0 / 3.0);
    double r43 = rho_s * r13;
    double g = sqrt(sigma_s);
    double x = g / r43;
    double as = asinh(x);
    double dn = 1.0 + 6.0 * beta * x * as;
    double term = beta * x * x / dn;
    double ddn = 6.0 * beta * (as + x / sqrt(1.0 + x * x));
    double dF = beta * (2.0 * x * dn - x * x * ddn) / (dn * dn);
    double dE_dx = -r43 * dF;
    o.eps = -term * r13;
    o.vsigma = dE_dx / (2.0 * r43 * g);
    o.vrho = (4.0 / 3.0) * (-(r43 * term) / rho_s) - (4.0 / 3.0) * dE_dx * (x / rho_s);
    return o;
}

/* dft_solver.cu:140-178 (closed shell; floor 1e-14 at :144) */
static xc_out lyp_c(double rho, double sigma)
{
    xc_out o = {0.0, 0.0, 0.0};
    if (rho < 1e-14) return o;
    const double a = 0.04918, b = 0.132, c = 0.2533, d = 0.349;
    const double CF = 2.87123400018819108;
    double r13 = pow(rho, 1.0 / 3.0);
    double rm13 = 1.0 / r13;
    double rm53 = rm13 * rm13 * rm13 * rm13 * rm13;
    double ex = exp(-c * rm13);
    double dn = 1.0 + d * rm13;
    double idn = 1.0 / dn;
    double G = ex * idn;
    double td = d * rm13 * idn;
    double delta = c * rm13 + td;
    double H = -a * rho * idn - a * b * CF * rho * G
             + (a * b / 72.0) * sigma * rm53 * G * (3.0 + 7.0 * delta);
    double d_rm13 = -(1.0 / 3.0) * rm13 / rho;
    double d_dn = d * d_rm13;
    double d_G = G * delta / (3.0 * rho);
    double d_td = d * (d_rm13 * idn - rm13 * idn * idn * d_dn);
    double d_delta = c * d_rm13 + d_td;
    double d_H1 = -a * (dn - rho * d_dn) * (idn * idn);
    double d_H2a = -a * b * CF * (G + rho * d_G);
    double br = 3.0 + 7.0 * delta;
    double tder = (-5.0 / (3.0 * rho)) * br + (delta / (3.0 * rho)) * br + 7.0 * d_delta;
    double d_H2b = (a * b / 72.0) * sigma * (rm53 * G) * tder;
    o.eps = H / rho;
    o.vrho = d_H1 + d_H2a + d_H2b;
    o.vsigma = (a * b / 72.0) * rm53 * G * br;
    return o;
}

/* ------------------------------------------------ routine-level combinations */

/* type: 0 LDA (Slater+VWN5), 1 GGA (PBE x + PBE c), 2 B3LYP local part
 * (0.80 Slater + 0.72 dB88 + 0.19 VWN-RPA + 0.81 LYP; :33-36, :476-479).
 * Output exactly what the device routines return, i.e. BEFORE the kernel-level
 * row gate and BEFORE the w, 1/2, 2x, 4x factors of the B rows. */
static xc_out functional_point(int type, int mode, double rho, double sigma)
{
    xc_out o = {0.0, 0.0, 0.0};
    if (type == 0) {
        xc_out x = slater_x(rho), c = vwn5_c(rho, mode);
        o.eps = x.eps + c.eps;
        o.vrho = x.vrho + c.vrho;
    } else if (type == 1) {
        xc_out x = pbe_x(rho, sigma), c = pbe_c(rho, sigma, mode);
        o.eps = x.eps + c.eps;
        o.vrho = x.vrho + c.vrho;
        o.vsigma = x.vsigma + c.vsigma;
    } else {
        xc_out s = slater_x(rho);
        xc_out bx = b88_dx(0.5 * rho, 0.25 * sigma);
        xc_out v = vwn_rpa_c(rho);
        xc_out l = lyp_c(rho, sigma);
        o.eps = 0.80 * s.eps + 0.72 * bx.eps + 0.19 * v.eps + 0.81 * l.eps;
        o.vrho = 0.80 * s.vrho + 0.72 * bx.vrho + 0.19 * v.vrho + 0.81 * l.vrho;
        o.vsigma = 0.72 * (0.5 * bx.vsigma) + 0.81 * l.vsigma; /* :468, :494-495 */
    }
    return o;
}

/* Batched pointwise evaluation (KAT tables, finite-difference tests, and the
 * numint-shaped CPU baseline in oracle/numint_port.py).
 * gate != 0 applies the kernel-level row gate rho < 1e-12 -> everything 0
 * (:318-324, :394-400, :447-453).  Outputs: exc = rho*eps, vrho, vsigma. */
void oracle_functional_points(int type, int mode, int gate, long n,
                              const double *rho, const double *sigma,
                              double *exc, double *vrho, double *vsigma)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        double r = rho[i], s = sigma ? sigma[i] : 0.0;
        xc_out o = {0.0, 0.0, 0.0};
        if (!(gate && r < ORACLE_RHO_FLOOR)) o = functional_point(type, mode, r, s);
        exc[i] = r * o.eps;
        if (vrho) vrho[i] = o.vrho;
        if (vsigma) vsigma[i] = o.vsigma;
    }
}

/* ------------------------------------------------------- densities on the grid */

/* rho, grad rho, sigma exactly as the reference's one-thread-per-point loops
 * (:294-307 and :346-380): plain double loop over (u,v), D used as given (no
 * symmetrisation).  grad may be NULL (LDA).  ao_grad is (3,ngrid,nao) planar. */
void oracle_density(long ngrid, int nao, const double *dm, const double *ao,
                    const double *ao_grad, double *rho, double *grad /* ngrid*3 */,
                    double *sigma)
{
    const double *gx = ao_grad, *gy = ao_grad ? ao_grad + (size_t)ngrid * nao : NULL,
                 *gz = ao_grad ? ao_grad + 2 * (size_t)ngrid * nao : NULL;
#pragma omp parallel for schedule(static)
    for (long g = 0; g < ngrid; ++g) {
        const double *p = ao + (size_t)g * nao;
        double r = 0.0, dx = 0.0, dy = 0.0, dz = 0.0;
        if (!ao_grad) {
            for (int u = 0; u < nao; ++u) {
                const double *drow = dm + (size_t)u * nao;
                double pu = p[u];
                for (int v = 0; v < nao; ++v) r += drow[v] * pu * p[v];
            }
            rho[g] = r;
            continue;
        }
        const double *px = gx + (size_t)g * nao, *py = gy + (size_t)g * nao,
                     *pz = gz + (size_t)g * nao;
        for (int u = 0; u < nao; ++u) {
            const double *drow = dm + (size_t)u * nao;
            double pu = p[u], xu = px[u], yu = py[u], zu = pz[u];
            for (int v = 0; v < nao; ++v) {
                double d = drow[v];
                r += d * p[v] * pu;
                dx += d * (xu * p[v] + pu * px[v]);
                dy += d * (yu * p[v] + pu * py[v]);
                dz += d * (zu * p[v] + pu * pz[v]);
            }
        }
        rho[g] = r;
        sigma[g] = dx * dx + dy * dy + dz * dz;
        grad[3 * g + 0] = dx;
        grad[3 * g + 1] = dy;
        grad[3 * g + 2] = dz;
    }
}

/* ------------------------------------------------------------ the whole path */

/* Restatement of {LDA,GGA,B3LYP}Solver::compute_xc (:559-584, :588-621, :625-672).
 * vxc receives the reference's RAW output convention (row-major nao x nao):
 *   LDA   : B^T Phi (symmetric)
 *   GGA   : B^T Phi, unsymmetrised, B = w(vrho Phi + 4 vsigma grad rho . grad Phi)
 *   B3LYP : M + M^T, M = B^T Phi, B = w(vrho/2 Phi + 2 vsigma grad rho . grad Phi)
 * Returns E_xc = sum_g w_g rho_g eps_g.  The consumer applies 1/2 (V + V^T)
 * (dft.py:212); parity is defined on that symmetrised matrix.
 * Optional outputs (may be NULL): rho_out[ngrid], sigma_out[ngrid]. */
double oracle_compute_xc(int type, int mode, long ngrid, int nao, const double *dm,
                         const double *ao, const double *ao_grad, const double *w,
                         double *vxc, double *rho_out, double *sigma_out)
{
    size_t n2 = (size_t)nao * nao;
    double *rho = (double *)malloc(sizeof(double) * ngrid);
    double *sigma = (double *)calloc(ngrid, sizeof(double));
    double *grad = (double *)calloc((size_t)ngrid * 3, sizeof(double));
    oracle_density(ngrid, nao, dm, ao, type == 0 ? NULL : ao_grad, rho, grad, sigma);
    if (rho_out) memcpy(rho_out, rho, sizeof(double) * ngrid);
    if (sigma_out) memcpy(sigma_out, sigma, sizeof(double) * ngrid);

    const double *gx = ao_grad, *gy = ao_grad ? ao_grad + (size_t)ngrid * nao : NULL,
                 *gz = ao_grad ? ao_grad + 2 * (size_t)ngrid * nao : NULL;
    double exc_total = 0.0;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    double *vpart = (double *)calloc(n2 * nthreads, sizeof(double));
    double *epart = (double *)calloc(nthreads, sizeof(double));
#pragma omp parallel
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double *vloc = vpart + n2 * tid;
        double *brow = (double *)malloc(sizeof(double) * nao);
        double e = 0.0;
#pragma omp for schedule(static)
        for (long g = 0; g < ngrid; ++g) {
            double r = rho[g];
            if (r < ORACLE_RHO_FLOOR) continue; /* exc = 0, B row = 0 */
            xc_out o = functional_point(type, mode, r, sigma[g]);
            e += w[g] * (r * o.eps);
            const double *p = ao + (size_t)g * nao;
            if (type == 0) {
                double f = w[g] * o.vrho;
                for (int i = 0; i < nao; ++i) brow[i] = f * p[i];
            } else {
                double cr = (type == 1) ? o.vrho : 0.5 * o.vrho;
                double cs = (type == 1) ? 4.0 * o.vsigma : 2.0 * o.vsigma;
                double ax = grad[3 * g], ay = grad[3 * g + 1], az = grad[3 * g + 2];
                const double *px = gx + (size_t)g * nao, *py = gy + (size_t)g * nao,
                             *pz = gz + (size_t)g * nao;
                for (int i = 0; i < nao; ++i) {
                    double dot = ax * px[i] + ay * py[i] + az * pz[i];
                    brow[i] = w[g] * (cr * p[i] + cs * dot);
                }
            }
            /* V[j][i] += B[g][j] * Phi[g][i]  (row-major view of the Dgemm at :580) */
            for (int j = 0; j < nao; ++j) {
                double bj = brow[j];
                double *vr = vloc + (size_t)j * nao;
                for (int i = 0; i < nao; ++i) vr[i] += bj * p[i];
            }
        }
        epart[tid] = e;
        free(brow);
    }
    memset(vxc, 0, sizeof(double) * n2);
    for (int t = 0; t < nthreads; ++t) {
        exc_total += epart[t];
        for (size_t k = 0; k < n2; ++k) vxc[k] += vpart[n2 * t + k];
    }
    if (type == 2) { /* symmetrize_matrix_kernel :515-527 */
        for (int r = 0; r < nao; ++r)
            for (int c = 0; c <= r; ++c) {
                double s = vxc[(size_t)r * nao + c] + vxc[(size_t)c * nao + r];
                vxc[(size_t)r * nao + c] = s;
                vxc[(size_t)c * nao + r] = s;
            }
    }
    free(vpart); free(epart); free(rho); free(sigma); free(grad);
    return exc_total;
}

/* Coulomb J_ij = sum_kl (ij|kl) D_kl: the Dgemv at :550-555 with a row-major
 * (nao^2, nao^2) ERI (cublas column-major, no transpose, symmetric super-matrix
 * assumed by the reference: y = A_cm x where A_cm[r][c] = eri[c*N2 + r]). */
void oracle_coulomb(int nao, const double *eri, const double *dm, double *J)
{
    long N2 = (long)nao * nao;
#pragma omp parallel for schedule(static)
    for (long r = 0; r < N2; ++r) {
        double s = 0.0;
        for (long c = 0; c < N2; ++c) s += eri[(size_t)c * N2 + r] * dm[c];
        J[r] = s;
    }
}

/* ----------------------------------------------------------- AO evaluation */

/* Values (and first derivatives) of contracted real Gaussian s and p shells on
 * grid points: the CPU statement of what PySCF's numint.eval_ao returns for the
 * reference at grid.py:30,38 (deriv=0 -> (ngrid,nao); deriv=1 -> value + planar
 * (3,ngrid,nao) gradient, dft.py:136-142,155,172).  PySCF itself is absent from
 * this environment; conventions are restated in SURVEY.md Appendix B:
 *   s: N e^{-a r^2},  p_j: N r_j e^{-a r^2}; coefficients passed in already
 *   include the primitive normalisation; AO order is the shell order given.
 * A primitive is dropped when a*r^2 > exp_cutoff (PySCF drops tiny exponentials
 * too); both sides of every parity test use the same rule. */
void oracle_eval_ao(long ngrid, const double *coords /* ngrid*3 */, int nshell,
                    const double *shell_xyz /* nshell*3 */, const int *shell_l,
                    const int *shell_ao_off, const int *shell_prim_off,
                    const int *shell_nprim, const double *prim_exp,
                    const double *prim_coef, int nao, double exp_cutoff, int deriv,
                    double *ao, double *ao_grad)
{
    size_t plane = (size_t)ngrid * nao;
#pragma omp parallel for schedule(static)
    for (long g = 0; g < ngrid; ++g) {
        double x = coords[3 * g], y = coords[3 * g + 1], z = coords[3 * g + 2];
        for (int s = 0; s < nshell; ++s) {
            double d[3] = {x - shell_xyz[3 * s], y - shell_xyz[3 * s + 1], z - shell_xyz[3 * s + 2]};
            double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            double e0 = 0.0, e1 = 0.0; /* sum c e^{-a r2},  sum c (-2a) e^{-a r2} */
            for (int k = 0; k < shell_nprim[s]; ++k) {
                double a = prim_exp[shell_prim_off[s] + k];
                if (a * r2 > exp_cutoff) continue;
                double t = prim_coef[shell_prim_off[s] + k] * exp(-a * r2);
                e0 += t;
                e1 += -2.0 * a * t;
            }
            size_t o = (size_t)g * nao + shell_ao_off[s];
            if (shell_l[s] == 0) {
                ao[o] = e0;
                if (deriv)
                    for (int c = 0; c < 3; ++c) ao_grad[c * plane + o] = e1 * d[c];
            } else {
                for (int j = 0; j < 3; ++j) {
                    ao[o + j] = d[j] * e0;
                    if (deriv)
                        for (int c = 0; c < 3; ++c)
                            ao_grad[c * plane + o + j] = d[j] * d[c] * e1 + (j == c ? e0 : 0.0);
                }
            }
        }
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
