// ao_eval.cu -- atomic-orbital values and gradients on the grid (subsystem (a) of the north_star).
//
// Replaces, on the GPU, what the reference obtains on the host from PySCF
// (numint.eval_ao(mol, coords, deriv=0|1), grid.py:30,38) and then uploads (dft.py:155,172).
// Output layouts are exactly what DFT_ComputeXC consumes: ao (ngrid,nao) and the planar
// gradient (3,ngrid,nao).
//
// Design (HBM-bound by nature: 8*ngrid*nao*P bytes written, P = 1 or 4).  Earlier versions (one warp
// per point, lanes over shells, a staged row per point) were bound by shared-memory bank conflicts and
// then by instruction issue (ncu: 2500 warp instructions per point, exp() only 10 % of them).  This
// one minimises instructions per output value:
//   * shells on the same centre with the same exponents (the s and p parts of an STO-3G "sp" shell)
//     form one GROUP whose exponentials are evaluated once -- 1/3 fewer exp() than shell by shell;
//   * all tables are structure-of-arrays in shared memory, staged once per CTA;
//   * a CTA takes blocks of 16 consecutive grid points (two or three CTAs per SM overlap each other's
//     phases);
//   * phase 1, lanes over POINTS: a half-warp evaluates one group for the 16 points -- the group's data
//     is uniform over the half-warp (broadcast loads), neighbouring points agree on which primitives
//     fall beyond the cutoff (AO screening without divergence or compaction; a primitive is dropped
//     when exp*r^2 > cutoff exactly as in the CPU statement; a group whose most diffuse primitive is
//     beyond the cutoff is skipped outright) -- and stores the radial sums
//     (e0, e1) = sum c (1, -2a) exp(-a r^2) of each member shell in a [point][shell] array whose
//     point pitch is odd in 16-byte units (conflict-free both ways);
//   * phase 2, lanes over AOs: a warp takes one point, lane i combines its shell's (e0, e1) with the
//     distance vector and writes the value and the three gradient components STRAIGHT to global memory:
//     consecutive lanes -> consecutive AOs, 256 contiguous bytes per warp store, no staging row.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/dft_b200_ext.h"
#include "engine.h"

namespace xc {
namespace ao {

constexpr int MAXP = 4;       // primitives per group evaluated from registers (more: recomputed per member)
// Block shapes (grid points per block G, warps per CTA NW): 32 points x 16 warps when at least two such
// CTAs fit in an SM's shared memory, 16 points x 8 warps otherwise.

// exp(t) for t in [-700, 0] (here t = -a r^2 >= -cutoff): Cody-Waite reduction t = n ln2 + r, |r| <= ln2/2,
// Taylor polynomial to r^13 (truncation 4e-18), scaling by 2^n through the exponent field.  About 1 ulp;
// half the instructions of the general-purpose exp(), which this kernel's phase 1 is made of.
__device__ __forceinline__ double exp_neg(double t) {
#ifdef DFT_AO_LIBM_EXP
    return exp(t);
#else
    const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer
    const double u = fma(t, 1.4426950408889634074, SHIFT);
    const int n = __double2loint(u);
    const double nf = u - SHIFT;
    double r = fma(nf, -6.93147180369123816490e-01, t);
    r = fma(nf, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821614599e-10;            // 1/13!
    p = fma(p, r, 2.0876756987868098979e-09);        // 1/12!
    p = fma(p, r, 2.5052108385441718775e-08);        // 1/11!
    p = fma(p, r, 2.7557319223985890653e-07);        // 1/10!
    p = fma(p, r, 2.7557319223985892511e-06);        // 1/9!
    p = fma(p, r, 2.4801587301587301566e-05);        // 1/8!
    p = fma(p, r, 1.9841269841269841253e-04);        // 1/7!
    p = fma(p, r, 1.3888888888888889419e-03);        // 1/6!
    p = fma(p, r, 8.3333333333333332177e-03);        // 1/5!
    p = fma(p, r, 4.1666666666666664354e-02);        // 1/4!
    p = fma(p, r, 1.6666666666666665741e-01);        // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
#endif
}

// Device tables, one block of memory (offsets in bytes from the start; all 8-byte aligned)
struct Tables {
    const unsigned char* base;
    size_t bytes;
    int ngroup, nshell, nmember, nexp, ncoef, nao;
    // doubles: gx gy gz gamin [ngroup] | sx sy sz [nshell] | exps [nexp] | coefs [ncoef]
    // ints:    gprim gnprim gmem gnmem [ngroup] | mshell mcoef [nmember] | ao_shell ao_comp [nao]
    size_t o_gx, o_gy, o_gz, o_gamin, o_sx, o_sy, o_sz, o_exp, o_coef;
    size_t o_gprim, o_gnprim, o_gmem, o_gnmem, o_mshell, o_mcoef, o_aoshell, o_aocomp;
};

template <bool DERIV, int G, int NWARPS, bool VEC>
__global__ void __launch_bounds__(NWARPS * 32)
eval_kernel(int ngrid, const double* __restrict__ coords, Tables t, int epitch, double cutoff, double* __restrict__ ao,
            double* __restrict__ gout) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    {   // stage the tables (8-byte words)
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(t.base);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(smem_raw);
        for (size_t i = threadIdx.x; i < t.bytes / 8; i += blockDim.x) dst[i] = src[i];
    }
    const double* gx = reinterpret_cast<const double*>(smem_raw + t.o_gx);
    const double* gy = reinterpret_cast<const double*>(smem_raw + t.o_gy);
    const double* gz = reinterpret_cast<const double*>(smem_raw + t.o_gz);
    const double* gamin = reinterpret_cast<const double*>(smem_raw + t.o_gamin);
    const double* sx = reinterpret_cast<const double*>(smem_raw + t.o_sx);
    const double* sy = reinterpret_cast<const double*>(smem_raw + t.o_sy);
    const double* sz = reinterpret_cast<const double*>(smem_raw + t.o_sz);
    const double* s_exp = reinterpret_cast<const double*>(smem_raw + t.o_exp);
    const double* s_coef = reinterpret_cast<const double*>(smem_raw + t.o_coef);
    const int* gprim = reinterpret_cast<const int*>(smem_raw + t.o_gprim);
    const int* gnprim = reinterpret_cast<const int*>(smem_raw + t.o_gnprim);
    const int* gmem = reinterpret_cast<const int*>(smem_raw + t.o_gmem);
    const int* gnmem = reinterpret_cast<const int*>(smem_raw + t.o_gnmem);
    const int* mshell = reinterpret_cast<const int*>(smem_raw + t.o_mshell);
    const int* mcoef = reinterpret_cast<const int*>(smem_raw + t.o_mcoef);
    const int* ao_meta = reinterpret_cast<const int*>(smem_raw + t.o_aoshell);  // shell | (comp + 1) << 24
    double* s_pts2 = reinterpret_cast<double*>(smem_raw + t.bytes);              // [2][G][3]: this block's points, the next block's
    double2* e01 = reinterpret_cast<double2*>(s_pts2 + 6 * G);                   // [G][epitch]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ng = t.ngroup, nao = t.nao;
    const int pt = tid & (G - 1), slot = tid / G;
    constexpr int SLOTS = NWARPS * 32 / G;
    const size_t plane = (size_t)ngrid * nao;
    const int nblk = (ngrid + G - 1) / G;
    // The coordinates of a block are fetched while the previous block is in phase 1 (double-buffered), so a
    // block costs two CTA barriers and no exposed global-memory latency.
    if ((int)blockIdx.x < nblk) {
        const int p0 = blockIdx.x * G, np = min(G, ngrid - p0);
        if (tid < 3 * np) s_pts2[tid] = __ldg(coords + 3 * (size_t)p0 + tid);
    }
    int buf = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x, buf ^= 1) {
        const int p0 = blk * G;
        const int np = min(G, ngrid - p0);
        const double* s_pts = s_pts2 + buf * 3 * G;
        __syncthreads();  // tables and this block's points staged / previous block's phase 2 done with e01
        {   // next block's points into the other buffer (last read by the previous block's phase 2)
            const int nb = blk + gridDim.x;
            if (nb < nblk) {
                const int q0 = nb * G, nq = min(G, ngrid - q0);
                if (tid < 3 * nq) s_pts2[(buf ^ 1) * 3 * G + tid] = __ldg(coords + 3 * (size_t)q0 + tid);
            }
        }
        // ---- phase 1: radial sums; a half-warp = one group x 16 points
        if (pt < np) {
            const double x = s_pts[3 * pt], y = s_pts[3 * pt + 1], z = s_pts[3 * pt + 2];
            double2* erow = e01 + (size_t)pt * epitch;
            for (int gi = slot; gi < ng; gi += SLOTS) {
                const double dx = x - gx[gi], dy = y - gy[gi], dz = z - gz[gi];
                const double r2 = dx * dx + dy * dy + dz * dz;
                const int m0 = gmem[gi], nm = gnmem[gi];
                // AO screening at the group level: beyond the cutoff of the group's most diffuse primitive every
                // primitive is dropped (a >= amin => a r^2 > cutoff), so its shells are exact zeros here -- most
                // (point, group) pairs of a large molecule.  Neighbouring points and the shells of one centre
                // decide alike, so the warp rarely diverges.
                if (gamin[gi] * r2 > cutoff) {
                    for (int m = 0; m < nm; ++m) erow[mshell[m0 + m]] = make_double2(0.0, 0.0);
                    continue;
                }
                const int q0 = gprim[gi], nq = gnprim[gi];
                if (nq <= MAXP) {
                    double v[MAXP], av[MAXP];
#pragma unroll
                    for (int k = 0; k < MAXP; ++k) {
                        v[k] = 0.0; av[k] = 0.0;
                        if (k < nq) {
                            const double al = s_exp[q0 + k];
                            const double ar2 = al * r2;
                            if (!(ar2 > cutoff)) { v[k] = exp_neg(-ar2); av[k] = al; }
                        }
                    }
                    for (int m = 0; m < nm; ++m) {
                        const int c0 = mcoef[m0 + m];
                        double e0 = 0.0, e1 = 0.0;
#pragma unroll
                        for (int k = 0; k < MAXP; ++k) {
                            if (k < nq && av[k] != 0.0) {
                                const double tv = s_coef[c0 + k] * v[k];
                                e0 += tv;
                                e1 = fma(-2.0 * av[k], tv, e1);
                            }
                        }
                        erow[mshell[m0 + m]] = make_double2(e0, e1);
                    }
                } else {
                    for (int m = 0; m < nm; ++m) {
                        const int c0 = mcoef[m0 + m];
                        double e0 = 0.0, e1 = 0.0;
                        for (int k = 0; k < nq; ++k) {
                            const double al = s_exp[q0 + k];
                            const double ar2 = al * r2;
                            if (ar2 > cutoff) continue;
                            const double tv = s_coef[c0 + k] * exp_neg(-ar2);
                            e0 += tv;
                            e1 = fma(-2.0 * al, tv, e1);
                        }
                        erow[mshell[m0 + m]] = make_double2(e0, e1);
                    }
                }
            }
        }
        __syncthreads();
        // ---- phase 2: a warp per point, lanes over AOs, results straight to global memory (coalesced)
        if (VEC) {
            // 16-byte stores: a lane owns the AO pair (i0, i0 + 1) whose first element sits on a 16-byte boundary of
            // its row -- row (p0 + r) starts at element (p0 + r) nao, so for odd nao the pairs of odd rows start one
            // AO earlier (i0 = -1: only the second element exists).  512 contiguous bytes per warp store, half the
            // store instructions and loop trips of the scalar form.  The y-gradient plane of an odd x odd problem
            // has the opposite parity: that one plane falls back to two 8-byte stores per pair.
            for (int r = warp; r < np; r += NWARPS) {
                const double x = s_pts[3 * r], y = s_pts[3 * r + 1], z = s_pts[3 * r + 2];
                const double2* erow = e01 + (size_t)r * epitch;
                const size_t row0 = (size_t)(p0 + r) * nao;
                const int a = (int)(row0 & 1);
                const bool vy = ((plane & 1) == 0);                     // gy pairs are 16-byte aligned too
                for (int i0 = 2 * lane - a; i0 < nao; i0 += 64) {
                    const bool has0 = i0 >= 0, has1 = i0 + 1 < nao;
                    double v[2], vx[2], vyv[2], vz[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        int i = i0 + h;
                        i = i < 0 ? 0 : (i >= nao ? nao - 1 : i);
                        const int meta = ao_meta[i];
                        const int sh = meta & 0xffffff, comp = (meta >> 24) - 1;
                        const double2 e = erow[sh];
                        const double dx = x - sx[sh], dy = y - sy[sh], dz = z - sz[sh];
                        double dj = dz;
                        dj = comp == 1 ? dy : dj;
                        dj = comp == 0 ? dx : dj;
                        dj = comp < 0 ? 1.0 : dj;
                        v[h] = dj * e.x;
                        if (DERIV) {
                            const double dj1 = dj * e.y;
                            vx[h] = fma(dj1, dx, comp == 0 ? e.x : 0.0);
                            vyv[h] = fma(dj1, dy, comp == 1 ? e.x : 0.0);
                            vz[h] = fma(dj1, dz, comp == 2 ? e.x : 0.0);
                        }
                    }
                    double* d0 = ao + row0 + i0;
                    if (has0 && has1) {
                        *reinterpret_cast<double2*>(d0) = make_double2(v[0], v[1]);
                        if (DERIV) {
                            double* d1 = gout + row0 + i0;
                            *reinterpret_cast<double2*>(d1) = make_double2(vx[0], vx[1]);
                            if (vy) *reinterpret_cast<double2*>(d1 + plane) = make_double2(vyv[0], vyv[1]);
                            else { d1[plane] = vyv[0]; d1[plane + 1] = vyv[1]; }
                            *reinterpret_cast<double2*>(d1 + 2 * plane) = make_double2(vz[0], vz[1]);
                        }
                    } else {
                        const int h = has0 ? 0 : 1;
                        d0[h] = has0 ? v[0] : v[1];
                        if (DERIV) {
                            double* d1 = gout + row0 + i0 + h;
                            d1[0] = has0 ? vx[0] : vx[1]; d1[plane] = has0 ? vyv[0] : vyv[1]; d1[2 * plane] = has0 ? vz[0] : vz[1];
                        }
                    }
                }
            }
        } else {
        for (int r = warp; r < np; r += NWARPS) {
            const double x = s_pts[3 * r], y = s_pts[3 * r + 1], z = s_pts[3 * r + 2];
            const double2* erow = e01 + (size_t)r * epitch;
            double* d0 = ao + (size_t)(p0 + r) * nao + lane;
            double* d1 = DERIV ? gout + (size_t)(p0 + r) * nao + lane : nullptr;
            double* d2 = DERIV ? d1 + plane : nullptr;
            double* d3 = DERIV ? d2 + plane : nullptr;
            // two AOs per lane and trip (i, i + 32), all their shared-memory loads issued before the first use: the
            // chain meta -> (e0, e1), centre -> value is two shared-memory latencies long, and with one AO per trip
            // it was exposed twelve times per point (ncu: short-scoreboard was the top stall of this loop)
            for (int i = lane; i < nao; i += 64) {
                const bool h1 = i + 32 < nao;
                const int meta0 = ao_meta[i], meta1 = ao_meta[h1 ? i + 32 : i];
                const int sh0 = meta0 & 0xffffff, comp0 = (meta0 >> 24) - 1;  // comp: -1 s, 0..2 p_x p_y p_z
                const int sh1 = meta1 & 0xffffff, comp1 = (meta1 >> 24) - 1;
                const double2 e0 = erow[sh0], e1 = erow[sh1];
                const double dx0 = x - sx[sh0], dy0 = y - sy[sh0], dz0 = z - sz[sh0];
                const double dx1 = x - sx[sh1], dy1 = y - sy[sh1], dz1 = z - sz[sh1];
                // (a chain of selects, not nested conditionals: lanes hold different components, and the
                // nested form compiles to divergent branches)
                double dj0 = dz0, dj1 = dz1;
                dj0 = comp0 == 1 ? dy0 : dj0; dj1 = comp1 == 1 ? dy1 : dj1;
                dj0 = comp0 == 0 ? dx0 : dj0; dj1 = comp1 == 0 ? dx1 : dj1;
                dj0 = comp0 < 0 ? 1.0 : dj0;  dj1 = comp1 < 0 ? 1.0 : dj1;
                d0[0] = dj0 * e0.x;
                if (h1) d0[32] = dj1 * e1.x;
                if (DERIV) {
                    // s: e1 d;  p_j: d_j e1 d + e0 delta_j
                    const double t0 = dj0 * e0.y, t1 = dj1 * e1.y;
                    d1[0] = fma(t0, dx0, comp0 == 0 ? e0.x : 0.0);
                    d2[0] = fma(t0, dy0, comp0 == 1 ? e0.x : 0.0);
                    d3[0] = fma(t0, dz0, comp0 == 2 ? e0.x : 0.0);
                    if (h1) {
                        d1[32] = fma(t1, dx1, comp1 == 0 ? e1.x : 0.0);
                        d2[32] = fma(t1, dy1, comp1 == 1 ? e1.x : 0.0);
                        d3[32] = fma(t1, dz1, comp1 == 2 ? e1.x : 0.0);
                    }
                    d1 += 64; d2 += 64; d3 += 64;
                }
                d0 += 64;
            }
        }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Direct kernel (round 2, option ao_shape 1): lanes over AOs from start to finish, no staging of radial sums, no CTA barrier.
// ------------------------------------------------------------------------------------------------
// An experiment with a negative result, kept selectable.  The two-phase kernel above shares the exponentials of a group
// through shared memory, which costs it 63 KB of [point][shell] staging per CTA (16 resident warps per SM), two CTA
// barriers per 16 points and a phase 1 whose half-warps diverge on the cutoff (ncu: stalls on barriers and
// shared-memory latency, 0.76 of the copy bandwidth).  Here a lane owns one AO of one point and evaluates its shell's
// radial sums itself -- the three AOs of a p shell and the s AO of the same sp group repeat the same exponentials,
// twice the exp() work overall -- so the only shared memory is the per-AO tables (18 KB at C5), 48 warps are resident
// per SM, nothing waits on a barrier, and every warp store is still 256 contiguous bytes.  Same arithmetic in the
// same order: results are bit-identical.  Measured (profiles/r2_u12_ao_direct_kernel.txt): C5 4.31 ms against 3.48,
// C4 1.02 against 0.75, C2 0.121 against 0.067 -- the doubled exponentials and the mixed live / dead lanes of a
// 32-AO chunk cost more than the barriers and the occupancy did.
struct DirectTables {
    const unsigned char* base;
    size_t bytes;
    int nao, nprim;
    // doubles: cx cy cz amin [nao] | exps coefs [nprim];  ints: poff npr comp [nao]
    size_t o_cx, o_cy, o_cz, o_amin, o_exp, o_coef, o_poff, o_npr, o_comp;
};

constexpr int DIRECT_PW = 4;          // consecutive points per warp step
constexpr int DIRECT_WARPS = 8;

template <bool DERIV>
__global__ void __launch_bounds__(DIRECT_WARPS * 32)
eval_direct_kernel(int ngrid, const double* __restrict__ coords, DirectTables t, double cutoff, double* __restrict__ ao,
                   double* __restrict__ gout) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(t.base);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(smem_raw);
        for (size_t i = threadIdx.x; i < t.bytes / 8; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();   // (the only one)
    const double* cx = reinterpret_cast<const double*>(smem_raw + t.o_cx);
    const double* cy = reinterpret_cast<const double*>(smem_raw + t.o_cy);
    const double* cz = reinterpret_cast<const double*>(smem_raw + t.o_cz);
    const double* amin = reinterpret_cast<const double*>(smem_raw + t.o_amin);
    const double* s_exp = reinterpret_cast<const double*>(smem_raw + t.o_exp);
    const double* s_coef = reinterpret_cast<const double*>(smem_raw + t.o_coef);
    const int* poff = reinterpret_cast<const int*>(smem_raw + t.o_poff);
    const int* npr = reinterpret_cast<const int*>(smem_raw + t.o_npr);
    const int* compt = reinterpret_cast<const int*>(smem_raw + t.o_comp);

    const int lane = threadIdx.x & 31;
    const int nao = t.nao;
    const size_t plane = (size_t)ngrid * nao;
    const long nstep = ((long)ngrid + DIRECT_PW - 1) / DIRECT_PW;
    const long wid = (long)blockIdx.x * DIRECT_WARPS + (threadIdx.x >> 5), nw = (long)gridDim.x * DIRECT_WARPS;
    // coordinates of a step: lanes 0 .. 3 PW - 1 hold one double each (coalesced), the next step's are in flight
    auto load_pts = [&](long st) {
        const long e = st * (3 * DIRECT_PW) + lane;
        return (st < nstep && lane < 3 * DIRECT_PW && e < 3 * (long)ngrid) ? __ldg(coords + e) : 0.0;
    };
    double pts_next = load_pts(wid);
    for (long st = wid; st < nstep; st += nw) {
        const double pts = pts_next;
        pts_next = load_pts(st + nw);
        const long p0 = st * DIRECT_PW;
#pragma unroll 1
        for (int r = 0; r < DIRECT_PW; ++r) {
            if (p0 + r >= ngrid) break;
            const double x = __shfl_sync(0xffffffffu, pts, 3 * r), y = __shfl_sync(0xffffffffu, pts, 3 * r + 1),
                         z = __shfl_sync(0xffffffffu, pts, 3 * r + 2);
            double* d0 = ao + (size_t)(p0 + r) * nao + lane;
            double* d1 = DERIV ? gout + (size_t)(p0 + r) * nao + lane : nullptr;
            for (int i = lane; i - lane < nao; i += 32, d0 += 32, d1 += 32) {
                const bool valid = i < nao;
                const int ii = valid ? i : nao - 1;
                const double dx = x - cx[ii], dy = y - cy[ii], dz = z - cz[ii];
                const double r2 = dx * dx + dy * dy + dz * dz;
                const int comp = compt[ii];
                // beyond the cutoff of the shell's most diffuse primitive every primitive is dropped: exact zeros
                const bool live = valid && !(amin[ii] * r2 > cutoff);
                double e0 = 0.0, e1 = 0.0;
                if (__any_sync(0xffffffffu, live)) {
                    const int q0 = poff[ii], nq = live ? npr[ii] : 0;
                    int nqmax = nq;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) nqmax = max(nqmax, __shfl_xor_sync(0xffffffffu, nqmax, o));
                    for (int k = 0; k < nqmax; ++k) {
                        if (k < nq) {
                            const double al = s_exp[q0 + k];
                            const double ar2 = al * r2;
                            if (!(ar2 > cutoff)) {
                                const double tv = s_coef[q0 + k] * exp_neg(-ar2);
                                e0 += tv;
                                e1 = fma(-2.0 * al, tv, e1);
                            }
                        }
                    }
                }
                if (!valid) continue;
                double dj = dz;
                dj = comp == 1 ? dy : dj;
                dj = comp == 0 ? dx : dj;
                dj = comp < 0 ? 1.0 : dj;
                *d0 = dj * e0;
                if (DERIV) {
                    const double dj1 = dj * e1;
                    d1[0] = fma(dj1, dx, comp == 0 ? e0 : 0.0);
                    d1[plane] = fma(dj1, dy, comp == 1 ? e0 : 0.0);
                    d1[2 * plane] = fma(dj1, dz, comp == 2 ? e0 : 0.0);
                }
            }
        }
    }
}

}  // namespace ao
}  // namespace xc

extern "C" int DFT_EvalAO(XCSolver* solver, int ngrid, unsigned long long d_coords_ptr, int nshell,
                          const double* shell_xyz, const int* shell_l, const int* shell_ao_off,
                          const int* shell_prim_off, const int* shell_nprim, int nprim_total,
                          const double* prim_exp, const double* prim_coef, int nao, int deriv, double exp_cutoff,
                          unsigned long long d_ao_ptr, unsigned long long d_ao_grad_ptr) {
    using namespace xc::ao;
    if (!solver || !d_coords_ptr || !d_ao_ptr || nshell <= 0 || nao <= 0 || nprim_total <= 0) return 1;
    if (deriv && !d_ao_grad_ptr) return 1;
    if (ngrid <= 0) return 0;
    CublasHandleWrapper* ctx = solver->context();
    DeviceGuard guard(ctx->device);
    ctx->failed = false;
    if (exp_cutoff <= 0.0) exp_cutoff = 60.0;
    if (exp_cutoff > 700.0) exp_cutoff = 700.0;  // exp(-700) ~ 1e-304: nothing beyond contributes (and exp_neg stays in range)

    // ---- group the shells: same centre, same exponents -> one group (exponentials evaluated once)
    struct HGroup { double x, y, z, amin; int prim_off, nprim, mem_off, nmem; };
    std::vector<HGroup> groups;
    std::vector<double> exps, coefs;
    std::vector<int> shell_group(nshell, -1), ao_shell(nao, -1), ao_comp(nao, -1);
    for (int s = 0; s < nshell; ++s) {
        if (shell_l[s] < 0 || shell_l[s] > 1) return 3;  // s and p shells only
        if (shell_nprim[s] <= 0 || shell_prim_off[s] < 0 || shell_prim_off[s] + shell_nprim[s] > nprim_total) return 3;
        if (shell_ao_off[s] < 0 || shell_ao_off[s] + (shell_l[s] ? 3 : 1) > nao) return 3;
        for (int j = 0; j < (shell_l[s] ? 3 : 1); ++j) {
            ao_shell[shell_ao_off[s] + j] = s;
            ao_comp[shell_ao_off[s] + j] = shell_l[s] ? j : -1;
        }
        int found = -1;
        for (int g = 0; g < (int)groups.size() && found < 0; ++g) {
            const HGroup& q = groups[g];
            if (q.x != shell_xyz[3 * s] || q.y != shell_xyz[3 * s + 1] || q.z != shell_xyz[3 * s + 2]) continue;
            if (q.nprim != shell_nprim[s]) continue;
            bool same = true;
            for (int k = 0; k < q.nprim && same; ++k) same = exps[q.prim_off + k] == prim_exp[shell_prim_off[s] + k];
            if (same) found = g;
        }
        if (found < 0) {
            HGroup q;
            q.x = shell_xyz[3 * s]; q.y = shell_xyz[3 * s + 1]; q.z = shell_xyz[3 * s + 2];
            q.prim_off = (int)exps.size(); q.nprim = shell_nprim[s]; q.mem_off = 0; q.nmem = 0;
            q.amin = prim_exp[shell_prim_off[s]];
            for (int k = 0; k < q.nprim; ++k) {
                const double a = prim_exp[shell_prim_off[s] + k];
                exps.push_back(a);
                if (a < q.amin) q.amin = a;
            }
            found = (int)groups.size();
            groups.push_back(q);
        }
        shell_group[s] = found;
        groups[found].nmem++;
    }
    for (int i = 0; i < nao; ++i)
        if (ao_shell[i] < 0) return 3;  // every AO must belong to a shell
    const int ngroup = (int)groups.size();
    if (ngroup > 65535) return 5;
    double* ao_out = reinterpret_cast<double*>(d_ao_ptr);
    double* g_out = reinterpret_cast<double*>(d_ao_grad_ptr);
    const double* coords = reinterpret_cast<const double*>(d_coords_ptr);
    if (ctx->num_sms <= 0) cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, ctx->device);

    if (ctx->ao_shape == 1) {
        // ---- direct kernel (option ao_shape 1; measured slower than the two-phase kernel, see its header): per-AO tables only
        DirectTables t;
        memset(&t, 0, sizeof(t));
        t.nao = nao; t.nprim = nprim_total;
        size_t off = 0;
        auto take = [&](size_t nbytes) { const size_t o = off; off += (nbytes + 7) & ~(size_t)7; return o; };
        t.o_cx = take(8 * (size_t)nao); t.o_cy = take(8 * (size_t)nao); t.o_cz = take(8 * (size_t)nao); t.o_amin = take(8 * (size_t)nao);
        t.o_exp = take(8 * (size_t)nprim_total); t.o_coef = take(8 * (size_t)nprim_total);
        t.o_poff = take(4 * (size_t)nao); t.o_npr = take(4 * (size_t)nao); t.o_comp = take(4 * (size_t)nao);
        const size_t bytes = (off + 15) & ~(size_t)15;
        if (bytes <= 96 * 1024) {   // (larger bases: the two-phase kernel's shell tables are smaller)
            std::vector<unsigned char> h(bytes, 0);
            double* hd = reinterpret_cast<double*>(h.data());
            for (int i = 0; i < nao; ++i) {
                const int sh = ao_shell[i];
                reinterpret_cast<double*>(h.data() + t.o_cx)[i] = shell_xyz[3 * sh];
                reinterpret_cast<double*>(h.data() + t.o_cy)[i] = shell_xyz[3 * sh + 1];
                reinterpret_cast<double*>(h.data() + t.o_cz)[i] = shell_xyz[3 * sh + 2];
                reinterpret_cast<double*>(h.data() + t.o_amin)[i] = groups[shell_group[sh]].amin;
                reinterpret_cast<int*>(h.data() + t.o_poff)[i] = shell_prim_off[sh];
                reinterpret_cast<int*>(h.data() + t.o_npr)[i] = shell_nprim[sh];
                reinterpret_cast<int*>(h.data() + t.o_comp)[i] = ao_comp[i];
            }
            (void)hd;
            memcpy(h.data() + t.o_exp, prim_exp, 8 * (size_t)nprim_total);
            memcpy(h.data() + t.o_coef, prim_coef, 8 * (size_t)nprim_total);
            unsigned char* d = (unsigned char*)ctx->scratch.ensure(bytes, &ctx->failed);
            if (ctx->failed) return 4;
            DFT_CUDA_CHECK(ctx, cudaMemcpyAsync(d, h.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
            DFT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // h is pageable: the copy is done now
            t.base = d; t.bytes = bytes;
            int per_sm = (int)((227 * 1024) / (bytes + 1024));
            per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);      // 8 warps per CTA: up to 48 resident warps per SM
            const long nstep = ((long)ngrid + DIRECT_PW - 1) / DIRECT_PW;
            const long want = (nstep + DIRECT_WARPS - 1) / DIRECT_WARPS;
            const int grid = (int)(want < (long)ctx->num_sms * per_sm ? want : (long)ctx->num_sms * per_sm);
            xc::resolve_times(ctx);
            if (ctx->timing) cudaEventRecord(ctx->ev[0], ctx->stream);
            if (deriv) {
                auto k = eval_direct_kernel<true>;
                DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
                k<<<grid, DIRECT_WARPS * 32, bytes, ctx->stream>>>(ngrid, coords, t, exp_cutoff, ao_out, g_out);
            } else {
                auto k = eval_direct_kernel<false>;
                DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
                k<<<grid, DIRECT_WARPS * 32, bytes, ctx->stream>>>(ngrid, coords, t, exp_cutoff, ao_out, g_out);
            }
            if (ctx->timing) cudaEventRecord(ctx->ev[1], ctx->stream);
            DFT_CUDA_CHECK(ctx, cudaGetLastError());
            DFT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
            if (ctx->timing && !ctx->failed) cudaEventElapsedTime(&ctx->stats.ao_ms, ctx->ev[0], ctx->ev[1]);
            return ctx->failed ? 6 : 0;
        }
    }
    // Order of the groups in phase 1.  The two half-warps of a warp work on CONSECUTIVE groups, and a warp pays for the
    // long path (three exponentials) whenever either of them is inside its cutoff: pair groups that decide alike -- the
    // same reach sqrt(cutoff / amin) (core shells reach 4 bohr, valence shells 13-19) and neighbouring centres (sorted
    // along the molecule's longest axis within a reach class).  Shell by shell in input order, a carbon's 1s group sat
    // next to its own 2sp group, which almost never agree.
    if (!ctx->ao_input_order) {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (const HGroup& q : groups) {
            const double c[3] = {q.x, q.y, q.z};
            for (int d = 0; d < 3; ++d) { lo[d] = c[d] < lo[d] ? c[d] : lo[d]; hi[d] = c[d] > hi[d] ? c[d] : hi[d]; }
        }
        int ax = 0;
        for (int d = 1; d < 3; ++d) if (hi[d] - lo[d] > hi[ax] - lo[ax]) ax = d;
        std::vector<int> order(ngroup), inv(ngroup);
        for (int g = 0; g < ngroup; ++g) order[g] = g;
        auto reach = [&](const HGroup& q) { return (long)(sqrt(exp_cutoff / q.amin) + 0.5); };   // bohr, rounded
        auto pos = [&](const HGroup& q) { return ax == 0 ? q.x : (ax == 1 ? q.y : q.z); };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            const long ra = reach(groups[a]), rb = reach(groups[b]);
            if (ra != rb) return ra > rb;
            return pos(groups[a]) < pos(groups[b]);
        });
        std::vector<HGroup> sorted(ngroup);
        for (int g = 0; g < ngroup; ++g) { sorted[g] = groups[order[g]]; inv[order[g]] = g; }
        groups.swap(sorted);
        for (int s = 0; s < nshell; ++s) shell_group[s] = inv[shell_group[s]];
    }
    std::vector<int> mshell(nshell), mcoef(nshell);
    {
        int off = 0;
        for (auto& q : groups) { q.mem_off = off; off += q.nmem; q.nmem = 0; }
        for (int s = 0; s < nshell; ++s) {
            HGroup& q = groups[shell_group[s]];
            mshell[q.mem_off + q.nmem] = s;
            mcoef[q.mem_off + q.nmem] = (int)coefs.size();
            q.nmem++;
            for (int k = 0; k < shell_nprim[s]; ++k) coefs.push_back(prim_coef[shell_prim_off[s] + k]);
        }
    }
    const int nexp = (int)exps.size(), ncoef = (int)coefs.size();

    // ---- one structure-of-arrays staging block
    Tables t;
    memset(&t, 0, sizeof(t));
    t.ngroup = ngroup; t.nshell = nshell; t.nmember = nshell; t.nexp = nexp; t.ncoef = ncoef; t.nao = nao;
    size_t off = 0;
    auto take = [&](size_t nbytes) { const size_t o = off; off += (nbytes + 7) & ~(size_t)7; return o; };
    t.o_gx = take(8 * (size_t)ngroup); t.o_gy = take(8 * (size_t)ngroup); t.o_gz = take(8 * (size_t)ngroup);
    t.o_gamin = take(8 * (size_t)ngroup);
    t.o_sx = take(8 * (size_t)nshell); t.o_sy = take(8 * (size_t)nshell); t.o_sz = take(8 * (size_t)nshell);
    t.o_exp = take(8 * (size_t)nexp); t.o_coef = take(8 * (size_t)ncoef);
    t.o_gprim = take(4 * (size_t)ngroup); t.o_gnprim = take(4 * (size_t)ngroup);
    t.o_gmem = take(4 * (size_t)ngroup); t.o_gnmem = take(4 * (size_t)ngroup);
    t.o_mshell = take(4 * (size_t)nshell); t.o_mcoef = take(4 * (size_t)nshell);
    t.o_aoshell = take(4 * (size_t)nao); t.o_aocomp = take(4 * (size_t)nao);
    off = (off + 15) & ~(size_t)15;   // the per-warp double2 scratch follows
    const size_t bytes = off;
    std::vector<unsigned char> h(bytes, 0);
    auto dptr = [&](size_t o) { return reinterpret_cast<double*>(h.data() + o); };
    auto iptr = [&](size_t o) { return reinterpret_cast<int*>(h.data() + o); };
    for (int g = 0; g < ngroup; ++g) {
        dptr(t.o_gx)[g] = groups[g].x; dptr(t.o_gy)[g] = groups[g].y; dptr(t.o_gz)[g] = groups[g].z;
        dptr(t.o_gamin)[g] = groups[g].amin;
        iptr(t.o_gprim)[g] = groups[g].prim_off; iptr(t.o_gnprim)[g] = groups[g].nprim;
        iptr(t.o_gmem)[g] = groups[g].mem_off; iptr(t.o_gnmem)[g] = groups[g].nmem;
    }
    for (int s = 0; s < nshell; ++s) {
        dptr(t.o_sx)[s] = shell_xyz[3 * s]; dptr(t.o_sy)[s] = shell_xyz[3 * s + 1]; dptr(t.o_sz)[s] = shell_xyz[3 * s + 2];
        iptr(t.o_mshell)[s] = mshell[s]; iptr(t.o_mcoef)[s] = mcoef[s];
    }
    memcpy(dptr(t.o_exp), exps.data(), 8 * (size_t)nexp);
    memcpy(dptr(t.o_coef), coefs.data(), 8 * (size_t)ncoef);
    if (nshell >= (1 << 24)) return 5;
    for (int i = 0; i < nao; ++i) iptr(t.o_aoshell)[i] = ao_shell[i] | ((ao_comp[i] + 1) << 24);
    memcpy(iptr(t.o_aocomp), ao_comp.data(), 4 * (size_t)nao);
    unsigned char* d = (unsigned char*)ctx->scratch.ensure(bytes, &ctx->failed);
    if (ctx->failed) return 4;
    DFT_CUDA_CHECK(ctx, cudaMemcpyAsync(d, h.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    DFT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // h is pageable: the copy is done now
    t.base = d;
    t.bytes = bytes;

    // ---- shared memory: tables + the block's points + radial sums [G][epitch] (epitch odd)
    const int epitch = nshell | 1;
    const size_t smem_max = 227 * 1024;
    auto smem_for = [&](int G) { return bytes + sizeof(double) * 6 * G + sizeof(double2) * (size_t)G * epitch + 64; };
    // ao_shape: 8 | 16 | 32 points per block with 8 | 8 | 16 warps, 17 = 16 points with 16 warps (twice the resident
    // warps on the same shared memory).  ao_vec_stores: 16-byte stores in phase 2 (a lane owns an aligned AO pair).
    // Measured at C5 / C4 (profiles/r2_u3_ao_eval_variants.txt): the kernel is bound by instruction issue and
    // shared-memory latency, not by the number of store instructions -- 3.59 / 0.83 ms with 16-byte stores against
    // 3.52 / 0.73 ms with 8-byte ones (a warp's 8-byte stores already fill whole 256-byte runs) -- so 8-byte is the default.
    int shape = ctx->ao_shape;
    bool vec = ctx->ao_vec_stores;
    // 16-byte stores need 16-byte aligned output bases (cudaMalloc gives 256)
    if ((d_ao_ptr & 15ull) || (deriv && (d_ao_grad_ptr & 15ull))) vec = false;
    if (shape != 8 && shape != 16 && shape != 17 && shape != 32) shape = 2 * (smem_for(32) + 1024) <= smem_max ? 32 : 16;
    const int G = shape == 17 ? 16 : shape, NW = (shape == 32 || shape == 17) ? 16 : 8;
    const size_t smem = smem_for(G);
    if (smem > smem_max) {
        fprintf(stderr, "[dft_b200] DFT_EvalAO: basis too large for the shared-memory staging (%zu B)\n", smem);
        return 5;
    }
    int per_sm = (int)(smem_max / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 64 / NW) per_sm = 64 / NW;
    const long nblk = ((long)ngrid + G - 1) / G;
    const int grid = (int)(nblk < (long)ctx->num_sms * per_sm ? nblk : (long)ctx->num_sms * per_sm);
    xc::resolve_times(ctx);   // (this call reuses the first two timing events)
    if (ctx->timing) cudaEventRecord(ctx->ev[0], ctx->stream);
#define DFT_AO_LAUNCH(D_, G_, NW_)                                                                                      \
    do {                                                                                                                \
        if (vec) {                                                                                                      \
            auto k = eval_kernel<D_, G_, NW_, true>;                                                                    \
            DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
            k<<<grid, NW_ * 32, smem, ctx->stream>>>(ngrid, coords, t, epitch, exp_cutoff, ao_out, g_out);              \
        } else {                                                                                                        \
            auto k = eval_kernel<D_, G_, NW_, false>;                                                                   \
            DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
            k<<<grid, NW_ * 32, smem, ctx->stream>>>(ngrid, coords, t, epitch, exp_cutoff, ao_out, g_out);              \
        }                                                                                                               \
    } while (0)
    if (deriv) { if (G == 32) DFT_AO_LAUNCH(true, 32, 16); else if (NW == 16) DFT_AO_LAUNCH(true, 16, 16); else if (G == 16) DFT_AO_LAUNCH(true, 16, 8); else DFT_AO_LAUNCH(true, 8, 8); }
    else { if (G == 32) DFT_AO_LAUNCH(false, 32, 16); else if (NW == 16) DFT_AO_LAUNCH(false, 16, 16); else if (G == 16) DFT_AO_LAUNCH(false, 16, 8); else DFT_AO_LAUNCH(false, 8, 8); }
#undef DFT_AO_LAUNCH
    if (ctx->timing) cudaEventRecord(ctx->ev[1], ctx->stream);
    DFT_CUDA_CHECK(ctx, cudaGetLastError());
    DFT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->timing && !ctx->failed) cudaEventElapsedTime(&ctx->stats.ao_ms, ctx->ev[0], ctx->ev[1]);
    return ctx->failed ? 6 : 0;
}
