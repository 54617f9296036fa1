// xc_tma.cu -- the fast XC path: TMA-fed, mbarrier-pipelined FP64 tensor-core (DMMA) kernels.
//
// Two persistent, warp-specialised kernels per XC build.  One CTA per SM, 384 threads: two consumer
// warpgroups (8 warps in a 2 x 4 grid, warp tile 64 x 32 at the widest) and one producer warpgroup
// whose first lane drives TMA; `setmaxnreg` moves registers from the producer to the consumers
// (40 / 232 per thread) so that 64 x 32 FP64 accumulator tiles fit without spills.
//
//   density_tma_kernel   subsystem (b)+(c): for each block of 128 grid points
//        C = Phi_blk . Dsym                      DMMA tiles, operands streamed by TMA (SWIZZLE_128B)
//        rho = rowsum(C o Phi), grad rho = 2 rowsum(C o dPhi)   fused epilogue, C never stored
//        pointwise functional once per point -> (a, b) coefficients + E_xc partial
//     replaces get_rho_kernel / get_rho_sigma_kernel_planar (dft_solver.cu:294-307, :346-380), both
//     passes of the *_fused_kernel's (:309-513) and reduce_sum_kernel (:285-292).
//
//   vxc_tma_kernel       subsystem (d): for each (output tile, grid slice)
//        B = a o Phi + b . grad Phi               built on the fly in shared memory (no (ngrid,nao) B
//        M += B^T Phi                              matrix in HBM), DMMA tiles, split over grid slices
//     replaces the B matrix (:577,:613,:655) and cublasDgemm (:580,:616,:663).
//
//   finalize_tma_kernel  out = M + M^T over slices in a fixed order (replaces :515-527) + E_xc.
//
// Shared-memory operand tiles are written by TMA with the 128-byte swizzle; fragment rows (density
// kernel) or reduction rows (V kernel) are permuted so that every 64-bit fragment load is
// bank-conflict free ("swizzle and fragment permutation").
//
// Odd nao.  The caller's AO rows are 8*nao bytes apart; TMA needs 16-byte aligned rows and box
// starts.  For odd nao only every second row is aligned, so the grid is split into two
// SUB-PROBLEMS that are each a clean 2-D tensor with pitch 2*nao:
//     E: even rows g = 2j,   columns 0..nao-1, base = ptr
//     O: odd rows  g = 2j+1, base = ptr + (nao-1)*8 (16-byte aligned), columns 0..nao where column 0
//        is the last element of the previous row (finite junk) and column c >= 1 is AO index c-1.
// Sub-problem O therefore works with every AO index shifted by one: it multiplies by a copy of
// Dsym shifted by (1,1) (row/column 0 zero, so the junk column is annihilated) and accumulates a
// V matrix shifted by (1,1), which the finalize kernel un-shifts.  Blocks, tiles and slices are
// always uniform in parity, so the inner loops do not know about any of this.
// Inputs TMA cannot address at all (misaligned base pointers, odd nao with odd ngrid) take the
// generic path (xc_generic.cu).
#include <cuda.h>

#include <cstdio>
#include <cstring>

#include "dmma.cuh"
#include "engine.h"
#include "tma.cuh"
#include "xc_functionals.cuh"

namespace xc {
namespace tmapath {

constexpr int NCW = 8;                      // V kernel consumer warps, 2 (m) x 4 (n)
constexpr int NCONS = NCW * 32;             // 256 consumer threads = 2 warpgroups
constexpr int NTHREADS = NCONS + 128;       // + 1 producer warpgroup
constexpr int REGS_CONSUMER = 232;          // 1 CTA/SM of 384 threads: 384*168 = 256*232 + 128*40
constexpr int REGS_PRODUCER = 40;
#ifdef DFT_DEBUG_REGS2
constexpr int REGS_CONSUMER_2CTA = DFT_DEBUG_REGS2;
#else
constexpr int REGS_CONSUMER_2CTA = 216;     // 2 CTAs/SM of 256 threads: 256*128 = 128*216 + 128*40
#endif

// Density kernel shape, selected by WM (number of 64-row warp rows):
//   WM = 2: one CTA per SM, 8 consumer warps (2 x 4), 128-row blocks, 384 threads;
//   WM = 1: two CTAs per SM, 4 consumer warps (1 x 4) each, 64-row blocks, 256 threads -- the two
//           CTAs drift out of phase, so one CTA's fused epilogue (global gathers, no tensor work)
//           overlaps the other's k-loop and the DMMA pipe stays busy.
template <int WM>
struct DShape {
    static constexpr int MB = 64 * WM;            // grid rows per block
    static constexpr int NCW = 4 * WM;            // consumer warps
    static constexpr int NCONS = NCW * 32;
    static constexpr int NTHREADS = NCONS + 128;  // + producer warpgroup
#ifdef DFT_DEBUG_WM1_SINGLE   // diagnostic build: 64-row shape but one CTA per SM
    static constexpr int CTAS_PER_SM = 1;
    static constexpr int REGS_CONS = WM == 1 ? DFT_DEBUG_WM1_SINGLE : REGS_CONSUMER;
#else
    static constexpr int CTAS_PER_SM = WM == 1 ? 2 : 1;
    static constexpr int REGS_CONS = WM == 1 ? REGS_CONSUMER_2CTA : REGS_CONSUMER;
#endif
    static constexpr int STAGES = WM == 1 ? 4 : 5;
    static constexpr int A_TILE_BYTES = MB * 128;  // MB rows x 16 doubles
};
constexpr int VP_STAGES = 2;                // V kernel: plane ring (TMA -> builder warps)
constexpr int VN_STAGES = 3;                // V kernel: Phi column-tile ring (TMA -> MMA warps)
constexpr int VB_STAGES = 2;                // V kernel: B tile double buffer (built by the consumer warps)
constexpr int D_PREFETCH_LEAD = 8;          // density: k-chunks before the epilogue at which grad tiles are L2-prefetched
constexpr int VK = 16;                      // grid rows per V chunk

struct SubProblem {
    int rows;    // rows of this sub-problem
    int gmul;    // grid point of row j: g = gmul * j + gadd
    int gadd;
    int shift;   // AO index = column - shift (0 or 1)
    int blk0;    // first density block of this sub-problem
    int coef0;   // first row of this sub-problem in the coefficient array (padded to 128 rows)
};

struct DensityParams {
    CUtensorMap map_a[2];   // Phi of each sub-problem, box 16 x 128
    CUtensorMap map_g[2][3];  // grad Phi planes, box 16 x 128 (L2 prefetch of the epilogue tiles only)
    CUtensorMap map_d;      // [2][NP][KP] symmetrised density (plain, shifted), box 16 x NT
    SubProblem sub[2];
    int nsub, ngrid, nao, xc_mode, nblocks, ntiles, nk, NP, l2_prefetch;
    const double *ao, *gx, *gy, *gz, *w;
    double* coef;
    double* exc_part;
};

struct VxcParams {
    CUtensorMap map_p[2][4];  // planes of each sub-problem, box 16 x 16
    SubProblem sub[2];
    int nsub, ntiles, lda_half, rows_per_slice, slices_per_sub, NP;
    const double* coef;
    double* vpart;
};

__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64x2(uint32_t addr, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
template <int N>
__device__ __forceinline__ void reg_inc() {
#ifndef DFT_DEBUG_NO_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
#endif
}
template <int N>
__device__ __forceinline__ void reg_dec() {
#ifndef DFT_DEBUG_NO_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
#endif
}

__device__ __forceinline__ xcfun::PointCoef eval_mode(int mode, double rho, double gx, double gy, double gz, double w) {
    switch (mode) {
        case 0: return xcfun::evaluate_point<0, false>(rho, gx, gy, gz, w);
        case 1: return xcfun::evaluate_point<0, true>(rho, gx, gy, gz, w);
        case 2: return xcfun::evaluate_point<1, false>(rho, gx, gy, gz, w);
        case 3: return xcfun::evaluate_point<1, true>(rho, gx, gy, gz, w);
        default: return xcfun::evaluate_point<2, false>(rho, gx, gy, gz, w);
    }
}

// ------------------------------------------------------------------------------------------------
// density kernel
// ------------------------------------------------------------------------------------------------
template <int NF, int WM>
struct DensitySmem {
    using S = DShape<WM>;
    static constexpr int NT = 32 * NF;
    static constexpr int B_TILE_BYTES = NT * 128;
    static constexpr int STAGE_BYTES = S::A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int RED_OFF = S::STAGES * STAGE_BYTES;             // double red[MB][4][4]
    static constexpr int BAR_OFF = RED_OFF + S::MB * 4 * 4 * 8;         // full[STAGES], empty[STAGES]
    static constexpr int ESUM_OFF = BAR_OFF + 2 * S::STAGES * 8;
    static constexpr int TOTAL = ESUM_OFF + 8 * 8 + 1024;               // + alignment slack
};

template <int NF, int NPL, int WM>
__global__ void __launch_bounds__(DShape<WM>::NTHREADS, DShape<WM>::CTAS_PER_SM)
density_tma_kernel(const __grid_constant__ DensityParams P) {
    using L = DensitySmem<NF, WM>;
    using S = DShape<WM>;
    constexpr int NT = L::NT;
    constexpr int MB = S::MB, NCW = S::NCW, NCONS = S::NCONS, D_STAGES = S::STAGES;
    constexpr int A_TILE_BYTES = S::A_TILE_BYTES;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (tma::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - tma::smem_u32(smem_raw));
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + L::BAR_OFF);
    uint64_t* empty = full + D_STAGES;
    double* red = reinterpret_cast<double*>(sm + L::RED_OFF);
    double* esum = reinterpret_cast<double*>(sm + L::ESUM_OFF);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < D_STAGES; ++s) {
            tma::mbar_init(&full[s], 1);
            tma::mbar_init(&empty[s], NCW);
        }
        tma::fence_barrier_init();
    }
    __syncthreads();

    const int nblocks = P.nblocks, ntiles = P.ntiles, nk = P.nk;

    if (warp >= NCW) {
        // ===================== producer warpgroup: one elected lane drives TMA =====================
        reg_dec<REGS_PRODUCER>();
        if (warp == NCW && lane == 0) {
            tma::prefetch_map(&P.map_a[0]);
            if (P.nsub > 1) tma::prefetch_map(&P.map_a[1]);
            tma::prefetch_map(&P.map_d);
            uint32_t it = 0;
            for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
                const int si = (P.nsub > 1 && b >= P.sub[1].blk0) ? 1 : 0;
                const int blk = b - P.sub[si].blk0;
                const int drow0 = P.sub[si].shift * P.NP;
                for (int nt = 0; nt < ntiles; ++nt) {
                    const int kpf = nk > D_PREFETCH_LEAD ? nk - D_PREFETCH_LEAD : 0;
                    for (int kc = 0; kc < nk; ++kc, ++it) {
                        if (NPL == 4 && kc == kpf && P.l2_prefetch) {
                            // the epilogue of this tile gathers grad Phi[blk rows][nt columns] with plain
                            // loads: pull those tiles into L2 a few chunks ahead (Phi itself is L2-hot, it
                            // is this block's A operand)
                            for (int pl = 0; pl < 3; ++pl)
                                for (int bx = 0; bx < NT / 16; ++bx)
                                    tma::prefetch_2d(&P.map_g[si][pl], nt * NT + 16 * bx, blk * MB);
                        }
                        const uint32_t s = it % D_STAGES, ph = (it / D_STAGES) & 1u;
                        tma::mbar_wait(&empty[s], ph ^ 1u);
                        unsigned char* st = sm + s * L::STAGE_BYTES;
                        tma::mbar_arrive_expect_tx(&full[s], L::STAGE_BYTES);
                        tma::load_2d(st, &P.map_a[si], kc * 16, blk * MB, &full[s]);
                        tma::load_2d(st + A_TILE_BYTES, &P.map_d, kc * 16, drow0 + nt * NT, &full[s]);
                    }
                }
            }
        }
        // Park the whole producer warpgroup until the consumers are done instead of exiting: warps
        // that exit right after setmaxnreg.dec hand their registers back to the SM while a co-resident
        // CTA is still re-allocating, which corrupted consumer registers with two CTAs per SM.
        __syncwarp();
        tma::named_bar_sync(2, S::NTHREADS);
        return;
    }

    // ===================== consumers: 8 warps, warp tile 64 x (8 NF) =====================
    reg_inc<S::REGS_CONS>();
    const int wm = warp >> 2, wn = warp & 3;  // wm in [0, WM)
    const int q = lane >> 2, qcol = lane & 3;
    const int perm = 2 * (q & 3) + (q >> 2);  // fragment row -> tile row: conflict-free with SWIZZLE_128B
    // per-lane byte offsets inside a 128-byte-row tile for k-step ks: chunk = (2ks + qcol/2) ^ perm
    uint32_t koff[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) koff[ks] = ((((2 * ks + (qcol >> 1)) ^ perm) & 7) << 4) + ((qcol & 1) << 3);
    const uint32_t a_row = (uint32_t)(wm * 64 + perm) * 128u;
    const uint32_t b_row = (uint32_t)(wn * 8 * NF + perm) * 128u;
    // accumulator column j = 2 qcol + e of an n-fragment is tile column perm_j(j)
    int ncol[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int j = 2 * qcol + e;
        ncol[e] = 2 * (j & 3) + (j >> 2);
    }
    const int nao = P.nao;
    const double* __restrict__ ao = P.ao;
    const double* __restrict__ gx = P.gx;
    const double* __restrict__ gy = P.gy;
    const double* __restrict__ gz = P.gz;

    double e_acc = 0.0;
    uint32_t it = 0;
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const int si = (P.nsub > 1 && b >= P.sub[1].blk0) ? 1 : 0;
        const int blk = b - P.sub[si].blk0;
        const int rows = P.sub[si].rows, gmul = P.sub[si].gmul, gadd = P.sub[si].gadd, shift = P.sub[si].shift;
        for (int nt = 0; nt < ntiles; ++nt) {
            double acc[8][NF][2];
#pragma unroll
            for (int mf = 0; mf < 8; ++mf)
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;

            for (int kc = 0; kc < nk; ++kc, ++it) {
                const uint32_t s = it % D_STAGES, ph = (it / D_STAGES) & 1u;
                tma::mbar_wait(&full[s], ph);
                const uint32_t a_base = base + s * L::STAGE_BYTES + a_row;
                const uint32_t b_base = base + s * L::STAGE_BYTES + A_TILE_BYTES + b_row;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    double a[8], bf[NF];
#pragma unroll
                    for (int mf = 0; mf < 8; ++mf) a[mf] = lds_f64(a_base + mf * 1024 + koff[ks]);
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) bf[nf] = lds_f64(b_base + nf * 1024 + koff[ks]);
#pragma unroll
                    for (int mf = 0; mf < 8; ++mf)
#pragma unroll
                        for (int nf = 0; nf < NF; ++nf) dmma::mma8x8x4(acc[mf][nf], a[mf], bf[nf]);
                }
                __syncwarp();
                if (lane == 0) tma::mbar_arrive(&empty[s]);
            }
            // ---- fused epilogue: row-dots of C with Phi and grad Phi (global loads, L2-hot for Phi);
            //      the row sums of this column tile are folded into shared memory right away so that
            //      no row accumulator stays live across the k-loop (each red[] entry has one owner lane)
            const int nbase = nt * NT + wn * 8 * NF - shift;
            // column indices / validity of this lane's 2 NF accumulator columns (same for every row)
            int ncl[NF][2];
            bool nok[NF][2];
#pragma unroll
            for (int nf = 0; nf < NF; ++nf)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int n = nbase + nf * 8 + ncol[e];
                    nok[nf][e] = (n >= 0) && (n < nao);
                    ncl[nf][e] = min(max(n, 0), nao - 1);
                }
#pragma unroll
            for (int mf = 0; mf < 8; ++mf) {
                const int r = wm * 64 + mf * 8 + perm;
                const int j = blk * MB + r;
                const bool rok = j < rows;
                const size_t rowoff = (size_t)((long)gmul * (rok ? j : 0) + gadd) * nao;
                // all loads of one fragment row are issued back to back (unconditional, clamped
                // addresses) so that 8 NF x NPL independent requests are in flight per lane
                double pv[NPL][NF][2];
#pragma unroll
                for (int nf = 0; nf < NF; ++nf)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const size_t o = rowoff + ncl[nf][e];
                        pv[0][nf][e] = __ldg(ao + o);
                        if (NPL == 4) {
                            pv[1][nf][e] = __ldg(gx + o);
                            pv[2][nf][e] = __ldg(gy + o);
                            pv[3][nf][e] = __ldg(gz + o);
                        }
                    }
                double rs[NPL];
#pragma unroll
                for (int p = 0; p < NPL; ++p) rs[p] = 0.0;
#pragma unroll
                for (int nf = 0; nf < NF; ++nf)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double cv = (rok && nok[nf][e]) ? acc[mf][nf][e] : 0.0;
#pragma unroll
                        for (int p = 0; p < NPL; ++p) rs[p] = fma(cv, pv[p][nf][e], rs[p]);
                    }
#pragma unroll
                for (int p = 0; p < NPL; ++p) {
                    double v = rs[p];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if (qcol == 0) {
                        double* dst = red + (r * 4 + wn) * 4 + p;
                        *dst = (nt == 0) ? v : *dst + v;
                    }
                }
            }
        }
        tma::named_bar_sync(1, NCONS);
        if (tid < MB) {
            const int r = tid;
            const int j = blk * MB + r;
            double2 c01 = make_double2(0.0, 0.0), c23 = make_double2(0.0, 0.0);
            if (j < rows) {
                const long g = (long)gmul * j + gadd;
                const double* rr = red + r * 16;
                const double rho = (rr[0] + rr[4]) + (rr[8] + rr[12]);
                double dx = 0.0, dy = 0.0, dz = 0.0;
                if (NPL == 4) {
                    dx = 2.0 * ((rr[1] + rr[5]) + (rr[9] + rr[13]));
                    dy = 2.0 * ((rr[2] + rr[6]) + (rr[10] + rr[14]));
                    dz = 2.0 * ((rr[3] + rr[7]) + (rr[11] + rr[15]));
                }
                const xcfun::PointCoef pc = eval_mode(P.xc_mode, rho, dx, dy, dz, __ldg(P.w + g));
                c01 = make_double2(pc.a, pc.bx);
                c23 = make_double2(pc.by, pc.bz);
                e_acc += pc.exc;
            }
            // coefficient rows are stored per sub-problem, padded to whole blocks (zeros past the end)
            double2* cp = reinterpret_cast<double2*>(P.coef) + 2 * ((size_t)P.sub[si].coef0 + j);
            cp[0] = c01;
            cp[1] = c23;
        }
        tma::named_bar_sync(1, NCONS);
    }
    // ---- per-CTA E_xc partial (fixed order)
    if (warp < MB / 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e_acc += __shfl_xor_sync(0xffffffffu, e_acc, o);
        if (lane == 0) esum[warp] = e_acc;
    }
    tma::named_bar_sync(1, NCONS);
    if (tid == 0) {
        double e = esum[0] + esum[1];
        if (MB / 32 == 4) e += esum[2] + esum[3];
        P.exc_part[blockIdx.x] = e;
    }
    tma::named_bar_sync(2, S::NTHREADS);  // releases the parked producer warpgroup
}

// ------------------------------------------------------------------------------------------------
// V kernel
// ------------------------------------------------------------------------------------------------
template <int NF, int NPL>
struct VxcSmem {
    static constexpr int NT = 32 * NF;
    static constexpr int TILE_BYTES = VK * NT * 8;                       // one plane tile: 2NF boxes of 2 KB
    static constexpr int P_STAGE_BYTES = NPL * TILE_BYTES + 1024;        // planes + 16 x (a,bx,by,bz)
    static constexpr int COEF_OFF = NPL * TILE_BYTES;
    static constexpr int N_OFF = VP_STAGES * P_STAGE_BYTES;              // Phi column-tile ring
    static constexpr int BPITCH = NT + 4;                                // doubles; (NT+4) mod 16 == 4
    static constexpr int BS_BYTES = ((VK * BPITCH * 8 + 1023) / 1024) * 1024;
    static constexpr int BS_OFF = N_OFF + VN_STAGES * TILE_BYTES;
    static constexpr int BAR_OFF = BS_OFF + VB_STAGES * BS_BYTES;
    static constexpr int NBAR = 2 * (VP_STAGES + VN_STAGES);
    static constexpr int TOTAL = BAR_OFF + NBAR * 8 + 1024;
};

// Roles (384 threads): warps 0..7 build the B tile of each chunk and run the MMAs (2 x 4, warp tile
// 16NF x 8NF); warp 8 lane 0 drives TMA.  Two TMA rings with independent lifetimes: the plane ring
// is released as soon as the B tile is built (so the next chunks' planes stream in during the MMAs),
// the Phi column-tile ring is released after the MMAs.
template <int NF, int NPL>
__global__ void __launch_bounds__(NTHREADS, 1)
vxc_tma_kernel(const __grid_constant__ VxcParams P) {
    using L = VxcSmem<NF, NPL>;
    constexpr int NT = L::NT;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (tma::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - tma::smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::BAR_OFF);
    uint64_t* full_p = bars;
    uint64_t* empty_p = full_p + VP_STAGES;
    uint64_t* full_n = empty_p + VP_STAGES;
    uint64_t* empty_n = full_n + VN_STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // output tile of this CTA
    const int ntiles = P.ntiles;
    int tm, tn;
    if (P.lda_half) {  // upper-triangular tile pairs, row-major
        int t = blockIdx.x;
        tm = 0;
        while (t >= ntiles - tm) { t -= ntiles - tm; ++tm; }
        tn = tm + t;
    } else {
        tm = blockIdx.x / ntiles;
        tn = blockIdx.x % ntiles;
    }
    const int m0 = tm * NT, n0 = tn * NT;
    // grid slice of this CTA: rows [jbeg, jend) of sub-problem si
    const int si = blockIdx.y / P.slices_per_sub;
    const int sl = blockIdx.y % P.slices_per_sub;
    const int jbeg = sl * P.rows_per_slice;
    const int jend = min(P.sub[si].rows, jbeg + P.rows_per_slice);
    const int nchunks = jend > jbeg ? (jend - jbeg + VK - 1) / VK : 0;

    if (tid == 0) {
        for (int s = 0; s < VP_STAGES; ++s) { tma::mbar_init(&full_p[s], 1); tma::mbar_init(&empty_p[s], NCW); }
        for (int s = 0; s < VN_STAGES; ++s) { tma::mbar_init(&full_n[s], 1); tma::mbar_init(&empty_n[s], NCW); }
        tma::fence_barrier_init();
    }
    __syncthreads();

    if (warp >= NCW) {
        reg_dec<REGS_PRODUCER>();
        // ===================== TMA producer =====================
        if (warp == NCW && lane == 0) {
            for (int p = 0; p < NPL; ++p) tma::prefetch_map(&P.map_p[si][p]);
            const double* coef = P.coef + 4 * (size_t)P.sub[si].coef0;
            for (int c = 0; c < nchunks; ++c) {
                const int j0 = jbeg + c * VK;
                {
                    const uint32_t s = c % VP_STAGES, ph = (c / VP_STAGES) & 1u;
                    tma::mbar_wait(&empty_p[s], ph ^ 1u);
                    unsigned char* st = sm + s * L::P_STAGE_BYTES;
                    tma::mbar_arrive_expect_tx(&full_p[s], (uint32_t)(NPL * L::TILE_BYTES + VK * 32));
                    for (int p = 0; p < NPL; ++p)
                        for (int b = 0; b < 2 * NF; ++b)
                            tma::load_2d(st + p * L::TILE_BYTES + b * 2048, &P.map_p[si][p], m0 + 16 * b, j0, &full_p[s]);
                    tma::load_1d(st + L::COEF_OFF, coef + 4 * (size_t)j0, VK * 32, &full_p[s]);
                }
                {
                    const uint32_t s = c % VN_STAGES, ph = (c / VN_STAGES) & 1u;
                    tma::mbar_wait(&empty_n[s], ph ^ 1u);
                    unsigned char* st = sm + L::N_OFF + s * L::TILE_BYTES;
                    tma::mbar_arrive_expect_tx(&full_n[s], (uint32_t)L::TILE_BYTES);
                    for (int b = 0; b < 2 * NF; ++b)
                        tma::load_2d(st + b * 2048, &P.map_p[si][0], n0 + 16 * b, j0, &full_n[s]);
                }
            }
        }
        return;
    }

    // ===================== MMA warps: 2 x 4, warp tile (16 NF) x (8 NF) =====================
    reg_inc<REGS_CONSUMER>();
    const int wm = warp >> 2, wn = warp & 3;
    const int q = lane >> 2, qcol = lane & 3;
    constexpr int MF = 2 * NF;
    double acc[MF][NF][2];
#pragma unroll
    for (int mf = 0; mf < MF; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;

    // B-fragment (Phi n-tile) addressing: reduction row of k-step ks is 8(ks/2) + 2 qcol + (ks&1)
    uint32_t boff[4][NF];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int r = 8 * (ks >> 1) + 2 * qcol + (ks & 1);
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
            const int n = wn * 8 * NF + nf * 8 + q;
            boff[ks][nf] = (uint32_t)((n >> 4) * 2048) + tma::swz128((uint32_t)r, (uint32_t)(n & 15));
        }
    }
    const uint32_t aoff = (uint32_t)((qcol * L::BPITCH + wm * 16 * NF + q) * 8);

    // ---- B-tile builder, software-pipelined under the MMAs of the previous chunk.
    // B = a Phi + bx dxPhi + by dyPhi + bz dzPhi for one chunk of 16 grid points: 256 NF tasks of one
    // 16-byte chunk each, NF per thread, handled in batches of TB = 2 (loads first, math + store later
    // so that the shared-memory latency hides behind 32 DMMAs).
    constexpr int TB = NF >= 2 ? 2 : 1;
    constexpr int NBATCH = (NF + TB - 1) / TB;  // 1 or 2
    double2 bca[TB], bcb[TB], bv[TB][NPL];
    uint32_t bdst[TB];
    auto build_load = [&](int batch, uint32_t st, uint32_t bs) {
#pragma unroll
        for (int u = 0; u < TB; ++u) {
            const int t = (batch * TB + u < NF) ? batch * TB + u : NF - 1;
            const int task = tid + t * NCONS;
            const int j = task & 7, rb = task >> 3;
            const int r = rb & 15, b = rb >> 4;
            const uint32_t off = (uint32_t)(b * 2048 + r * 128 + (((j ^ r) & 7) << 4));
            bca[u] = lds_f64x2(st + L::COEF_OFF + r * 32);
            if (NPL == 4) bcb[u] = lds_f64x2(st + L::COEF_OFF + r * 32 + 16);
#pragma unroll
            for (int p = 0; p < NPL; ++p) bv[u][p] = lds_f64x2(st + p * L::TILE_BYTES + off);
            const int rho_idx = 8 * (r >> 3) + 4 * (r & 1) + ((r & 7) >> 1);  // MMA order of tile row r
            bdst[u] = bs + (uint32_t)((rho_idx * L::BPITCH + b * 16 + 2 * j) * 8);
        }
    };
    auto build_store = [&](int batch) {
#pragma unroll
        for (int u = 0; u < TB; ++u) {
            if (batch * TB + u >= NF) continue;
            double2 o = make_double2(bca[u].x * bv[u][0].x, bca[u].x * bv[u][0].y);
            if (NPL == 4) {
                o.x = fma(bca[u].y, bv[u][1].x, o.x); o.y = fma(bca[u].y, bv[u][1].y, o.y);
                o.x = fma(bcb[u].x, bv[u][2].x, o.x); o.y = fma(bcb[u].x, bv[u][2].y, o.y);
                o.x = fma(bcb[u].y, bv[u][3].x, o.x); o.y = fma(bcb[u].y, bv[u][3].y, o.y);
            }
            sts_f64x2(bdst[u], o);
        }
    };
    auto mma_step = [&](int ks, uint32_t bs, uint32_t phin) {
        double a[MF], bf[NF];
#pragma unroll
        for (int mf = 0; mf < MF; ++mf) a[mf] = lds_f64(bs + aoff + (uint32_t)((4 * ks * L::BPITCH + mf * 8) * 8));
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) bf[nf] = lds_f64(phin + boff[ks][nf]);
#pragma unroll
        for (int mf = 0; mf < MF; ++mf)
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) dmma::mma8x8x4(acc[mf][nf], a[mf], bf[nf]);
    };

    // prologue: B tile of chunk 0
    if (nchunks > 0) {
        tma::mbar_wait(&full_p[0], 0);
#pragma unroll
        for (int bt = 0; bt < NBATCH; ++bt) {
            build_load(bt, base, base + L::BS_OFF);
            build_store(bt);
        }
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(&empty_p[0]);
        tma::named_bar_sync(1, NCONS);
    }
    for (int c = 0; c < nchunks; ++c) {
        const uint32_t sn = c % VN_STAGES, phn = (c / VN_STAGES) & 1u;
        const uint32_t bs = base + L::BS_OFF + (c & 1) * L::BS_BYTES;
        const uint32_t phin = base + L::N_OFF + sn * L::TILE_BYTES;
        const bool has_next = c + 1 < nchunks;
        const uint32_t sp1 = (c + 1) % VP_STAGES, php1 = ((c + 1) / VP_STAGES) & 1u;
        const uint32_t st1 = base + sp1 * L::P_STAGE_BYTES;
        const uint32_t bs1 = base + L::BS_OFF + ((c + 1) & 1) * L::BS_BYTES;
        tma::mbar_wait(&full_n[sn], phn);
        if (has_next) {
            tma::mbar_wait(&full_p[sp1], php1);
            build_load(0, st1, bs1);
        }
        mma_step(0, bs, phin);
        if (has_next) {
            build_store(0);
            if (NBATCH > 1) build_load(1, st1, bs1);
        }
        mma_step(1, bs, phin);
        if (has_next) {
            if (NBATCH > 1) build_store(1);
            __syncwarp();
            if (lane == 0) tma::mbar_arrive(&empty_p[sp1]);  // plane stage of chunk c+1 is free again
        }
        mma_step(2, bs, phin);
        mma_step(3, bs, phin);
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(&empty_n[sn]);
        // B tile of chunk c+1 complete, and everyone is done reading the B tile of chunk c
        tma::named_bar_sync(1, NCONS);
    }
    // ---- partial tile out
    const int NP = P.NP;
    double* out = P.vpart + (size_t)blockIdx.y * NP * NP;
#pragma unroll
    for (int mf = 0; mf < MF; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
            const int r = m0 + wm * 16 * NF + mf * 8 + q;
            const int cc = n0 + wn * 8 * NF + nf * 8 + 2 * qcol;
            *reinterpret_cast<double2*>(out + (size_t)r * NP + cc) = make_double2(acc[mf][nf][0], acc[mf][nf][1]);
        }
}

// out[i][j] = sum over slices of T(i+s, j+s) + T(j+s, i+s), s = column shift of the slice's
// sub-problem; T = M where the tile was computed (lda_half: the mirror tile otherwise).
// Fixed summation order -> bit-reproducible and exactly symmetric.
__global__ void finalize_tma_kernel(int nao, int NP, int NT, int nsub, int slices_per_sub, int shift1, int lda_half,
                                    const double* __restrict__ vpart, double* __restrict__ vxc, int nepart,
                                    const double* __restrict__ epart, double* __restrict__ d_exc) {
    const size_t n2 = (size_t)nao * nao;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n2) {
        const int i0 = (int)(idx / nao), j0 = (int)(idx % nao);
        double s = 0.0;
        for (int su = 0; su < nsub; ++su) {
            const int sh = su ? shift1 : 0;
            const int i = i0 + sh, j = j0 + sh;
            size_t o1 = (size_t)i * NP + j, o2 = (size_t)j * NP + i;
            if (lda_half) {
                if (i / NT > j / NT) o1 = o2;
                else if (j / NT > i / NT) o2 = o1;
            }
            for (int sl = 0; sl < slices_per_sub; ++sl) {
                const double* p = vpart + (size_t)(su * slices_per_sub + sl) * NP * NP;
                s += p[o1] + p[o2];
            }
        }
        vxc[idx] = s;
    }
    if (blockIdx.x == 0) {
        __shared__ double sh[256];
        double e = 0.0;
        for (int k = threadIdx.x; k < nepart; k += blockDim.x) e += epart[k];
        sh[threadIdx.x] = e;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) *d_exc = sh[0];
    }
}

// dsym[s][i][j] = 1/2 (D[i-s][j-s] + D[j-s][i-s]) inside the matrix, 0 elsewhere; s = 0..nshift-1
__global__ void symmetrize_pad_tma_kernel(int nao, int ld, int rows, int nshift, const double* __restrict__ dm,
                                          double* __restrict__ dsym) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y % rows, s = blockIdx.y / rows;
    if (j >= ld || s >= nshift) return;
    double v = 0.0;
    const int ii = i - s, jj = j - s;
    if (ii >= 0 && jj >= 0 && ii < nao && jj < nao) v = 0.5 * (dm[(size_t)ii * nao + jj] + dm[(size_t)jj * nao + ii]);
    dsym[((size_t)s * rows + i) * ld + j] = v;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            fprintf(stderr, "[dft_b200] cuTensorMapEncodeTiled not available from the driver\n");
    }
    return fn;
}

// 2-D f64 map over a row-major (rows x cols) array with row pitch `pitch_elems`, box = 16 x box_rows,
// 128-byte swizzle, zero fill out of bounds.
static bool make_map(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t pitch_elems,
                     uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {pitch_elems * 8};
    cuuint32_t box[2] = {16, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "[dft_b200] cuTensorMapEncodeTiled failed (%d): cols=%llu rows=%llu pitch=%llu box_rows=%u\n",
                (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch_elems, box_rows);
        return false;
    }
    return true;
}

// Tensor map of plane `ptr` for sub-problem (parity) si of an (ngrid x nao) array: see file header.
static bool make_sub_map(CUtensorMap* m, const double* ptr, int ngrid, int nao, bool split, int si, uint32_t box_rows) {
    if (!split) return make_map(m, ptr, (uint64_t)nao, (uint64_t)ngrid, (uint64_t)nao, box_rows);
    if (si == 0) return make_map(m, ptr, (uint64_t)nao, (uint64_t)(ngrid + 1) / 2, 2ull * nao, box_rows);
    return make_map(m, ptr + (nao - 1), (uint64_t)nao + 1, (uint64_t)ngrid / 2, 2ull * nao, box_rows);
}

template <int NF, int NPL, int WM>
static void launch(CublasHandleWrapper* ctx, const Problem& p, int nsm) {
    using DS = DShape<WM>;
    constexpr int MB = DS::MB;
    cudaStream_t st = ctx->stream;
    const int ngrid = p.ngrid, nao = p.nao;
    constexpr int NT = 32 * NF;
    const bool split = (nao & 1) != 0;
    const int nsub = (split && ngrid >= 2) ? 2 : 1;
    const int ncols = nao + (split ? 1 : 0);            // widest sub-problem
    const int ntiles = (ncols + NT - 1) / NT;
    const int NP = ntiles * NT;
    const int KP = ((ncols + 15) / 16) * 16;

    DensityParams dp;
    VxcParams vp;
    memset(&dp, 0, sizeof(dp));
    memset(&vp, 0, sizeof(vp));
    SubProblem sub[2];
    memset(sub, 0, sizeof(sub));
    if (!split) {
        sub[0] = SubProblem{ngrid, 1, 0, 0, 0, 0};
    } else {
        sub[0] = SubProblem{(ngrid + 1) / 2, 2, 0, 0, 0, 0};
        sub[1] = SubProblem{ngrid / 2, 2, 1, 1, 0, 0};
    }
    int nblocks = 0, coef_rows = 0;
    for (int s = 0; s < nsub; ++s) {
        sub[s].blk0 = nblocks;
        sub[s].coef0 = coef_rows;
        const int nb = (sub[s].rows + MB - 1) / MB;
        nblocks += nb;
        coef_rows += nb * MB;
    }
    const int grid1 = nblocks < nsm * DS::CTAS_PER_SM ? nblocks : nsm * DS::CTAS_PER_SM;

    double* dsym = (double*)ctx->dsym.ensure(sizeof(double) * (size_t)nsub * NP * KP, &ctx->failed);
    double* coef = (double*)ctx->coef.ensure(sizeof(double) * 4 * (size_t)coef_rows, &ctx->failed);
    double* epart = (double*)ctx->epart.ensure(sizeof(double) * grid1, &ctx->failed);
    if (ctx->failed) return;

    const double* planes[4] = {p.ao, p.gx, p.gy, p.gz};
    bool ok = make_map(&dp.map_d, dsym, (uint64_t)KP, (uint64_t)nsub * NP, (uint64_t)KP, NT);
    for (int s = 0; s < nsub; ++s) {
        ok = ok && make_sub_map(&dp.map_a[s], p.ao, ngrid, nao, split, s, MB);
        for (int i = 0; i < 3; ++i)
            ok = ok && make_sub_map(&dp.map_g[s][i], NPL == 4 ? planes[i + 1] : p.ao, ngrid, nao, split, s, MB);
        for (int i = 0; i < 4; ++i) ok = ok && make_sub_map(&vp.map_p[s][i], planes[i < NPL ? i : 0], ngrid, nao, split, s, VK);
    }
    if (!ok) { ctx->failed = true; return; }

    dp.sub[0] = sub[0]; dp.sub[1] = sub[1];
    dp.nsub = nsub; dp.ngrid = ngrid; dp.nao = nao;
    dp.xc_mode = p.xc_type == 2 ? 4 : p.xc_type * 2 + (ctx->exact_functionals ? 1 : 0);
    dp.nblocks = nblocks; dp.ntiles = ntiles; dp.nk = KP / 16; dp.NP = NP; dp.l2_prefetch = ctx->l2_prefetch ? 1 : 0;
    dp.ao = p.ao; dp.gx = p.gx; dp.gy = p.gy; dp.gz = p.gz; dp.w = p.w;
    dp.coef = coef; dp.exc_part = epart;

    const int lda_half = (NPL == 1) ? 1 : 0;
    const int tiles = lda_half ? ntiles * (ntiles + 1) / 2 : ntiles * ntiles;
    const int maxrows = sub[0].rows;
    int nsl = nsm / (tiles * nsub);
    if (nsl < 1) nsl = 1;
    const int max_slices = (maxrows + VK - 1) / VK;
    if (nsl > max_slices) nsl = max_slices;
    int rows_per_slice = (maxrows + nsl - 1) / nsl;
    rows_per_slice = ((rows_per_slice + VK - 1) / VK) * VK;
    nsl = (maxrows + rows_per_slice - 1) / rows_per_slice;
    double* vpart = (double*)ctx->vpart.ensure(sizeof(double) * (size_t)nsub * nsl * NP * NP, &ctx->failed);
    if (ctx->failed) return;

    vp.sub[0] = sub[0]; vp.sub[1] = sub[1];
    vp.nsub = nsub; vp.ntiles = ntiles; vp.lda_half = lda_half; vp.rows_per_slice = rows_per_slice;
    vp.slices_per_sub = nsl; vp.NP = NP; vp.coef = coef; vp.vpart = vpart;

    using DL = DensitySmem<NF, WM>;
    using VL = VxcSmem<NF, NPL>;
    auto dk = density_tma_kernel<NF, NPL, WM>;
    auto vk = vxc_tma_kernel<NF, NPL>;
    DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(dk, cudaFuncAttributeMaxDynamicSharedMemorySize, DL::TOTAL));
    DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(vk, cudaFuncAttributeMaxDynamicSharedMemorySize, VL::TOTAL));

    if (ctx->timing) cudaEventRecord(ctx->ev[0], st);
    symmetrize_pad_tma_kernel<<<dim3((KP + 127) / 128, NP * nsub), 128, 0, st>>>(nao, KP, NP, nsub, p.dm, dsym);
    dk<<<grid1, DS::NTHREADS, DL::TOTAL, st>>>(dp);
    if (ctx->timing) cudaEventRecord(ctx->ev[1], st);
    vk<<<dim3(tiles, nsl * nsub), NTHREADS, VL::TOTAL, st>>>(vp);
    if (ctx->timing) cudaEventRecord(ctx->ev[2], st);
    const size_t n2 = (size_t)nao * nao;
    finalize_tma_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(nao, NP, NT, nsub, nsl, split ? 1 : 0, lda_half, vpart,
                                                                      p.vxc, grid1, epart, p.d_exc);
    if (ctx->timing) cudaEventRecord(ctx->ev[3], st);
    ctx->stats.launches = 4;
    ctx->stats.path = PATH_TMA;
    DFT_CUDA_CHECK(ctx, cudaGetLastError());
}

}  // namespace tmapath

bool tma_compatible(const Problem& p) {
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (p.nao < 1 || p.ngrid < 1) return false;
    if (!al16(p.ao)) return false;
    if (p.xc_type != 0 && !(al16(p.gx) && al16(p.gy) && al16(p.gz))) return false;
    if (p.nao > 128 * 16) return false;  // keep the padded D and the slice partials modest
    return tmapath::encode_fn() != nullptr;
}

void run_tma(CublasHandleWrapper* ctx, const Problem& p) {
    using namespace tmapath;
    const int ncols = p.nao + (p.nao & 1);
    const int ntiles = (ncols + 127) / 128;
    const int NF = (ncols + 32 * ntiles - 1) / (32 * ntiles);  // 1..4 -> column tile 32 NF
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    const bool two_ctas = ctx->density_ctas_per_sm != 1;
#define DFT_LAUNCH(NF_)                                                                                    \
    do {                                                                                                   \
        if (two_ctas) { if (p.xc_type == 0) launch<NF_, 1, 1>(ctx, p, nsm); else launch<NF_, 4, 1>(ctx, p, nsm); } \
        else { if (p.xc_type == 0) launch<NF_, 1, 2>(ctx, p, nsm); else launch<NF_, 4, 2>(ctx, p, nsm); }          \
    } while (0)
    switch (NF) {
        case 1: DFT_LAUNCH(1); break;
        case 2: DFT_LAUNCH(2); break;
        case 3: DFT_LAUNCH(3); break;
        default: DFT_LAUNCH(4); break;
    }
#undef DFT_LAUNCH
}

}  // namespace xc
