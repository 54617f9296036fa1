// xc_tma.cu -- the fast XC path: TMA-fed, mbarrier-pipelined FP64 tensor-core (DMMA) kernels.
//
// Two persistent, warp-specialised kernels per XC build (one CTA per SM, 8 consumer warps in a
// 2 x 4 grid + 1 TMA producer warp):
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
// bank-conflict free (see DESIGN.md "swizzle and fragment permutation").
//
// Row pitch of the caller's AO arrays is 8*nao bytes.  TMA needs 16-byte multiples, so for odd nao
// the arrays are addressed as (ngrid/2) x (2 nao) "row pairs" and each tile is fetched with two box
// loads (even rows, odd rows); rows of a tile are then a fixed permutation of grid points, which is
// harmless for both contractions.  Inputs TMA cannot address at all (misaligned base pointers, odd
// nao with odd ngrid) take the generic path (xc_generic.cu).
#include <cuda.h>

#include <cstdio>

#include "dmma.cuh"
#include "engine.h"
#include "tma.cuh"
#include "xc_functionals.cuh"

namespace xc {
namespace tmapath {

constexpr int MB = 128;                     // grid rows per block (density kernel)
constexpr int NCW = 8;                      // consumer warps, 2 (m) x 4 (n); warp tile 64 x 8NF
constexpr int NCONS = NCW * 32;             // 256 consumer threads (<= 224 registers each)
constexpr int NTHREADS = NCONS + 32;        // + 1 producer warp
constexpr int D_STAGES = 5;                 // density pipeline depth
constexpr int V_STAGES = 2;                 // V pipeline depth (stages are 5 planes wide)
constexpr int VK = 16;                      // grid rows per V chunk
constexpr int A_TILE_BYTES = MB * 128;      // 128 rows x 16 doubles

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
template <int NF>
struct DensitySmem {
    static constexpr int NT = 32 * NF;
    static constexpr int B_TILE_BYTES = NT * 128;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int RED_OFF = D_STAGES * STAGE_BYTES;            // double red[128][4][4]
    static constexpr int BAR_OFF = RED_OFF + MB * 4 * 4 * 8;          // full[D_STAGES], empty[D_STAGES]
    static constexpr int ESUM_OFF = BAR_OFF + 2 * D_STAGES * 8;
    static constexpr int TOTAL = ESUM_OFF + 8 * 8 + 1024;             // + alignment slack
};

template <int NF, int NPL>
__global__ void __maxnreg__(224)
density_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_d,
                   int ngrid, int nao, int paired, int xc_mode, int nblocks, int ntiles, int nk,
                   const double* __restrict__ ao, const double* __restrict__ gx, const double* __restrict__ gy,
                   const double* __restrict__ gz, const double* __restrict__ w, double* __restrict__ coef,
                   double* __restrict__ exc_part) {
    using L = DensitySmem<NF>;
    constexpr int NT = L::NT;
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

    if (warp == NCW) {
        // ===================== TMA producer (one elected lane) =====================
        if (lane == 0) {
            tma::prefetch_map(&map_a);
            tma::prefetch_map(&map_d);
            uint32_t it = 0;
            for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
                for (int nt = 0; nt < ntiles; ++nt) {
                    for (int kc = 0; kc < nk; ++kc, ++it) {
                        const uint32_t s = it % D_STAGES, ph = (it / D_STAGES) & 1u;
                        tma::mbar_wait(&empty[s], ph ^ 1u);
                        unsigned char* st = sm + s * L::STAGE_BYTES;
                        tma::mbar_arrive_expect_tx(&full[s], L::STAGE_BYTES);
                        if (!paired) {
                            tma::load_2d(st, &map_a, kc * 16, blk * MB, &full[s]);
                        } else {  // even rows -> tile rows 0..63, odd rows -> 64..127
                            tma::load_2d(st, &map_a, kc * 16, blk * (MB / 2), &full[s]);
                            tma::load_2d(st + A_TILE_BYTES / 2, &map_a, nao + kc * 16, blk * (MB / 2), &full[s]);
                        }
                        tma::load_2d(st + A_TILE_BYTES, &map_d, kc * 16, nt * NT, &full[s]);
                    }
                }
            }
        }
        return;
    }

    // ===================== consumers: 8 warps, warp tile 64 x (8 NF) =====================
    const int wm = warp >> 2, wn = warp & 3;
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

    double e_acc = 0.0;
    uint32_t it = 0;
    for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
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
                    double a[8], b[NF];
#pragma unroll
                    for (int mf = 0; mf < 8; ++mf) a[mf] = lds_f64(a_base + mf * 1024 + koff[ks]);
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) b[nf] = lds_f64(b_base + nf * 1024 + koff[ks]);
#pragma unroll
                    for (int mf = 0; mf < 8; ++mf)
#pragma unroll
                        for (int nf = 0; nf < NF; ++nf) dmma::mma8x8x4(acc[mf][nf], a[mf], b[nf]);
                }
                __syncwarp();
                if (lane == 0) tma::mbar_arrive(&empty[s]);
            }
            // ---- fused epilogue: row-dots of C with Phi and grad Phi (global loads, L2-hot for Phi);
            //      the row sums of this column tile are folded into shared memory right away so that
            //      no row accumulator stays live across the k-loop (each red[] entry has one owner lane)
            const int nbase = nt * NT + wn * 8 * NF;
#pragma unroll
            for (int mf = 0; mf < 8; ++mf) {
                const int r = wm * 64 + mf * 8 + perm;
                const long g = (long)blk * MB + (paired ? (r < MB / 2 ? 2 * r : 2 * (r - MB / 2) + 1) : r);
                double rs[NPL];
#pragma unroll
                for (int p = 0; p < NPL; ++p) rs[p] = 0.0;
                if (g < ngrid) {
                    const size_t rowoff = (size_t)g * nao;
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int n = nbase + nf * 8 + ncol[e];
                            if (n >= nao) continue;
                            const double cv = acc[mf][nf][e];
                            rs[0] = fma(cv, __ldg(ao + rowoff + n), rs[0]);
                            if (NPL == 4) {
                                rs[1] = fma(cv, __ldg(gx + rowoff + n), rs[1]);
                                rs[2] = fma(cv, __ldg(gy + rowoff + n), rs[2]);
                                rs[3] = fma(cv, __ldg(gz + rowoff + n), rs[3]);
                            }
                        }
                    }
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
            const long g = (long)blk * MB + (paired ? (r < MB / 2 ? 2 * r : 2 * (r - MB / 2) + 1) : r);
            double2 c01 = make_double2(0.0, 0.0), c23 = make_double2(0.0, 0.0);
            if (g < ngrid) {
                const double* rr = red + r * 16;
                const double rho = (rr[0] + rr[4]) + (rr[8] + rr[12]);
                double dx = 0.0, dy = 0.0, dz = 0.0;
                if (NPL == 4) {
                    dx = 2.0 * ((rr[1] + rr[5]) + (rr[9] + rr[13]));
                    dy = 2.0 * ((rr[2] + rr[6]) + (rr[10] + rr[14]));
                    dz = 2.0 * ((rr[3] + rr[7]) + (rr[11] + rr[15]));
                }
                const xcfun::PointCoef pc = eval_mode(xc_mode, rho, dx, dy, dz, __ldg(w + g));
                c01 = make_double2(pc.a, pc.bx);
                c23 = make_double2(pc.by, pc.bz);
                e_acc += pc.exc;
            }
            // coef rows are padded to nblocks*128: rows beyond ngrid are written as zeros
            double2* cp = reinterpret_cast<double2*>(coef) + 2 * (size_t)g;
            cp[0] = c01;
            cp[1] = c23;
        }
        tma::named_bar_sync(1, NCONS);
    }
    // ---- per-CTA E_xc partial (fixed order)
    if (warp < 4) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e_acc += __shfl_xor_sync(0xffffffffu, e_acc, o);
        if (lane == 0) esum[warp] = e_acc;
    }
    tma::named_bar_sync(1, NCONS);
    if (tid == 0) exc_part[blockIdx.x] = (esum[0] + esum[1]) + (esum[2] + esum[3]);
}

// ------------------------------------------------------------------------------------------------
// V kernel
// ------------------------------------------------------------------------------------------------
template <int NF, int NPL>
struct VxcSmem {
    static constexpr int NT = 32 * NF;
    static constexpr int TILE_BYTES = VK * NT * 8;                       // one plane tile: 2NF boxes of 2 KB
    static constexpr int COEF_OFF = (NPL + 1) * TILE_BYTES;              // 16 x (a,bx,by,bz)
    static constexpr int STAGE_BYTES = COEF_OFF + 1024;
    static constexpr int BPITCH = NT + 4;                                // doubles; (NT+4) mod 16 == 4
    static constexpr int BS_OFF = V_STAGES * STAGE_BYTES;
    static constexpr int BS_BYTES = VK * BPITCH * 8;
    static constexpr int BAR_OFF = BS_OFF + 2 * BS_BYTES;
    static constexpr int TOTAL = BAR_OFF + 2 * V_STAGES * 8 + 1024;
};

template <int NF, int NPL>
__global__ void __maxnreg__(224)
vxc_tma_kernel(const __grid_constant__ CUtensorMap map_p0, const __grid_constant__ CUtensorMap map_px,
               const __grid_constant__ CUtensorMap map_py, const __grid_constant__ CUtensorMap map_pz,
               int ngrid, int nao, int paired, int ntiles, int lda_half, int rows_per_slice, int NP,
               const double* __restrict__ coef, double* __restrict__ vpart) {
    using L = VxcSmem<NF, NPL>;
    constexpr int NT = L::NT;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (tma::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - tma::smem_u32(smem_raw));
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + L::BAR_OFF);
    uint64_t* empty = full + V_STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // output tile of this CTA
    int tm, tn;
    if (lda_half) {  // upper-triangular tile pairs, row-major
        int t = blockIdx.x;
        tm = 0;
        while (t >= ntiles - tm) { t -= ntiles - tm; ++tm; }
        tn = tm + t;
    } else {
        tm = blockIdx.x / ntiles;
        tn = blockIdx.x % ntiles;
    }
    const bool diag = (tm == tn);
    const int m0 = tm * NT, n0 = tn * NT;
    const long gbeg = (long)blockIdx.y * rows_per_slice;
    const long gend = min((long)ngrid, gbeg + rows_per_slice);
    const int nchunks = gend > gbeg ? (int)((gend - gbeg + VK - 1) / VK) : 0;

    if (tid == 0) {
        for (int s = 0; s < V_STAGES; ++s) {
            tma::mbar_init(&full[s], 1);
            tma::mbar_init(&empty[s], NCW);
        }
        tma::fence_barrier_init();
    }
    __syncthreads();

    if (warp == NCW) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const CUtensorMap* maps[4] = {&map_p0, &map_px, &map_py, &map_pz};
            for (int p = 0; p < NPL; ++p) tma::prefetch_map(maps[p]);
            const uint32_t stage_tx = (uint32_t)((NPL + (diag ? 0 : 1)) * L::TILE_BYTES + VK * 32);
            for (int c = 0; c < nchunks; ++c) {
                const uint32_t s = c % V_STAGES, ph = (c / V_STAGES) & 1u;
                tma::mbar_wait(&empty[s], ph ^ 1u);
                unsigned char* st = sm + s * L::STAGE_BYTES;
                const long g0 = gbeg + (long)c * VK;
                tma::mbar_arrive_expect_tx(&full[s], stage_tx);
                for (int p = 0; p < NPL + (diag ? 0 : 1); ++p) {
                    const CUtensorMap* mp = (p < NPL) ? maps[p] : &map_p0;
                    const int col0 = (p < NPL) ? m0 : n0;
                    unsigned char* dst = st + (p < NPL ? p : NPL) * L::TILE_BYTES;
                    for (int b = 0; b < 2 * NF; ++b) {
                        if (!paired) {
                            tma::load_2d(dst + b * 2048, mp, col0 + 16 * b, (int)g0, &full[s]);
                        } else {
                            tma::load_2d(dst + b * 2048, mp, col0 + 16 * b, (int)(g0 >> 1), &full[s]);
                            tma::load_2d(dst + b * 2048 + 1024, mp, nao + col0 + 16 * b, (int)(g0 >> 1), &full[s]);
                        }
                    }
                }
                tma::load_1d(st + L::COEF_OFF, coef + 4 * g0, VK * 32, &full[s]);
            }
        }
        return;
    }

    // ===================== consumers: 2 x 4 warps, warp tile (16 NF) x (8 NF) =====================
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
    const uint32_t bs_base = base + L::BS_OFF;
    const uint32_t aoff = (uint32_t)((qcol * L::BPITCH + wm * 16 * NF + q) * 8);

    for (int c = 0; c < nchunks; ++c) {
        const uint32_t s = c % V_STAGES, ph = (c / V_STAGES) & 1u;
        const uint32_t st = base + s * L::STAGE_BYTES;
        const uint32_t bs = bs_base + (c & 1) * L::BS_BYTES;
        tma::mbar_wait(&full[s], ph);
        // ---- build B rows for this chunk: B = a Phi + bx dxPhi + by dyPhi + bz dzPhi
        for (int task = tid; task < 256 * NF; task += NCONS) {
            const int j = task & 7, rb = task >> 3;
            const int r = rb & 15, b = rb >> 4;
            const uint32_t off = (uint32_t)(b * 2048 + r * 128 + (((j ^ r) & 7) << 4));
            const int gi = paired ? (r < 8 ? 2 * r : 2 * (r - 8) + 1) : r;
            const double2 ca = lds_f64x2(st + L::COEF_OFF + gi * 32);
            const double2 v0 = lds_f64x2(st + off);
            double2 o = make_double2(ca.x * v0.x, ca.x * v0.y);
            if (NPL == 4) {
                const double2 cb = lds_f64x2(st + L::COEF_OFF + gi * 32 + 16);
                const double2 v1 = lds_f64x2(st + L::TILE_BYTES + off);
                const double2 v2 = lds_f64x2(st + 2 * L::TILE_BYTES + off);
                const double2 v3 = lds_f64x2(st + 3 * L::TILE_BYTES + off);
                o.x = fma(ca.y, v1.x, o.x); o.y = fma(ca.y, v1.y, o.y);
                o.x = fma(cb.x, v2.x, o.x); o.y = fma(cb.x, v2.y, o.y);
                o.x = fma(cb.y, v3.x, o.x); o.y = fma(cb.y, v3.y, o.y);
            }
            const int rho_idx = 8 * (r >> 3) + 4 * (r & 1) + ((r & 7) >> 1);  // MMA order of tile row r
            sts_f64x2(bs + (uint32_t)((rho_idx * L::BPITCH + b * 16 + 2 * j) * 8), o);
        }
        tma::named_bar_sync(1, NCONS);
        // ---- M += B^T Phi
        const uint32_t phin = st + (diag ? 0 : NPL * L::TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double a[MF], b[NF];
#pragma unroll
            for (int mf = 0; mf < MF; ++mf) a[mf] = lds_f64(bs + aoff + (uint32_t)((4 * ks * L::BPITCH + mf * 8) * 8));
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) b[nf] = lds_f64(phin + boff[ks][nf]);
#pragma unroll
            for (int mf = 0; mf < MF; ++mf)
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) dmma::mma8x8x4(acc[mf][nf], a[mf], b[nf]);
        }
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(&empty[s]);
    }
    // ---- partial tile out
    double* out = vpart + (size_t)blockIdx.y * NP * NP;
#pragma unroll
    for (int mf = 0; mf < MF; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
            const int r = m0 + wm * 16 * NF + mf * 8 + q;
            const int cc = n0 + wn * 8 * NF + nf * 8 + 2 * qcol;
            *reinterpret_cast<double2*>(out + (size_t)r * NP + cc) = make_double2(acc[mf][nf][0], acc[mf][nf][1]);
        }
}

// out[i][j] = sum_s (T_s(i,j) + T_s(j,i)), T = M where the tile was computed (lda_half: the mirror
// tile otherwise).  Fixed summation order -> bit-reproducible and exactly symmetric.
__global__ void finalize_tma_kernel(int nao, int NP, int NT, int nslices, int lda_half,
                                    const double* __restrict__ vpart, double* __restrict__ vxc, int nepart,
                                    const double* __restrict__ epart, double* __restrict__ d_exc) {
    const size_t n2 = (size_t)nao * nao;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n2) {
        const int i = (int)(idx / nao), j = (int)(idx % nao);
        size_t o1 = (size_t)i * NP + j, o2 = (size_t)j * NP + i;
        if (lda_half) {
            if (i / NT > j / NT) o1 = o2;
            else if (j / NT > i / NT) o2 = o1;
        }
        double s = 0.0;
        for (int sl = 0; sl < nslices; ++sl) {
            const double* p = vpart + (size_t)sl * NP * NP;
            s += p[o1] + p[o2];
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

__global__ void symmetrize_pad_tma_kernel(int nao, int ld, int rows, const double* __restrict__ dm,
                                          double* __restrict__ dsym) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= ld || i >= rows) return;
    double v = 0.0;
    if (i < nao && j < nao) v = 0.5 * (dm[(size_t)i * nao + j] + dm[(size_t)j * nao + i]);
    dsym[(size_t)i * ld + j] = v;
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

static bool make_plane_map(CUtensorMap* m, const double* ptr, int ngrid, int nao, bool paired, uint32_t box_rows) {
    if (!paired) return make_map(m, ptr, (uint64_t)nao, (uint64_t)ngrid, (uint64_t)nao, box_rows);
    return make_map(m, ptr, 2ull * nao, (uint64_t)ngrid / 2, 2ull * nao, box_rows);
}

template <int NF, int NPL>
static void launch(CublasHandleWrapper* ctx, const Problem& p, bool paired, int ntiles, int NP, int KP, int nsm,
                   double* dsym, double* coef, double* epart, int nblocks, int grid1, int xc_mode) {
    cudaStream_t st = ctx->stream;
    const int ngrid = p.ngrid, nao = p.nao;
    constexpr int NT = 32 * NF;
    CUtensorMap map_a, map_d, mp[4];
    bool ok = make_plane_map(&map_a, p.ao, ngrid, nao, paired, paired ? MB / 2 : MB);
    ok = ok && make_map(&map_d, dsym, (uint64_t)KP, (uint64_t)NP, (uint64_t)KP, NT);
    const double* planes[4] = {p.ao, p.gx, p.gy, p.gz};
    for (int i = 0; i < 4; ++i)
        ok = ok && make_plane_map(&mp[i], planes[i < NPL ? i : 0], ngrid, nao, paired, paired ? VK / 2 : VK);
    if (!ok) { ctx->failed = true; return; }

    using DL = DensitySmem<NF>;
    using VL = VxcSmem<NF, NPL>;
    auto dk = density_tma_kernel<NF, NPL>;
    auto vk = vxc_tma_kernel<NF, NPL>;
    DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(dk, cudaFuncAttributeMaxDynamicSharedMemorySize, DL::TOTAL));
    DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(vk, cudaFuncAttributeMaxDynamicSharedMemorySize, VL::TOTAL));

    if (ctx->timing) cudaEventRecord(ctx->ev[0], st);
    symmetrize_pad_tma_kernel<<<dim3((KP + 127) / 128, NP), 128, 0, st>>>(nao, KP, NP, p.dm, dsym);
    dk<<<grid1, NTHREADS, DL::TOTAL, st>>>(map_a, map_d, ngrid, nao, paired ? 1 : 0, xc_mode, nblocks, ntiles, KP / 16,
                                           p.ao, p.gx, p.gy, p.gz, p.w, coef, epart);
    if (ctx->timing) cudaEventRecord(ctx->ev[1], st);

    const int lda_half = (NPL == 1) ? 1 : 0;
    const int tiles = lda_half ? ntiles * (ntiles + 1) / 2 : ntiles * ntiles;
    int nslices = nsm / tiles;
    if (nslices < 1) nslices = 1;
    const int max_slices = (ngrid + VK - 1) / VK;
    if (nslices > max_slices) nslices = max_slices;
    int rows_per_slice = (ngrid + nslices - 1) / nslices;
    rows_per_slice = ((rows_per_slice + VK - 1) / VK) * VK;
    nslices = (ngrid + rows_per_slice - 1) / rows_per_slice;
    double* vpart = (double*)ctx->vpart.ensure(sizeof(double) * (size_t)nslices * NP * NP, &ctx->failed);
    if (ctx->failed) return;
    vk<<<dim3(tiles, nslices), NTHREADS, VL::TOTAL, st>>>(mp[0], mp[1], mp[2], mp[3], ngrid, nao, paired ? 1 : 0, ntiles,
                                                          lda_half, rows_per_slice, NP, coef, vpart);
    if (ctx->timing) cudaEventRecord(ctx->ev[2], st);
    const size_t n2 = (size_t)nao * nao;
    finalize_tma_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(nao, NP, NT, nslices, lda_half, vpart, p.vxc, grid1,
                                                                      epart, p.d_exc);
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
    if ((p.nao & 1) && (p.ngrid & 1)) return false;  // row pairs need an even number of rows
    if (p.nao > 128 * 16) return false;               // keep the padded D and partials modest
    return tmapath::encode_fn() != nullptr;
}

void run_tma(CublasHandleWrapper* ctx, const Problem& p) {
    using namespace tmapath;
    const int ngrid = p.ngrid, nao = p.nao;
    const bool paired = (nao & 1) != 0;
    const int ntiles = (nao + 127) / 128;
    const int NF = (nao + 32 * ntiles - 1) / (32 * ntiles);  // 1..4
    const int NT = 32 * NF, NP = ntiles * NT;
    const int KP = ((nao + 15) / 16) * 16;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    const int nblocks = (ngrid + MB - 1) / MB;
    const int grid1 = nblocks < nsm ? nblocks : nsm;

    double* dsym = (double*)ctx->dsym.ensure(sizeof(double) * (size_t)NP * KP, &ctx->failed);
    double* coef = (double*)ctx->coef.ensure(sizeof(double) * 4 * (size_t)nblocks * MB, &ctx->failed);
    double* epart = (double*)ctx->epart.ensure(sizeof(double) * grid1, &ctx->failed);
    if (ctx->failed) return;
    const int xc_mode = p.xc_type == 2 ? 4 : p.xc_type * 2 + (ctx->exact_functionals ? 1 : 0);

#define DFT_LAUNCH(NF_) \
    (p.xc_type == 0 ? launch<NF_, 1>(ctx, p, paired, ntiles, NP, KP, nsm, dsym, coef, epart, nblocks, grid1, xc_mode) \
                    : launch<NF_, 4>(ctx, p, paired, ntiles, NP, KP, nsm, dsym, coef, epart, nblocks, grid1, xc_mode))
    switch (NF) {
        case 1: DFT_LAUNCH(1); break;
        case 2: DFT_LAUNCH(2); break;
        case 3: DFT_LAUNCH(3); break;
        default: DFT_LAUNCH(4); break;
    }
#undef DFT_LAUNCH
}

}  // namespace xc
