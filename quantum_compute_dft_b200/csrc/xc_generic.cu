// xc_generic.cu -- alignment-agnostic XC path (any nao, any pointer alignment).
//
// Same mathematics and the same FP64 tensor-core (DMMA) inner product as the TMA-fed path in
// xc_tma.cu, but operands are staged into padded shared-memory tiles with ordinary guarded
// loads, so it accepts inputs the TMA engine cannot address (row pitch or plane offsets that
// are not multiples of 16 bytes).  It is the engine's fallback and the first correct CUDA path.
//
// What of the reference it replaces (file:line into /root/reference/src/dft_solver.cu):
//   get_rho_kernel :294-307, get_rho_sigma_kernel_planar :346-380
//        -> density_kernel: C = Phi_blk . Dsym as DMMA tiles, rho/grad-rho as row-dots fused in
//           the epilogue (C is never stored), functional evaluated once per point in the same
//           kernel (replaces both passes of lda/gga/b3lyp_fused_kernel :309-513)
//   reduce_sum_kernel :285-292 (65 536 same-address atomics)
//        -> per-CTA partial sums + fixed-order final sum (deterministic)
//   B matrix in global memory (:577,:613,:655) + cublasDgemm (:580,:616,:663)
//        -> vxc_kernel: B rows are formed on the fly in shared memory from (a,b) coefficients,
//           M = B^T Phi accumulated as DMMA tiles, split over grid-row slices
//   symmetrize_matrix_kernel :515-527 -> folded into finalize_kernel (out = M + M^T)
#include "dmma.cuh"
#include "engine.h"
#include "xc_functionals.cuh"

namespace xc {
namespace generic {

constexpr int MB = 64;         // grid rows per CTA in the density kernel
constexpr int NT = 64;         // AO columns per tile
constexpr int KC = 16;         // reduction chunk
constexpr int APITCH = KC + 4; // 20 doubles: (r*20 + c) mod 16 distinct for r,c in 0..3 -> conflict-free
constexpr int VPITCH = NT + 4; // 68 doubles: (k*68 + m) mod 16 = 4k + m -> conflict-free
constexpr int THREADS = 256;

__global__ void symmetrize_pad_kernel(int nao, int ld, int rows, const double* __restrict__ dm,
                                      double* __restrict__ dsym) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= ld || i >= rows) return;
    double v = 0.0;
    if (i < nao && j < nao) v = 0.5 * (dm[(size_t)i * nao + j] + dm[(size_t)j * nao + i]);
    dsym[(size_t)i * ld + j] = v;
}

template <int XC, bool EXACT>
__global__ void __launch_bounds__(THREADS, 2)
density_kernel(int ngrid, int nao, int ld, const double* __restrict__ dsym,
               const double* __restrict__ ao, const double* __restrict__ gx,
               const double* __restrict__ gy, const double* __restrict__ gz,
               const double* __restrict__ w, double* __restrict__ coef,
               double* __restrict__ exc_part) {
    constexpr int NPL = (XC == 0) ? 1 : 4;
    __shared__ double As[MB * APITCH];
    __shared__ double Bs[NT * APITCH];
    __shared__ double red[MB][2][NPL];
    __shared__ double esum[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int qrow = lane >> 2, qcol = lane & 3;
    const long g0 = (long)blockIdx.x * MB;
    const int ntiles = (nao + NT - 1) / NT, nk = (nao + KC - 1) / KC;

    double racc[2][NPL];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int p = 0; p < NPL; ++p) racc[mf][p] = 0.0;

    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    for (int nt = 0; nt < ntiles; ++nt) {
        double acc[2][4][2];
#pragma unroll
        for (int mf = 0; mf < 2; ++mf)
#pragma unroll
            for (int nf = 0; nf < 4; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;

        for (int kc = 0; kc < nk; ++kc) {
            __syncthreads();
            {
                const long g = g0 + lrow;
                const int k = kc * KC + lk;
                const double* src = ao + (size_t)g * nao + k;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    As[lrow * APITCH + lk + i] = (g < ngrid && k + i < nao) ? __ldg(src + i) : 0.0;
                const double2* bsrc =
                    reinterpret_cast<const double2*>(dsym + (size_t)(nt * NT + lrow) * ld + kc * KC + lk);
                const double2 b0 = __ldg(bsrc), b1 = __ldg(bsrc + 1);
                Bs[lrow * APITCH + lk + 0] = b0.x;
                Bs[lrow * APITCH + lk + 1] = b0.y;
                Bs[lrow * APITCH + lk + 2] = b1.x;
                Bs[lrow * APITCH + lk + 3] = b1.y;
            }
            __syncthreads();
#pragma unroll
            for (int ks = 0; ks < KC / 4; ++ks) {
                double a[2], b[4];
#pragma unroll
                for (int mf = 0; mf < 2; ++mf)
                    a[mf] = As[(wm * 16 + mf * 8 + qrow) * APITCH + ks * 4 + qcol];
#pragma unroll
                for (int nf = 0; nf < 4; ++nf)
                    b[nf] = Bs[(wn * 32 + nf * 8 + qrow) * APITCH + ks * 4 + qcol];
#pragma unroll
                for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                    for (int nf = 0; nf < 4; ++nf) dmma::mma8x8x4(acc[mf][nf], a[mf], b[nf]);
            }
        }
        // row-dot epilogue for this column tile: rho += C.phi, grad += C.dphi
#pragma unroll
        for (int mf = 0; mf < 2; ++mf) {
            const long g = g0 + wm * 16 + mf * 8 + qrow;
            if (g >= ngrid) continue;
#pragma unroll
            for (int nf = 0; nf < 4; ++nf) {
                const int c = nt * NT + wn * 32 + nf * 8 + 2 * qcol;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (c + e >= nao) continue;
                    const size_t idx = (size_t)g * nao + c + e;
                    const double cv = acc[mf][nf][e];
                    racc[mf][0] = fma(cv, __ldg(ao + idx), racc[mf][0]);
                    if (NPL == 4) {
                        racc[mf][1] = fma(cv, __ldg(gx + idx), racc[mf][1]);
                        racc[mf][2] = fma(cv, __ldg(gy + idx), racc[mf][2]);
                        racc[mf][3] = fma(cv, __ldg(gz + idx), racc[mf][3]);
                    }
                }
            }
        }
    }
    // quad reduce, then across the two column-warps through shared memory
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int p = 0; p < NPL; ++p) {
            double v = racc[mf][p];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (qcol == 0) red[wm * 16 + mf * 8 + qrow][wn][p] = v;
        }
    __syncthreads();
    double e = 0.0;
    if (tid < MB) {
        const long g = g0 + tid;
        if (g < ngrid) {
            const double rho = red[tid][0][0] + red[tid][1][0];
            double dx = 0.0, dy = 0.0, dz = 0.0;
            if (NPL == 4) {
                dx = 2.0 * (red[tid][0][1] + red[tid][1][1]);
                dy = 2.0 * (red[tid][0][2] + red[tid][1][2]);
                dz = 2.0 * (red[tid][0][3] + red[tid][1][3]);
            }
            const xcfun::PointCoef pc = xcfun::evaluate_point<XC, EXACT>(rho, dx, dy, dz, w[g]);
            double4 c4 = make_double4(pc.a, pc.bx, pc.by, pc.bz);
            reinterpret_cast<double4*>(coef)[g] = c4;
            e = pc.exc;
        }
    }
    if (warp < 2) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if (lane == 0) esum[warp] = e;
    }
    __syncthreads();
    if (tid == 0) exc_part[blockIdx.x] = esum[0] + esum[1];
}

template <int NPL>
__global__ void __launch_bounds__(THREADS, 2)
vxc_kernel(int ngrid, int nao, int tiles_n, int rows_per_slice, int NP,
           const double* __restrict__ ao, const double* __restrict__ gx,
           const double* __restrict__ gy, const double* __restrict__ gz,
           const double* __restrict__ coef, double* __restrict__ vpart) {
    __shared__ double Bs[KC * VPITCH];
    __shared__ double Ps[KC * VPITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int qrow = lane >> 2, qcol = lane & 3;
    const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
    const int m0 = tm * NT, n0 = tn * NT;
    const long gbeg = (long)blockIdx.y * rows_per_slice;
    const long gend = min((long)ngrid, gbeg + rows_per_slice);

    double acc[2][4][2];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;

    const int lk = tid >> 4, lc = (tid & 15) * 4;
    for (long gc = gbeg; gc < gend; gc += KC) {
        __syncthreads();
        {
            const long g = gc + lk;
            const bool valid = g < gend;
            double4 c4 = make_double4(0.0, 0.0, 0.0, 0.0);
            if (valid) {
                const double2 lo = __ldg(reinterpret_cast<const double2*>(coef) + 2 * g);
                const double2 hi = __ldg(reinterpret_cast<const double2*>(coef) + 2 * g + 1);
                c4 = make_double4(lo.x, lo.y, hi.x, hi.y);
            }
            const size_t row = (size_t)g * nao;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int jm = m0 + lc + i, jn = n0 + lc + i;
                double bv = 0.0, pv = 0.0;
                if (valid && jm < nao) {
                    bv = c4.x * __ldg(ao + row + jm);
                    if (NPL == 4) {
                        bv = fma(c4.y, __ldg(gx + row + jm), bv);
                        bv = fma(c4.z, __ldg(gy + row + jm), bv);
                        bv = fma(c4.w, __ldg(gz + row + jm), bv);
                    }
                }
                if (valid && jn < nao) pv = __ldg(ao + row + jn);
                Bs[lk * VPITCH + lc + i] = bv;
                Ps[lk * VPITCH + lc + i] = pv;
            }
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < KC / 4; ++ks) {
            double a[2], b[4];
#pragma unroll
            for (int mf = 0; mf < 2; ++mf)
                a[mf] = Bs[(ks * 4 + qcol) * VPITCH + wm * 16 + mf * 8 + qrow];
#pragma unroll
            for (int nf = 0; nf < 4; ++nf)
                b[nf] = Ps[(ks * 4 + qcol) * VPITCH + wn * 32 + nf * 8 + qrow];
#pragma unroll
            for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                for (int nf = 0; nf < 4; ++nf) dmma::mma8x8x4(acc[mf][nf], a[mf], b[nf]);
        }
    }
    double* out = vpart + (size_t)blockIdx.y * NP * NP;
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) {
            const int r = m0 + wm * 16 + mf * 8 + qrow;
            const int c = n0 + wn * 32 + nf * 8 + 2 * qcol;
            *reinterpret_cast<double2*>(out + (size_t)r * NP + c) =
                make_double2(acc[mf][nf][0], acc[mf][nf][1]);
        }
}

// out[i][j] = sum_s (M_s[i][j] + M_s[j][i])  (fixed order -> bit-reproducible and exactly
// symmetric); block 0 additionally reduces the per-CTA E_xc partials in a fixed order.
__global__ void finalize_kernel(int nao, int NP, int nslices, int raw, const double* __restrict__ vpart,
                                double* __restrict__ vxc, int nepart,
                                const double* __restrict__ epart, double* __restrict__ d_exc) {
    const size_t n2 = (size_t)nao * nao;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n2) {
        const int i = (int)(idx / nao), j = (int)(idx % nao);
        double s = 0.0;
        for (int sl = 0; sl < nslices; ++sl) {
            const double* p = vpart + (size_t)sl * NP * NP;
            s += raw ? 2.0 * p[(size_t)i * NP + j] : p[(size_t)i * NP + j] + p[(size_t)j * NP + i];
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

template <int XC>
static void launch_density(CublasHandleWrapper* ctx, const Problem& p, int ld, const double* dsym,
                           double* coef, double* epart, int nblk) {
    if (ctx->exact_functionals)
        density_kernel<XC, true><<<nblk, THREADS, 0, ctx->stream>>>(p.ngrid, p.nao, ld, dsym, p.ao, p.gx, p.gy,
                                                                    p.gz, p.w, coef, epart);
    else
        density_kernel<XC, false><<<nblk, THREADS, 0, ctx->stream>>>(p.ngrid, p.nao, ld, dsym, p.ao, p.gx, p.gy,
                                                                     p.gz, p.w, coef, epart);
}

}  // namespace generic

void run_generic(CublasHandleWrapper* ctx, const Problem& p) {
    using namespace generic;
    const int ngrid = p.ngrid, nao = p.nao;
    const int NP = ((nao + NT - 1) / NT) * NT;  // padded matrix dimension (multiple of 64)
    const int tiles = NP / NT;
    const int nblk = (ngrid + MB - 1) / MB;

    int nslices = (148 * 4) / (tiles * tiles);
    const int max_slices = (ngrid + KC * 4 - 1) / (KC * 4);
    if (nslices > max_slices) nslices = max_slices;
    if (nslices < 1) nslices = 1;
    int rows_per_slice = (ngrid + nslices - 1) / nslices;
    rows_per_slice = ((rows_per_slice + KC - 1) / KC) * KC;
    nslices = (ngrid + rows_per_slice - 1) / rows_per_slice;

    double* dsym = (double*)ctx->dsym.ensure(sizeof(double) * NP * NP, &ctx->failed);
    double* coef = (double*)ctx->coef.ensure(sizeof(double) * 4 * (size_t)ngrid, &ctx->failed);
    double* epart = (double*)ctx->epart.ensure(sizeof(double) * nblk, &ctx->failed);
    double* vpart = (double*)ctx->vpart.ensure(sizeof(double) * (size_t)nslices * NP * NP, &ctx->failed);
    if (ctx->failed) return;
    cudaStream_t st = ctx->stream;

    if (ctx->timing) cudaEventRecord(ctx->ev[0], st);
    symmetrize_pad_kernel<<<dim3((NP + 127) / 128, NP), 128, 0, st>>>(nao, NP, NP, p.dm, dsym);
    if (p.xc_type == 0) launch_density<0>(ctx, p, NP, dsym, coef, epart, nblk);
    else if (p.xc_type == 1) launch_density<1>(ctx, p, NP, dsym, coef, epart, nblk);
    else launch_density<2>(ctx, p, NP, dsym, coef, epart, nblk);
    if (ctx->timing) cudaEventRecord(ctx->ev[1], st);
    if (p.xc_type == 0)
        vxc_kernel<1><<<dim3(tiles * tiles, nslices), THREADS, 0, st>>>(ngrid, nao, tiles, rows_per_slice, NP, p.ao,
                                                                        nullptr, nullptr, nullptr, coef, vpart);
    else
        vxc_kernel<4><<<dim3(tiles * tiles, nslices), THREADS, 0, st>>>(ngrid, nao, tiles, rows_per_slice, NP, p.ao,
                                                                        p.gx, p.gy, p.gz, coef, vpart);
    if (ctx->timing) cudaEventRecord(ctx->ev[2], st);
    const size_t n2 = (size_t)nao * nao;
    finalize_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(nao, NP, nslices, (ctx->raw_convention && p.xc_type == 1) ? 1 : 0, vpart, p.vxc, nblk, epart,
                                                                  p.d_exc);
    if (ctx->timing) cudaEventRecord(ctx->ev[3], st);
    ctx->stats.launches = 4;
    ctx->stats.path = PATH_GENERIC;
    DFT_CUDA_CHECK(ctx, cudaGetLastError());
}

}  // namespace xc
