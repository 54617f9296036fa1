// ao_eval.cu -- atomic-orbital values and gradients on the grid (subsystem (a) of the north_star).
//
// Replaces, on the GPU, what the reference obtains on the host from PySCF
// (numint.eval_ao(mol, coords, deriv=0|1), grid.py:30,38) and then uploads (dft.py:155,172).
// Output layouts are exactly what DFT_ComputeXC consumes: ao (ngrid,nao) and the planar
// gradient (3,ngrid,nao).
//
// Design (HBM-bound: 8*ngrid*nao*P bytes written, P = 1 or 4):
//   * shell tables (centres, exponents, contraction coefficients) are staged once per CTA in
//     shared memory;
//   * one warp owns one grid point at a time: lanes evaluate different shells into a per-warp
//     shared-memory row (all P planes), then the warp streams the finished rows out with fully
//     coalesced stores (consecutive lanes -> consecutive AOs of the same point);
//   * primitives with exp*r^2 > cutoff are skipped (PySCF-like screening) which removes most of
//     the exp() work for core functions far from their atom and produces exact zeros that later
//     block screening can exploit.
#include <cstdlib>
#include <cstring>

#include "../../include/dft_b200_ext.h"
#include "engine.h"

namespace xc {
namespace ao {

struct Tables {
    const double* shell_xyz;
    const double* prim_exp;
    const double* prim_coef;
    const int* shell_meta;  // [nshell][4] = l, ao_off, prim_off, nprim
};

template <bool DERIV>
__global__ void __launch_bounds__(512)
eval_kernel(int ngrid, const double* __restrict__ coords, Tables t, int nshell, int nprim, int nao,
            double cutoff, double* __restrict__ ao, double* __restrict__ gout) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int P = DERIV ? 4 : 1;
    double* s_xyz = reinterpret_cast<double*>(smem_raw);
    double* s_exp = s_xyz + 3 * nshell;
    double* s_coef = s_exp + nprim;
    double* s_rows = s_coef + nprim;  // [nwarps][P][nao]
    const int nwarps = blockDim.x >> 5;
    int* s_meta = reinterpret_cast<int*>(s_rows + (size_t)nwarps * P * nao);

    for (int i = threadIdx.x; i < 3 * nshell; i += blockDim.x) s_xyz[i] = t.shell_xyz[i];
    for (int i = threadIdx.x; i < nprim; i += blockDim.x) {
        s_exp[i] = t.prim_exp[i];
        s_coef[i] = t.prim_coef[i];
    }
    for (int i = threadIdx.x; i < 4 * nshell; i += blockDim.x) s_meta[i] = t.shell_meta[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* row = s_rows + (size_t)warp * P * nao;
    const size_t plane = (size_t)ngrid * nao;
    const long stride = (long)gridDim.x * nwarps;
    for (long g = (long)blockIdx.x * nwarps + warp; g < ngrid; g += stride) {
        const double x = __ldg(coords + 3 * g), y = __ldg(coords + 3 * g + 1), z = __ldg(coords + 3 * g + 2);
        for (int s = lane; s < nshell; s += 32) {
            const double dx = x - s_xyz[3 * s], dy = y - s_xyz[3 * s + 1], dz = z - s_xyz[3 * s + 2];
            const double r2 = dx * dx + dy * dy + dz * dz;
            const int l = s_meta[4 * s], off = s_meta[4 * s + 1], p0 = s_meta[4 * s + 2], np = s_meta[4 * s + 3];
            double e0 = 0.0, e1 = 0.0;
            for (int k = p0; k < p0 + np; ++k) {
                const double a = s_exp[k];
                const double ar2 = a * r2;
                if (ar2 > cutoff) continue;
                const double v = s_coef[k] * exp(-ar2);
                e0 += v;
                e1 = fma(-2.0 * a, v, e1);
            }
            if (l == 0) {
                row[off] = e0;
                if (DERIV) {
                    row[nao + off] = e1 * dx;
                    row[2 * nao + off] = e1 * dy;
                    row[3 * nao + off] = e1 * dz;
                }
            } else {
                const double d[3] = {dx, dy, dz};
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    row[off + j] = d[j] * e0;
                    if (DERIV) {
                        const double dj1 = d[j] * e1;
                        row[nao + off + j] = fma(dj1, dx, j == 0 ? e0 : 0.0);
                        row[2 * nao + off + j] = fma(dj1, dy, j == 1 ? e0 : 0.0);
                        row[3 * nao + off + j] = fma(dj1, dz, j == 2 ? e0 : 0.0);
                    }
                }
            }
        }
        __syncwarp();
        double* dst = ao + (size_t)g * nao;
        for (int i = lane; i < nao; i += 32) dst[i] = row[i];
        if (DERIV) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double* gd = gout + c * plane + (size_t)g * nao;
                const double* src = row + (c + 1) * nao;
                for (int i = lane; i < nao; i += 32) gd[i] = src[i];
            }
        }
        __syncwarp();
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
    ctx->failed = false;
    if (exp_cutoff <= 0.0) exp_cutoff = 60.0;

    // pack the host tables into one staging block: xyz | exp | coef | meta
    const size_t nd = (size_t)3 * nshell + 2 * (size_t)nprim_total;
    const size_t bytes = nd * sizeof(double) + (size_t)4 * nshell * sizeof(int);
    unsigned char* h = (unsigned char*)malloc(bytes);
    if (!h) return 2;
    double* hd = reinterpret_cast<double*>(h);
    memcpy(hd, shell_xyz, sizeof(double) * 3 * nshell);
    memcpy(hd + 3 * nshell, prim_exp, sizeof(double) * nprim_total);
    memcpy(hd + 3 * nshell + nprim_total, prim_coef, sizeof(double) * nprim_total);
    int* hm = reinterpret_cast<int*>(hd + nd);
    for (int s = 0; s < nshell; ++s) {
        if (shell_l[s] < 0 || shell_l[s] > 1) { free(h); return 3; }  // s and p shells only
        hm[4 * s] = shell_l[s];
        hm[4 * s + 1] = shell_ao_off[s];
        hm[4 * s + 2] = shell_prim_off[s];
        hm[4 * s + 3] = shell_nprim[s];
    }
    unsigned char* d = (unsigned char*)ctx->scratch.ensure(bytes, &ctx->failed);
    if (ctx->failed) { free(h); return 4; }
    DFT_CUDA_CHECK(ctx, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    DFT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // h is pageable: copy is done now
    free(h);

    Tables t;
    t.shell_xyz = reinterpret_cast<const double*>(d);
    t.prim_exp = t.shell_xyz + 3 * nshell;
    t.prim_coef = t.prim_exp + nprim_total;
    t.shell_meta = reinterpret_cast<const int*>(t.shell_xyz + nd);

    const int P = deriv ? 4 : 1;
    const size_t table_bytes = nd * sizeof(double) + (size_t)4 * nshell * sizeof(int);
    const size_t row_bytes = (size_t)P * nao * sizeof(double);
    const size_t smem_max = 227 * 1024;
    if (table_bytes + row_bytes > smem_max) {
        fprintf(stderr, "[dft_b200] DFT_EvalAO: basis too large for the shared-memory staging (%zu B)\n",
                table_bytes + row_bytes);
        return 5;
    }
    int nwarps = (int)((smem_max - table_bytes) / row_bytes);
    if (nwarps > 16) nwarps = 16;
    const size_t smem = table_bytes + nwarps * row_bytes;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    int per_sm = (int)(smem_max / smem);
    if (per_sm < 1) per_sm = 1;
    if (per_sm * nwarps > 64) per_sm = 64 / nwarps;
    long want = ((long)ngrid + nwarps - 1) / nwarps;
    int grid = (int)(want < (long)nsm * per_sm ? want : (long)nsm * per_sm);
    double* ao_out = reinterpret_cast<double*>(d_ao_ptr);
    double* g_out = reinterpret_cast<double*>(d_ao_grad_ptr);
    const double* coords = reinterpret_cast<const double*>(d_coords_ptr);
    if (deriv) {
        DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_kernel<true><<<grid, nwarps * 32, smem, ctx->stream>>>(ngrid, coords, t, nshell, nprim_total, nao,
                                                                  exp_cutoff, ao_out, g_out);
    } else {
        DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_kernel<false><<<grid, nwarps * 32, smem, ctx->stream>>>(ngrid, coords, t, nshell, nprim_total, nao,
                                                                   exp_cutoff, ao_out, g_out);
    }
    DFT_CUDA_CHECK(ctx, cudaGetLastError());
    DFT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return ctx->failed ? 6 : 0;
}
