// linalg.cu -- the two dense helpers the reference delegates to cuBLAS, hand-written.
//
//   coulomb_gemv   : J = ERI . vec(D), the cublasDgemv at dft_solver.cu:550-555 (column-major,
//                    no transpose, N2 x N2 with N2 = nao^2).  HBM-bound: every ERI element is
//                    read exactly once (8*nao^4 bytes), coalesced along the contiguous index,
//                    split over column chunks for parallelism and reduced in a fixed order.
//   dgemm_colmajor : definition behind XCSolver::safe_cublas_dgemm (dft_solver.h:25-27,
//                    dft_solver.cu:541-548) so code written against the reference header links.
//                    Not on the engine's own hot path (V_xc is built by the fused kernels).
#include <cstdint>
#include <cstdio>

#include "engine.h"

namespace xc {
namespace {

constexpr int GEMV_THREADS = 256;

// partial[s][r] = sum_{c in chunk s} A[c*N2 + r] * x[c]
// VEC = 2: a thread owns two adjacent rows and streams 16-byte loads (N2 even: every column start is
// 16-byte aligned); 8 columns' loads are issued before the first FMA so that 128 bytes per thread are in
// flight.  ld.global.nc + no reuse: the ERI passes through L2 once.
template <int VEC>
__global__ void __launch_bounds__(GEMV_THREADS)
gemv_partial_kernel(long N2, int cols_per_chunk, const double* __restrict__ A, const double* __restrict__ x,
                    double* __restrict__ partial) {
    const long r = ((long)blockIdx.x * GEMV_THREADS + threadIdx.x) * VEC;
    const long c0 = (long)blockIdx.y * cols_per_chunk;
    const long c1 = min(N2, c0 + cols_per_chunk);
    if (r >= N2) return;
    constexpr int U = 8;
    double acc[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[u][v] = 0.0;
    long c = c0;
    for (; c + U <= c1; c += U) {
        double a[U][VEC], xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double* p = A + (size_t)(c + u) * N2 + r;
            if (VEC == 2) {
                const double2 t = __ldg(reinterpret_cast<const double2*>(p));
                a[u][0] = t.x; a[u][VEC - 1] = t.y;
            } else {
                a[u][0] = __ldg(p);
            }
            xv[u] = __ldg(x + c + u);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[u][v] = fma(a[u][v], xv[u], acc[u][v]);
    }
    for (; c < c1; ++c) {
        const double xc = __ldg(x + c);
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[0][v] = fma(__ldg(A + (size_t)c * N2 + r + v), xc, acc[0][v]);
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const double s = ((acc[0][v] + acc[1][v]) + (acc[2][v] + acc[3][v])) + ((acc[4][v] + acc[5][v]) + (acc[6][v] + acc[7][v]));
        partial[(size_t)blockIdx.y * N2 + r + v] = s;
    }
}

__global__ void gemv_reduce_kernel(long N2, int nchunks, const double* __restrict__ partial,
                                   double* __restrict__ y) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N2) return;
    double s = 0.0;
    for (int k = 0; k < nchunks; ++k) s += partial[(size_t)k * N2 + r];
    y[r] = s;
}

// C(m x n) = op(A) op(B), column-major, alpha = 1, beta = 0
__global__ void gemm_simple_kernel(bool ta, bool tb, int m, int n, int k, const double* __restrict__ A, int lda,
                                   const double* __restrict__ B, int ldb, double* __restrict__ C, int ldc) {
    __shared__ double As[16][17], Bs[16][17];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int row = blockIdx.x * 16 + tx, col = blockIdx.y * 16 + ty;
    double acc = 0.0;
    for (int k0 = 0; k0 < k; k0 += 16) {
        const int ka = k0 + ty, kb = k0 + tx;
        As[ty][tx] = (row < m && ka < k) ? (ta ? A[(size_t)row * lda + ka] : A[(size_t)ka * lda + row]) : 0.0;
        Bs[tx][ty] = (kb < k && col < n) ? (tb ? B[(size_t)kb * ldb + col] : B[(size_t)col * ldb + kb]) : 0.0;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) acc = fma(As[kk][tx], Bs[kk][ty], acc);
        __syncthreads();
    }
    if (row < m && col < n) C[(size_t)col * ldc + row] = acc;
}

// ---- fused Coulomb + exchange: one pass over the ERI -------------------------------------------------
// J[i,j] = sum_kl (ij|kl) D[k,l]   (the gemv above)          K[i,k] = sum_jl (ij|kl) D[j,l]
// K is what the reference's driver computes for B3LYP with cupy.einsum('ijkl,jl->ik', eri, dm)
// (dft.py:218), a second full pass over the 8 nao^4-byte ERI right after the J gemv.  Here every ERI
// element is loaded once and used twice.  CTA (i, chunk of KC values of k): warp w takes rows j = w, w+8, ..;
// row (i,j), columns (k0..k0+KC, all l) are KC*nao contiguous doubles.  Per lane: accJ (this row, reduced
// over the warp once per row) and accK[KC] (summed over this warp's rows, reduced once at the end);
// every J/K element has one owner and a fixed summation order -> bit-reproducible, no atomics.
constexpr int JK_KC = 8;
constexpr int JK_WARPS = 8;

__global__ void __launch_bounds__(JK_WARPS * 32)
coulomb_exchange_kernel(int n, const double* __restrict__ A, const double* __restrict__ D, double* __restrict__ jpart,
                        double* __restrict__ K) {
    extern __shared__ double sm[];
    double* s_dk = sm;                       // [KC][n]: rows k0.. of D
    double* s_k = sm + JK_KC * n;            // [JK_WARPS][KC]: per-warp K partials
    const int i = blockIdx.x, chunk = blockIdx.y, k0 = chunk * JK_KC;
    const int kc = min(JK_KC, n - k0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t n2 = (size_t)n * n;
    for (int t = threadIdx.x; t < JK_KC * n; t += blockDim.x) s_dk[t] = (t / n) < kc ? D[(size_t)(k0 + t / n) * n + t % n] : 0.0;
    __syncthreads();
    double acck[JK_KC];
#pragma unroll
    for (int kk = 0; kk < JK_KC; ++kk) acck[kk] = 0.0;
    for (int j = warp; j < n; j += JK_WARPS) {
        const double* row = A + ((size_t)i * n + j) * n2 + (size_t)k0 * n;
        const double* dj = D + (size_t)j * n;
        double accj = 0.0;
        // 4 x KC = 32 loads per lane are issued before the first FMA (the loop is latency-bound otherwise)
        for (int l0 = lane; l0 < n; l0 += 128) {
            double a[4][JK_KC], djl[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int l = l0 + 32 * u;
                const bool ok = l < n;
                djl[u] = ok ? __ldg(dj + l) : 0.0;
#pragma unroll
                for (int kk = 0; kk < JK_KC; ++kk) a[u][kk] = (ok && kk < kc) ? __ldg(row + (size_t)kk * n + l) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int l = min(l0 + 32 * u, n - 1);
#pragma unroll
                for (int kk = 0; kk < JK_KC; ++kk) {
                    accj = fma(a[u][kk], s_dk[kk * n + l], accj);
                    acck[kk] = fma(a[u][kk], djl[u], acck[kk]);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) accj += __shfl_xor_sync(0xffffffffu, accj, o);
        if (lane == 0) jpart[(size_t)chunk * n2 + (size_t)i * n + j] = accj;
    }
#pragma unroll
    for (int kk = 0; kk < JK_KC; ++kk) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acck[kk] += __shfl_xor_sync(0xffffffffu, acck[kk], o);
        if (lane == 0) s_k[warp * JK_KC + kk] = acck[kk];
    }
    __syncthreads();
    if (threadIdx.x < kc) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < JK_WARPS; ++w) t += s_k[w * JK_KC + threadIdx.x];
        K[(size_t)i * n + k0 + threadIdx.x] = t;
    }
}

// ---- device-resident Fock assembly and SCF energy terms ------------------------
// F = Hcore + J + 1/2 (V + V^T) - 1/2 c_hf K            (dft.py:212, :221-223)
__global__ void fock_kernel(int n, const double* __restrict__ h, const double* __restrict__ J,
                            const double* __restrict__ v, const double* __restrict__ K, double c_hf,
                            double* __restrict__ F) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const int i = idx / n, j = idx % n;
    double f = h[idx] + J[idx] + 0.5 * (v[idx] + v[(size_t)j * n + i]);
    if (K) f -= 0.5 * c_hf * K[idx];
    F[idx] = f;
}

// out[0] = sum D o Hcore, out[1] = 1/2 sum D o J, out[2] = -1/4 c_hf sum D o K   (dft.py:230-236)
// one CTA, fixed-order tree: bit-reproducible
__global__ void scf_energy_kernel(int n2, const double* __restrict__ D, const double* __restrict__ h,
                                  const double* __restrict__ J, const double* __restrict__ K, double c_hf,
                                  double* __restrict__ out) {
    __shared__ double sh[3][256];
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < n2; i += 256) {
        const double d = D[i];
        a = fma(d, h[i], a);
        b = fma(d, J[i], b);
        if (K) c = fma(d, K[i], c);
    }
    sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b; sh[2][threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int t = 0; t < 3; ++t) sh[t][threadIdx.x] += sh[t][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = sh[0][0];
        out[1] = 0.5 * sh[1][0];
        out[2] = -0.25 * c_hf * sh[2][0];
    }
}

}  // namespace

void coulomb_gemv(CublasHandleWrapper* ctx, int nao, const double* eri, const double* dm, double* J) {
    if (!ctx || nao <= 0 || !eri || !dm || !J) return;
    const long N2 = (long)nao * nao;
    const bool vec2 = (N2 % 2 == 0) && ((reinterpret_cast<uintptr_t>(eri) & 15u) == 0);
    const long rows_per_block = (long)GEMV_THREADS * (vec2 ? 2 : 1);
    const int rblocks = (int)((N2 + rows_per_block - 1) / rows_per_block);
    // enough CTAs to saturate HBM: aim for >= 8 resident CTAs per SM
    if (ctx->num_sms <= 0) cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, ctx->device);
    int nchunks = (ctx->num_sms * 8 + rblocks - 1) / rblocks;
    if (nchunks > N2 / 8) nchunks = (int)(N2 / 8);
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 65535) nchunks = 65535;
    int cols = (int)((N2 + nchunks - 1) / nchunks);
    cols = ((cols + 7) / 8) * 8;  // whole unrolled groups
    nchunks = (int)((N2 + cols - 1) / cols);
    double* partial = (double*)ctx->vpart.ensure(sizeof(double) * (size_t)nchunks * N2, &ctx->failed);
    if (ctx->failed) return;
    if (vec2) gemv_partial_kernel<2><<<dim3(rblocks, nchunks), GEMV_THREADS, 0, ctx->stream>>>(N2, cols, eri, dm, partial);
    else gemv_partial_kernel<1><<<dim3(rblocks, nchunks), GEMV_THREADS, 0, ctx->stream>>>(N2, cols, eri, dm, partial);
    gemv_reduce_kernel<<<(int)((N2 + GEMV_THREADS - 1) / GEMV_THREADS), GEMV_THREADS, 0, ctx->stream>>>(N2, nchunks, partial, J);
    DFT_CUDA_CHECK(ctx, cudaGetLastError());
}

void coulomb_exchange(CublasHandleWrapper* ctx, int nao, const double* eri, const double* dm, double* J, double* K) {
    if (!ctx || nao <= 0 || !eri || !dm || !J || !K) return;
    const long N2 = (long)nao * nao;
    const int nchunks = (nao + JK_KC - 1) / JK_KC;
    double* partial = (double*)ctx->vpart.ensure(sizeof(double) * (size_t)nchunks * N2, &ctx->failed);
    if (ctx->failed) return;
    const size_t smem = sizeof(double) * (size_t)(JK_KC * nao + JK_WARPS * JK_KC);
    if (smem > 200 * 1024) { fprintf(stderr, "[dft_b200] coulomb_exchange: nao too large\n"); ctx->failed = true; return; }
    DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(coulomb_exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    coulomb_exchange_kernel<<<dim3(nao, nchunks), JK_WARPS * 32, smem, ctx->stream>>>(nao, eri, dm, partial, K);
    gemv_reduce_kernel<<<(int)((N2 + GEMV_THREADS - 1) / GEMV_THREADS), GEMV_THREADS, 0, ctx->stream>>>(N2, nchunks, partial, J);
    DFT_CUDA_CHECK(ctx, cudaGetLastError());
}

void build_fock(CublasHandleWrapper* ctx, int nao, const double* hcore, const double* J, const double* vxc,
                const double* K, double c_hf, double* F) {
    if (!ctx || nao <= 0 || !hcore || !J || !vxc || !F) return;
    const int n2 = nao * nao;
    fock_kernel<<<(n2 + 255) / 256, 256, 0, ctx->stream>>>(nao, hcore, J, vxc, K, c_hf, F);
    DFT_CUDA_CHECK(ctx, cudaGetLastError());
}

void scf_energies(CublasHandleWrapper* ctx, int nao, const double* dm, const double* hcore, const double* J,
                  const double* K, double c_hf, double* out3_host) {
    if (!ctx || nao <= 0 || !dm || !hcore || !J || !out3_host) return;
    double* d_out = (double*)ctx->result.ensure(sizeof(double) * ((size_t)nao * nao + 4), &ctx->failed);
    if (ctx->failed) return;
    scf_energy_kernel<<<1, 256, 0, ctx->stream>>>(nao * nao, dm, hcore, J, K, c_hf, d_out);
    DFT_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_scalar + 8, d_out, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DFT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int t = 0; t < 3; ++t) out3_host[t] = ctx->h_scalar[8 + t];
}

void dgemm_colmajor(CublasHandleWrapper* ctx, bool transA, bool transB, int m, int n, int k, const double* A,
                    int lda, const double* B, int ldb, double* C, int ldc) {
    if (!ctx || m <= 0 || n <= 0) return;
    dim3 grid((m + 15) / 16, (n + 15) / 16), block(16, 16);
    gemm_simple_kernel<<<grid, block, 0, ctx->stream>>>(transA, transB, m, n, k, A, lda, B, ldb, C, ldc);
    DFT_CUDA_CHECK(ctx, cudaGetLastError());
}

}  // namespace xc
