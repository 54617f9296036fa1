// xc_small.cu -- the small-basis XC path (nao <= 48: H2O 7, benzene 36): ONE pass over the AO planes.
//
// At these sizes the whole call is HBM / latency bound (4 nao^2 flop against 8 (P nao + 1) bytes per
// point puts the ridge near nao ~ 100), and the tensor-tile machinery of xc_tma.cu -- two full passes over the planes,
// five launches, 64- and 128-wide DMMA tiles that are mostly padding at nao 7 -- is the wrong tool.  Here a
// block of 32 grid points goes through EVERYTHING while it sits in shared memory, so every plane byte is read from
// HBM exactly once per XC build (the algorithmic figure of SURVEY.md 8d):
//
//   bulk copies (double-buffered)  tiles of 8 grid points, NPL planes x [8 rows][nao]: one cp.async.bulk per plane against
//                                  an mbarrier per warp and buffer (small nao: whole 32-point super-blocks, see the kernel)
//   step 1  C = Phi_tile . Dsym                       DMMA m8n8k4, a warp per tile
//   step 2  rho, grad rho / 2 = rowsum(C o plane)     from the accumulator fragments, quad shuffles
//   step 3  the functional, ONCE per point            per super-block of 4 tiles: all 32 lanes on distinct points
//   step 4  B = a Phi + b . grad Phi                  built in registers as DMMA A fragments
//   step 5  M += B^T Phi                              DMMA, K = the tile's 8 rows; each warp keeps a private
//                                                     (NP x NP) accumulator over all the tiles it sees
//   end     warp accumulators -> CTA partial (fixed order) -> global;  xc_small_finalize (programmatic dependent
//           launch: its launch latency overlaps the main kernel) sums the CTA partials in a fixed order, writes
//           M + M^T and E_xc.  Bit-reproducible, two launches per build.
//
// CTAs are 4 or 8 warps; several are resident per SM (2 x 4 warps at benzene/GGA, where a warp's two tile buffers are
// 18 KB) so that one warp's functional evaluation runs under another's tensor work.  Any alignment and any ngrid work:
// a ragged last tile, a plane base that is not 16-byte aligned fall back to 8-byte cp.async with zero fill (zero
// weight -> zero coefficients).
//
// Replaces, for small nao, the same reference code as xc_tma.cu: get_rho[_sigma]_kernel (dft_solver.cu:294-380),
// *_fused_kernel (:309-513), reduce_sum_kernel (:285-292), cublasDgemm (:580,:616,:663), symmetrize (:515-527).
#include <mutex>
#include <cstdio>
#include <cstring>

#include "dmma.cuh"
#include "engine.h"
#include "tma.cuh"
#include "xc_functionals.cuh"

namespace xc {
namespace smallpath {

constexpr int TR = 8;             // grid points per warp step (one DMMA row fragment)
constexpr int MAXW = 8;           // warps per CTA (fewer where shared memory does not hold 8 double-buffered tiles)

__device__ __forceinline__ xcfun::PointCoef eval_mode(int mode, double rho, double gx, double gy, double gz, double w) {
    switch (mode) {
        case 0: return xcfun::evaluate_point<0, false>(rho, gx, gy, gz, w);
        case 1: return xcfun::evaluate_point<0, true>(rho, gx, gy, gz, w);
        case 2: return xcfun::evaluate_point<1, false>(rho, gx, gy, gz, w);
        case 3: return xcfun::evaluate_point<1, true>(rho, gx, gy, gz, w);
        default: return xcfun::evaluate_point<2, false>(rho, gx, gy, gz, w);
    }
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// leading dimension of Dsym in shared memory: = 4 mod 16 doubles, so that the B-operand loads of step 1
// (4 consecutive k rows x 4 consecutive columns per half-warp) hit 16 distinct bank pairs
__host__ __device__ constexpr int ldd_for(int NP) { return ((NP + 11) / 16) * 16 + 4; }

struct SmallParams {
    int ngrid, nao, xc_mode, nblocks, vec16;
    int resident;    // a warp's buffer holds a whole super-block (4 tiles), loaded once and used by both passes (small nao)
    const double* dm;
    const double* plane[4];
    const double* w;
    double* vpart;   // [gridDim.x][NP * NP]
    double* epart;   // [gridDim.x]
};

// Every warp is autonomous: it streams its own tiles of 8 grid points (bulk copies, double-buffered, private shared
// memory), works on SUPER-BLOCKS of 4 tiles = 32 points without a single CTA barrier, and keeps a private (NP x NP)
// accumulator in registers:
//   pass A over the 4 tiles: steps 1 + 2; lane (q, qc) keeps the row sums of row q of tile qc
//   step 3: the functional, once, every lane its own point (a long dependent FP64 chain: done per tile, with the four
//           lanes of a row evaluating redundantly, it WAS the kernel's time at 8 warps per SM)
//   pass B over the same 4 tiles, streamed a second time (from L2: they were read a few microseconds ago): steps 4 + 5,
//           the coefficients of the points a lane supplies to the fragments arrive by shuffle
// HBM sees every plane byte once; rho, coefficients and E_xc never leave the registers.
// shared-memory layout (doubles): Dsym[NP][LDD] | per warp: 2 x planes[NPL][8 * nao] (resident mode: [NPL][32 * nao]) |
// 8 doubles of slack | per warp: 2 mbarriers
template <int NF, int NPL>
__global__ void __launch_bounds__(MAXW * 32)
xc_small_kernel(const SmallParams P) {
    constexpr int NP = 8 * NF, LDD = ldd_for(NP);
    extern __shared__ double smd[];
    // programmatic dependent launch: the finalize kernel may be scheduled now; it waits (griddepcontrol.wait) until this
    // whole grid has completed and its partials are visible, so only its launch latency overlaps this kernel
    asm volatile("griddepcontrol.launch_dependents;");
    const int nao = P.nao;
    const int tile_d = TR * nao;                         // doubles per plane tile (even)
    // RESIDENT mode (small nao: H2O): a buffer holds the four tiles of a whole super-block, plane by plane -- one bulk copy
    // per plane and super-block, double-buffered across super-blocks -- and pass B works from the same buffer.  The
    // streaming mode (a tile per buffer, pass B's tiles streamed a second time from L2) pays eight load latencies per
    // super-block in sequence, which at H2O size, where a warp sees one super-block, was a third of the kernel.
    const bool res = P.resident != 0;
    constexpr int SBT = 4;                               // tiles per super-block
    const int ps = res ? SBT * tile_d : tile_d;          // plane stride inside a buffer (doubles)
    const int buf_d = NPL * ps;                          // doubles per buffer
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int q = lane >> 2, qc = lane & 3;
    double* dsym = smd;
    double* buf0 = dsym + NP * LDD + (size_t)warp * 2 * buf_d;
    // one mbarrier per warp and buffer, behind the last warp's buffers and their slack: whole tiles arrive as four bulk
    // copies (one per plane, 64 nao contiguous bytes each) issued by lane 0 -- round 2's first version issued them as
    // 16-byte cp.async from all lanes, 18 per lane and tile plus their address arithmetic, and ncu put a quarter of the
    // kernel's samples there
    uint64_t* bars = reinterpret_cast<uint64_t*>(dsym + NP * LDD + (size_t)nwarp * 2 * buf_d + 8) + 2 * warp;
    if (lane == 0) { tma::mbar_init(&bars[0], 1); tma::mbar_init(&bars[1], 1); tma::fence_barrier_init(); }
    uint32_t bar_phase = 0;      // bit b: phase of buffer b's barrier
    uint32_t bar_bulk = 0;       // bit b: the load in flight into buffer b is a bulk copy (else cp.async)

    // ---- Dsym = 1/2 (D + D^T), zero-padded (replaces symmetrize_pad; nao^2 doubles from L2 per CTA).  The buffers
    // are cleared once: the fragment loads below read up to 7 columns past the end of a row (into the next row, the
    // next plane or the next buffer) instead of testing every column against nao.  What they pick up there is
    // multiplied by the zero padding of Dsym (steps 1, 2) or lands in rows / columns >= nao of the accumulator that
    // nobody reads (step 5) -- but it has to be FINITE, which stale tile data is and uninitialised memory is not.
    for (int i = tid; i < NP * LDD; i += blockDim.x) {
        const int r = i / LDD, c = i - r * LDD;
        double v = 0.0;
        if (r < nao && c < nao) v = 0.5 * (__ldg(P.dm + (size_t)r * nao + c) + __ldg(P.dm + (size_t)c * nao + r));
        dsym[i] = v;
    }
    for (int i = lane; i < 2 * buf_d + (warp == nwarp - 1 ? 8 : 0); i += 32) buf0[i] = 0.0;   // (+ the slack behind the last buffer)
    __syncwarp();

    // ---- asynchronous tile loads: 8 * nao contiguous doubles per plane; rows past the grid are zero-filled
    const uint32_t lane16 = (uint32_t)lane * 16u;
    auto issue_tile = [&](long g0, int b) {
        double* buf = buf0 + b * buf_d;
        const uint32_t sb = (uint32_t)__cvta_generic_to_shared(buf);
        if (P.vec16 && g0 + TR <= (long)P.ngrid) {   // the common case: whole tile, one bulk copy per plane
            const uint32_t bytes = (uint32_t)tile_d * 8u;
            if (lane == 0) {
                tma::mbar_arrive_expect_tx(&bars[b], NPL * bytes);
#pragma unroll
                for (int p = 0; p < NPL; ++p)
                    tma::load_1d(buf + (size_t)p * ps, P.plane[p] + g0 * nao, bytes, &bars[b]);
            }
            bar_bulk |= 1u << b;
        } else {
            long valid = ((long)P.ngrid - g0) * nao;                          // doubles of this tile that exist
            valid = valid < 0 ? 0 : (valid > tile_d ? tile_d : valid);
            for (int p = 0; p < NPL; ++p) {
                const double* src = P.plane[p] + (valid > 0 ? g0 * nao : 0);
                for (int i = lane; i < tile_d; i += 32) cp_async_8(sb + (uint32_t)(p * ps + i) * 8u, src + i, i < valid ? 8 : 0);
            }
            cp_async_commit();
            bar_bulk &= ~(1u << b);
        }
    };
    // resident mode: the whole super-block sblk (32 points = 4 tiles, contiguous per plane) into buffer b
    auto issue_super = [&](int sblk, int b) {
        double* buf = buf0 + b * buf_d;
        const long gs = (long)sblk * SBT * TR;
        if (P.vec16 && gs + SBT * TR <= (long)P.ngrid) {
            const uint32_t bytes = (uint32_t)(SBT * tile_d) * 8u;
            if (lane == 0) {
                tma::mbar_arrive_expect_tx(&bars[b], NPL * bytes);
#pragma unroll
                for (int p = 0; p < NPL; ++p) tma::load_1d(buf + (size_t)p * ps, P.plane[p] + gs * nao, bytes, &bars[b]);
            }
            bar_bulk |= 1u << b;
        } else {
            const uint32_t sb = (uint32_t)__cvta_generic_to_shared(buf);
            long valid = ((long)P.ngrid - gs) * nao;
            valid = valid < 0 ? 0 : (valid > (long)SBT * tile_d ? (long)SBT * tile_d : valid);
            for (int p = 0; p < NPL; ++p) {
                const double* src = P.plane[p] + (valid > 0 ? gs * nao : 0);
                for (int i = lane; i < SBT * tile_d; i += 32) cp_async_8(sb + (uint32_t)(p * ps + i) * 8u, src + i, i < valid ? 8 : 0);
            }
            cp_async_commit();
            bar_bulk &= ~(1u << b);
        }
    };
    // the load into buffer b has landed for every lane (and every lane is done with what it read before)
    auto wait_tile = [&](int b) {
        if (bar_bulk & (1u << b)) {
            tma::mbar_wait(&bars[b], (bar_phase >> b) & 1u);
            bar_phase ^= 1u << b;
        } else {
            cp_async_wait_all();
        }
        __syncwarp();
    };

    double acc[NF][NF][2];
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double e_acc = 0.0;
    const int nk = (nao + 3) / 4;    // k-steps of step 1 that contain real columns
    const int first = blockIdx.x * nwarp + warp, stride = gridDim.x * nwarp;
    constexpr int SB = 4;            // tiles per super-block
    const int nsuper = (P.nblocks + SB - 1) / SB;

    // L2 prefetch of a whole super-block (its 4 tiles are contiguous: 32 nao doubles per plane).  A warp has ONE tile
    // (9 KB at benzene size) in flight in shared memory, an SM 8 of them -- far fewer bytes than the HBM latency needs
    // (v4 of this kernel ran latency-bound at 6 us per tile and warp); prefetching the NEXT super-block into L2 while
    // the current one is processed costs no shared memory and turns its tile loads into L2 hits.
    auto prefetch_super = [&](int sblk) {
        const long gs = (long)sblk * SB * TR;
        long n = ((long)P.ngrid - gs) * nao;                               // doubles that exist
        n = n > (long)SB * tile_d ? (long)SB * tile_d : n;
        if (P.vec16 && n == (long)SB * tile_d) {
            if (lane < NPL) tma::prefetch_1d(P.plane[lane] + gs * nao, (uint32_t)(n * 8));
        } else {
            for (int p = 0; p < NPL; ++p) {
                const char* src = reinterpret_cast<const char*>(P.plane[p] + gs * nao);
                for (long o = (long)lane * 128; o < n * 8; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + o));
            }
        }
    };
    tma::fence_proxy_async();        // the cleared buffers (generic proxy) are ordered before the bulk copies (async proxy) into them
    __syncwarp();                    // (and the barriers are initialised before lane 0 arms one)
    if (first < nsuper) { if (res) issue_super(first, 0); else issue_tile((long)first * SB * TR, 0); }
    if (first + stride < nsuper) prefetch_super(first + stride);
    __syncthreads();                 // Dsym is complete and every buffer is cleared (the only CTA barrier before the final reduction)
    int it = 0;                      // tile loads consumed so far (selects the buffer)
    for (int sblk = first; sblk < nsuper; sblk += stride) {
        const long gs = (long)sblk * SB * TR;                          // first grid point of the super-block
        if (sblk + 2 * stride < nsuper) prefetch_super(sblk + 2 * stride);
        const double* sbuf = nullptr;
        if (res) {   // the whole super-block: wait for it, request the next one into the other buffer
            sbuf = buf0 + (it & 1) * buf_d;
            wait_tile(it & 1);
            if (sblk + stride < nsuper) issue_super(sblk + stride, (it + 1) & 1);
            ++it;
        }
        double keep[NPL];
#pragma unroll
        for (int p = 0; p < NPL; ++p) keep[p] = 0.0;
        // ---------------- pass A: C = Phi . Dsym and its row dots, tile by tile
#pragma unroll 1
        for (int m = 0; m < SB; ++m) {
            const double* phi;
            if (res) {
                phi = sbuf + m * tile_d;
            } else {
                phi = buf0 + (it & 1) * buf_d;
                wait_tile(it & 1);       // the tile has landed for every lane; every lane is done with the other buffer
                issue_tile(gs + (long)(m + 1 < SB ? m + 1 : 0) * TR, (it + 1) & 1);   // next of pass A, or pass B's first
                ++it;
            }
            const double* my_row = phi + (size_t)q * nao;             // fragment row of this lane
            double c[NF][2];
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) c[nf][0] = c[nf][1] = 0.0;
            for (int ks = 0; ks < nk; ++ks) {
                const int k = 4 * ks + qc;
                const double a = my_row[k];                           // (k >= nao: finite, times a zero row of Dsym)
                const double* drow = dsym + k * LDD + q;              // Dsym[k][8 nf + q]
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) dmma::mma8x8x4(c[nf], a, drow[8 * nf]);
            }
            // row sums of C o plane; lane holds columns 8 nf + 2 qc + {0, 1} of row q (columns >= nao: C is zero)
            double s[NPL];
#pragma unroll
            for (int p = 0; p < NPL; ++p) s[p] = 0.0;
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) {
                const int col = 8 * nf + 2 * qc;
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int p = 0; p < NPL; ++p) s[p] = fma(c[nf][e], my_row[p * ps + col + e], s[p]);
            }
#pragma unroll
            for (int p = 0; p < NPL; ++p) {
                s[p] += __shfl_xor_sync(0xffffffffu, s[p], 1);
                s[p] += __shfl_xor_sync(0xffffffffu, s[p], 2);
                if (m == qc) keep[p] = s[p];                          // lane (q, qc) keeps row q of tile qc
            }
        }
        // ---------------- step 3: the functional; lane (q, qc) evaluates point gs + 8 qc + q
        const long g = gs + 8 * qc + q;
        const double wgt = g < (long)P.ngrid ? __ldg(P.w + g) : 0.0;
        const xcfun::PointCoef pc = NPL == 4 ? eval_mode(P.xc_mode, keep[0], 2.0 * keep[1], 2.0 * keep[2], 2.0 * keep[3], wgt)
                                             : eval_mode(P.xc_mode, keep[0], 0.0, 0.0, 0.0, wgt);
        e_acc += pc.exc;
        // ---------------- pass B: M += B^T Phi, tile by tile (two k-steps of 4 points each)
#pragma unroll 1
        for (int m = 0; m < SB; ++m) {
            const double* phi;
            if (res) {
                phi = sbuf + m * tile_d;
            } else {
                phi = buf0 + (it & 1) * buf_d;
                wait_tile(it & 1);
                if (m + 1 < SB) issue_tile(gs + (long)(m + 1) * TR, (it + 1) & 1);
                else if (sblk + stride < nsuper) issue_tile((long)(sblk + stride) * SB * TR, (it + 1) & 1);
                ++it;
            }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int pt = 4 * ks + qc;                           // the row of this tile the lane supplies
                const int src = 4 * pt + m;                           // its coefficients live in lane (q = pt, qc = m)
                const double ca = __shfl_sync(0xffffffffu, pc.a, src);
                double cbx = 0.0, cby = 0.0, cbz = 0.0;
                if (NPL == 4) {
                    cbx = __shfl_sync(0xffffffffu, pc.bx, src);
                    cby = __shfl_sync(0xffffffffu, pc.by, src);
                    cbz = __shfl_sync(0xffffffffu, pc.bz, src);
                }
                const double* prow = phi + (size_t)pt * nao;
                double ph[NF], bb[NF];
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const int col = 8 * f + q;   // (columns >= nao end up in rows / columns of M that nobody reads)
                    ph[f] = prow[col];
                    double v = ca * ph[f];
                    if (NPL == 4) {
                        v = fma(cbx, prow[ps + col], v);
                        v = fma(cby, prow[2 * ps + col], v);
                        v = fma(cbz, prow[3 * ps + col], v);
                    }
                    bb[f] = v;
                }
#pragma unroll
                for (int mf = 0; mf < NF; ++mf)
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) dmma::mma8x8x4(acc[mf][nf], bb[mf], ph[nf]);
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- warp accumulators -> CTA partial, fixed order (warp 0 + warp 1 + ...), through Dsym's and the tiles' space
    double* red = smd;
    for (int w = 0; w < nwarp; ++w) {
        if (warp == w) {
#pragma unroll
            for (int mf = 0; mf < NF; ++mf)
#pragma unroll
                for (int nf = 0; nf < NF; ++nf)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double* d = red + (8 * mf + q) * NP + 8 * nf + 2 * qc + e;
                        *d = (w == 0 ? 0.0 : *d) + acc[mf][nf][e];
                    }
        }
        __syncthreads();
    }
    double* out = P.vpart + (size_t)blockIdx.x * NP * NP;
    for (int i = tid; i < NP * NP; i += blockDim.x) out[i] = red[i];
    // E_xc: lanes -> warps -> CTA in a fixed order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e_acc += __shfl_xor_sync(0xffffffffu, e_acc, o);
    __syncthreads();
    if (lane == 0) red[NP * NP + warp] = e_acc;
    __syncthreads();
    if (tid == 0) {
        double e = 0.0;
        for (int w = 0; w < nwarp; ++w) e += red[NP * NP + w];
        P.epart[blockIdx.x] = e;
    }
}

// out[i][j] = sum over CTAs of M[i][j] + M[j][i] (raw: 2 M[i][j]); E_xc = sum of the CTA partials.  One WARP per
// output element: the lanes stride over the CTA partials (all loads in flight at once -- a serial loop over 296
// partials took 90 us at benzene size), then a shuffle tree.  Fixed order -> bit-reproducible, exactly symmetric.
// `host_slot` (optional): two doubles of MAPPED PINNED host memory.  The last CTA to finish -- counted with a device
// atomic after every CTA has fenced its part of V_xc -- writes E_xc and then the call's sequence number there, so the
// blocking entry point can return E_xc the moment the whole result exists, without a D2H copy and a stream synchronise
// (at H2O size those two were a fifth of the call).
__global__ void __launch_bounds__(256)
xc_small_finalize(int nao, int NP, int ncta, int raw, const double* __restrict__ vpart, const double* __restrict__ epart,
                  double* __restrict__ vxc, double* __restrict__ d_exc, unsigned int* __restrict__ done_counter,
                  volatile double* host_slot, double seq) {
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * 8 + (threadIdx.x >> 5);
    asm volatile("griddepcontrol.wait;" ::: "memory");   // (a no-op when launched without the PDL attribute)
    if (idx < nao * nao) {
        const int i = idx / nao, j = idx - i * nao;
        const double* a = vpart + (size_t)i * NP + j;
        const double* b = raw ? a : vpart + (size_t)j * NP + i;
        const size_t ss = (size_t)NP * NP;
        double s = 0.0;
        for (int c = lane; c < ncta; c += 32) s += __ldg(a + c * ss) + __ldg(b + c * ss);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) vxc[idx] = s;
    }
    __shared__ double sh[256];
    __shared__ bool last;
    if (blockIdx.x == 0) {
        double e = 0.0;
        for (int k = threadIdx.x; k < ncta; k += 256) e += epart[k];
        sh[threadIdx.x] = e;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) *d_exc = sh[0];
    }
    if (host_slot) {
        __threadfence();               // this CTA's V_xc (and E_xc) stores are visible device-wide ...
        __syncthreads();
        if (threadIdx.x == 0) last = atomicAdd(done_counter, 1u) == gridDim.x - 1;   // ... before it is counted
        __syncthreads();
        if (last && threadIdx.x == 0) {
            *done_counter = 0u;        // (self-resetting: ready for the next call)
            __threadfence();
            host_slot[0] = *reinterpret_cast<volatile double*>(d_exc);
            __threadfence_system();
            host_slot[1] = seq;
        }
    }
}

template <int NF, int NPL>
static void launch(CublasHandleWrapper* ctx, const Problem& p, int nsm) {
    constexpr int NP = 8 * NF, LDD = ldd_for(NP);
    const int nao = p.nao;
    // resident super-blocks (4 tiles per buffer) where two of them stay within 16 KB per warp: nao <= 8 for GGA, 32 for LDA
    const bool resident = 2 * (size_t)NPL * 4 * TR * nao * sizeof(double) <= 16384 && !ctx->small_streaming;
    const size_t warp_d = 2 * (size_t)NPL * TR * nao * (resident ? 4 : 1);   // doubles of one warp's two buffers
    auto k = xc_small_kernel<NF, NPL>;
    // launch shape of this (instance, nao), worked out ONCE: warps per CTA so that the CTA's shared memory (Dsym + the
    // warps' buffers, at least the NP^2 + 8 doubles of the final reduction) fits, resident CTAs per SM from the
    // occupancy calculator -- at H2O size the whole call is ~20 us and two driver queries per call were a third of it
    // (function attributes are per device: one slot per device ordinal)
    struct Shape { int nao, nwarp, per_sm, resident; size_t smem; };
    static Shape cache[16] = {};
    static std::mutex cache_mu;   // (solvers on several host threads -- the fan-out's workers -- may share a device slot)
    std::lock_guard<std::mutex> cache_lock(cache_mu);
    Shape& sh = cache[ctx->device & 15];
    if (sh.nao != nao || sh.per_sm <= 0 || sh.resident != (resident ? 1 : 0)) {
        int nwarp = MAXW;
        auto smem_for = [&](int nw) {
            size_t d = (size_t)NP * LDD + (size_t)nw * warp_d + 8 + 2 * (size_t)nw;   // (+ two mbarriers per warp)
            if (d < (size_t)NP * NP + 8) d = (size_t)NP * NP + 8;
            return d * sizeof(double);
        };
        while (nwarp > 1 && smem_for(nwarp) > 110 * 1024) nwarp >>= 1;   // aim at two resident CTAs per SM
        const size_t smem = smem_for(nwarp);
        int per_sm = 0;
        DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DFT_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, nwarp * 32, smem));
        if (per_sm < 1) { ctx->failed = true; return; }
        if (per_sm * nwarp > 16) per_sm = 16 / nwarp > 0 ? 16 / nwarp : 1;   // (more resident warps only add partials)
        sh.nao = nao; sh.nwarp = nwarp; sh.per_sm = per_sm; sh.smem = smem; sh.resident = resident ? 1 : 0;
    }
    const int nblocks = (p.ngrid + TR - 1) / TR;
    const int nsuper = (nblocks + 3) / 4;          // the unit of work of a warp: 4 tiles = 32 points
    int grid = nsm * sh.per_sm;
    if (grid > (nsuper + sh.nwarp - 1) / sh.nwarp) grid = (nsuper + sh.nwarp - 1) / sh.nwarp;
    double* vpart = (double*)ctx->vpart.ensure(sizeof(double) * (size_t)grid * NP * NP, &ctx->failed);
    double* epart = (double*)ctx->epart.ensure(sizeof(double) * grid, &ctx->failed);
    if (ctx->failed) return;

    SmallParams sp;
    sp.ngrid = p.ngrid; sp.nao = nao; sp.nblocks = nblocks; sp.resident = resident ? 1 : 0;
    sp.xc_mode = p.xc_type == 2 ? 4 : p.xc_type * 2 + (ctx->exact_functionals ? 1 : 0);
    sp.dm = p.dm; sp.w = p.w;
    sp.plane[0] = p.ao; sp.plane[1] = p.gx; sp.plane[2] = p.gy; sp.plane[3] = p.gz;
    // 16-byte copies need 16-byte aligned plane bases (a tile starts 64 nao bytes into a plane per 8 points)
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    sp.vec16 = al16(p.ao) && (NPL == 1 || (al16(p.gx) && al16(p.gy) && al16(p.gz))) ? 1 : 0;
    sp.vpart = vpart; sp.epart = epart;

    cudaStream_t st = ctx->stream;
    if (ctx->timing) cudaEventRecord(ctx->ev[0], st);
    k<<<grid, sh.nwarp * 32, sh.smem, st>>>(sp);
    if (ctx->timing) { cudaEventRecord(ctx->ev[1], st); cudaEventRecord(ctx->ev[2], st); }
    const int raw = (ctx->raw_convention && p.xc_type == 1) ? 1 : 0;
    // zero-copy return of E_xc (blocking single-GPU calls only: capi.cu sets `host_exc_slot` and then polls it)
    unsigned int* done = nullptr;
    if (p.host_exc_slot) {
        const bool fresh = ctx->counters.ptr == nullptr;
        done = reinterpret_cast<unsigned int*>(ctx->counters.ensure(COUNTERS_BYTES, &ctx->failed));
        if (ctx->failed) return;
        if (fresh) cudaMemsetAsync(done, 0, COUNTERS_HEAD_BYTES, st);
        done += 10;   // (the last 8 bytes of the buffer: the TMA path's statistics and work counter use the first 40)
    }
    {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((nao * nao + 7) / 8); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = ctx->timing ? 0 : 1;   // (with the timing events between the two launches there is nothing to overlap)
        DFT_CUDA_CHECK(ctx, cudaLaunchKernelEx(&cfg, xc_small_finalize, nao, NP, grid, raw, (const double*)vpart, (const double*)epart,
                                               p.vxc, p.d_exc, done, p.host_exc_slot, p.host_exc_seq));
    }
    if (ctx->timing) cudaEventRecord(ctx->ev[3], st);
    ctx->stats.launches = 2;
    ctx->stats.path = PATH_SMALL;
    DFT_CUDA_CHECK(ctx, cudaGetLastError());
}

}  // namespace smallpath

bool small_compatible(const Problem& p) { return p.nao >= 1 && p.nao <= 48 && p.ngrid >= 1; }

void run_small(CublasHandleWrapper* ctx, const Problem& p) {
    using namespace smallpath;
    if (ctx->num_sms <= 0) cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const int nf = (p.nao + 7) / 8;
#define DFT_SMALL(NF_) do { if (p.xc_type == 0) launch<NF_, 1>(ctx, p, ctx->num_sms); else launch<NF_, 4>(ctx, p, ctx->num_sms); } while (0)
    switch (nf) {
        case 1: DFT_SMALL(1); break;
        case 2: DFT_SMALL(2); break;
        case 3: DFT_SMALL(3); break;
        case 4: DFT_SMALL(4); break;
        case 5: DFT_SMALL(5); break;
        default: DFT_SMALL(6); break;
    }
#undef DFT_SMALL
}

}  // namespace xc
