// xc_tma.cu -- the fast XC path: TMA-fed, mbarrier-pipelined FP64 tensor-core (DMMA) kernels.
//
// Two persistent-style, warp-specialised contraction kernels per XC build (plus the pointwise kernel between
// them).  One CTA per SM, 384 threads: two consumer warpgroups (8 warps) and one producer warpgroup whose
// elected lanes drive TMA; `setmaxnreg` moves registers from the producer to the consumers (40 / 232 per
// thread) so that 64 x 40 FP64 accumulator tiles fit without spills.  Neither kernel has a CTA-wide barrier
// in its main loop: everything a consumer warp touches arrives through a TMA ring (full/empty mbarriers),
// so the DMMA pipe never drains between tiles.
//
//   density_tma_kernel   subsystem (b): for each block of 64 grid points and column tile nt
//        C = Phi_blk . Dsym[:, nt]                DMMA, operands streamed by TMA (SWIZZLE_128B)
//        rho += rowsum(C o Phi), grad rho += 2 rowsum(C o dPhi)
//                                                 the Phi / dPhi tiles of the row-dots come through
//                                                 the SAME ring as 32 KB "pieces" right behind the
//                                                 k-chunks (no global gathers)
//     two ping-pong consumer groups with a ring each; blocks handed out dynamically; k-steps whose Phi
//     fragment is exactly zero are skipped (AO screening).
//     replaces get_rho_kernel / get_rho_sigma_kernel_planar (dft_solver.cu:294-307, :346-380).
//
//   xc_point_kernel      subsystem (c): the functional once per point -> (a, b) coefficients + E_xc partials
//     replaces both passes of the *_fused_kernel's (:309-513) and reduce_sum_kernel (:285-292).
//
//   vxc_tma_kernel       subsystem (d): for each (output tile, grid slice)
//        M += B^T Phi,  B = a o Phi + b . grad Phi
//     B never exists in memory: each warp combines the four plane tiles with the point's
//     coefficients directly into its DMMA A fragments (4 shared loads + 4 FP64 ops per fragment).
//     replaces the B matrix (:577,:613,:655) and cublasDgemm (:580,:616,:663).
//
//   finalize_tma_kernel  out = M + M^T over slices in a fixed order (replaces :515-527) + E_xc.
//
// Shared-memory tiles are written by TMA with the 128-byte swizzle; fragment rows (density kernel)
// or reduction rows (V kernel) are permuted so that every 64-bit fragment load -- including the
// epilogue's -- is bank-conflict free (tools/check_smem_maps.py proves the maps on the host).
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
// Odd nao with odd ngrid puts the y-gradient plane at 8 mod 16: run_tma copies that one plane to aligned scratch per
// call (see tma_compatible).  Inputs TMA cannot address at all (misaligned base pointers) take the generic path
// (xc_generic.cu).
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "dmma.cuh"
#include "engine.h"
#include "tma.cuh"
#include "xc_functionals.cuh"

namespace xc {
namespace tmapath {

constexpr int NCW = 8;                      // consumer warps
constexpr int NCONS = NCW * 32;             // 256 consumer threads = 2 warpgroups
constexpr int NTHREADS = NCONS + 128;       // + 1 producer warpgroup
constexpr int REGS_CONSUMER = 232;          // 384 threads x 168 = 256 x 232 + 128 x 40
constexpr int REGS_PRODUCER = 40;
static_assert(256 * REGS_CONSUMER + 128 * REGS_PRODUCER <= 384 * 168, "setmaxnreg budget");
constexpr int MB = 64;                      // grid rows per density block (one consumer group's)

struct SubProblem {
    int rows;    // rows of this sub-problem
    int gmul;    // grid point of row j: g = gmul * j + gadd
    int gadd;
    int shift;   // AO index = column - shift (0 or 1)
    int blk0;    // first density block of this sub-problem
    int coef0;   // first row of this sub-problem in the coefficient array (padded to 128 rows)
};

struct DensityParams {
    CUtensorMap map_a[2];     // Phi of each sub-problem, box 16 x 64 (k-loop A operand)
    CUtensorMap map_e[2][4];  // the four planes, box 16 x 64 (epilogue pieces)
    CUtensorMap map_d;        // [2][NP][KP] symmetrised density (plain, shifted), box 16 x NT
    SubProblem sub[2];
    int nsub, nblocks, ntiles, nk, NP, l2_prefetch, coef_rows, zero_skip;
    int per_tile;      // unit of work: 0 = a 64-point block (all column tiles), 1 = one column tile of a block
    int stagger_min;   // group 1 starts half a tile period late when the CTA has more than this many blocks to do
    int wait_ns;       // producer threads sleep this long between polls of an `empty` barrier (0: poll back to back)
    int debug_nodmma;  // diagnostic: treat every k-step as zero (measures the operand-delivery floor; results are wrong)
    int producers2;    // two TMA-issuing threads per consumer group (primary + helper) instead of one
    int unit_stride;   // visit order of the blocks: k-th draw -> block (k * unit_stride) mod nblocks (<= 1: grid order)
    unsigned long long* counters;  // [2]: k-steps executed, k-steps total (AO screening statistics)
    unsigned int* sched;           // next density block to hand out (dynamic scheduling; reset to 0 before the launch), or null
    double* rho;       // [2 warp columns][coef_rows][4]: partial (rho, drho/2) row sums, summed by the point kernel
    long long* phase;  // DFT_PHASE_TIMING builds: [CTA][warp][4] cycles in {k-loop, piece wait, piece math, block tail}
};

// V kernel operand maps.  A plane is also described as a 3-D tensor {16 columns, rows, 16-column
// blocks} over its WHOLE 16-column blocks, so that one TMA instruction brings all the blocks of a tile
// (box {16, VK, tile/16}) in the [block][row][16] layout the fragment loads expect; the partial last
// block (nao mod 16 columns, zero-filled past the end) still comes through the 2-D map.
struct VxcParams {
    CUtensorMap m3[2][4];    // M side: box {16, VK, MT / 16}
    CUtensorMap m3l[2][4];   // M side, last (partial) tile: box {16, VK, whole blocks in that tile}
    CUtensorMap n3[2];       // N side (Phi): box {16, VK, NT / 16}
    CUtensorMap n3l[2];
    CUtensorMap p2[2][4];    // 2-D, box 16 x VK
    SubProblem sub[2];
    int nfull[2];            // whole 16-column blocks of each sub-problem
    int rem[2];              // columns in the partial block (0: none)
    int use3d, zero_skip;
    int producers;           // TMA-issuing threads per CTA (1..4)
    int prefetch;            // L2 prefetch distance of the producer in ring stages (0: none)
    int wait_ns;             // producer / scanner threads sleep this long between barrier polls
    int debug_nodmma;        // diagnostic: skip every DMMA (operand-delivery floor; results are wrong)
    int chunk_stride[2];     // per sub-problem: multiplier (coprime to the chunk count) that scatters consecutive stages over the grid
    int nsub, tiles_m, tiles_n, lda_half, rows_per_slice, slices_per_sub, ldv, mpv;
    const double* coef;
    double* vpart;
    unsigned long long* counters;  // [2]: (box, k-step) units executed / total (box-bit instances)
    const unsigned char* fmap;     // [tiles_m][16]: 8 x 1 per-warp-vote instances: warp w owns 8-column fragments fmap[tm][2w], [2w+1] of the M tile
    unsigned int* fstat;           // [tiles_m][16]: live (fragment, k-step) units found by those instances (drives the next call's fmap)
    long long* phase;  // DFT_PHASE_TIMING builds: [CTA][9 warps][4] cycles {wait, work, tail, total}
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
template <int N>
__device__ __forceinline__ void reg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
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
// The 8 consumer warps form TWO independent groups of 4 (2 x 2 warps, warp tile 32 x (8 NF2)); each
// group works through its own blocks of 64 grid points with its own TMA ring (producer warps 8 and
// 9), and group 1 starts half a column-tile period after group 0.  While one group is in its
// epilogue (no tensor work) the other one is in its k-loop, so the DMMA pipe stays busy: a
// "ping-pong" of two CTAs' worth of work inside one CTA.
//
// Ring contents per (block, column tile): nk k-chunks {A: Phi[64 rows][16 k], D: Dsym[NT cols][16 k]}
// followed by the tile's epilogue pieces.  A GGA piece is {Phi, dxPhi, dyPhi, dzPhi}[64 rows][16 cols]
// (4 boxes of 8 KB); an LDA piece is Phi[64 rows][16 cols].  Every warp of the group waits for and
// releases every stage; only the two warps whose accumulator columns fall into a piece read it.
//
// A lane owns 4 fragment rows, and its partial row sums rs[4][NPL] stay in registers for the WHOLE
// block (all column tiles, all pieces).  A piece costs its owners 64 shared loads + 64 FMAs in 16
// independent chains; the cross-lane reduction happens once per block.
constexpr int GW = 4;                        // warps per consumer group

// WIDE (round 2, option density_wide): ONE consumer group of 8 warps (4 x 2) on blocks of 128 grid points with one
// ring, instead of two ping-pong groups of 4 warps on 64-point blocks with a ring each.  The Dsym chunk -- two thirds of
// the k-loop's operand bytes -- is then delivered once per 128 rows instead of once per 64 (a third fewer k-loop bytes),
// and the ring is seven stages of 32 KB deep instead of three; the price is that nobody works on the tensor pipe
// while the group is in its epilogue.  The per-warp code (fragment maps, votes, row-dot epilogue) is the same.
template <int NF2, int NPL, bool WIDE = false>
struct DensitySmem {
    static_assert(NF2 % 2 == 0, "a warp's columns are whole 16-column groups");
    static constexpr int NT = 16 * NF2;
    static constexpr int GRPS = WIDE ? 1 : 2;                             // consumer groups (rings) per CTA
    static constexpr int GWW = NCW / GRPS;                                // warps per group
    static constexpr int RB = WIDE ? 2 * MB : MB;                         // grid rows per group block
    static constexpr int A_BYTES = RB * 128;
    static constexpr int D_BYTES = NT * 128;
    static constexpr int K_BYTES = A_BYTES + D_BYTES;
    static constexpr int PLANE_BYTES = RB * 128;                          // one plane of a piece
    // A piece may travel as PSPLIT stages of NPL / PSPLIT planes each.  For the two-group kernel PSPLIT = 2 (16 KB stages
    // fit the 24 KB slot of a k-chunk: a 4-stage ring instead of 3 stages of 32 KB) was measured in round 2 and is SLOWER
    // -- C5 9.30 against 9.03 ms, C4 1.235 against 1.186 ms (profiles/r2_u4_density_half_pieces.txt): twice the stage
    // hand-shakes in the epilogue cost more than the deeper ring gains -- so there a piece stays one stage.  The wide
    // kernel's 128-row GGA piece is 64 KB and travels as two stages of two planes.
    static constexpr int PSPLIT = (WIDE && NPL == 4) ? 2 : 1;             // stages per piece
    static constexpr int PPL = NPL / PSPLIT;                              // planes per stage of a piece
    static constexpr int PIECE_BYTES = PPL * PLANE_BYTES;
    static constexpr int STAGE_BYTES = K_BYTES > PIECE_BYTES ? K_BYTES : PIECE_BYTES;
    static constexpr int FIXED_BYTES = 640 + 1024;                        // barriers, slots, unit queue, alignment slack
    static constexpr int CAP = WIDE ? 8 : 6;
    static constexpr int STAGES = (232448 - FIXED_BYTES) / (GRPS * STAGE_BYTES) < CAP ? (232448 - FIXED_BYTES) / (GRPS * STAGE_BYTES) : CAP;
    static_assert(STAGES >= 2, "ring depth");
    static constexpr int NCG = NT / 16;                                   // 16-column groups per tile (= NF2)
    static constexpr int NPIECES = NCG * PSPLIT;                          // piece stages per tile
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;               // one group's ring
    static constexpr int BAR_OFF = GRPS * RING_BYTES;                     // [groups]{full[STAGES], empty[STAGES]}
    static constexpr int BLK_OFF = BAR_OFF + 2 * 2 * STAGES * 8;          // [2 groups][STAGES] block id carried by a stage
    static constexpr int UQ_OFF = ((BLK_OFF + 2 * STAGES * 4 + 7) / 8) * 8; // [2 groups][4] (sequence << 32 | unit): primary -> helper issuing thread
    static constexpr int TOTAL = UQ_OFF + 2 * 4 * 8 + 8 + 1024;           // + alignment slack
    static_assert(TOTAL <= 232448, "shared memory");

    // piece pc -> column group; consecutive pieces go to different warp columns
    __host__ __device__ static constexpr int piece_cg(int pc) { return (pc & 1) * (NCG / 2) + (pc >> 1); }
    // piece STAGE ps -> (column group, first plane)
    __host__ __device__ static constexpr int stage_cg(int ps) { return piece_cg(ps / PSPLIT); }
    __host__ __device__ static constexpr int stage_pl0(int ps) { return (ps % PSPLIT) * PPL; }
};

template <int NF2, int NPL, bool WIDE>
__global__ void __launch_bounds__(NTHREADS, 1)
density_tma_kernel(const __grid_constant__ DensityParams P) {
    using L = DensitySmem<NF2, NPL, WIDE>;
    constexpr int NT = L::NT;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (tma::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - tma::smem_u32(smem_raw));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::BAR_OFF);
        for (int g = 0; g < L::GRPS; ++g)
            for (int s = 0; s < L::STAGES; ++s) {
                tma::mbar_init(&bars[g * 2 * L::STAGES + s], P.producers2 ? 2 : 1);   // full: every issuing thread of the group arrives
                tma::mbar_init(&bars[g * 2 * L::STAGES + L::STAGES + s], L::GWW);
            }
        tma::fence_barrier_init();
        for (int i = 0; i < 8; ++i) reinterpret_cast<volatile unsigned long long*>(sm + L::UQ_OFF)[i] = 0ull;
    }
    __syncthreads();

    const int nblocks = P.nblocks, ntiles = P.ntiles, nk = P.nk;
    // consumer group / ring of this warp: warps 0-3 -> 0, 4-7 -> 1, producer warps 8 -> 0, 9 -> 1
    const int grp = WIDE ? 0 : (warp < NCW ? warp >> 2 : (warp - NCW) & 1);
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + L::BAR_OFF) + grp * 2 * L::STAGES;
    uint64_t* empty = full + L::STAGES;
    unsigned char* ring = sm + grp * L::RING_BYTES;
    const uint32_t ring_u32 = base + grp * L::RING_BYTES;
    const int b_first = L::GRPS * blockIdx.x + grp, b_step = L::GRPS * gridDim.x;
    // Blocks are handed out dynamically (one global counter): a block of points near many atoms costs several
    // times one out in the tail, so a static deal leaves the last groups running alone.  The producer draws
    // the block ids (the next one while the current block streams) and passes each to its consumers in a
    // shared-memory slot that travels with the block's first ring stage; a stage carrying -1 ends the group.
    // The unit of work is one column tile of a block (a third of a block at C5: a shorter tail, which is what a
    // multi-GPU shard with ~10 blocks per group notices); the point kernel adds one partial per tile and warp
    // column.  With `density_unit` 1 it is a whole block, whose partial row sums then stay in registers over
    // all column tiles.
    const bool dyn = P.sched != nullptr;
    const uint32_t blk_slot = base + L::BLK_OFF + grp * L::STAGES * 4;
    const bool per_tile = P.per_tile != 0;
    const int nunits = per_tile ? nblocks * ntiles : nblocks;
    // Order in which the units are visited (option density_scatter): the k-th draw is block (k' * stride) mod nblocks,
    // stride coprime to nblocks (golden ratio), the column tiles of a block still back to back.  In grid order the 296
    // consumer groups of the GPU work on ~100 neighbouring blocks -- one atom's region -- at any time, so all SMs
    // are in a sparse (delivery-bound) or a dense (tensor-bound) stretch TOGETHER and their demand on the L2 comes in
    // chip-wide bursts; scattered, the stretches average out across SMs.  MEASURED: no difference (C5 8.85 against 8.84 ms,
    // profiles/r2_u16_density_scatter.txt) -- the stage waits are not chip-wide L2 bursts either.  Default: grid order.
    const int ntl = per_tile ? ntiles : 1;
    auto unit_of = [&](int draw) {
        if (P.unit_stride <= 1) return draw;
        const int bb = draw / ntl, tt = draw - bb * ntl;
        return (int)(((long long)bb * P.unit_stride) % nblocks) * ntl + tt;
    };

    if (warp >= NCW) {
        // ===================== producer warpgroup: warps 8 and 9, one elected lane each =====================
        reg_dec<REGS_PRODUCER>();
        // Optionally two issuing threads per group (option density_producers 2; round 2).  One thread streams only ~21
        // bytes per clock into shared memory, and the kernel's delivery floor with every DMMA off, 43 B/clk/SM, is
        // exactly two such threads, while ncu shows the consumers spending a fifth of their samples waiting for stages.
        // The PRIMARY (lane 0 of warps 8 | 9) draws the units, arms the barriers and brings the Phi chunk / the first
        // half of a piece's planes; the HELPER (lane 0 of warps 10 | 11) follows the same unit sequence -- handed over
        // through a 4-entry queue in shared memory when the deal is dynamic -- and brings the Dsym chunk / the second
        // half of the planes.  (A transfer that lands before the primary has armed the barrier only makes its tx-count
        // transiently negative.)  MEASURED: no gain -- C5 8.83 against 8.82 ms, C4 1.179 against 1.183
        // (profiles/r2_u13_density_two_issuers.txt): the waits are not the issue rate.  Default: one thread.
        const bool two = P.producers2 != 0;
        volatile unsigned long long* uq = reinterpret_cast<volatile unsigned long long*>(sm + L::UQ_OFF) + 4 * grp;
        if (warp >= NCW + 2 && warp < NCW + 2 + L::GRPS && lane == 0 && two) {
            // ===================== helper issuing thread =====================
            tma::prefetch_map(&P.map_d);
            uint32_t it = 0;
            unsigned int kseq = 0;
            int u = b_first;
            auto next_unit = [&]() {
                if (!dyn) { if (kseq) u += b_step; ++kseq; return; }
                ++kseq;
                unsigned long long e;
                uint32_t spins = 0;
                do { e = uq[(kseq - 1) & 3]; if ((++spins & 1023u) == 0) __nanosleep(64); } while ((unsigned int)(e >> 32) != kseq);
                u = (int)(unsigned int)(e & 0xffffffffull);
            };
            next_unit();
            while (u < nunits) {
                const int ue = unit_of(u);
                const int bu = per_tile ? ue / ntiles : ue;                    // unit block (RB rows)
                const int nt0 = per_tile ? ue - bu * ntiles : 0, nt1 = per_tile ? nt0 + 1 : ntiles;
                const int b = WIDE ? 2 * bu : bu;                              // its first 64-row block
                const int si = (P.nsub > 1 && b >= P.sub[1].blk0) ? 1 : 0;
                const int blk = b - P.sub[si].blk0;
                const int drow0 = P.sub[si].shift * P.NP;
                for (int nt = nt0; nt < nt1; ++nt) {
                    for (int kc = 0; kc < nk; ++kc, ++it) {
                        const uint32_t s = it % L::STAGES, ph = (it / L::STAGES) & 1u;
                        tma::mbar_wait_relaxed(&empty[s], ph ^ 1u, (uint32_t)P.wait_ns);
                        tma::load_2d(ring + s * L::STAGE_BYTES + L::A_BYTES, &P.map_d, kc * 16, drow0 + nt * NT, &full[s]);
                        tma::mbar_arrive(&full[s]);
                    }
                    // The helper ARRIVES on every stage, also where it has nothing to bring (an LDA piece is one plane):
                    // waits are by phase parity, and a thread that sat out a stage could be lapped twice on a slot
                    // between two of its polls and then mistake the slot's previous phase for the one it waits for.
                    for (int pc = 0; pc < L::NPIECES; ++pc, ++it) {
                        const uint32_t s = it % L::STAGES, ph = (it / L::STAGES) & 1u;
                        tma::mbar_wait_relaxed(&empty[s], ph ^ 1u, (uint32_t)P.wait_ns);
                        unsigned char* st = ring + s * L::STAGE_BYTES;
                        const int c0 = nt * NT + 16 * L::stage_cg(pc);
                        if (L::PPL >= 2)
                            for (int p = L::PPL / 2; p < L::PPL; ++p)
                                tma::load_2d(st + p * L::PLANE_BYTES, &P.map_e[si][L::stage_pl0(pc) + p], c0, blk * MB, &full[s]);
                        tma::mbar_arrive(&full[s]);
                    }
                }
                next_unit();
            }
            return;
        }
        if (warp < NCW + L::GRPS && lane == 0) {
            unsigned int kseq = 0;
            auto publish = [&](int uu) {   // (dynamic deal: the helper reads the unit sequence from here)
                ++kseq;
                if (two && dyn) uq[(kseq - 1) & 3] = ((unsigned long long)kseq << 32) | (unsigned int)uu;
            };
            tma::prefetch_map(&P.map_a[0]);
            if (P.nsub > 1) tma::prefetch_map(&P.map_a[1]);
            tma::prefetch_map(&P.map_d);
            uint32_t it = 0;
            // Short-range L2 prefetch: the ring is only STAGES deep, which covers L2 latency but not DRAM
            // latency when this group has the tensor pipe to itself.  A tiles are requested PF k-chunks
            // ahead, the pieces of a tile during its last k-chunks -- a few microseconds ahead, a few MB in
            // flight over the whole GPU (a prefetch a whole k-loop ahead was evicted before use and
            // doubled the DRAM traffic).
            constexpr int PF = 6;
            const bool pf = P.l2_prefetch != 0;
            int u = dyn ? (int)atomicAdd(P.sched, 1u) : b_first;
            publish(u);
            while (u < nunits) {
                const int ue = unit_of(u);
                const int bu = per_tile ? ue / ntiles : ue;                    // unit block (RB rows)
                const int nt0 = per_tile ? ue - bu * ntiles : 0, nt1 = per_tile ? nt0 + 1 : ntiles;
                const int b = WIDE ? 2 * bu : bu;                              // its first 64-row block
                const int si = (P.nsub > 1 && b >= P.sub[1].blk0) ? 1 : 0;
                const int blk = b - P.sub[si].blk0;
                const int drow0 = P.sub[si].shift * P.NP;
                const int un = dyn ? (int)atomicAdd(P.sched, 1u) : u + b_step;  // this group's next unit
                publish(un);
                const int une = un < nunits ? unit_of(un) : un;
                const int bn = (WIDE ? 2 : 1) * (per_tile ? une / ntiles : une);  // (its block: L2 prefetch only)
                const int sin = (P.nsub > 1 && bn >= P.sub[1].blk0) ? 1 : 0;
                const int blkn = bn - P.sub[sin].blk0;
                for (int nt = nt0; nt < nt1; ++nt) {
                    const int pc0 = nk > L::NPIECES ? nk - L::NPIECES : 0;  // chunk at which piece prefetch starts
                    for (int kc = 0; kc < nk; ++kc, ++it) {
                        if (pf) {
                            const int ka = kc + PF;
                            if (ka < nk) tma::prefetch_2d(&P.map_a[si], ka * 16, blk * MB);
                            else if (nt + 1 < nt1) { if (ka - nk < nk) tma::prefetch_2d(&P.map_a[si], (ka - nk) * 16, blk * MB); }
                            else if (une < nunits && ka - nk < nk) tma::prefetch_2d(&P.map_a[sin], (ka - nk) * 16, blkn * MB);
                            if (kc >= pc0) {
                                // pieces [lo, hi) of this tile; all of them by the last chunk
                                const int span = nk - pc0;
                                const int lo = ((kc - pc0) * L::NPIECES) / span, hi = ((kc - pc0 + 1) * L::NPIECES) / span;
                                for (int pc = lo; pc < hi; ++pc)
                                    for (int p = 0; p < L::PPL; ++p)
                                        tma::prefetch_2d(&P.map_e[si][L::stage_pl0(pc) + p], nt * NT + 16 * L::stage_cg(pc), blk * MB);
                            }
                        }
                        const uint32_t s = it % L::STAGES, ph = (it / L::STAGES) & 1u;
                        tma::mbar_wait_relaxed(&empty[s], ph ^ 1u, (uint32_t)P.wait_ns);
                        unsigned char* st = ring + s * L::STAGE_BYTES;
                        if (nt == nt0 && kc == 0) asm volatile("st.shared.s32 [%0], %1;" ::"r"(blk_slot + 4 * s), "r"(ue) : "memory");
                        tma::mbar_arrive_expect_tx(&full[s], L::K_BYTES);
                        tma::load_2d(st, &P.map_a[si], kc * 16, blk * MB, &full[s]);
                        if (!two) tma::load_2d(st + L::A_BYTES, &P.map_d, kc * 16, drow0 + nt * NT, &full[s]);
                    }
                    for (int pc = 0; pc < L::NPIECES; ++pc, ++it) {
                        const uint32_t s = it % L::STAGES, ph = (it / L::STAGES) & 1u;
                        tma::mbar_wait_relaxed(&empty[s], ph ^ 1u, (uint32_t)P.wait_ns);
                        unsigned char* st = ring + s * L::STAGE_BYTES;
                        tma::mbar_arrive_expect_tx(&full[s], L::PIECE_BYTES);
                        const int c0 = nt * NT + 16 * L::stage_cg(pc);
                        const int pend = (two && L::PPL >= 2) ? L::PPL / 2 : L::PPL;   // (the helper brings the other half)
                        for (int p = 0; p < pend; ++p)
                            tma::load_2d(st + p * L::PLANE_BYTES, &P.map_e[si][L::stage_pl0(pc) + p], c0, blk * MB, &full[s]);
                    }
                }
                u = un;
            }
            if (dyn) {  // no units left: one empty stage carrying -1 tells the consumers to stop
                const uint32_t s = it % L::STAGES, ph = (it / L::STAGES) & 1u;
                tma::mbar_wait_relaxed(&empty[s], ph ^ 1u, (uint32_t)P.wait_ns);
                asm volatile("st.shared.s32 [%0], %1;" ::"r"(blk_slot + 4 * s), "r"(-1) : "memory");
                tma::mbar_arrive(&full[s]);
                if (two) tma::mbar_arrive(&full[s]);   // (the helper never sees this stage)
            }
        }
        return;
    }

    // ===================== consumers: 2 groups x (2 x 2 warps), warp tile 32 x (8 NF2) =====================
    reg_inc<REGS_CONSUMER>();
    const int gw = WIDE ? warp : (warp & 3);
    const int wm = gw >> 1, wn = gw & 1;   // (wide: 4 x 2 warps over 128 rows; else 2 x 2 over 64)
    const int q = lane >> 2, qcol = lane & 3;
    // fragment row -> tile row.  A side: {0,3,4,7 | 1,2,5,6}; B side: {0,2,4,6 | 1,3,5,7}.  Both make
    // the k-loop loads conflict-free under SWIZZLE_128B, and together they make the epilogue loads
    // conflict-free as well (tools/check_smem_maps.py).
    const int rho = (0x65217430u >> (4 * q)) & 7;
    const int perm = 2 * (q & 3) + (q >> 2);
    uint32_t koff_a[4], koff_b[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        koff_a[ks] = ((((2 * ks + (qcol >> 1)) ^ rho) & 7) << 4) + ((qcol & 1) << 3);
        koff_b[ks] = ((((2 * ks + (qcol >> 1)) ^ perm) & 7) << 4) + ((qcol & 1) << 3);
    }
    const uint32_t a_row = (uint32_t)(wm * 32 + rho) * 128u;
    const uint32_t b_row = (uint32_t)(wn * 8 * NF2 + perm) * 128u;
    // accumulator column j = 2 qcol + e of an n-fragment is tile column 2 (j & 3) + (j >> 2); eoff[e][s]
    // is the byte offset of that column in a piece row for the fragment in the lower (s = 0) / upper
    // (s = 1) 8 columns of the 16-column group
    uint32_t eoff[2][2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int j = 2 * qcol + e;
        const int nc = 2 * (j & 3) + (j >> 2);
#pragma unroll
        for (int sx = 0; sx < 2; ++sx)
            eoff[e][sx] = (uint32_t)(rho * 128 + ((((4 * sx + (nc >> 1)) ^ rho) & 7) << 4) + ((nc & 1) << 3));
    }
    const int prow0 = wm * 32;  // this warp's rows inside a piece
    // Per-lane fragment offsets, kept in registers behind an opaque move.  (Left to itself the compiler re-derives
    // them from threadIdx in front of the loads of EVERY k-step -- ten dependent integer instructions between the vote
    // and the B-fragment loads, on the critical path of each executed k-step; SASS of round 1's instance.)  The four
    // k-steps of a chunk differ only in the 16-byte chunk index, (2 ks + ..) ^ row: koff[ks] = koff[0] ^ (ks << 5),
    // and every other term of an operand address is a multiple of 128, so one XOR per k-step does it.
    uint32_t a_lane = a_row + koff_a[0], b_lane = b_row + koff_b[0];
    asm volatile("mov.u32 %0, %1;" : "=r"(a_lane) : "r"(a_lane));
    asm volatile("mov.u32 %0, %1;" : "=r"(b_lane) : "r"(b_lane));
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int sx = 0; sx < 2; ++sx) asm volatile("mov.u32 %0, %1;" : "=r"(eoff[e][sx]) : "r"(eoff[e][sx]));
    const bool no_skip = P.zero_skip == 0;
#ifdef DFT_DIAGNOSTICS
    const bool dbg_off = P.debug_nodmma != 0;
#else
    constexpr bool dbg_off = false;
#endif

    // group 1 starts half a column-tile period late (only worth it when there are several blocks to do)
    if (!WIDE && grp == 1 && nblocks > P.stagger_min * (int)gridDim.x) {
        const long long t_start = clock64(), delay = (long long)nk * 2048;
        while (clock64() - t_start < delay) __nanosleep(2000);
    }

    uint32_t it = 0;
    unsigned int n_ks_done = 0, n_ks_total = 0;
#ifdef DFT_PHASE_TIMING
    long long t_k = 0, t_w = 0, t_m = 0, t_t = 0, t0 = clock64(), t1;
    int n_iv = 0;
#define PHASE_MARK(acc_) do { t1 = clock64(); acc_ += t1 - t0; t0 = t1; } while (0)
#else
#define PHASE_MARK(acc_) do { } while (0)
#endif
    for (int u = b_first;; u += b_step) {
        if (dyn) {  // the unit id travels with the unit's first stage
            const uint32_t s = it % L::STAGES, ph = (it / L::STAGES) & 1u;
            tma::mbar_wait(&full[s], ph);
            asm volatile("ld.shared.s32 %0, [%1];" : "=r"(u) : "r"(blk_slot + 4 * s) : "memory");
            if (u < 0) break;
        } else if (u >= nunits) {
            break;
        }
        const int ue = dyn ? u : unit_of(u);   // (the dynamic deal hands over the unit itself)
        const int bu = per_tile ? ue / ntiles : ue;
        const int nt0 = per_tile ? ue - bu * ntiles : 0, nt1 = per_tile ? nt0 + 1 : ntiles;
        const int b = WIDE ? 2 * bu : bu;
        const int si = (P.nsub > 1 && b >= P.sub[1].blk0) ? 1 : 0;
        const int blk = b - P.sub[si].blk0;
        double rs[4][NPL];  // per-lane partial row sums of the whole block
#pragma unroll
        for (int mf = 0; mf < 4; ++mf)
#pragma unroll
            for (int p = 0; p < NPL; ++p) rs[mf][p] = 0.0;

        for (int nt = nt0; nt < nt1; ++nt) {
            double acc[4][NF2][2];
#pragma unroll
            for (int mf = 0; mf < 4; ++mf)
#pragma unroll
                for (int nf = 0; nf < NF2; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
#ifdef DFT_PHASE_TIMING
            if (gw == 0 && lane == 0 && P.phase && n_iv > 0 && n_iv <= 64)
                P.phase[65536 + ((size_t)blockIdx.x * 2 + grp) * 128 + 2 * (n_iv - 1) + 1] = clock64();
            ++n_iv;
#endif

            n_ks_total += 4 * nk;
            for (int kc = 0; kc < nk; ++kc, ++it) {
                const uint32_t s = it % L::STAGES, ph = (it / L::STAGES) & 1u;
                tma::mbar_wait(&full[s], ph);
                const uint32_t a_base = ring_u32 + s * L::STAGE_BYTES + a_lane;              // (k-step 0; ks: ^ (ks << 5))
                const uint32_t b_base = ring_u32 + s * L::STAGE_BYTES + L::A_BYTES + b_lane;
                // AO screening at the tensor-core tile level: far from an atom its AOs are EXACT zeros (the
                // evaluator drops primitives beyond the cutoff), so a k-step whose 32 x 4 Phi fragment is all
                // zero adds nothing to C and its 32 DMMAs (and B loads) are branched around -- a warp-uniform
                // vote, no mask to build or keep valid.  (Predicating the DMMAs off instead does NOT help: a
                // predicated-off DMMA still pays its statically scheduled issue stall.)  The four warps of a
                // group see the same AO columns for neighbouring points, so they skip together; the next
                // k-step's A fragments are in flight meanwhile.
                double a[2][4];
#pragma unroll
                for (int mf = 0; mf < 4; ++mf) a[0][mf] = lds_f64(a_base + mf * 1024);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const int cur = ks & 1;
                    if (ks < 3) {
#pragma unroll
                        for (int mf = 0; mf < 4; ++mf) a[cur ^ 1][mf] = lds_f64((a_base ^ (uint32_t)((ks + 1) << 5)) + mf * 1024);
                    }
                    const bool nz = ((a[cur][0] != 0.0) | (a[cur][1] != 0.0)) | ((a[cur][2] != 0.0) | (a[cur][3] != 0.0));
                    if (__any_sync(0xffffffffu, nz | no_skip) && !dbg_off) {
                        double bf[NF2];
#pragma unroll
                        for (int nf = 0; nf < NF2; ++nf) bf[nf] = lds_f64((b_base ^ (uint32_t)(ks << 5)) + nf * 1024);
#pragma unroll
                        for (int mf = 0; mf < 4; ++mf)
#pragma unroll
                            for (int nf = 0; nf < NF2; ++nf) dmma::mma8x8x4(acc[mf][nf], a[cur][mf], bf[nf]);
                        ++n_ks_done;
                    }
                }
                __syncwarp();
                if (lane == 0) tma::mbar_arrive(&empty[s]);
            }
            PHASE_MARK(t_k);
#ifdef DFT_PHASE_TIMING
            if (gw == 0 && lane == 0 && P.phase && n_iv >= 1 && n_iv <= 64) P.phase[65536 + ((size_t)blockIdx.x * 2 + grp) * 128 + 2 * (n_iv - 1)] = t0;
#endif
            // ---- epilogue: row-dots of C with the plane tiles, straight from the ring
            for (int pcg = 0; pcg < L::NCG; ++pcg) {
                const int cg = L::piece_cg(pcg);
#pragma unroll
                for (int hs = 0; hs < L::PSPLIT; ++hs, ++it) {   // the stages of this piece: planes hs PPL .. hs PPL + PPL - 1
                    const uint32_t s = it % L::STAGES, ph = (it / L::STAGES) & 1u;
                    tma::mbar_wait(&full[s], ph);
                    PHASE_MARK(t_w);
                    if (cg / (NF2 / 2) == wn) {
                        // plain shared-memory loads (not asm volatile): the compiler batches them freely
                        const unsigned char* pbase = ring + s * L::STAGE_BYTES + prow0 * 128;
                        const int nfp = cg % (NF2 / 2);
#pragma unroll
                        for (int np = 0; np < NF2 / 2; ++np) {
                            if (np == nfp) {
#pragma unroll
                                for (int sx = 0; sx < 2; ++sx)
#pragma unroll
                                    for (int e = 0; e < 2; ++e)
#pragma unroll
                                        for (int mf = 0; mf < 4; ++mf) {
                                            const double cv = acc[mf][2 * np + sx][e];
                                            const unsigned char* ad = pbase + mf * 1024 + eoff[e][sx];
#pragma unroll
                                            for (int p = 0; p < L::PPL; ++p)
                                                rs[mf][hs * L::PPL + p] = fma(cv, *reinterpret_cast<const double*>(ad + p * L::PLANE_BYTES),
                                                                              rs[mf][hs * L::PPL + p]);
                                        }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) tma::mbar_arrive(&empty[s]);
                    PHASE_MARK(t_m);
                }
            }
        }
        // ---- once per block: reduce the partial row sums over the 4 lanes of a fragment row and hand them
        // to the point kernel.  GGA: transpose-reduce (3 shuffles per row) after which lane qcol holds plane
        // qcol; the 4 lanes of a row write 32 contiguous bytes.  No barrier: the two warp columns' partial
        // sums are added by the point kernel.
        double* rho_mine = P.rho + ((size_t)(2 * nt0 + wn) * P.coef_rows + P.sub[si].coef0 + (size_t)blk * MB + wm * 32 + rho) * 4 + qcol;
        if (NPL == 4) {
            const bool b0 = (qcol & 1) != 0, b1 = (qcol & 2) != 0;
#pragma unroll
            for (int mf = 0; mf < 4; ++mf) {
                // round 1 (lane ^ 1): keep planes {b0, 2 + b0}, hand over the other two
                const double k0 = b0 ? rs[mf][1] : rs[mf][0], g0v = b0 ? rs[mf][0] : rs[mf][1];
                const double k1 = b0 ? rs[mf][NPL - 1] : rs[mf][NPL / 2], g1v = b0 ? rs[mf][NPL / 2] : rs[mf][NPL - 1];
                const double s0 = k0 + __shfl_xor_sync(0xffffffffu, g0v, 1);
                const double s1 = k1 + __shfl_xor_sync(0xffffffffu, g1v, 1);
                // round 2 (lane ^ 2): keep plane 2 b1 + b0 = qcol
                const double k = b1 ? s1 : s0, g = b1 ? s0 : s1;
                rho_mine[mf * 32] = k + __shfl_xor_sync(0xffffffffu, g, 2);
            }
        } else {
#pragma unroll
            for (int mf = 0; mf < 4; ++mf) {
                double v = rs[mf][0];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                rho_mine[mf * 32] = qcol == 0 ? v : 0.0;
            }
        }
        PHASE_MARK(t_t);
    }
    if (lane == 0 && P.counters) {
        atomicAdd(&P.counters[0], (unsigned long long)n_ks_done);
        atomicAdd(&P.counters[1], (unsigned long long)n_ks_total);
    }
#ifdef DFT_PHASE_TIMING
    if (lane == 0 && P.phase) {
        long long* o = P.phase + ((size_t)blockIdx.x * NCW + warp) * 4;
        o[0] = t_k; o[1] = t_w; o[2] = t_m; o[3] = t_t;
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// point kernel -- subsystem (c)
// ------------------------------------------------------------------------------------------------
// One thread per grid point (coefficient row): adds the two partial row sums the density kernel left,
// evaluates the functional ONCE (the reference evaluates it twice, dft_solver.cu:309-513) and writes the
// point's coefficients (a, bx, by, bz) for the V kernel; E_xc = sum w rho eps is reduced with warp shuffles
// and a fixed-order block sum into one partial per CTA (replaces reduce_sum_kernel's 65 536 same-address
// atomics, :285-292).  HBM-bound: 64 + 8 bytes in, 32 bytes out per point, at full occupancy -- inside the
// density kernel only 64 of 384 threads could evaluate while the tensor pipe waited.
struct PointParams {
    SubProblem sub[2];
    int nsub, xc_mode, coef_rows;
    int nparts;          // partial row sums per point: 2 (warp columns), times the column tiles in per-tile mode
    const double* rho;   // [nparts][coef_rows][4]
    const double* w;
    double* coef;        // [coef_rows][4]
    double* exc_part;    // [gridDim.x]
};

constexpr int POINT_THREADS = 256;

__global__ void __launch_bounds__(POINT_THREADS)
xc_point_kernel(const PointParams P) {
    const int row = blockIdx.x * POINT_THREADS + threadIdx.x;
    double e = 0.0;
    if (row < P.coef_rows) {
        const int si = (P.nsub > 1 && row >= P.sub[1].coef0) ? 1 : 0;
        const int j = row - P.sub[si].coef0;
        double2 c01 = make_double2(0.0, 0.0), c23 = make_double2(0.0, 0.0);
        if (j < P.sub[si].rows) {  // rows past the end of a sub-problem are padding: zero coefficients
            const double2* r0 = reinterpret_cast<const double2*>(P.rho) + 2 * (size_t)row;
            double2 s0 = __ldg(r0), s1 = __ldg(r0 + 1);
            for (int part = 1; part < P.nparts; ++part) {   // fixed order: bit-reproducible
                const double2* rp = r0 + 2 * (size_t)P.coef_rows * part;
                const double2 t0 = __ldg(rp), t1 = __ldg(rp + 1);
                s0.x += t0.x; s0.y += t0.y; s1.x += t1.x; s1.y += t1.y;
            }
            const long g = (long)P.sub[si].gmul * j + P.sub[si].gadd;
            const xcfun::PointCoef pc = eval_mode(P.xc_mode, s0.x, 2.0 * s0.y, 2.0 * s1.x, 2.0 * s1.y, __ldg(P.w + g));
            c01 = make_double2(pc.a, pc.bx);
            c23 = make_double2(pc.by, pc.bz);
            e = pc.exc;
        }
        double2* cp = reinterpret_cast<double2*>(P.coef) + 2 * (size_t)row;
        cp[0] = c01;
        cp[1] = c23;
    }
    __shared__ double wsum[POINT_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < POINT_THREADS / 32; ++k) t += wsum[k];
        P.exc_part[blockIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// V kernel
// ------------------------------------------------------------------------------------------------
// Output tile MT x NT_ = (WM MF 8) x (WN NFN 8), consumer warps WM x WN, warp tile (8 MF) x (8 NFN).
// One ring stage holds VK grid rows: the NPL plane tiles of the MT columns, the Phi tile of the NT_
// columns and the VK coefficient rows.
template <int MF, int NFN, int WM, int WN, int NPL, int VK, int STAGES>
struct VxcCfg {
    static_assert(WM * WN == NCW, "8 consumer warps");
    static constexpr int MT = WM * MF * 8;
    static constexpr int NT_ = WN * NFN * 8;
    static_assert(MT % 16 == 0 && NT_ % 16 == 0, "tiles are whole 16-column boxes");
    static constexpr int KS = VK / 4;
    static constexpr int BOXB = VK * 128;                                // one 16-column box
    static constexpr int PLANE_BYTES = (MT / 16) * BOXB;
    static constexpr int N_OFF = NPL * PLANE_BYTES;
    static constexpr int COEF_OFF = N_OFF + (NT_ / 16) * BOXB;
    static constexpr int TX_BYTES = COEF_OFF + VK * 32;
    static constexpr int STAGE_BYTES = ((TX_BYTES + 1023) / 1024) * 1024;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;                 // full, empty [STAGES] (8 B each)
    static constexpr int TOTAL = BAR_OFF + 3 * STAGES * 8 + STAGES * 4 + 4 + 1024;
    static_assert(TOTAL <= 232448, "shared memory");
};

// SKIP selects the AO-screening scheme of the instance:
//   0  none (branch-free; the fastest on dense operands)
//   1  per-fragment votes on the M side: every warp votes on its own built B fragments (round 1's scheme, kept
//      for comparison: the warps skip different fragments and then wait for each other on the shared ring)
// (Round 1 also carried N-side "box bit" instances fed by a scanner warp and a 96 x 192 tile; both measured no
// better than SKIP 1 -- profiles/r1_s53_vxc_instances_C5.txt, r1_s57_vxc_tile_96x192_C5.txt -- and were removed
// when the staged-B kernel below replaced them.)
template <int MF, int NFN, int WM, int WN, int NPL, int VK, int STAGES, int SKIP>
__global__ void __launch_bounds__(NTHREADS, 1)
vxc_tma_kernel(const __grid_constant__ VxcParams P) {
    using L = VxcCfg<MF, NFN, WM, WN, NPL, VK, STAGES>;
    constexpr int KS = L::KS;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (tma::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - tma::smem_u32(smem_raw));
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + L::BAR_OFF);
    uint64_t* empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // output tile of this CTA
    int tm, tn;
    if (P.lda_half) {  // upper-triangular tile pairs, row-major (square tiles only)
        int t = blockIdx.x;
        tm = 0;
        while (t >= P.tiles_n - tm) { t -= P.tiles_n - tm; ++tm; }
        tn = tm + t;
    } else {
        tm = blockIdx.x / P.tiles_n;
        tn = blockIdx.x % P.tiles_n;
    }
    const int m0 = tm * L::MT, n0 = tn * L::NT_;
    // grid slice of this CTA: chunks sl, sl + nsl, sl + 2 nsl, .. (of VK rows) of sub-problem si, dealt
    // round-robin so that every slice of a tile sees the same mix of dense and (skipped) zero regions
    const int si = blockIdx.y / P.slices_per_sub;
    const int sl = blockIdx.y % P.slices_per_sub;
    const int total_chunks = (P.sub[si].rows + VK - 1) / VK;
    const int nchunks = total_chunks > sl ? (total_chunks - sl + P.slices_per_sub - 1) / P.slices_per_sub : 0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tma::mbar_init(&full[s], 1);
            tma::mbar_init(&empty[s], NCW);
        }
        tma::fence_barrier_init();
    }
    // 16-column blocks past the edge of the matrix are never written by TMA: clear the ring once
    for (int o = tid * 16; o < STAGES * L::STAGE_BYTES; o += NTHREADS * 16)
        *reinterpret_cast<double2*>(sm + o) = make_double2(0.0, 0.0);
    tma::fence_proxy_async();
    __syncthreads();

    if (warp >= NCW) {
        reg_dec<REGS_PRODUCER>();
        // ===================== TMA producer =====================
        // Issuing threads: lane 0 of warps 8, 10, 11 (and 9 where it is not the scanner).  One thread streams at most
        // ~23 bytes/clk into shared memory (measured with the DMMAs switched off), well below what the SM can
        // take, so the stage's transfers are dealt to up to four threads; thread 0 also arms the barrier.  (A
        // transfer that completes before the barrier is armed only makes its tx-count transiently negative; the
        // phase cannot complete before thread 0 has arrived.)
        const int nprod = P.producers < 1 ? 1 : (P.producers > 4 ? 4 : P.producers);
        const int pid = warp - NCW;                                    // 0..3
        if (lane == 0 && pid < nprod) {
            // transfers of a stage: bits 0..3 = M-side planes, 4 = Phi tile (N side), 5 = coefficients
            unsigned ops;
            if (NPL == 4) {
                constexpr unsigned T[4][4] = {{0x3f, 0, 0, 0}, {0x23, 0x1c, 0, 0}, {0x03, 0x0c, 0x30, 0}, {0x21, 0x12, 0x04, 0x08}};
                ops = T[nprod - 1][pid];
            } else {
                ops = nprod == 1 ? 0x31u : (pid == 0 ? 0x21u : (pid == 1 ? 0x10u : 0u));
            }
            const double* coef = P.coef + 4 * (size_t)P.sub[si].coef0;
            // blocks of this CTA's M / N tile: fbm / fbn whole ones (one 3-D load), then possibly the partial one
            const int nfull = P.nfull[si], rem = P.rem[si];
            constexpr int NBM = L::MT / 16, NBN = L::NT_ / 16;
            const int bm0 = m0 / 16, bn0 = n0 / 16;
            const int fbm = min(max(nfull - bm0, 0), NBM), fbn = min(max(nfull - bn0, 0), NBN);
            const bool pm = rem > 0 && nfull >= bm0 && nfull - bm0 < NBM;
            const bool pn = rem > 0 && nfull >= bn0 && nfull - bn0 < NBN;
            const uint32_t tx = (uint32_t)(NPL * (fbm + (pm ? 1 : 0)) * L::BOXB + (fbn + (pn ? 1 : 0)) * L::BOXB + VK * 32);
#ifdef DFT_PHASE_TIMING
            long long t_w = 0, t_c = 0, t_start = clock64(), t0 = t_start, t1;
#endif
            // chunk of stage c: (sl + c nsl) stride mod total, kept incrementally (no 64-bit division per stage)
            int cidx = (int)(((long long)sl * P.chunk_stride[si]) % total_chunks);
            const int cstep = (int)(((long long)P.slices_per_sub * P.chunk_stride[si]) % total_chunks);
            uint32_t s = 0, ph = 0;
            const int nloop = ops ? nchunks : 0;  // (a thread without transfers has nothing to wait for)
            // L2 prefetch `P.prefetch` stages ahead (option vxc_prefetch): with the stages scattered over the grid every
            // chunk is a fresh DRAM access, and a slot that frees up waits the whole DRAM + L2 latency for its refill
            // while the ring is only five stages deep
            const int pf = P.prefetch;
            int pidx = cidx;
            if (pf > 0 && P.use3d)
                for (int k = 0; k < pf; ++k) { pidx += cstep; if (pidx >= total_chunks) pidx -= total_chunks; }
            for (int c = 0; c < nloop; ++c) {
                const int j0 = cidx * VK;
                cidx += cstep;
                if (cidx >= total_chunks) cidx -= total_chunks;
                if (pf > 0 && P.use3d && c + pf < nloop) {
                    const int jp = pidx * VK;
                    pidx += cstep;
                    if (pidx >= total_chunks) pidx -= total_chunks;
#pragma unroll
                    for (int p = 0; p < NPL; ++p)
                        if ((ops & (1u << p)) && fbm == NBM) tma::prefetch_3d(&P.m3[si][p], 0, jp, bm0);
                    if ((ops & 0x10u) && fbn == NBN) tma::prefetch_3d(&P.n3[si], 0, jp, bn0);
                }
#ifdef DFT_PHASE_TIMING
                t1 = clock64(); t_c += t1 - t0; t0 = t1;
#endif
                tma::mbar_wait_relaxed(&empty[s], ph ^ 1u, (uint32_t)P.wait_ns);
#ifdef DFT_PHASE_TIMING
                t1 = clock64(); t_w += t1 - t0; t0 = t1;
#endif
                unsigned char* st = sm + s * L::STAGE_BYTES;
                uint64_t* fb = &full[s];
                if (++s == STAGES) { s = 0; ph ^= 1u; }
                if (pid == 0) tma::mbar_arrive_expect_tx(fb, tx);
#pragma unroll
                for (int p = 0; p < NPL; ++p) {
                    if (!(ops & (1u << p))) continue;
                    unsigned char* dst = st + p * L::PLANE_BYTES;
                    if (P.use3d) {
                        if (fbm == NBM) tma::load_3d(dst, &P.m3[si][p], 0, j0, bm0, fb);
                        else if (fbm > 0) tma::load_3d(dst, &P.m3l[si][p], 0, j0, bm0, fb);
                    } else {
                        for (int b = 0; b < fbm; ++b) tma::load_2d(dst + b * L::BOXB, &P.p2[si][p], m0 + 16 * b, j0, fb);
                    }
                    if (pm) tma::load_2d(dst + fbm * L::BOXB, &P.p2[si][p], m0 + 16 * fbm, j0, fb);
                }
                if (ops & 0x10u) {
                    unsigned char* dst = st + L::N_OFF;
                    if (P.use3d) {
                        if (fbn == NBN) tma::load_3d(dst, &P.n3[si], 0, j0, bn0, fb);
                        else if (fbn > 0) tma::load_3d(dst, &P.n3l[si], 0, j0, bn0, fb);
                    } else {
                        for (int b = 0; b < fbn; ++b) tma::load_2d(dst + b * L::BOXB, &P.p2[si][0], n0 + 16 * b, j0, fb);
                    }
                    if (pn) tma::load_2d(dst + fbn * L::BOXB, &P.p2[si][0], n0 + 16 * fbn, j0, fb);
                }
                if (ops & 0x20u) tma::load_1d(st + L::COEF_OFF, coef + 4 * (size_t)j0, VK * 32, fb);
            }
#ifdef DFT_PHASE_TIMING
            if (P.phase && pid == 0) {
                long long* o = P.phase + (((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (NCW + 1) + NCW) * 4;
                t1 = clock64();
                o[0] = t_w; o[1] = t_c + (t1 - t0); o[2] = 0; o[3] = t1 - t_start;
            }
#endif
        }
        return;
    }

    // ===================== MMA warps =====================
    reg_inc<REGS_CONSUMER>();
    const int wm = warp / WN, wn = warp % WN;
    const int q = lane >> 2, qcol = lane & 3;
    double acc[MF][NFN][2];
#pragma unroll
    for (int mf = 0; mf < MF; ++mf)
#pragma unroll
        for (int nf = 0; nf < NFN; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;

    const bool no_skip = P.zero_skip == 0;
    // Fragment addressing.  A DMMA k-step ks contracts 4 grid rows; lane (q, qcol) supplies row
    // krow(ks, qcol) -- rows {0,2,4,6} / {1,3,5,7} of an 8-row group, so that with the 128-byte swizzle
    // the 16 lanes of a load phase hit 16 distinct bank pairs -- and column (8-column group G) * 8 + q.
    // Group G lives in box G/2 at columns 8 (G & 1) + q, whose swizzled offset is c0 ^ ((G & 1) << 6).
    const int ga0 = wm * MF, gb0 = wn * NFN;  // first 8-column group of this warp in the M / N tile
    // M-side 8-column group of fragment mf.  Normally consecutive; in the zero-skipping 8 x 1 instance warp w
    // takes groups w and 15 - w instead of 2w and 2w + 1: two halves of one 16-column box (one or two atoms)
    // are zero or non-zero together, while groups from opposite ends of the tile average out, which evens the
    // work of the eight warps that share the ring.
    constexpr bool MIRROR = ((SKIP >= 1 && SKIP <= 3) || SKIP == 7) && WM == 8 && WN == 1 && MF == 2;
    // 4 x 2 warps on the 128 x 128 tile (round 2): M-group wm takes fragments wm, wm + 4, wm + 8, wm + 12 -- four
    // samples spread over the whole tile instead of two, so the groups' live counts per stage are closer together
    // (CPU census at C5: a stage costs max over groups = 0.75 of dense, against 0.86 for the mirrored pairs; the
    // busiest group's total is 0.65 against 0.76), and the two warps that share an SM sub-partition's tensor pipe
    // (w, w + 4) hold all the even or all the odd fragments between them.
    constexpr bool INTERLEAVE = (SKIP >= 1 && SKIP <= 3) && WM == 4 && WN == 2 && MF == 4;
    // (MIRROR: the deal comes from a table the host refreshes from the previous call's live counts -- heaviest fragment
    // with lightest, heavy pairs on the same SM sub-partition as light ones; it starts as w, 15 - w.  Which warp owns a
    // fragment does not change a single rounding: every output element is still summed over the stages in order.)
    int mg[MF];
#pragma unroll
    for (int mf = 0; mf < MF; ++mf) mg[mf] = MIRROR ? (mf == 0 ? wm : 15 - wm) : (INTERLEAVE ? wm + 4 * mf : ga0 + mf);
    if (MIRROR && P.fmap) {
#pragma unroll
        for (int mf = 0; mf < MF; ++mf) mg[mf] = (int)P.fmap[tm * 16 + 2 * wm + mf] & 15;
    }
    auto mgroup = [&](int mf) { return mg[mf]; };
    unsigned int n_live[MF];
#pragma unroll
    for (int mf = 0; mf < MF; ++mf) n_live[mf] = 0;
    uint32_t a_off[KS][MF], b_even[KS], b_odd[KS], c_off[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const int row = (VK == 16 ? 8 * (ks >> 1) : 0) + 2 * qcol + (ks & 1);
        const uint32_t c0 = (uint32_t)(row * 128 + ((((q >> 1) ^ row) & 7) << 4) + ((q & 1) << 3));
        const uint32_t pb = (uint32_t)(gb0 & 1);
#pragma unroll
        for (int mf = 0; mf < MF; ++mf) {
            const int G = mgroup(mf);
            a_off[ks][mf] = (uint32_t)((G >> 1) * L::BOXB) + (c0 ^ ((uint32_t)(G & 1) << 6));
        }
        b_even[ks] = (uint32_t)(L::N_OFF + (gb0 >> 1) * L::BOXB) + (c0 ^ (pb << 6));
        b_odd[ks] = (uint32_t)(L::N_OFF + ((gb0 >> 1) + (int)pb) * L::BOXB) + (c0 ^ ((pb ^ 1u) << 6));
        c_off[ks] = (uint32_t)(L::COEF_OFF + row * 32);
    }
    // The offsets are laundered through an opaque move.  Left to itself the compiler does not keep them: it re-derives
    // all of them from threadIdx at the top of EVERY ring stage (an S2R and ~30 integer instructions between the
    // barrier wait and the first fragment load -- on the critical path of the warp that skips least; SASS of round 1's
    // instances).  Ten registers are cheaper.
    if constexpr (SKIP != 0 || true) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int mf = 0; mf < MF; ++mf) asm volatile("mov.u32 %0, %1;" : "=r"(a_off[ks][mf]) : "r"(a_off[ks][mf]));
            asm volatile("mov.u32 %0, %1;" : "=r"(b_even[ks]) : "r"(b_even[ks]));
            asm volatile("mov.u32 %0, %1;" : "=r"(b_odd[ks]) : "r"(b_odd[ks]));
            asm volatile("mov.u32 %0, %1;" : "=r"(c_off[ks]) : "r"(c_off[ks]));
        }
    }

#ifdef DFT_PHASE_TIMING
    long long t_w = 0, t_c = 0, t_start = clock64(), t0 = t_start, t1;
#endif
    if constexpr (SKIP == 7) {
        // Software-pipelined across ring stages (round 2).  The warp that skips least is the CTA's critical path: it never
        // waits for data (its next stage is always full already), its partner on the SM sub-partition is usually
        // ahead and idle, so every latency between its DMMA bursts -- plane loads, the FP64 chain, the votes, the Phi
        // loads -- is exposed.  Here the raw operands of stage c + 1 are REQUESTED before the DMMAs of stage c issue and
        // COMBINED after them, k-step by k-step (12 doubles in flight at a time: the register file has no room for a
        // whole stage's 24), whenever a non-blocking test finds stage c + 1 already delivered; otherwise the warp
        // falls back to the blocking wait + build of the batched variant (SKIP 2).
        static_assert(KS == 2, "pipelined V instance: two k-steps per stage");
        auto build_stage = [&](uint32_t sbx, double (&ax)[KS][MF]) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const double2 ca = lds_f64x2(sbx + c_off[ks]);
                double2 cb = make_double2(0.0, 0.0);
                if (NPL == 4) cb = lds_f64x2(sbx + c_off[ks] + 16);
#pragma unroll
                for (int mf = 0; mf < MF; ++mf) {
                    const uint32_t ad = sbx + a_off[ks][mf];
                    double v = ca.x * lds_f64(ad);
                    if (NPL == 4) {
                        v = fma(ca.y, lds_f64(ad + L::PLANE_BYTES), v);
                        v = fma(cb.x, lds_f64(ad + 2 * L::PLANE_BYTES), v);
                        v = fma(cb.y, lds_f64(ad + 3 * L::PLANE_BYTES), v);
                    }
                    ax[ks][mf] = v;
                }
            }
        };
        auto vote_stage = [&](const double (&ax)[KS][MF]) {
            unsigned lv = 0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int mf = 0; mf < MF; ++mf)
                    lv |= (__any_sync(0xffffffffu, (ax[ks][mf] != 0.0) | no_skip) ? 1u : 0u) << (ks * MF + mf);
            return lv;
        };
        double a[KS][MF];
        unsigned live = 0;
        if (nchunks > 0) {
            tma::mbar_wait(&full[0], 0);
            build_stage(base, a);
            live = vote_stage(a);
        }
        for (int c = 0; c < nchunks; ++c) {
            const uint32_t s = c % STAGES;
            const uint32_t sb = base + s * L::STAGE_BYTES;
            const bool more = c + 1 < nchunks;
            const uint32_t s1 = (c + 1) % STAGES, ph1 = ((c + 1) / STAGES) & 1u;
            const uint32_t sb1 = base + s1 * L::STAGE_BYTES;
            // (all lanes must have seen the phase complete: each lane's own test is its acquire)
            const bool pre = more && __all_sync(0xffffffffu, tma::mbar_test_wait(&full[s1], ph1));
            double an[KS][MF];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                // raw operands of k-step ks of the NEXT stage: requested now, combined after this k-step's DMMAs
                double2 rca = make_double2(0.0, 0.0), rcb = make_double2(0.0, 0.0);
                double raw[MF][NPL];
                if (pre) {
                    rca = lds_f64x2(sb1 + c_off[ks]);
                    if (NPL == 4) rcb = lds_f64x2(sb1 + c_off[ks] + 16);
#pragma unroll
                    for (int mf = 0; mf < MF; ++mf)
#pragma unroll
                        for (int pl = 0; pl < NPL; ++pl) raw[mf][pl] = lds_f64(sb1 + a_off[ks][mf] + pl * L::PLANE_BYTES);
                }
                const unsigned lk = (live >> (ks * MF)) & ((1u << MF) - 1u);
                if (lk) {
                    double bf[NFN];
#pragma unroll
                    for (int nf = 0; nf < NFN; ++nf)
                        bf[nf] = lds_f64(sb + ((nf & 1) ? b_odd[ks] + (uint32_t)(((nf - 1) / 2) * L::BOXB)
                                                        : b_even[ks] + (uint32_t)((nf / 2) * L::BOXB)));
#pragma unroll
                    for (int mf = 0; mf < MF; ++mf) {
                        if (lk & (1u << mf)) {
#pragma unroll
                            for (int nf = 0; nf < NFN; ++nf) dmma::mma8x8x4(acc[mf][nf], a[ks][mf], bf[nf]);
                        }
                    }
                }
                if (pre) {
#pragma unroll
                    for (int mf = 0; mf < MF; ++mf) {
                        double v = rca.x * raw[mf][0];
                        if (NPL == 4) {
                            v = fma(rca.y, raw[mf][1], v);
                            v = fma(rcb.x, raw[mf][2], v);
                            v = fma(rcb.y, raw[mf][3], v);
                        }
                        an[ks][mf] = v;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) tma::mbar_arrive(&empty[s]);
            if (more) {
                if (!pre) {
                    tma::mbar_wait(&full[s1], ph1);
                    build_stage(sb1, an);
                }
                live = vote_stage(an);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                    for (int mf = 0; mf < MF; ++mf) a[ks][mf] = an[ks][mf];
            }
        }
    } else {
    // (ptxas software-pipelines the k-steps of a stage by itself: the fragment loads of k-step ks+1 are
    // interleaved with the DMMAs of k-step ks.  Pipelining by hand across stages costs registers and spills.)
    for (int c = 0; c < nchunks; ++c) {
        const uint32_t s = c % STAGES, ph = (c / STAGES) & 1u;
        tma::mbar_wait(&full[s], ph);
#ifdef DFT_PHASE_TIMING
        t1 = clock64(); t_w += t1 - t0; t0 = t1;
#endif
        const uint32_t sb = base + s * L::STAGE_BYTES;
        if constexpr (SKIP == 2 || SKIP == 3) {
            // Batched variant of the per-fragment votes: the A fragments of ALL the stage's k-steps are built before
            // the first vote, so the shared loads and FP64 chains of a k-step that turns out to be skipped overlap
            // with its neighbours' instead of sitting, exposed, between two branches (the loads are `asm volatile`
            // and never move across a branch by themselves).
            double a[KS][MF];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const double2 ca = lds_f64x2(sb + c_off[ks]);
                double2 cb = make_double2(0.0, 0.0);
                if (NPL == 4) cb = lds_f64x2(sb + c_off[ks] + 16);
#pragma unroll
                for (int mf = 0; mf < MF; ++mf) {
                    const uint32_t ad = sb + a_off[ks][mf];
                    double v = ca.x * lds_f64(ad);
                    if (NPL == 4) {
                        v = fma(ca.y, lds_f64(ad + L::PLANE_BYTES), v);
                        v = fma(cb.x, lds_f64(ad + 2 * L::PLANE_BYTES), v);
                        v = fma(cb.y, lds_f64(ad + 3 * L::PLANE_BYTES), v);
                    }
                    a[ks][mf] = v;
                }
            }
            // SKIP 3: the Phi fragments of the stage's FIRST k-step are requested before the votes as well (they are
            // needed whenever any of its fragments is live, three stages in four at C5), so its DMMAs can issue the
            // moment the votes are in
            double bf0[NFN];
            if constexpr (SKIP == 3) {
#pragma unroll
                for (int nf = 0; nf < NFN; ++nf)
                    bf0[nf] = lds_f64(sb + ((nf & 1) ? b_odd[0] + (uint32_t)(((nf - 1) / 2) * L::BOXB)
                                                     : b_even[0] + (uint32_t)((nf / 2) * L::BOXB)));
            }
            unsigned live = 0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int mf = 0; mf < MF; ++mf)
                    live |= (__any_sync(0xffffffffu, (a[ks][mf] != 0.0) | no_skip) ? 1u : 0u) << (ks * MF + mf);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int mf = 0; mf < MF; ++mf) n_live[mf] += (live >> (ks * MF + mf)) & 1u;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const unsigned lk = (live >> (ks * MF)) & ((1u << MF) - 1u);
                if (lk) {
                    double bf[NFN];
                    if (SKIP == 3 && ks == 0) {
#pragma unroll
                        for (int nf = 0; nf < NFN; ++nf) bf[nf] = bf0[nf];
                    } else {
#pragma unroll
                    for (int nf = 0; nf < NFN; ++nf)
                        bf[nf] = lds_f64(sb + ((nf & 1) ? b_odd[ks] + (uint32_t)(((nf - 1) / 2) * L::BOXB)
                                                        : b_even[ks] + (uint32_t)((nf / 2) * L::BOXB)));
                    }
#pragma unroll
                    for (int mf = 0; mf < MF; ++mf) {
                        if (lk & (1u << mf)) {
#pragma unroll
                            for (int nf = 0; nf < NFN; ++nf) dmma::mma8x8x4(acc[mf][nf], a[ks][mf], bf[nf]);
                        }
                    }
                }
            }
        } else {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            // A fragments: B[k][m] = a Phi + bx dxPhi + by dyPhi + bz dzPhi, built in registers
            const double2 ca = lds_f64x2(sb + c_off[ks]);
            double2 cb = make_double2(0.0, 0.0);
            if (NPL == 4) cb = lds_f64x2(sb + c_off[ks] + 16);
            double a[MF], bf[NFN];
#pragma unroll
            for (int mf = 0; mf < MF; ++mf) {
                const uint32_t ad = sb + a_off[ks][mf];
                double v = ca.x * lds_f64(ad);
                if (NPL == 4) {
                    v = fma(ca.y, lds_f64(ad + L::PLANE_BYTES), v);
                    v = fma(cb.x, lds_f64(ad + 2 * L::PLANE_BYTES), v);
                    v = fma(cb.y, lds_f64(ad + 3 * L::PLANE_BYTES), v);
                }
                a[mf] = v;
            }
            // AO screening at the tile level: if this warp's B fragment of the k-step (its MF x 8 columns x 4 grid
            // rows) is all zero -- AOs far from the points are exact zeros, zero-weight points have zero
            // coefficients -- the k-step adds nothing to this warp's rows of M: skip its Phi loads and DMMAs.
            // (The eight warps share the ring, so a warp that skips mostly waits for the one that cannot: the
            // gain here is small, unlike in the density kernel where a group's warps skip together.)
            // SKIP = false is the branch-free instance (ptxas interleaves the next k-step's loads with this one's
            // DMMAs, ~6 % faster on dense operands); the host picks it when the density kernel of the previous
            // call found (almost) nothing to skip.
            if (!SKIP) {
#pragma unroll
                for (int nf = 0; nf < NFN; ++nf)
                    bf[nf] = lds_f64(sb + ((nf & 1) ? b_odd[ks] + (uint32_t)(((nf - 1) / 2) * L::BOXB)
                                                    : b_even[ks] + (uint32_t)((nf / 2) * L::BOXB)));
#pragma unroll
                for (int mf = 0; mf < MF; ++mf)
#pragma unroll
                    for (int nf = 0; nf < NFN; ++nf) dmma::mma8x8x4(acc[mf][nf], a[mf], bf[nf]);
            } else {
                // one vote per 8-column fragment: its NFN DMMAs are skipped when it is all zero
                unsigned live = 0;
#pragma unroll
                for (int mf = 0; mf < MF; ++mf) live |= (__any_sync(0xffffffffu, (a[mf] != 0.0) | no_skip) ? 1u : 0u) << mf;
#pragma unroll
                for (int mf = 0; mf < MF; ++mf) n_live[mf] += (live >> mf) & 1u;
                if (live) {
#pragma unroll
                    for (int nf = 0; nf < NFN; ++nf)
                        bf[nf] = lds_f64(sb + ((nf & 1) ? b_odd[ks] + (uint32_t)(((nf - 1) / 2) * L::BOXB)
                                                        : b_even[ks] + (uint32_t)((nf / 2) * L::BOXB)));
#pragma unroll
                    for (int mf = 0; mf < MF; ++mf) {
                        if (live & (1u << mf)) {
#pragma unroll
                            for (int nf = 0; nf < NFN; ++nf) dmma::mma8x8x4(acc[mf][nf], a[mf], bf[nf]);
                        }
                    }
                }
            }
        }
        }   // (SKIP 0 | 1)
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(&empty[s]);
#ifdef DFT_PHASE_TIMING
        t1 = clock64(); t_c += t1 - t0; t0 = t1;
#endif
    }
    }   // (SKIP != 7)
    if (MIRROR && P.fstat && lane == 0) {
#pragma unroll
        for (int mf = 0; mf < MF; ++mf) atomicAdd(&P.fstat[tm * 16 + mgroup(mf)], n_live[mf]);
    }
    if (SKIP != 0 && P.counters && lane == 0) {   // (fragment, k-step) units executed / total: the V kernel's screening statistic
        unsigned long long done = 0;
#pragma unroll
        for (int mf = 0; mf < MF; ++mf) done += n_live[mf];
        atomicAdd(&P.counters[0], done);
        atomicAdd(&P.counters[1], (unsigned long long)nchunks * KS * MF);
    }
    // ---- partial tile out
    double* out = P.vpart + (size_t)blockIdx.y * P.mpv * P.ldv;
#pragma unroll
    for (int mf = 0; mf < MF; ++mf)
#pragma unroll
        for (int nf = 0; nf < NFN; ++nf) {
            const int r = m0 + mgroup(mf) * 8 + q;
            const int cc = n0 + (gb0 + nf) * 8 + 2 * qcol;
            *reinterpret_cast<double2*>(out + (size_t)r * P.ldv + cc) = make_double2(acc[mf][nf][0], acc[mf][nf][1]);
        }
#ifdef DFT_PHASE_TIMING
    if (lane == 0 && P.phase) {
        long long* o = P.phase + (((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (NCW + 1) + warp) * 4;
        t1 = clock64();
        o[0] = t_w; o[1] = t_c; o[2] = t1 - t0; o[3] = t1 - t_start;
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// V kernel, staged-B instance (the default where AO screening finds zeros to skip)
// ------------------------------------------------------------------------------------------------
// Same mathematics and the same 128 x 128 output tile per CTA as vxc_tma_kernel, organised so that the zero
// skipping is UNIFORM over the eight MMA warps:
//
//   TMA  -> raw ring (RS stages): the NPL plane tiles of the M columns [8 boxes][8 rows][16] + the 8 coefficient rows
//   BUILDERS (the four warps of the producer warpgroup, 80 registers each) combine the planes with the
//        points' coefficients ONCE per CTA, B = a Phi + b . grad Phi, 16 bytes per thread and step, into the MMA
//        ring in the same swizzled box layout, and publish one bit per (k-step, 8-column fragment): is anything
//        in that 8 x 4 piece of B non-zero?  (zero AOs, zero-weight and density-gated points all end up there)
//   TMA  -> MMA ring (MS stages): the Phi tile of the N columns goes straight in beside the built B
//   MMA warps, 1 x 8: warp w owns ALL 128 M columns and the 16 N columns of box w (64 accumulator doubles per
//        lane).  Every warp reads the same A fragments -- one shared load each instead of 4 loads + 4 FP64 ops
//        (18 shared loads per 32 DMMAs instead of 26) -- so every warp skips exactly the same fragments: no warp
//        waits on the ring for another one that could not skip (round 1's per-warp votes made the stage time the
//        maximum over the warps, i.e. almost nothing was gained: 11.8 ms with 45 % of the DMMAs skipped).  A warp
//        whose own N box is all zero in a stage leaves the whole stage out, which only frees the tensor pipe for
//        the warp it shares an SM sub-partition with.
// One fragment bit = 2 DMMAs behind a uniform branch (the mask is made a uniform register with a warp
// reduction), loads batched per k-step.  Nothing is cached between calls; a skipped DMMA would have added
// exact zeros, so results are bit-identical to the branch-free instance.
template <int NPL>
struct StagedCfg {
    static constexpr int VK = 8, MT = 128, NT_ = 128;
    static constexpr int BOXB = VK * 128;                    // one 16-column box: 1 KB
    static constexpr int PLANE_BYTES = (MT / 16) * BOXB;     // 8 KB
    // ring depths: the raw ring carries the memory latency (RS - 1 stages of 33 KB in flight; with 3 stages the
    // kernel ran at 22.8 ms at C5, waiting for HBM), the MMA ring only decouples builders and MMA warps
    static constexpr int RS = NPL == 4 ? 5 : 8;              // raw ring depth
    static constexpr int MS = NPL == 4 ? 3 : 6;              // MMA ring depth
    static constexpr int COEF_OFF = NPL * PLANE_BYTES;
    static constexpr int RAW_BYTES = ((COEF_OFF + VK * 32 + 1023) / 1024) * 1024;
    static constexpr int NB_OFF = PLANE_BYTES;               // MMA stage: [B 8 KB][Phi tile 8 KB][4 mask words]
    static constexpr int MASK_OFF = 2 * PLANE_BYTES;
    static constexpr int MMA_BYTES = 2 * PLANE_BYTES + 1024;
    static constexpr int MMA_OFF = RS * RAW_BYTES;
    static constexpr int BAR_OFF = MMA_OFF + MS * MMA_BYTES;  // raw_full[RS], raw_empty[RS], mma_full[MS], mma_empty[MS]
    static constexpr int TOTAL = BAR_OFF + (2 * RS + 2 * MS) * 8 + 1024;
    static_assert(TOTAL <= 232448, "shared memory");
    // 16 warps: 0-7 MMA, 8-11 builders, 12-15 TMA issue (lane 0 of each; an issuing thread must not share its warp
    // with lanes that sit in barrier waits).  setmaxnreg only redistributes what the CTA was LAUNCHED with:
    // 512 threads x 128 registers = 65536; a set that asks for more never gets it and setmaxnreg.inc waits forever.
    static constexpr int THREADS = 512;
    static constexpr int REGS_MMA = 216, REGS_BUILD = 56, REGS_ISSUE = 24;
    static_assert(NCONS * REGS_MMA + 128 * REGS_BUILD + 128 * REGS_ISSUE <= THREADS * 128, "setmaxnreg budget");
    static_assert(REGS_MMA % 8 == 0 && REGS_BUILD % 8 == 0 && REGS_ISSUE % 8 == 0, "setmaxnreg granularity");
};

__device__ __forceinline__ double2 lds_f64x2_plain(const unsigned char* p) { return *reinterpret_cast<const double2*>(p); }

template <int NPL>
__global__ void __launch_bounds__(StagedCfg<NPL>::THREADS, 1)
vxc_staged_kernel(const __grid_constant__ VxcParams P) {
    using L = StagedCfg<NPL>;
    constexpr int VK = L::VK, RS = L::RS, MS = L::MS;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (tma::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - tma::smem_u32(smem_raw));
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(sm + L::BAR_OFF);
    uint64_t* raw_empty = raw_full + RS;
    uint64_t* mma_full = raw_empty + RS;
    uint64_t* mma_empty = mma_full + MS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int tm, tn;
    if (P.lda_half) {
        int t = blockIdx.x;
        tm = 0;
        while (t >= P.tiles_n - tm) { t -= P.tiles_n - tm; ++tm; }
        tn = tm + t;
    } else {
        tm = blockIdx.x / P.tiles_n;
        tn = blockIdx.x % P.tiles_n;
    }
    const int m0 = tm * L::MT, n0 = tn * L::NT_;
    const int si = blockIdx.y / P.slices_per_sub;
    const int sl = blockIdx.y % P.slices_per_sub;
    const int total_chunks = (P.sub[si].rows + VK - 1) / VK;
    const int nchunks = total_chunks > sl ? (total_chunks - sl + P.slices_per_sub - 1) / P.slices_per_sub : 0;

    constexpr int NBUILD = 4;     // builder warps 8-11
    constexpr int NISSUE = NPL;   // issuing threads of the raw ring: lane 0 of warp 12 + p brings plane p
    if (tid == 0) {
        for (int s = 0; s < RS; ++s) { tma::mbar_init(&raw_full[s], NISSUE); tma::mbar_init(&raw_empty[s], NBUILD); }
        for (int s = 0; s < MS; ++s) { tma::mbar_init(&mma_full[s], NBUILD + 1); tma::mbar_init(&mma_empty[s], NCW); }
        tma::fence_barrier_init();
    }
    // boxes past the edge of the matrix are never written by TMA: clear both rings once
    for (int o = tid * 16; o < L::BAR_OFF; o += L::THREADS * 16) *reinterpret_cast<double2*>(sm + o) = make_double2(0.0, 0.0);
    tma::fence_proxy_async();
    __syncthreads();

    if (warp >= NCW + 4) {
        // ===================== TMA issue: lane 0 of warps 12-15, one plane each =====================
        // (Round 2's first version let lane 0 of each builder warp issue the loads: the issuing lane then shares a warp
        // with 31 lanes that sit in mbarrier.try_wait, and it ran at 1.6 us per stage -- the try_wait time limit --
        // with the DMMAs switched off.  One issuing thread for everything ran at 9.8 ms: ~63 cycles per TMA
        // instruction + a cycle per ~38 bytes.)
        reg_dec<L::REGS_ISSUE>();
        const int p = warp - NCW - 4;             // plane of this thread
        if (lane != 0 || p >= NISSUE) return;
        const int nfull = P.nfull[si], rem = P.rem[si];
        constexpr int NB = L::MT / 16;
        const int bm0 = m0 / 16, bn0 = n0 / 16;
        const int fbm = min(max(nfull - bm0, 0), NB), fbn = min(max(nfull - bn0, 0), NB);
        const bool pm = rem > 0 && nfull >= bm0 && nfull - bm0 < NB;
        const bool pn = rem > 0 && nfull >= bn0 && nfull - bn0 < NB;
        const bool with_coef = p == 0, with_n = p == NISSUE - 1;   // the Phi tile of the N columns rides with the last plane
        const uint32_t tx_raw = (uint32_t)((fbm + (pm ? 1 : 0)) * L::BOXB + (with_coef ? VK * 32 : 0));
        const uint32_t tx_n = (uint32_t)((fbn + (pn ? 1 : 0)) * L::BOXB);
        const double* coef = P.coef + 4 * (size_t)P.sub[si].coef0;
        // chunk of step c: (sl + c nsl) stride mod total, kept incrementally
        const int cstep = (int)(((long long)P.slices_per_sub * P.chunk_stride[si]) % total_chunks);
        int cidx_raw = (int)(((long long)sl * P.chunk_stride[si]) % total_chunks);
        int cidx_n = cidx_raw;
        auto next_chunk = [&](int& ci) { const int j0 = ci * VK; ci += cstep; if (ci >= total_chunks) ci -= total_chunks; return j0; };
        auto issue_raw = [&](int c) {   // plane p of raw stage c (+ the 8 coefficient rows with plane 0)
            const int j0 = next_chunk(cidx_raw);
            unsigned char* st = sm + (c % RS) * L::RAW_BYTES;
            uint64_t* fb = &raw_full[c % RS];
            tma::mbar_arrive_expect_tx(fb, tx_raw);
            unsigned char* dst = st + p * L::PLANE_BYTES;
            if (P.use3d) {
                if (fbm == NB) tma::load_3d(dst, &P.m3[si][p], 0, j0, bm0, fb);
                else if (fbm > 0) tma::load_3d(dst, &P.m3l[si][p], 0, j0, bm0, fb);
            } else {
                for (int b = 0; b < fbm; ++b) tma::load_2d(dst + b * L::BOXB, &P.p2[si][p], m0 + 16 * b, j0, fb);
            }
            if (pm) tma::load_2d(dst + fbm * L::BOXB, &P.p2[si][p], m0 + 16 * fbm, j0, fb);
            if (with_coef) tma::load_1d(st + L::COEF_OFF, coef + 4 * (size_t)j0, VK * 32, fb);
        };
        auto issue_n = [&](int c) {     // the Phi tile of the N columns goes straight into MMA stage c
            const int j0 = next_chunk(cidx_n);
            uint64_t* fb = &mma_full[c % MS];
            tma::mbar_arrive_expect_tx(fb, tx_n);
            unsigned char* dst = sm + L::MMA_OFF + (c % MS) * L::MMA_BYTES + L::NB_OFF;
            if (P.use3d) {
                if (fbn == NB) tma::load_3d(dst, &P.n3[si], 0, j0, bn0, fb);
                else if (fbn > 0) tma::load_3d(dst, &P.n3l[si], 0, j0, bn0, fb);
            } else {
                for (int b = 0; b < fbn; ++b) tma::load_2d(dst + b * L::BOXB, &P.p2[si][0], n0 + 16 * b, j0, fb);
            }
            if (pn) tma::load_2d(dst + fbn * L::BOXB, &P.p2[si][0], n0 + 16 * fbn, j0, fb);
        };
        // Both rings are kept as full as their consumers allow: whichever slot is free gets its load (non-blocking
        // tests -- a blocking wait on one ring would hold up the other)
        int c_raw = 0, c_n = with_n ? 0 : nchunks;
        uint32_t idle = 0;
        uint64_t t0 = 0;
        while (c_raw < nchunks || c_n < nchunks) {
            bool did = false;
            if (c_n < nchunks && tma::mbar_test_wait(&mma_empty[c_n % MS], ((c_n / MS) & 1) ^ 1u)) { issue_n(c_n++); did = true; }
            if (c_raw < nchunks && tma::mbar_test_wait(&raw_empty[c_raw % RS], ((c_raw / RS) & 1) ^ 1u)) { issue_raw(c_raw++); did = true; }
            if (did) { idle = 0; continue; }
            if (P.wait_ns) __nanosleep(P.wait_ns);
            if ((++idle & 4095u) == 0) {   // wall-time bound, as in tma::mbar_wait
                const uint64_t t = tma::globaltimer_ns();
                if (idle == 4096u) t0 = t;
                else if (t - t0 > DFT_MBAR_TIMEOUT_NS) tma::mbar_timeout_trap(0);
            }
        }
        return;
    }
    if (warp >= NCW) {
        // ===================== builders: warps 8-11 =====================
        reg_dec<L::REGS_BUILD>();
        const int bt = tid - NCONS;               // 0..127
        const int bw = warp - NCW;                // 0..3
        // the 16-byte granule of this thread in step i: offset (i * 128 + bt) * 16 of a plane tile, i.e. box
        // 2 i + (bt >> 6), box row (bt >> 3) & 7 -- the same row for all four, so the coefficients are loaded once
        const int row = (bt >> 3) & 7;
        const int half = (((bt & 7) ^ row) >> 2) & 1;     // un-swizzled 16-byte chunk >> 2: lower / upper 8 columns
        // k-step 0 takes box rows {0,2,4,6}, k-step 1 rows {1,3,5,7}
        const uint32_t bit0 = 1u << ((row & 1) * 16 + 2 * (bt >> 6) + half);
        for (int c = 0; c < nchunks; ++c) {
            const int rs = c % RS, ms = c % MS;
            tma::mbar_wait(&mma_empty[ms], ((c / MS) & 1) ^ 1u);
            tma::mbar_wait(&raw_full[rs], (c / RS) & 1);
            const unsigned char* rst = sm + rs * L::RAW_BYTES;
            unsigned char* mst = sm + L::MMA_OFF + ms * L::MMA_BYTES;
            const double2 ca = lds_f64x2_plain(rst + L::COEF_OFF + row * 32);
            double2 cb = make_double2(0.0, 0.0);
            if (NPL == 4) cb = lds_f64x2_plain(rst + L::COEF_OFF + row * 32 + 16);
            uint32_t bits = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int o = (i * 128 + bt) * 16;
                const double2 f0 = lds_f64x2_plain(rst + o);
                double2 v = make_double2(ca.x * f0.x, ca.x * f0.y);
                if (NPL == 4) {
                    const double2 f1 = lds_f64x2_plain(rst + L::PLANE_BYTES + o);
                    const double2 f2 = lds_f64x2_plain(rst + 2 * L::PLANE_BYTES + o);
                    const double2 f3 = lds_f64x2_plain(rst + 3 * L::PLANE_BYTES + o);
                    v.x = fma(ca.y, f1.x, v.x); v.y = fma(ca.y, f1.y, v.y);
                    v.x = fma(cb.x, f2.x, v.x); v.y = fma(cb.x, f2.y, v.y);
                    v.x = fma(cb.y, f3.x, v.x); v.y = fma(cb.y, f3.y, v.y);
                }
                *reinterpret_cast<double2*>(mst + o) = v;
                if ((v.x != 0.0) | (v.y != 0.0)) bits |= bit0 << (4 * i);
            }
            bits = __reduce_or_sync(0xffffffffu, bits);
            if (lane == 0) *reinterpret_cast<uint32_t*>(mst + L::MASK_OFF + 4 * bw) = bits;
            __syncwarp();
            if (lane == 0) {
                tma::mbar_arrive(&mma_full[ms]);
                tma::mbar_arrive(&raw_empty[rs]);
            }
        }
        return;
    }

    // ===================== MMA warps, 1 x 8 =====================
    reg_inc<L::REGS_MMA>();
    const int q = lane >> 2, qcol = lane & 3;
    double acc[16][2][2];
#pragma unroll
    for (int g = 0; g < 16; ++g)
#pragma unroll
        for (int nf = 0; nf < 2; ++nf) acc[g][nf][0] = acc[g][nf][1] = 0.0;
    // fragment addressing (same maps as vxc_tma_kernel): k-step ks contracts box rows 2 qcol + ks; lane (q, qcol)
    // supplies column 8 (G & 1) + q of box G >> 1, whose swizzled offset is c0 ^ ((G & 1) << 6)
    // (the four offsets are laundered through an opaque move: left to itself the compiler re-derives them from
    // threadIdx in front of every load -- three logic instructions per fragment)
    uint32_t c0[2][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const int row = 2 * qcol + ks;
        const uint32_t v = (uint32_t)(row * 128 + ((((q >> 1) ^ row) & 7) << 4) + ((q & 1) << 3));
        asm volatile("mov.u32 %0, %1;" : "=r"(c0[ks][0]) : "r"(v));
        asm volatile("mov.u32 %0, %1;" : "=r"(c0[ks][1]) : "r"(v ^ 64u));
    }
    const uint32_t nb_off = (uint32_t)(L::NB_OFF + warp * L::BOXB);
    const bool no_skip = P.zero_skip == 0;
    unsigned int n_done = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int ms = c % MS;
        tma::mbar_wait(&mma_full[ms], (c / MS) & 1);
        const uint32_t sb = base + L::MMA_OFF + ms * L::MMA_BYTES;
        uint32_t m0w, m1w, m2w, m3w;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(m0w), "=r"(m1w), "=r"(m2w), "=r"(m3w) : "r"(sb + L::MASK_OFF));
        uint32_t m = (m0w | m1w) | (m2w | m3w);
        // this warp's Phi fragments of both k-steps (B operand of the DMMAs)
        double bf[2][2];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            bf[ks][0] = lds_f64(sb + nb_off + c0[ks][0]);
            bf[ks][1] = lds_f64(sb + nb_off + c0[ks][1]);
        }
        const bool nz = ((bf[0][0] != 0.0) | (bf[0][1] != 0.0)) | ((bf[1][0] != 0.0) | (bf[1][1] != 0.0));
        if (no_skip) m = 0xffffffffu;
#ifdef DFT_DIAGNOSTICS
        if (P.debug_nodmma) m = 0u;              // delivery + build floor (results are wrong)
#endif
        m = __reduce_or_sync(0xffffffffu, m);   // (all lanes hold the same word: this makes it a UNIFORM register)
        if (__any_sync(0xffffffffu, nz | no_skip) && m != 0u) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint32_t mk = (m >> (16 * ks)) & 0xffffu;
                if (mk == 0u) continue;
                double a[16];
#pragma unroll
                for (int g = 0; g < 16; ++g)
                    if (mk & (1u << g)) a[g] = lds_f64(sb + (uint32_t)((g >> 1) * L::BOXB) + c0[ks][g & 1]);
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                    // A REAL branch around the fragment's two DMMAs: ptxas if-converts a plain `if` of this size into
                    // predicated DMMAs, and a predicated-off DMMA still pays its statically scheduled issue stall
                    // (measured in round 1: 15.3 against 12.4 ms).  A loop with a run-time trip count of 0 or 1
                    // cannot be if-converted.
                    int rep;   // (opaque to the compiler, or it proves rep <= 1 and turns the loop back into an `if`)
                    asm volatile("bfe.u32 %0, %1, %2, 1;" : "=r"(rep) : "r"(mk), "r"(g));
#pragma unroll 1
                    for (int r = 0; r < rep; ++r) {
                        dmma::mma8x8x4(acc[g][0], a[g], bf[ks][0]);
                        dmma::mma8x8x4(acc[g][1], a[g], bf[ks][1]);
                    }
                }
                n_done += __popc(mk);
            }
        }
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(&mma_empty[ms]);
    }
    // ---- partial tile out
    double* out = P.vpart + (size_t)blockIdx.y * P.mpv * P.ldv;
#pragma unroll
    for (int g = 0; g < 16; ++g)
#pragma unroll
        for (int nf = 0; nf < 2; ++nf) {
            const int r = m0 + g * 8 + q;
            const int cc = n0 + warp * 16 + nf * 8 + 2 * qcol;
            *reinterpret_cast<double2*>(out + (size_t)r * P.ldv + cc) = make_double2(acc[g][nf][0], acc[g][nf][1]);
        }
    if (lane == 0 && P.counters) {   // (fragment, k-step) units executed / total: the V kernel's screening statistic
        atomicAdd(&P.counters[0], (unsigned long long)n_done);
        atomicAdd(&P.counters[1], (unsigned long long)nchunks * 32ull);
    }
}

// out[i][j] = sum over slices of T(i+s, j+s) + T(j+s, i+s), s = column shift of the slice's
// sub-problem; T = M where the tile was computed (lda_half: the mirror tile otherwise).
// One CTA per 32 x 32 output tile: thread (row, tx) sums, over the slices in a fixed order, the partial it
// reads DIRECTLY, T(i0 + row, j0 + tx), and the one whose TRANSPOSE the tile needs, T(j0 + row, i0 + tx) --
// both are coalesced along tx -- and the transposed sums change hands through shared memory.  (Round 1 read
// T(j, i) with a warp per output element: 32 x 8 bytes from 32 different rows per request, 48 us at nao 377.)
// Fixed summation order -> bit-reproducible and exactly symmetric.  `raw` (GGA, option "raw_convention"):
// out = 2 sum T(i, j), the reference's own unsymmetrised B^T Phi (dft_solver.cu:616).
constexpr int FIN_THREADS = 256;
constexpr int FIN_TILE = 32;
__global__ void __launch_bounds__(FIN_THREADS)
finalize_tma_kernel(int nao, int ldv, int mpv, int NT, int nsub, int slices_per_sub, int shift1, int lda_half, int raw,
                    const double* __restrict__ vpart, double* __restrict__ vxc, int nepart,
                    const double* __restrict__ epart, double* __restrict__ d_exc) {
    __shared__ double tr[FIN_TILE][FIN_TILE + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int ntile = (nao + FIN_TILE - 1) / FIN_TILE;
    const int i0 = (blockIdx.x / ntile) * FIN_TILE, j0 = (blockIdx.x % ntile) * FIN_TILE;
    const size_t ss = (size_t)mpv * ldv;
    // sum over the slices of sub-problem `p` at element q, two chains, fixed order
    auto slice_sum = [&](const double* q) {
        double s0 = 0.0, s1 = 0.0;
        int sl = 0;
        for (; sl + 1 < slices_per_sub; sl += 2) { s0 += __ldg(q + sl * ss); s1 += __ldg(q + (sl + 1) * ss); }
        if (sl < slices_per_sub) s0 += __ldg(q + sl * ss);
        return s0 + s1;
    };
    // T(a, b) of a shifted index pair exists unless only the upper block triangle was computed and (a, b) is below it
    auto computed = [&](int a, int b) { return !(lda_half && a / NT > b / NT); };
    double out[4] = {0.0, 0.0, 0.0, 0.0};
    for (int su = 0; su < nsub; ++su) {
        const int sh = su ? shift1 : 0;
        const double* p = vpart + (size_t)su * slices_per_sub * ss;
        double d[4];
        if (su) __syncthreads();   // (tr is reused)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = ty + 8 * r;
            d[r] = 0.0;
            {   // direct: T(I + sh, J + sh), (I, J) = (i0 + row, j0 + tx)
                const int I = i0 + row + sh, J = j0 + tx + sh;
                if (i0 + row < nao && j0 + tx < nao && computed(I, J)) d[r] = slice_sum(p + (size_t)I * ldv + J);
            }
            double t = 0.0;
            if (!raw) {   // for the transpose: T(I' + sh, J' + sh), (I', J') = (j0 + row, i0 + tx)
                const int I = j0 + row + sh, J = i0 + tx + sh;
                if (j0 + row < nao && i0 + tx < nao && computed(I, J)) t = slice_sum(p + (size_t)I * ldv + J);
            }
            tr[row][tx] = t;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = ty + 8 * r;
            const int I = i0 + row + sh, J = j0 + tx + sh;
            const double tt = tr[tx][row];   // sum T(J, I) of this thread's output element (I, J)
            if (raw) out[r] += 2.0 * d[r];
            else if (computed(I, J) && computed(J, I)) out[r] += d[r] + tt;
            else out[r] += 2.0 * (d[r] + tt);   // (exactly one of the two exists; the other term is 0)
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = ty + 8 * r;
        if (i0 + row < nao && j0 + tx < nao) vxc[(size_t)(i0 + row) * nao + j0 + tx] = out[r];
    }
    if (blockIdx.x == 0) {
        __shared__ double sh[FIN_THREADS];
        double e = 0.0;
        for (int k = threadIdx.x; k < nepart; k += FIN_THREADS) e += epart[k];
        sh[threadIdx.x] = e;
        __syncthreads();
        for (int o = FIN_THREADS / 2; o > 0; o >>= 1) {
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
    // (function-local static with an initialiser: initialised once, thread-safely -- the fan-out's worker threads build
    // their devices' plans concurrently)
    static const EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeTiledFn>(p);
        fprintf(stderr, "[dft_b200] cuTensorMapEncodeTiled not available from the driver\n");
        return static_cast<EncodeTiledFn>(nullptr);
    }();
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

// 3-D f64 map {16 columns, rows, nblk 16-column blocks} over the whole blocks of a row-major array
// (strides: row pitch, 128 bytes), box {16, box_rows, box_blks}, 128-byte swizzle.
static bool make_map3(CUtensorMap* m, const void* ptr, uint64_t nblk, uint64_t rows, uint64_t pitch_elems,
                      uint32_t box_rows, uint32_t box_blks) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[3] = {16, rows, nblk};
    cuuint64_t gstride[2] = {pitch_elems * 8, 128};
    cuuint32_t box[3] = {16, box_rows, box_blks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Tensor map of plane `ptr` for sub-problem (parity) si of an (ngrid x nao) array: see file header.
static bool make_sub_map(CUtensorMap* m, const double* ptr, int ngrid, int nao, bool split, int si, uint32_t box_rows) {
    if (!split) return make_map(m, ptr, (uint64_t)nao, (uint64_t)ngrid, (uint64_t)nao, box_rows);
    if (si == 0) return make_map(m, ptr, (uint64_t)nao, (uint64_t)(ngrid + 1) / 2, 2ull * nao, box_rows);
    return make_map(m, ptr + (nao - 1), (uint64_t)nao + 1, (uint64_t)ngrid / 2, 2ull * nao, box_rows);
}
static bool make_sub_map3(CUtensorMap* m, const double* ptr, int ngrid, int nao, bool split, int si, uint32_t box_rows,
                          uint32_t box_blks) {
    if (!split) return make_map3(m, ptr, (uint64_t)nao / 16, (uint64_t)ngrid, (uint64_t)nao, box_rows, box_blks);
    if (si == 0) return make_map3(m, ptr, (uint64_t)nao / 16, (uint64_t)(ngrid + 1) / 2, 2ull * nao, box_rows, box_blks);
    return make_map3(m, ptr + (nao - 1), ((uint64_t)nao + 1) / 16, (uint64_t)ngrid / 2, 2ull * nao, box_rows, box_blks);
}

struct Geometry {
    bool split;
    int nsub, ncols;
    SubProblem sub[2];
    int nblocks, coef_rows;
};

static Geometry make_geometry(int ngrid, int nao) {
    Geometry g;
    memset(&g, 0, sizeof(g));
    g.split = (nao & 1) != 0;
    g.nsub = (g.split && ngrid >= 2) ? 2 : 1;
    g.ncols = nao + (g.split ? 1 : 0);  // widest sub-problem
    if (!g.split) {
        g.sub[0] = SubProblem{ngrid, 1, 0, 0, 0, 0};
    } else {
        g.sub[0] = SubProblem{(ngrid + 1) / 2, 2, 0, 0, 0, 0};
        g.sub[1] = SubProblem{ngrid / 2, 2, 1, 1, 0, 0};
    }
    for (int s = 0; s < g.nsub; ++s) {
        g.sub[s].blk0 = g.nblocks;
        g.sub[s].coef0 = g.coef_rows;
        // (an even number of 64-row blocks per sub-problem: the wide density kernel takes them in pairs, and a pair must
        // not straddle the two sub-problems; a padding block is all zero rows)
        const int nb = 2 * ((g.sub[s].rows + 2 * MB - 1) / (2 * MB));
        g.nblocks += nb;
        g.coef_rows += nb * MB;
    }
    return g;
}

// ---- launch plan -------------------------------------------------------------------------------
// Everything a call needs besides the data itself: geometry, ~40 encoded tensor maps, kernel
// instances and launch shapes.  An SCF loop calls with the same arrays every iteration
// (dft.py:155-176 keeps the AO planes resident), so the plan of the previous call is kept in the
// engine context and reused when pointers, shapes and options are unchanged; tensor maps only hold
// addresses, so a plan stays valid when the caller rewrites the CONTENTS of its arrays.
struct PlanKey {
    Problem prob;
    int exact, l2_prefetch, tma_3d, vxc_shape, vxc_vk, zero_skip, vxc_skip_on, vxc_skip_mode, vxc_scatter, vxc_producers, debug_nodmma, wait_ns, dyn_sched, stagger_min, density_unit, vxc_prefetch, vxc_rebalance, density_producers, density_scatter, density_wide;
    const void *dsym, *coef, *epart, *vpart, *rho;  // engine workspaces (grow-only: may move when they grow)
};

struct Plan {
    bool valid = false;
    PlanKey key;
    DensityParams dp;
    PointParams pp;
    VxcParams vp;
    const void* dfunc = nullptr;
    const void* vfunc = nullptr;
    int dgrid = 0, dsmem = 0, vsmem = 0, pgrid = 0, vthreads = NTHREADS, v_tiles_m = 0, dgroups = 2;
    dim3 vgrid;
    // symmetrize_pad / finalize arguments
    int KP = 0, NP = 0, nsub = 0, ldv = 0, mpv = 0, fin_nt = 0, nsl = 0, shift1 = 0, lda_half = 0;
    double *dsym = nullptr, *epart = nullptr, *vpart = nullptr;
};

template <int NF2, int NPL, bool WIDE>
static void plan_density(CublasHandleWrapper* ctx, const Problem& p, const Geometry& g, int nsm, double* coef, Plan& pl) {
    using DL = DensitySmem<NF2, NPL, WIDE>;
    constexpr int NT = DL::NT;
    const int ngrid = p.ngrid, nao = p.nao;
    const int ntiles = (g.ncols + NT - 1) / NT;
    const int NP = ntiles * NT;
    const int KP = ((g.ncols + 15) / 16) * 16;
    const int ublocks = WIDE ? g.nblocks / 2 : g.nblocks;                     // unit blocks of DL::RB rows
    const int grid1 = (ublocks + DL::GRPS - 1) / DL::GRPS < nsm ? (ublocks + DL::GRPS - 1) / DL::GRPS : nsm;  // GRPS consumer groups per CTA
    const int pgrid = (g.coef_rows + POINT_THREADS - 1) / POINT_THREADS;

    double* dsym = (double*)ctx->dsym.ensure(sizeof(double) * (size_t)g.nsub * NP * KP, &ctx->failed);
    double* epart = (double*)ctx->epart.ensure(sizeof(double) * pgrid, &ctx->failed);
    // unit of work (see the kernel): a column tile of a block whenever there are several tiles -- measured faster
    // at every size (C5: 9.35 -> 9.31 ms on the whole grid, 1.305 -> 1.265 ms on an eighth of it)
    const bool per_tile = ntiles > 1 && ctx->density_unit != 1;
    const int nparts = per_tile ? 2 * ntiles : 2;
    double* rho = (double*)ctx->rho.ensure(sizeof(double) * 4 * (size_t)nparts * g.coef_rows, &ctx->failed);
    if (ctx->failed) return;

    DensityParams& dp = pl.dp;
    memset(&dp, 0, sizeof(dp));
    const double* planes[4] = {p.ao, p.gx, p.gy, p.gz};
    bool ok = make_map(&dp.map_d, dsym, (uint64_t)KP, (uint64_t)g.nsub * NP, (uint64_t)KP, NT);
    for (int s = 0; s < g.nsub; ++s) {
        ok = ok && make_sub_map(&dp.map_a[s], p.ao, ngrid, nao, g.split, s, DL::RB);
        for (int i = 0; i < 4; ++i)
            ok = ok && make_sub_map(&dp.map_e[s][i], planes[i < NPL ? i : 0], ngrid, nao, g.split, s, DL::RB);
    }
    if (!ok) { ctx->failed = true; return; }
    dp.sub[0] = g.sub[0]; dp.sub[1] = g.sub[1];
    dp.nsub = g.nsub;
    dp.nblocks = ublocks; dp.ntiles = ntiles; dp.nk = KP / 16; dp.NP = NP;
    dp.l2_prefetch = ctx->l2_prefetch ? 1 : 0;
    dp.zero_skip = ctx->zero_skip ? 1 : 0;
    dp.debug_nodmma = ctx->debug_nodmma; dp.wait_ns = ctx->wait_ns; dp.stagger_min = ctx->stagger_min;
    dp.per_tile = per_tile ? 1 : 0;
    dp.producers2 = ctx->density_producers >= 2 ? 1 : 0;
    dp.unit_stride = 1;
    if (ctx->density_scatter && ublocks > 64) {
        long long st = (long long)(0.6180339887 * (double)ublocks) | 1;
        auto gcd = [](long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; };
        while (gcd(st, ublocks) != 1) st += 2;
        dp.unit_stride = (int)(st % ublocks);
    }
    dp.coef_rows = g.coef_rows; dp.rho = rho;
    // (the counters must exist before `sched` is derived from them: round 1 had these two statements the other
    // way round, so `sched` was always null and the dynamic deal never ran)
    dp.counters = reinterpret_cast<unsigned long long*>(ctx->counters.ensure(COUNTERS_BYTES, &ctx->failed));
    if (ctx->failed) return;
    dp.sched = ctx->dyn_sched ? reinterpret_cast<unsigned int*>(dp.counters + 4) : nullptr;

    PointParams& pp = pl.pp;
    memset(&pp, 0, sizeof(pp));
    pp.sub[0] = g.sub[0]; pp.sub[1] = g.sub[1];
    pp.nsub = g.nsub; pp.coef_rows = g.coef_rows; pp.nparts = nparts;
    pp.xc_mode = p.xc_type == 2 ? 4 : p.xc_type * 2 + (ctx->exact_functionals ? 1 : 0);
    pp.rho = rho; pp.w = p.w; pp.coef = coef; pp.exc_part = epart;
    pl.pgrid = pgrid;
#ifdef DFT_PHASE_TIMING
    dp.phase = (long long*)ctx->scratch.ensure(sizeof(long long) * (65536 + 160 * 256), &ctx->failed);
    if (dp.phase) cudaMemsetAsync(dp.phase, 0, sizeof(long long) * 8192, ctx->stream);
#endif

    pl.dgroups = DL::GRPS;
    auto dk = density_tma_kernel<NF2, NPL, WIDE>;
    DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(dk, cudaFuncAttributeMaxDynamicSharedMemorySize, DL::TOTAL));
    pl.dfunc = reinterpret_cast<const void*>(dk);
    pl.dgrid = grid1; pl.dsmem = DL::TOTAL;
    pl.KP = KP; pl.NP = NP; pl.nsub = g.nsub; pl.dsym = dsym; pl.epart = epart;
}

// ---- V plan: output tile (WM MF 8) x (WN NFN 8)
template <int MF, int NFN, int WM, int WN, int NPL, int VK, int STAGES, int SKIP>
static void plan_vxc(CublasHandleWrapper* ctx, const Problem& p, const Geometry& g, int nsm, const double* coef, Plan& pl) {
    using VL = VxcCfg<MF, NFN, WM, WN, NPL, VK, STAGES>;
    const int ngrid = p.ngrid, nao = p.nao;
    const int tiles_m = (g.ncols + VL::MT - 1) / VL::MT, tiles_n = (g.ncols + VL::NT_ - 1) / VL::NT_;
    const int lda_half = (NPL == 1 && VL::MT == VL::NT_) ? 1 : 0;
    const int tiles = lda_half ? tiles_n * (tiles_n + 1) / 2 : tiles_m * tiles_n;
    const int mpv = tiles_m * VL::MT, ldv = tiles_n * VL::NT_;

    const int maxrows = g.sub[0].rows;
    int nsl = nsm / (tiles * g.nsub);
    if (nsl < 1) nsl = 1;
    const int max_slices = (maxrows + VK - 1) / VK;
    if (nsl > max_slices) nsl = max_slices;
    int rows_per_slice = (maxrows + nsl - 1) / nsl;
    rows_per_slice = ((rows_per_slice + VK - 1) / VK) * VK;
    nsl = (maxrows + rows_per_slice - 1) / rows_per_slice;
    double* vpart = (double*)ctx->vpart.ensure(sizeof(double) * (size_t)g.nsub * nsl * mpv * ldv, &ctx->failed);
    if (ctx->failed) return;

    VxcParams& vp = pl.vp;
    memset(&vp, 0, sizeof(vp));
    const double* planes[4] = {p.ao, p.gx, p.gy, p.gz};
    bool ok = true, ok3 = ctx->tma_3d;
    for (int s = 0; s < g.nsub; ++s) {
        const int cols = g.split ? (s == 0 ? nao : nao + 1) : nao;
        vp.nfull[s] = cols / 16;
        vp.rem[s] = cols % 16;
        // whole blocks in the last M / N tile of this sub-problem (where that tile is partial)
        const int lm = vp.nfull[s] - (tiles_m - 1) * (VL::MT / 16), ln = vp.nfull[s] - (tiles_n - 1) * (VL::NT_ / 16);
        for (int i = 0; i < 4; ++i) {
            const double* plane = planes[i < NPL ? i : 0];
            ok = ok && make_sub_map(&vp.p2[s][i], plane, ngrid, nao, g.split, s, VK);
            if (ok3 && vp.nfull[s] > 0) {
                ok3 = ok3 && make_sub_map3(&vp.m3[s][i], plane, ngrid, nao, g.split, s, VK, VL::MT / 16);
                if (lm > 0 && lm < VL::MT / 16) ok3 = ok3 && make_sub_map3(&vp.m3l[s][i], plane, ngrid, nao, g.split, s, VK, lm);
            }
        }
        if (ok3 && vp.nfull[s] > 0) {
            ok3 = ok3 && make_sub_map3(&vp.n3[s], p.ao, ngrid, nao, g.split, s, VK, VL::NT_ / 16);
            if (ln > 0 && ln < VL::NT_ / 16) ok3 = ok3 && make_sub_map3(&vp.n3l[s], p.ao, ngrid, nao, g.split, s, VK, ln);
        }
    }
    if (!ok) { ctx->failed = true; return; }
    for (int s = 0; s < g.nsub; ++s) {
        // golden-ratio stride, made coprime to the chunk count: consecutive ring stages of a CTA then come from
        // different atoms' grids, which decorrelates the per-stage work of the eight warps when zero fragments
        // are skipped (neighbouring chunks have the same zero pattern)
        const long long T = (g.sub[s].rows + VK - 1) / VK;
        long long st = 1;
        if (ctx->zero_skip && ctx->vxc_skip_on && ctx->vxc_scatter && T > 16) {
            st = (long long)(0.6180339887 * (double)T) | 1;
            auto gcd = [](long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; };
            while (gcd(st, T) != 1) st += 2;
        }
        vp.chunk_stride[s] = (int)st;
    }
    vp.use3d = ok3 ? 1 : 0;
    vp.debug_nodmma = ctx->debug_nodmma; vp.wait_ns = ctx->wait_ns;
    vp.producers = ctx->vxc_producers;
    vp.prefetch = ctx->vxc_prefetch;
    vp.zero_skip = ctx->zero_skip ? 1 : 0;  // (a driver that rejects the 3-D form leaves the per-block 2-D loads)
    vp.sub[0] = g.sub[0]; vp.sub[1] = g.sub[1];
    vp.nsub = g.nsub; vp.tiles_m = tiles_m; vp.tiles_n = tiles_n; vp.lda_half = lda_half;
    vp.rows_per_slice = rows_per_slice; vp.slices_per_sub = nsl; vp.ldv = ldv; vp.mpv = mpv;
    vp.coef = coef; vp.vpart = vpart;
    vp.counters = pl.dp.counters ? pl.dp.counters + 2 : nullptr;
    if (pl.dp.counters && tiles_m <= 16) {   // fragment deal of the 8 x 1 per-warp-vote instances (see the kernel)
        unsigned char* cb = reinterpret_cast<unsigned char*>(pl.dp.counters);
        vp.fmap = cb + FMAP_OFF;
        vp.fstat = ctx->vxc_rebalance ? reinterpret_cast<unsigned int*>(cb + FSTAT_OFF) : nullptr;
    }
    pl.v_tiles_m = tiles_m;
#ifdef DFT_PHASE_TIMING
    {   // the density kernel's phase record occupies the first 148*8*4 entries of `scratch`
        long long* ph = (long long*)ctx->scratch.ensure(sizeof(long long) * (65536 + 160 * 256), &ctx->failed);
        vp.phase = ph ? ph + 8192 : nullptr;
    }
#endif

    if constexpr (SKIP == 4) {   // staged-B instance (128 x 128 tile, 8 rows per stage): same maps, its own kernel
        static_assert(VL::MT == 128 && VL::NT_ == 128 && VK == 8, "staged V kernel shape");
        auto vk = vxc_staged_kernel<NPL>;
        DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(vk, cudaFuncAttributeMaxDynamicSharedMemorySize, StagedCfg<NPL>::TOTAL));
        pl.vfunc = reinterpret_cast<const void*>(vk);
        pl.vsmem = StagedCfg<NPL>::TOTAL;
        pl.vthreads = StagedCfg<NPL>::THREADS;
    } else {
        auto vk = vxc_tma_kernel<MF, NFN, WM, WN, NPL, VK, STAGES, SKIP>;
        DFT_CUDA_CHECK(ctx, cudaFuncSetAttribute(vk, cudaFuncAttributeMaxDynamicSharedMemorySize, VL::TOTAL));
        pl.vfunc = reinterpret_cast<const void*>(vk);
        pl.vsmem = VL::TOTAL;
        pl.vthreads = NTHREADS;
    }
    pl.vgrid = dim3(tiles, nsl * g.nsub);
    pl.ldv = ldv; pl.mpv = mpv; pl.fin_nt = VL::NT_; pl.nsl = nsl; pl.shift1 = g.split ? 1 : 0; pl.lda_half = lda_half;
    pl.vpart = vpart;
}

template <int NPL>
static void build_plan(CublasHandleWrapper* ctx, const Problem& p, int nsm, Plan& pl) {
    const Geometry g = make_geometry(p.ngrid, p.nao);
    double* coef = (double*)ctx->coef.ensure(sizeof(double) * 4 * (size_t)g.coef_rows, &ctx->failed);
    if (ctx->failed) return;

    // density column tile: the NF in 1..5 with the least padded width, larger tiles on ties
    int best_nf = 1, best_np = 1 << 30;
    for (int nf = 1; nf <= 5; ++nf) {
        const int nt = 32 * nf, np = ((g.ncols + nt - 1) / nt) * nt;
        if (np <= best_np) { best_np = np; best_nf = nf; }
    }
#define DFT_PLAN_D(NF2_) do { if (ctx->density_wide) plan_density<NF2_, NPL, true>(ctx, p, g, nsm, coef, pl); \
                              else plan_density<NF2_, NPL, false>(ctx, p, g, nsm, coef, pl); } while (0)
    switch (best_nf) {
        case 1: DFT_PLAN_D(2); break;
        case 2: DFT_PLAN_D(4); break;
        case 3: DFT_PLAN_D(6); break;
        case 4: DFT_PLAN_D(8); break;
        default: DFT_PLAN_D(10); break;
    }
#undef DFT_PLAN_D
    if (ctx->failed) return;

    // V output tile: 64 x 64 for narrow matrices, else the cheaper of 128 x 128 and 160 x 80
    // (the latter pays ~15 % more shared-memory and FP64 work per flop)
    const int n = g.ncols;
    auto pad = [](int x, int t) { return ((x + t - 1) / t) * t; };
    const double cost128 = (double)pad(n, 128) * pad(n, 128);
    const double cost160 = 1.15 * (double)pad(n, 160) * pad(n, 80);
    int shape = ctx->vxc_shape;
    if (shape == 0) shape = n <= 64 ? 64 : (cost160 < cost128 ? 160 : 128);
    // zero-skipping V instance only where the density kernel of the previous call actually skipped work
    const bool vskip = ctx->zero_skip && ctx->vxc_skip_on;
#define DFT_PLAN_V(...) do { if (vskip) plan_vxc<__VA_ARGS__, 1>(ctx, p, g, nsm, coef, pl); \
                             else plan_vxc<__VA_ARGS__, 0>(ctx, p, g, nsm, coef, pl); } while (0)
    // 128 x 128 tile with zero skipping: the staged-B instance (uniform skipping; vxc_skip_mode 4, the default) or
    // round 1's per-warp M-side votes (vxc_skip_mode 1, kept for comparison)
#define DFT_PLAN_V128(VK_, ST_) do { \
        if (!vskip) plan_vxc<2, 16, 8, 1, NPL, VK_, ST_, 0>(ctx, p, g, nsm, coef, pl); \
        else plan_vxc<2, 16, 8, 1, NPL, VK_, ST_, 1>(ctx, p, g, nsm, coef, pl); } while (0)
    if (shape == 64) {
        DFT_PLAN_V(1, 8, 8, 1, NPL, 16, 4);
    } else if (shape == 160) {
        DFT_PLAN_V(5, 5, 4, 2, NPL, 16, 2);
    } else if (vskip && ctx->vxc_skip_mode == 4) {
        plan_vxc<2, 16, 8, 1, NPL, 8, 5, 4>(ctx, p, g, nsm, coef, pl);
    } else if (vskip && ctx->vxc_skip_mode == 2) {   // per-warp votes, built and voted on for a whole ring stage at once (default)
        plan_vxc<2, 16, 8, 1, NPL, 8, 5, 2>(ctx, p, g, nsm, coef, pl);
#ifdef DFT_V_EXPERIMENTS
    // Variants that were measured and are NOT faster (profiles/r2_u2_*, r2_u6_*): compiled only into the
    // diagnostic build (build.py --diag) so that the product library does not carry them.
    } else if (vskip && ctx->vxc_skip_mode == 7) {   // mode 2 software-pipelined across ring stages
        plan_vxc<2, 16, 8, 1, NPL, 8, 5, 7>(ctx, p, g, nsm, coef, pl);
    } else if (vskip && ctx->vxc_skip_mode == 3) {   // mode 2 + the first k-step's Phi fragments requested before the votes
        plan_vxc<2, 16, 8, 1, NPL, 8, 5, 3>(ctx, p, g, nsm, coef, pl);
    } else if (vskip && ctx->vxc_skip_mode == 5) {   // 4 x 2 warps, interleaved M fragments, vote per k-step
        plan_vxc<4, 8, 4, 2, NPL, 8, 5, 1>(ctx, p, g, nsm, coef, pl);
    } else if (vskip && ctx->vxc_skip_mode == 6) {   // 4 x 2 warps, interleaved M fragments, votes batched per stage
        plan_vxc<4, 8, 4, 2, NPL, 8, 5, 2>(ctx, p, g, nsm, coef, pl);
#endif
    } else {
        // rows per ring stage: 16 (2 stages, fewer barriers) on dense operands; 8 (5 stages) when zero fragments
        // are skipped -- with the stages scattered over the grid, the deeper ring lets the warps drift apart
        const int vk = ctx->vxc_vk ? ctx->vxc_vk : (vskip ? 8 : 16);
        if (vk == 16) DFT_PLAN_V128(16, 2);
        else DFT_PLAN_V128(8, 5);
    }
#undef DFT_PLAN_V128
#undef DFT_PLAN_V
}

static PlanKey make_key(const CublasHandleWrapper* ctx, const Problem& p) {
    PlanKey k;
    memset(&k, 0, sizeof(k));  // (padding bytes too: keys are compared with memcmp)
    k.prob.xc_type = p.xc_type; k.prob.ngrid = p.ngrid; k.prob.nao = p.nao;
    k.prob.dm = p.dm; k.prob.ao = p.ao; k.prob.gx = p.gx; k.prob.gy = p.gy; k.prob.gz = p.gz; k.prob.w = p.w;
    k.prob.vxc = p.vxc; k.prob.d_exc = p.d_exc;
    k.exact = ctx->exact_functionals; k.l2_prefetch = ctx->l2_prefetch; k.tma_3d = ctx->tma_3d;
    k.vxc_shape = ctx->vxc_shape; k.vxc_vk = ctx->vxc_vk; k.zero_skip = ctx->zero_skip; k.vxc_skip_on = ctx->vxc_skip_on; k.vxc_skip_mode = ctx->vxc_skip_mode; k.vxc_scatter = ctx->vxc_scatter;
    k.debug_nodmma = ctx->debug_nodmma; k.wait_ns = ctx->wait_ns; k.dyn_sched = ctx->dyn_sched; k.stagger_min = ctx->stagger_min; k.density_unit = ctx->density_unit; k.vxc_producers = ctx->vxc_producers; k.vxc_prefetch = ctx->vxc_prefetch; k.vxc_rebalance = ctx->vxc_rebalance; k.density_producers = ctx->density_producers; k.density_scatter = ctx->density_scatter; k.density_wide = ctx->density_wide;
    k.dsym = ctx->dsym.ptr; k.coef = ctx->coef.ptr; k.epart = ctx->epart.ptr; k.vpart = ctx->vpart.ptr;
    k.rho = ctx->rho.ptr;
    return k;
}

static void run_plan(CublasHandleWrapper* ctx, const Problem& p, Plan& pl) {
    cudaStream_t st = ctx->stream;
    if (ctx->timing) cudaEventRecord(ctx->ev[0], st);
    symmetrize_pad_tma_kernel<<<dim3((pl.KP + 127) / 128, pl.NP * pl.nsub), 128, 0, st>>>(p.nao, pl.KP, pl.NP, pl.nsub, p.dm,
                                                                                         pl.dsym);
    if (pl.dp.counters) {
        cudaMemsetAsync(pl.dp.counters, 0, COUNTERS_HEAD_BYTES + FSTAT_BYTES, st);
        unsigned char* h_fmap = reinterpret_cast<unsigned char*>(ctx->h_scalar) + HOST_FMAP_OFF;
        if (!ctx->fmap_valid) {   // the initial deal: fragments w and 15 - w (opposite ends of the tile)
            for (int t = 0; t < 16; ++t)
                for (int w = 0; w < 8; ++w) { h_fmap[t * 16 + 2 * w] = (unsigned char)w; h_fmap[t * 16 + 2 * w + 1] = (unsigned char)(15 - w); }
            ctx->fmap_valid = true;
            ctx->fmap_dirty = true;
        }
        if (ctx->fmap_dirty) {    // (pinned source; the blocking entry point synchronises before the host touches it again)
            cudaMemcpyAsync(reinterpret_cast<unsigned char*>(pl.dp.counters) + FMAP_OFF, h_fmap, FMAP_BYTES, cudaMemcpyHostToDevice, st);
            ctx->fmap_dirty = false;
        }
    }
    ctx->stats.v_tiles_m = pl.v_tiles_m;
    void* dargs[1] = {&pl.dp};
    DFT_CUDA_CHECK(ctx, cudaLaunchKernel(pl.dfunc, dim3(pl.dgrid), dim3(NTHREADS), dargs, (size_t)pl.dsmem, st));
    xc_point_kernel<<<pl.pgrid, POINT_THREADS, 0, st>>>(pl.pp);
    if (ctx->timing) cudaEventRecord(ctx->ev[1], st);
    void* vargs[1] = {&pl.vp};
    DFT_CUDA_CHECK(ctx, cudaLaunchKernel(pl.vfunc, pl.vgrid, dim3(pl.vthreads), vargs, (size_t)pl.vsmem, st));
    if (ctx->timing) cudaEventRecord(ctx->ev[2], st);
    const int fin_tiles = (p.nao + FIN_TILE - 1) / FIN_TILE;
    const int raw = (ctx->raw_convention && p.xc_type == 1) ? 1 : 0;
    finalize_tma_kernel<<<fin_tiles * fin_tiles, FIN_THREADS, 0, st>>>(p.nao, pl.ldv, pl.mpv, pl.fin_nt, pl.nsub, pl.nsl, pl.shift1, pl.lda_half, raw,
                                                                      pl.vpart, p.vxc, pl.pgrid, pl.epart, p.d_exc);
    if (ctx->timing) cudaEventRecord(ctx->ev[3], st);
    ctx->stats.launches = 5;
    ctx->stats.path = PATH_TMA;
    ctx->stats.density_units = pl.dp.per_tile ? pl.dp.nblocks * pl.dp.ntiles : pl.dp.nblocks;
    ctx->stats.density_groups = pl.dgroups * pl.dgrid;
    ctx->stats.dyn_units = 0.0;
    DFT_CUDA_CHECK(ctx, cudaGetLastError());
}

}  // namespace tmapath

bool tma_compatible(const Problem& p) {
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (p.nao < 1 || p.ngrid < 1) return false;
    if (!al16(p.ao)) return false;
    if (p.xc_type != 0) {
        if (!(al16(p.gx) && al16(p.gz))) return false;
        // Odd nao with odd ngrid puts the y-gradient plane (grad + ngrid nao) at 8 mod 16.  TMA needs every box to start
        // on a 16-byte boundary, and in such a plane the rows that do are the OPPOSITE parity of the other planes', so no
        // tensor map can line it up with them: run_tma copies that one plane to aligned scratch first (one D2D copy per
        // call, ~6 % of the step at C5 size, against 2.7x for the generic kernels round 1 fell back to).
        if (!al16(p.gy) && (reinterpret_cast<uintptr_t>(p.gy) & 7u)) return false;
    }
    if (p.nao > 128 * 16) return false;  // keep the padded D and the slice partials modest
    return tmapath::encode_fn() != nullptr;
}

void run_tma(CublasHandleWrapper* ctx, const Problem& p_in) {
    using namespace tmapath;
    Problem p = p_in;
    if (p.xc_type != 0 && (reinterpret_cast<uintptr_t>(p.gy) & 15u) != 0) {   // see tma_compatible
        const size_t bytes = sizeof(double) * (size_t)p.ngrid * p.nao;
        double* gy2 = (double*)ctx->repack.ensure(bytes, &ctx->failed);
        if (ctx->failed) return;
        DFT_CUDA_CHECK(ctx, cudaMemcpyAsync(gy2, p.gy, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        p.gy = gy2;
    }
    if (!ctx->tma_plan) ctx->tma_plan = new Plan();
    Plan& pl = *static_cast<Plan*>(ctx->tma_plan);
    PlanKey key = make_key(ctx, p);
    if (!pl.valid || memcmp(&key, &pl.key, sizeof(key)) != 0) {
        pl.valid = false;
        if (ctx->num_sms <= 0) cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, ctx->device);
        if (p.xc_type == 0) build_plan<1>(ctx, p, ctx->num_sms, pl);
        else build_plan<4>(ctx, p, ctx->num_sms, pl);
        if (ctx->failed) return;
        pl.key = make_key(ctx, p);  // (workspaces may have grown while planning)
        pl.valid = true;
        ctx->stats.plans_built++;
    }
    run_plan(ctx, p, pl);
}

// New deal of the V kernel's 8-column M fragments to its warps, per M tile row, from the live (fragment, k-step) counts
// of the build that just finished: the warp that skips least is the CTA's critical path (phase timing at C5: it waits
// 9 % of the time, the lightest warp 36 %), so pair the heaviest fragment with the lightest, and put the heaviest pair
// on the same SM sub-partition (warps w and w + 4 share a tensor pipe) as the lightest pair.  The AO planes do not
// change during an SCF, so the counts of one iteration describe the next.
void tma_rebalance(CublasHandleWrapper* ctx) {
    const int tiles_m = ctx->stats.v_tiles_m;
    if (!ctx->h_scalar || tiles_m < 1 || tiles_m > 16) return;
    const unsigned int* fstat = reinterpret_cast<const unsigned int*>(reinterpret_cast<const unsigned char*>(ctx->h_scalar) + HOST_FSTAT_OFF);
    unsigned char* h_fmap = reinterpret_cast<unsigned char*>(ctx->h_scalar) + HOST_FMAP_OFF;
    for (int t = 0; t < tiles_m; ++t) {
        const unsigned int* c = fstat + t * 16;
        unsigned long long total = 0;
        for (int i = 0; i < 16; ++i) total += c[i];
        if (total == 0) continue;   // (an instance that keeps no counts ran: leave the deal alone)
        int f[16];
        for (int i = 0; i < 16; ++i) f[i] = i;
        std::stable_sort(f, f + 16, [&](int a, int b) { return c[a] > c[b]; });
        int pa[8], pb[8], order[8];
        unsigned long long load[8];
        for (int k = 0; k < 8; ++k) { pa[k] = f[k]; pb[k] = f[15 - k]; load[k] = (unsigned long long)c[pa[k]] + c[pb[k]]; order[k] = k; }
        std::stable_sort(order, order + 8, [&](int a, int b) { return load[a] > load[b]; });
        unsigned char m[16];
        for (int s = 0; s < 4; ++s) {   // sub-partition s: warps s and s + 4
            const int heavy = order[s], light = order[7 - s];
            m[2 * s] = (unsigned char)pa[heavy]; m[2 * s + 1] = (unsigned char)pb[heavy];
            m[2 * (s + 4)] = (unsigned char)pa[light]; m[2 * (s + 4) + 1] = (unsigned char)pb[light];
        }
        if (memcmp(m, h_fmap + t * 16, 16) != 0) { memcpy(h_fmap + t * 16, m, 16); ctx->fmap_dirty = true; }
    }
}

void free_tma_plan(CublasHandleWrapper* ctx) {
    delete static_cast<tmapath::Plan*>(ctx->tma_plan);
    ctx->tma_plan = nullptr;
}

}  // namespace xc
