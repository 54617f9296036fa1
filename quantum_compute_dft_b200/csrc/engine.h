// engine.h -- engine context behind the XCSolver pimpl (`CublasHandleWrapper`).
//
// The reference defines `struct CublasHandleWrapper { cublasHandle_t handle; }`
// (dft_solver.cu:530-534) and allocates/frees 6-7 temporaries with cudaMalloc/cudaFree on
// every compute_xc call (:561-582, :590-619, :627-670).  Here the same opaque slot owns one
// stream, grow-only workspaces that live as long as the solver, the TMA descriptors, the
// run-time options and (optionally) an NCCL communicator.  No cuBLAS.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#define DFT_CUDA_CHECK(ctx, call)                                                              \
    do {                                                                                       \
        cudaError_t err__ = (call);                                                            \
        if (err__ != cudaSuccess) {                                                            \
            fprintf(stderr, "[dft_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(err__), \
                    __FILE__, __LINE__, cudaGetErrorString(err__));                            \
            if (ctx) (ctx)->failed = true;                                                     \
        }                                                                                      \
    } while (0)

// Makes `device` current for the scope and restores the caller's device afterwards: every extern "C" entry
// point and the context destructor run under one, so a process that drives several GPUs never launches on
// the engine stream, allocates or frees while another device is current.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

struct DeviceBuffer {
    void* ptr = nullptr;
    size_t capacity = 0;
    // grow-only; contents are NOT preserved across growth
    void* ensure(size_t bytes, bool* failed);
    void release();
};

enum XcPath { PATH_AUTO = 0, PATH_GENERIC = 1, PATH_TMA = 2, PATH_SMALL = 3 };

// `counters` workspace: [0,1] density k-steps executed / total, [2,3] V units executed / total, [4] dynamic work counter,
// [5] small path's finished-CTA counter (6 x 8 bytes, zeroed before every TMA build) | fstat[256] live k-step units per
// (M tile row, 8-column fragment) of the V kernel (u32, zeroed with the counters) | fmap[256] bytes: which fragments
// each warp of the V kernel owns (persists across calls)
constexpr size_t COUNTERS_HEAD_BYTES = 6 * sizeof(unsigned long long);
constexpr size_t FSTAT_OFF = COUNTERS_HEAD_BYTES, FSTAT_BYTES = 256 * sizeof(unsigned int);
constexpr size_t FMAP_OFF = FSTAT_OFF + FSTAT_BYTES, FMAP_BYTES = 256;
constexpr size_t COUNTERS_BYTES = FMAP_OFF + FMAP_BYTES;
// pinned host block: doubles [0..15] scalars (see capi.cu) | fstat copy (1 KB) | fmap staging (256 B)
constexpr size_t HOST_FSTAT_OFF = 128, HOST_FMAP_OFF = HOST_FSTAT_OFF + FSTAT_BYTES, HOST_BLOCK_BYTES = HOST_FMAP_OFF + FMAP_BYTES;

struct XcStats {
    float density_ms = 0.f, vxc_ms = 0.f, reduce_ms = 0.f, total_ms = 0.f;
    float ao_ms = 0.f;    // last DFT_EvalAO kernel
    int launches = 0;
    int path = 0;
    double skip_fraction = 0.0;  // TMA density kernel: fraction of k-steps skipped as exact zeros in the last call
    double vxc_skip_fraction = 0.0;  // TMA V kernel, box-bit instances: fraction of (box, k-step) units skipped
    int plans_built = 0;  // TMA path: launch plans (tensor maps, geometry) encoded so far; a steady SCF loop builds one
    int density_units = 0, density_groups = 0;  // TMA density kernel: units of work and consumer groups (2 per CTA) of the last launch
    int v_tiles_m = 0;    // TMA V kernel: M tile rows of the last launch (rows of the fragment map)
    double dyn_units = 0.0;  // TMA density kernel: draws from the dynamic work counter in the last call (units + consumer groups; 0 = static deal)
};

struct CublasHandleWrapper {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool failed = false;
    bool times_pending = false;   // the last XC build's event intervals have not been read into `stats` yet

    // options (DFT_SetOption)
    bool exact_functionals = false;
    int path = PATH_AUTO;
    bool timing = false;           // record CUDA events around the kernels of every call (DFT_GetStat "*_ms"); off by default: five event records are a tenth of a call at H2O size
    bool l2_prefetch = false;      // TMA density kernel: short-range L2 prefetch of A tiles and epilogue pieces (measured: no gain)
    int vxc_skip = -1;             // V kernel zero-skipping instance: -1 adaptive (default), 0 never, 1 always
    bool vxc_skip_on = true;       // (adaptive) the zero-skipping V instance is used while the density kernel finds zeros
    int vxc_skip_mode = 2;         // zero-skipping V instance (128 x 128 tile): per-warp M-side votes, 2 = built and voted on for a whole ring stage at once (default, 11.2 ms at C5), 1 = k-step by k-step (round 1, 11.8 ms), 3 = 2 + early Phi fragments, 5 | 6 = 4 x 2 warps with interleaved fragments (12.2-12.4 ms), 4 = staged B with uniform fragment skipping (12.5 ms)
    int vxc_scatter = 1;           // zero-skipping V instances: scatter consecutive ring stages over the grid (golden-ratio stride)
    int density_unit = 0;          // TMA density kernel, unit of work: 0 | 2 = one column tile of a 64-point block (default), 1 = a whole block
    int stagger_min = 8;           // TMA density kernel: consumer group 1 starts half a tile period late when a CTA has more blocks than this
    int dyn_sched = 1;             // TMA density kernel: hand the 64-point blocks out dynamically (one global counter)
    int wait_ns = 0;               // TMA kernels: producer / scanner threads sleep this long between barrier polls
    int debug_nodmma = 0;          // -DDFT_DIAGNOSTICS builds only: TMA kernels skip every DMMA (measures the operand-delivery floor)
    bool vxc_rebalance = true;     // TMA V kernel, per-warp-vote instances: re-deal the 8-column M fragments to the warps from the live counts of the previous call (heaviest with lightest, heavy pairs share an SM sub-partition with light ones)
    bool fmap_valid = false, fmap_dirty = false;   // (state of the fragment map: initialised / host copy newer than the device copy)
    int density_wide = 0;          // TMA density kernel: one consumer group of 8 warps on 128-point blocks with one ring (1) instead of two ping-pong groups of 4 on 64-point blocks (0)
    int density_scatter = 0;       // TMA density kernel: visit the 64-point blocks in a scattered order (golden-ratio stride) instead of grid order (measured: no gain)
    int density_producers = 1;     // TMA density kernel: TMA-issuing threads per consumer group (1 | 2; 2 measured no faster: 8.83 against 8.82 ms at C5)
    int vxc_prefetch = 0;          // TMA V kernel: L2 prefetch distance of the producer in ring stages (0: none)
    int vxc_producers = 2;         // TMA V kernel: TMA-issuing threads per CTA (1..4)
    bool raw_convention = false;   // GGA only: leave the reference's raw unsymmetrised B^T Phi in d_vxc (dft_solver.cu:616) instead of the symmetric matrix
    bool zero_skip = true;         // TMA kernels: skip k-steps whose operand fragment is all zero (exact: adds nothing)
    bool tma_3d = true;            // TMA V kernel: one 3-D TMA load per plane and stage instead of one per 16-column block
    bool small_streaming = false;  // small-basis kernel: always stream tile by tile (never keep whole super-blocks resident)
    int ao_shape = 0;              // DFT_EvalAO block shape: 0 = auto, 16 (points, 8 warps) | 32 (points, 16 warps)
    bool ao_vec_stores = false;    // DFT_EvalAO: 16-byte stores in phase 2 (measured slower than 8-byte ones: not the default)
    bool ao_input_order = false;   // DFT_EvalAO: keep the exponent-sharing groups in shell input order (round 1) instead of sorting them by reach and position
    int vxc_shape = 0;             // TMA V kernel output tile: 0 = auto, 64 | 128 | 160 (= 160 x 80)
    int vxc_vk = 0;                // TMA V kernel, 128 x 128 tile: grid rows per ring stage (8: 5 stages, 16: 2 stages, 0: auto)

    // workspaces
    DeviceBuffer dsym;     // symmetrised, zero-padded density matrix
    DeviceBuffer counters; // AO-screening statistics of the density kernel
    DeviceBuffer rho;      // per-point partial (rho, grad rho / 2) row sums of the density kernel's two warp columns
    DeviceBuffer coef;     // per-point (a,bx,by,bz)
    DeviceBuffer epart;    // per-CTA partial E_xc
    DeviceBuffer vpart;    // split-K partial V tiles
    DeviceBuffer result;   // packed [V_xc (nao*nao) | E_xc] for the all-reduce / async E
    DeviceBuffer scratch;  // DFT_EvalAO shell tables; phase-timing records
    DeviceBuffer repack;   // aligned copy of a y-gradient plane that sits at 8 mod 16 (odd nao x odd ngrid), refreshed every call
    double* h_scalar = nullptr;  // pinned (and, with unified addressing, device-accessible through the same pointer)
    double zc_seq = 0.0;         // sequence number of the last zero-copy E_xc return (h_scalar[10], [11])
    void* tma_plan = nullptr;    // cached launch plan of the TMA path (xc_tma.cu)
    int num_sms = 0;

    // multi-GPU (NCCL loaded lazily; see comm.cu)
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;

    // single-process multi-GPU behind the unmodified ABI (fanout.cu): per-device child engines, their cached AO shards
    void* fan = nullptr;
    bool is_fan_child = false;
    std::vector<std::pair<std::string, double>> option_log;   // every option set so far, replayed onto new child engines

    XcStats stats;

    CublasHandleWrapper();
    ~CublasHandleWrapper();
    size_t workspace_bytes() const;
};

// ---- kernels / launchers implemented across the .cu files --------------------------------
namespace xc {

struct Problem {
    int xc_type;  // 0 LDA, 1 GGA(PBE), 2 B3LYP
    int ngrid, nao;
    const double* dm;
    const double* ao;
    const double* gx;  // nullptr for LDA
    const double* gy;
    const double* gz;
    const double* w;
    double* vxc;     // (nao,nao) output
    double* d_exc;   // device scalar output
    volatile double* host_exc_slot = nullptr;   // small path, blocking calls: mapped pinned {E_xc, sequence number} (see xc_small_finalize)
    double host_exc_seq = 0.0;
};

// generic path (any alignment, any nao): xc_generic.cu
void run_generic(CublasHandleWrapper* ctx, const Problem& p);
// TMA-fed path: xc_tma.cu.  Returns false when the inputs are not TMA-compatible.
bool tma_compatible(const Problem& p);
void run_tma(CublasHandleWrapper* ctx, const Problem& p);
void free_tma_plan(CublasHandleWrapper* ctx);
// after a blocking TMA build: new fragment map for the V kernel from the live counts the build left in the pinned block
void tma_rebalance(CublasHandleWrapper* ctx);
// small-basis single-pass path (nao <= 48): xc_small.cu
bool small_compatible(const Problem& p);
void run_small(CublasHandleWrapper* ctx, const Problem& p);

// One XC build on ctx's own device and stream (capi.cu).  d_exc_out == nullptr: blocks and returns E_xc; otherwise E_xc
// stays on the device and the call returns 0 (NaN on failure) at once.  `defer_readback`: the caller reads the TMA
// path's counters back itself (enqueue_counter_readback / apply_counters) after its own synchronisation.
double run_build(CublasHandleWrapper* ctx, int xc_type, int ngrid, int nao, const double* d_dm, const double* d_ao,
                 const double* d_ao_grad, const double* d_w, double* d_vxc, double* d_exc_out);
bool enqueue_counter_readback(CublasHandleWrapper* ctx);   // async D2H of the TMA path's counters on ctx->stream
void apply_counters(CublasHandleWrapper* ctx);             // after the stream has drained: statistics, adaptive V instance, re-deal
int set_option(CublasHandleWrapper* ctx, const char* key, double value);

// single process, several GPUs (fanout.cu): the primary engine deals the caller's grid to one child engine per device
int fanout_configure(CublasHandleWrapper* ctx, int ndev, bool allow_virtual);   // ndev <= 1 tears the fan-out down
bool fanout_wants(CublasHandleWrapper* ctx, int ngrid, int nao);
double run_fanout(CublasHandleWrapper* ctx, int xc_type, int ngrid, int nao, const double* d_dm, const double* d_ao,
                  const double* d_ao_grad, const double* d_w, double* d_vxc, double* d_exc_out);
void fanout_destroy(CublasHandleWrapper* ctx);
int fanout_set_option(CublasHandleWrapper* ctx, const char* key, double value);   // 2 = not a fan-out key
double fanout_stat(CublasHandleWrapper* ctx, const char* key, bool* known);
void fanout_forward_option(CublasHandleWrapper* ctx, const char* key, double value);
void fanout_invalidate(CublasHandleWrapper* ctx);

// all-reduce of [V | E | failed ranks] over the communicator (comm.cu); no-op when nranks == 1
int allreduce_result(CublasHandleWrapper* ctx, double* d_packed, double* d_vxc, size_t n2);
void comm_destroy(CublasHandleWrapper* ctx);
// reads the last XC build's CUDA-event intervals into ctx->stats if that has not happened yet (capi.cu)
void resolve_times(CublasHandleWrapper* ctx);

void coulomb_gemv(CublasHandleWrapper* ctx, int nao, const double* eri, const double* dm, double* J);
// J and the exact-exchange matrix K[i,k] = sum_jl (ij|kl) D[j,l] in ONE pass over the ERI
void coulomb_exchange(CublasHandleWrapper* ctx, int nao, const double* eri, const double* dm, double* J, double* K);

// device-resident Fock assembly and SCF energy terms (no J / V_xc / K downloads per iteration)
void build_fock(CublasHandleWrapper* ctx, int nao, const double* hcore, const double* J, const double* vxc,
                const double* K, double c_hf, double* F);
void scf_energies(CublasHandleWrapper* ctx, int nao, const double* dm, const double* hcore, const double* J,
                  const double* K, double c_hf, double* out3_host);

void dgemm_colmajor(CublasHandleWrapper* ctx, bool transA, bool transB, int m, int n, int k,
                    const double* A, int lda, const double* B, int ldb, double* C, int ldc);

}  // namespace xc
