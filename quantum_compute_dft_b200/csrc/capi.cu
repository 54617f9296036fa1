// capi.cu -- solver classes and the C ABI (drop-in for the reference's weights/dft.so).
//
// Reference counterparts (file:line into /root/reference/src/dft_solver.cu):
//   XCSolver ctor/dtor :536-539, CublasHandleWrapper :530-534
//   LDASolver::compute_xc :559-584, GGASolver :588-621, B3LYPSolver :625-672
//   extern "C" block :675-719
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "../../include/dft_b200_ext.h"
#include "engine.h"

// ------------------------------------------------------------------------- context
void* DeviceBuffer::ensure(size_t bytes, bool* failed) {
    if (bytes <= capacity && ptr) return ptr;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    capacity = 0;
    size_t want = bytes + bytes / 8 + 256;  // a little slack so slowly growing sizes do not thrash
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        fprintf(stderr, "[dft_b200] cudaMalloc(%zu) failed: %s\n", want, cudaGetErrorString(e));
        ptr = nullptr;
        if (failed) *failed = true;
        return nullptr;
    }
    capacity = want;
    return ptr;
}

void DeviceBuffer::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    capacity = 0;
}

CublasHandleWrapper::CublasHandleWrapper() {
    DFT_CUDA_CHECK(this, cudaGetDevice(&device));
    // A *blocking* stream: implicitly ordered with the legacy default stream the reference's
    // driver uses (dft.py:178,201,207), so callers need no extra synchronisation.
    DFT_CUDA_CHECK(this, cudaStreamCreateWithFlags(&stream, cudaStreamDefault));
    for (auto& e : ev) DFT_CUDA_CHECK(this, cudaEventCreate(&e));
    DFT_CUDA_CHECK(this, cudaMallocHost(reinterpret_cast<void**>(&h_scalar), HOST_BLOCK_BYTES));
    if (h_scalar) { h_scalar[8] = 0.0; h_scalar[9] = 1.0; h_scalar[10] = 0.0; h_scalar[11] = 0.0; }  // [8],[9]: constant sources of the "this rank failed" element; [10],[11]: zero-copy {E_xc, seq}
}

CublasHandleWrapper::~CublasHandleWrapper() {
    DeviceGuard guard(device);
    xc::fanout_destroy(this);   // (child engines on the other devices first: they reduce into this one)
    if (stream) cudaStreamSynchronize(stream);
    xc::comm_destroy(this);
    xc::free_tma_plan(this);
    dsym.release(); counters.release(); rho.release(); coef.release(); epart.release(); vpart.release(); result.release(); scratch.release(); repack.release();
    if (h_scalar) cudaFreeHost(h_scalar);
    for (auto& e : ev)
        if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
}

size_t CublasHandleWrapper::workspace_bytes() const {
    return dsym.capacity + rho.capacity + coef.capacity + epart.capacity + vpart.capacity + result.capacity + scratch.capacity + repack.capacity;
}

namespace xc {
void resolve_times(CublasHandleWrapper* c) {
    if (!c || !c->times_pending) return;
    DeviceGuard guard(c->device);
    c->times_pending = false;
    if (cudaEventSynchronize(c->ev[4]) == cudaSuccess) {
        cudaEventElapsedTime(&c->stats.density_ms, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&c->stats.vxc_ms, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&c->stats.reduce_ms, c->ev[2], c->ev[4]);
        cudaEventElapsedTime(&c->stats.total_ms, c->ev[0], c->ev[4]);
    }
}
}  // namespace xc

// ------------------------------------------------------------------------- solver classes
XCSolver::XCSolver() : handle_wrapper(new CublasHandleWrapper()) {}
XCSolver::~XCSolver() = default;

void XCSolver::compute_coulomb(int nao, const double* d_eri, const double* d_dm, double* d_J) {
    DeviceGuard guard(handle_wrapper->device);
    xc::coulomb_gemv(handle_wrapper.get(), nao, d_eri, d_dm, d_J);
}

void XCSolver::safe_cublas_dgemm(bool transA, bool transB, int m, int n, int k, const double* A, int lda,
                                 const double* B, int ldb, double* C, int ldc) {
    DeviceGuard guard(handle_wrapper->device);
    xc::dgemm_colmajor(handle_wrapper.get(), transA, transB, m, n, k, A, lda, B, ldb, C, ldc);
}

namespace {
// result = failed ranks > 0 ? NaN : E  (multi-GPU, asynchronous variant: the host never sees the reduced failure count)
__global__ void publish_exc_kernel(const double* __restrict__ e_and_failed, double* __restrict__ out) {
    out[0] = e_and_failed[1] != 0.0 ? __longlong_as_double(0x7ff8000000000000ll) : e_and_failed[0];
}
}  // namespace

namespace xc {
bool enqueue_counter_readback(CublasHandleWrapper* ctx) {
    if (!(ctx->stats.path == PATH_TMA && ctx->counters.ptr)) return false;
    DFT_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_scalar + 2, ctx->counters.ptr, 5 * sizeof(unsigned long long),
                                        cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->vxc_rebalance)   // the V kernel's live counts per fragment: next call's deal of the fragments to the warps
        DFT_CUDA_CHECK(ctx, cudaMemcpyAsync(reinterpret_cast<unsigned char*>(ctx->h_scalar) + HOST_FSTAT_OFF,
                                            static_cast<unsigned char*>(ctx->counters.ptr) + FSTAT_OFF, FSTAT_BYTES,
                                            cudaMemcpyDeviceToHost, ctx->stream));
    return true;
}

void apply_counters(CublasHandleWrapper* ctx) {
    unsigned long long c[5];
    memcpy(c, ctx->h_scalar + 2, sizeof(c));
    ctx->stats.skip_fraction = c[1] ? 1.0 - (double)c[0] / (double)c[1] : 0.0;
    ctx->stats.vxc_skip_fraction = c[3] ? 1.0 - (double)c[2] / (double)c[3] : 0.0;
    ctx->stats.dyn_units = (double)(c[4] & 0xffffffffull);
    // adaptive: the zero-skipping V instance pays ~6 % on dense operands; use it only where the density
    // kernel just skipped a real share of its k-steps (the decision takes effect with the next call)
    if (ctx->vxc_skip < 0) ctx->vxc_skip_on = ctx->stats.skip_fraction >= 0.10;
    if (ctx->vxc_rebalance) xc::tma_rebalance(ctx);
}

// One XC build on the engine stream.  When `d_exc_out` is null the call blocks and returns E_xc
// (reference semantics); otherwise E_xc is left on the device and the call returns immediately.
//
// Multi-GPU: the packed buffer is [V_xc (nao^2) | E_xc | failed] and EVERY rank that has a usable `nao` reaches the
// collective, whatever happened locally -- a rank with bad arguments or a failed launch contributes zeros and
// failed = 1, and every rank returns NaN when the reduced count is non-zero (a rank that left early would leave
// the others blocked in ncclAllReduce; a rank that contributed an unwritten buffer would corrupt their sums).
double run_build(CublasHandleWrapper* ctx, int xc_type, int ngrid, int nao, const double* d_dm, const double* d_ao,
                 const double* d_ao_grad, const double* d_w, double* d_vxc, double* d_exc_out) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    if (!ctx) return nan;
    if (fanout_wants(ctx, ngrid, nao) && d_dm && d_ao && d_w && d_vxc && (xc_type == 0 || d_ao_grad))
        return run_fanout(ctx, xc_type, ngrid, nao, d_dm, d_ao, d_ao_grad, d_w, d_vxc, d_exc_out);
    DeviceGuard guard(ctx->device);   // (restores the caller's current device on return)
    ctx->failed = false;
    ctx->times_pending = false;       // (the events are about to be recorded again; unread intervals are dropped)
    const bool multi = ctx->nranks > 1 && ctx->nccl_comm;
    bool bad = ngrid < 0 || nao <= 0 || !d_dm || !d_ao || !d_w || !d_vxc;
    if (!bad && xc_type != 0 && !d_ao_grad) {
        fprintf(stderr, "[dft_b200] GGA/B3LYP need d_ao_grad (3,ngrid,nao)\n");
        bad = true;
    }
    if (bad && !(multi && nao > 0 && d_vxc)) return nan;

    const size_t n2 = (size_t)nao * nao;
    double* packed = (double*)ctx->result.ensure(sizeof(double) * (n2 + 2), &ctx->failed);
    if (!packed) return nan;

    xc::Problem p;
    p.xc_type = xc_type;
    p.ngrid = ngrid;
    p.nao = nao;
    p.dm = d_dm;
    p.ao = d_ao;
    const size_t plane = (size_t)(ngrid > 0 ? ngrid : 0) * nao;
    p.gx = xc_type ? d_ao_grad : nullptr;
    p.gy = xc_type ? d_ao_grad + plane : nullptr;
    p.gz = xc_type ? d_ao_grad + 2 * plane : nullptr;
    p.w = d_w;
    p.vxc = multi ? packed : d_vxc;
    p.d_exc = multi ? packed + n2 : (d_exc_out ? d_exc_out : packed + n2);

    if (bad) {
        ctx->failed = true;
    } else if (ngrid == 0) {
        cudaMemsetAsync(p.vxc, 0, sizeof(double) * n2, ctx->stream);
        cudaMemsetAsync(p.d_exc, 0, sizeof(double), ctx->stream);
        { XcStats z; z.plans_built = ctx->stats.plans_built; z.ao_ms = ctx->stats.ao_ms; z.skip_fraction = ctx->stats.skip_fraction; z.vxc_skip_fraction = ctx->stats.vxc_skip_fraction; ctx->stats = z; }
    } else {
        // auto: the single-pass small-basis kernel up to nao 48, the TMA / DMMA kernels where the inputs are
        // TMA-addressable, the generic kernels otherwise; "path" forces one (a forced path that cannot take the
        // inputs falls through to the next)
        const bool use_small = (ctx->path == PATH_AUTO || ctx->path == PATH_SMALL) && xc::small_compatible(p);
        if (use_small && !multi && !d_exc_out && ctx->h_scalar) {   // E_xc comes back through mapped pinned memory
            ctx->zc_seq += 1.0;
            p.host_exc_slot = ctx->h_scalar + 10;
            p.host_exc_seq = ctx->zc_seq;
        }
        const bool use_tma = !use_small && ctx->path != PATH_GENERIC && xc::tma_compatible(p);
        if (use_small) xc::run_small(ctx, p);
        else if (use_tma) xc::run_tma(ctx, p);
        else xc::run_generic(ctx, p);
    }
    if (multi) {
        const bool local_failed = ctx->failed;
        if (local_failed) cudaMemsetAsync(packed, 0, sizeof(double) * (n2 + 1), ctx->stream);
        cudaMemcpyAsync(packed + n2 + 1, ctx->h_scalar + (local_failed ? 9 : 8), sizeof(double), cudaMemcpyHostToDevice,
                        ctx->stream);
        // reduced V_xc lands in the caller's array directly (out of place); [E | failed] in place
        if (xc::allreduce_result(ctx, packed, d_vxc, n2) != 0) ctx->failed = true;
        if (d_exc_out) publish_exc_kernel<<<1, 1, 0, ctx->stream>>>(packed + n2, d_exc_out);
    }
    if (ctx->timing) cudaEventRecord(ctx->ev[4], ctx->stream);
    if (d_exc_out) {
        DFT_CUDA_CHECK(ctx, cudaGetLastError());
        return ctx->failed ? nan : 0.0;
    }
    if (p.host_exc_slot && !ctx->failed) {
        // The finalize kernel's last CTA writes {E_xc, seq} once the whole result is in place (device-wide fence before
        // its count); everything the caller enqueues next is stream-ordered behind that kernel anyway.  Poll the slot;
        // look at the stream now and then so that a faulted launch cannot leave the host spinning.
        volatile double* slot = ctx->h_scalar + 10;
        bool seen = false;
        for (unsigned long long spins = 1; !seen; ++spins) {
            seen = slot[1] == ctx->zc_seq;
            if (!seen && (spins & 0x3fff) == 0) {
                const cudaError_t q = cudaStreamQuery(ctx->stream);
                if (q != cudaErrorNotReady) {       // finished (or failed) without publishing: one last look
                    seen = slot[1] == ctx->zc_seq;
                    if (!seen) { DFT_CUDA_CHECK(ctx, q); ctx->failed = true; }
                    break;
                }
            }
        }
        DFT_CUDA_CHECK(ctx, cudaGetLastError());
        ctx->times_pending = ctx->timing && !ctx->failed;
        return (ctx->failed || !seen) ? nan : slot[0];
    }
    DFT_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_scalar, multi ? packed + n2 : p.d_exc, (multi ? 2 : 1) * sizeof(double),
                                        cudaMemcpyDeviceToHost, ctx->stream));
    const bool have_counters = !bad && ngrid > 0 && enqueue_counter_readback(ctx);
    DFT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (multi && ctx->h_scalar[1] != 0.0) {
        if (!ctx->failed) fprintf(stderr, "[dft_b200] rank %d: %d rank(s) failed in this XC build\n", ctx->rank, (int)ctx->h_scalar[1]);
        ctx->failed = true;
    }
    if (have_counters && !ctx->failed) apply_counters(ctx);
    DFT_CUDA_CHECK(ctx, cudaGetLastError());
    // (the four event intervals are read when DFT_GetStat asks for one: four driver calls per XC build are real money
    // at H2O size, where the whole call is ~20 us)
    ctx->times_pending = ctx->timing && !bad && ngrid > 0 && !ctx->failed;
    return ctx->failed ? nan : ctx->h_scalar[0];
}

}  // namespace xc




LDASolver::LDASolver() : XCSolver() {}
double LDASolver::compute_xc(int ngrid, int nao, const double* d_dm, const double* d_ao, const double* d_ao_grad,
                             const double* d_weights, double* d_vxc) {
    (void)d_ao_grad;  // ignored, as in the reference (dft.py:75 passes 0)
    return xc::run_build(handle_wrapper.get(), 0, ngrid, nao, d_dm, d_ao, nullptr, d_weights, d_vxc, nullptr);
}

GGASolver::GGASolver() : XCSolver() {}
double GGASolver::compute_xc(int ngrid, int nao, const double* d_dm, const double* d_ao, const double* d_ao_grad,
                             const double* d_weights, double* d_vxc) {
    return xc::run_build(handle_wrapper.get(), 1, ngrid, nao, d_dm, d_ao, d_ao_grad, d_weights, d_vxc, nullptr);
}

B3LYPSolver::B3LYPSolver() : XCSolver() {}
double B3LYPSolver::compute_xc(int ngrid, int nao, const double* d_dm, const double* d_ao, const double* d_ao_grad,
                               const double* d_weights, double* d_vxc) {
    return xc::run_build(handle_wrapper.get(), 2, ngrid, nao, d_dm, d_ao, d_ao_grad, d_weights, d_vxc, nullptr);
}

namespace {
int solver_type(XCSolver* s) {
    if (dynamic_cast<LDASolver*>(s)) return 0;
    if (dynamic_cast<GGASolver*>(s)) return 1;
    if (dynamic_cast<B3LYPSolver*>(s)) return 2;
    return -1;
}
}  // namespace

// ------------------------------------------------------------------------- C ABI
extern "C" {

XCSolver* DFT_CreateSolver(int type) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fprintf(stderr, "[dft_b200] no CUDA device: this library has no CPU fallback\n");
        return nullptr;
    }
    XCSolver* s = nullptr;
    if (type == SOLVER_LDA) s = new LDASolver();
    else if (type == SOLVER_GGA) s = new GGASolver();
    else if (type == SOLVER_B3LYP) s = new B3LYPSolver();
    if (!s) return nullptr;
    // DFT_B200_DEVICES=<n>|all: the unmodified driver (dft.py:107-116 knows nothing of options) uses n GPUs of the
    // box from its single process; the same as DFT_SetOption(solver, "devices", n).  DFT_B200_VIRTUAL_DEVICES=<n>
    // (n children dealt over the visible devices: tests on a one-GPU box) and DFT_B200_DEVICES_MIN_WORK=<x> are the
    // environment forms of the options "virtual_devices" and "devices_min_work"; DFT_B200_VERBOSE=1 reports every
    // shard cut on stderr.
    const char* env = getenv("DFT_B200_DEVICES");
    const char* venv = getenv("DFT_B200_VIRTUAL_DEVICES");
    if (env || venv) {
        int want = venv ? atoi(venv) : (!strcmp(env, "all") ? ndev : atoi(env));
        if (!venv && want > ndev) want = ndev;
        if (want > 1) {
            if (xc::fanout_configure(s->context(), want, venv != nullptr) != 0)
                fprintf(stderr, "[dft_b200] %s=%s: multi-GPU fan-out not available, using one device\n",
                        venv ? "DFT_B200_VIRTUAL_DEVICES" : "DFT_B200_DEVICES", venv ? venv : env);
            else if (const char* mw = getenv("DFT_B200_DEVICES_MIN_WORK"))
                xc::fanout_set_option(s->context(), "devices_min_work", atof(mw));
        }
    }
    return s;
}

void DFT_DestroySolver(XCSolver* solver) {
    if (solver) delete solver;   // (~CublasHandleWrapper makes the solver's device current while it frees)
}

double DFT_ComputeXC(XCSolver* solver, int ngrid, int nao, unsigned long long d_dm_ptr,
                     unsigned long long d_ao_ptr, unsigned long long d_ao_grad_ptr,
                     unsigned long long d_weights_ptr, unsigned long long d_vxc_ptr) {
    if (!solver) return 0.0;
    return solver->compute_xc(ngrid, nao, reinterpret_cast<const double*>(d_dm_ptr),
                              reinterpret_cast<const double*>(d_ao_ptr),
                              reinterpret_cast<const double*>(d_ao_grad_ptr),
                              reinterpret_cast<const double*>(d_weights_ptr), reinterpret_cast<double*>(d_vxc_ptr));
}

void DFT_ComputeCoulomb(XCSolver* solver, int nao, unsigned long long d_eri_ptr, unsigned long long d_dm_ptr,
                        unsigned long long d_J_ptr) {
    if (!solver) return;
    solver->compute_coulomb(nao, reinterpret_cast<const double*>(d_eri_ptr),
                            reinterpret_cast<const double*>(d_dm_ptr), reinterpret_cast<double*>(d_J_ptr));
}

int DFT_ComputeCoulombExchange(XCSolver* solver, int nao, unsigned long long d_eri_ptr, unsigned long long d_dm_ptr,
                               unsigned long long d_J_ptr, unsigned long long d_K_ptr) {
    if (!solver || nao <= 0 || !d_eri_ptr || !d_dm_ptr || !d_J_ptr || !d_K_ptr) return 1;
    CublasHandleWrapper* ctx = solver->context();
    DeviceGuard guard(ctx->device);
    ctx->failed = false;
    xc::coulomb_exchange(ctx, nao, reinterpret_cast<const double*>(d_eri_ptr), reinterpret_cast<const double*>(d_dm_ptr),
                         reinterpret_cast<double*>(d_J_ptr), reinterpret_cast<double*>(d_K_ptr));
    return ctx->failed ? 2 : 0;
}

int DFT_BuildFock(XCSolver* solver, int nao, unsigned long long d_hcore_ptr, unsigned long long d_J_ptr,
                  unsigned long long d_vxc_ptr, unsigned long long d_K_ptr, double c_hf, unsigned long long d_F_ptr) {
    if (!solver || nao <= 0 || !d_hcore_ptr || !d_J_ptr || !d_vxc_ptr || !d_F_ptr) return 1;
    CublasHandleWrapper* ctx = solver->context();
    DeviceGuard guard(ctx->device);
    ctx->failed = false;
    xc::build_fock(ctx, nao, reinterpret_cast<const double*>(d_hcore_ptr), reinterpret_cast<const double*>(d_J_ptr),
                   reinterpret_cast<const double*>(d_vxc_ptr), reinterpret_cast<const double*>(d_K_ptr), c_hf,
                   reinterpret_cast<double*>(d_F_ptr));
    return ctx->failed ? 2 : 0;
}

int DFT_SCFEnergies(XCSolver* solver, int nao, unsigned long long d_dm_ptr, unsigned long long d_hcore_ptr,
                    unsigned long long d_J_ptr, unsigned long long d_K_ptr, double c_hf, double* out3) {
    if (!solver || nao <= 0 || !d_dm_ptr || !d_hcore_ptr || !d_J_ptr || !out3) return 1;
    CublasHandleWrapper* ctx = solver->context();
    DeviceGuard guard(ctx->device);
    ctx->failed = false;
    xc::scf_energies(ctx, nao, reinterpret_cast<const double*>(d_dm_ptr), reinterpret_cast<const double*>(d_hcore_ptr),
                     reinterpret_cast<const double*>(d_J_ptr), reinterpret_cast<const double*>(d_K_ptr), c_hf, out3);
    return ctx->failed ? 2 : 0;
}

int DFT_ComputeXCAsync(XCSolver* solver, int ngrid, int nao, unsigned long long d_dm_ptr,
                       unsigned long long d_ao_ptr, unsigned long long d_ao_grad_ptr,
                       unsigned long long d_weights_ptr, unsigned long long d_vxc_ptr,
                       unsigned long long d_exc_ptr) {
    if (!solver || !d_exc_ptr) return 1;
    const int t = solver_type(solver);
    if (t < 0) return 2;
    double r = xc::run_build(solver->context(), t, ngrid, nao, reinterpret_cast<const double*>(d_dm_ptr),
                      reinterpret_cast<const double*>(d_ao_ptr),
                      t ? reinterpret_cast<const double*>(d_ao_grad_ptr) : nullptr,
                      reinterpret_cast<const double*>(d_weights_ptr), reinterpret_cast<double*>(d_vxc_ptr),
                      reinterpret_cast<double*>(d_exc_ptr));
    return std::isnan(r) ? 3 : 0;
}

int DFT_StreamSynchronize(XCSolver* solver) {
    if (!solver) return 1;
    DeviceGuard guard(solver->context()->device);
    return cudaStreamSynchronize(solver->context()->stream) == cudaSuccess ? 0 : 2;
}

unsigned long long DFT_GetStream(XCSolver* solver) {
    if (!solver) return 0ull;
    return reinterpret_cast<unsigned long long>(solver->context()->stream);
}

int DFT_SetOption(XCSolver* solver, const char* key, double value) {
    if (!solver || !key) return 1;
    return xc::set_option(solver->context(), key, value);
}

}  // extern "C"

namespace {
int set_engine_option(CublasHandleWrapper* c, const char* key, double value) {
    if (!strcmp(key, "exact_functionals")) { c->exact_functionals = value != 0.0; return 0; }
    if (!strcmp(key, "path")) { c->path = (int)value; return 0; }
    if (!strcmp(key, "timing")) { c->timing = value != 0.0; return 0; }
    if (!strcmp(key, "l2_prefetch")) { c->l2_prefetch = value != 0.0; return 0; }
    if (!strcmp(key, "small_streaming")) { c->small_streaming = value != 0.0; return 0; }
    if (!strcmp(key, "ao_shape")) { c->ao_shape = (int)value; return 0; }
    if (!strcmp(key, "ao_vec_stores")) { c->ao_vec_stores = value != 0.0; return 0; }
    if (!strcmp(key, "ao_input_order")) { c->ao_input_order = value != 0.0; return 0; }
    if (!strcmp(key, "vxc_skip")) { c->vxc_skip = (int)value; if (c->vxc_skip >= 0) c->vxc_skip_on = c->vxc_skip != 0; return 0; }
    if (!strcmp(key, "density_unit")) { const int v = (int)value; if (v < 0 || v > 2) return 3; c->density_unit = v; return 0; }
    if (!strcmp(key, "stagger_min")) { c->stagger_min = value < 0.0 ? 0 : (int)value; return 0; }
    if (!strcmp(key, "dyn_sched")) { c->dyn_sched = value != 0.0; return 0; }
    if (!strcmp(key, "wait_ns")) { c->wait_ns = value < 0.0 ? 0 : (int)value; return 0; }
#ifdef DFT_DIAGNOSTICS
    if (!strcmp(key, "debug_nodmma")) { c->debug_nodmma = value != 0.0; return 0; }   // results are WRONG: delivery floor
#endif
    if (!strcmp(key, "vxc_scatter")) { c->vxc_scatter = value != 0.0; return 0; }
    if (!strcmp(key, "density_wide")) { c->density_wide = value != 0.0; return 0; }
    if (!strcmp(key, "density_scatter")) { c->density_scatter = value != 0.0; return 0; }
    if (!strcmp(key, "density_producers")) { c->density_producers = value >= 2.0 ? 2 : 1; return 0; }
    if (!strcmp(key, "vxc_prefetch")) { c->vxc_prefetch = value < 0.0 ? 0 : (value > 64.0 ? 64 : (int)value); return 0; }
    if (!strcmp(key, "vxc_rebalance")) { c->vxc_rebalance = value != 0.0; c->fmap_valid = false; return 0; }
    if (!strcmp(key, "vxc_producers")) { c->vxc_producers = value < 1.0 ? 1 : (value > 4.0 ? 4 : (int)value); return 0; }
    if (!strcmp(key, "vxc_skip_mode")) {
        const int v = (int)value;
#ifdef DFT_V_EXPERIMENTS
        if (v < 1 || v > 7) return 3;
#else
        if (v != 1 && v != 2 && v != 4) return 3;   // (3, 5, 6, 7: measured variants, diagnostic builds only)
#endif
        c->vxc_skip_mode = v;
        return 0;
    }
    if (!strcmp(key, "raw_convention")) { c->raw_convention = value != 0.0; return 0; }
    if (!strcmp(key, "zero_skip")) { c->zero_skip = value != 0.0; return 0; }
    if (!strcmp(key, "tma_3d")) { c->tma_3d = value != 0.0; return 0; }
    if (!strcmp(key, "vxc_shape")) { const int v = (int)value; if (v != 0 && v != 64 && v != 128 && v != 160) return 3; c->vxc_shape = v; return 0; }
    if (!strcmp(key, "vxc_vk")) { c->vxc_vk = value == 16.0 ? 16 : (value == 8.0 ? 8 : 0); return 0; }
    if (!strcmp(key, "deterministic")) { return value != 0.0 ? 0 : 3; }  // reductions are always fixed-order
    return 2;
}
}  // namespace

namespace xc {
// Options of the fan-out itself ("devices", ...) are handled by fanout.cu; every engine option that was accepted is
// remembered and forwarded to the child engines, so that all devices of a fan-out run the same instances.
int set_option(CublasHandleWrapper* c, const char* key, double value) {
    if (!c || !key) return 1;
    if (!c->is_fan_child) {
        const int rf = fanout_set_option(c, key, value);
        if (rf != 2) return rf;
    }
    const int rc = set_engine_option(c, key, value);
    if (rc == 0 && !c->is_fan_child) {
        c->option_log.emplace_back(key, value);
        fanout_forward_option(c, key, value);
    }
    return rc;
}
}  // namespace xc

extern "C" {

double DFT_GetStat(XCSolver* solver, const char* key) {
    if (!solver || !key) return -1.0;
    CublasHandleWrapper* c = solver->context();
    {
        bool known = false;
        const double v = xc::fanout_stat(c, key, &known);
        if (known) return v;
    }
    if (strstr(key, "_ms") && strcmp(key, "ao_ms")) xc::resolve_times(c);
    if (!strcmp(key, "density_ms")) return c->stats.density_ms;
    if (!strcmp(key, "vxc_ms")) return c->stats.vxc_ms;
    if (!strcmp(key, "reduce_ms")) return c->stats.reduce_ms;
    if (!strcmp(key, "total_ms")) return c->stats.total_ms;
    if (!strcmp(key, "ao_ms")) return c->stats.ao_ms;
    if (!strcmp(key, "launches")) return c->stats.launches;
    if (!strcmp(key, "plans_built")) return c->stats.plans_built;
    if (!strcmp(key, "dyn_units")) return c->stats.dyn_units;
    if (!strcmp(key, "density_units")) return c->stats.density_units;
    if (!strcmp(key, "density_groups")) return c->stats.density_groups;
    if (!strcmp(key, "skip_fraction")) return c->stats.skip_fraction;
    if (!strcmp(key, "vxc_skip_fraction")) return c->stats.vxc_skip_fraction;
    if (!strcmp(key, "path")) return c->stats.path;
    if (!strcmp(key, "workspace_bytes")) return (double)c->workspace_bytes();
    if (!strcmp(key, "nranks")) return c->nranks;
    return -1.0;
}

#ifdef DFT_DIAGNOSTICS
// Diagnostic builds only (build.py --diag -> weights/dft_diag.so; declared in include/dft_b200_ext.h under the
// same macro): copy an engine workspace to the host.
int DFT_DebugRead(XCSolver* solver, const char* what, void* dst, unsigned long long nbytes) {
    if (!solver || !what || !dst) return 1;
    CublasHandleWrapper* c = solver->context();
    DeviceBuffer* b = !strcmp(what, "coef") ? &c->coef : !strcmp(what, "epart") ? &c->epart
                    : !strcmp(what, "dsym") ? &c->dsym : !strcmp(what, "vpart") ? &c->vpart
                    : !strcmp(what, "scratch") ? &c->scratch : !strcmp(what, "rho") ? &c->rho : nullptr;
    if (!b || !b->ptr) return 2;
    if (nbytes > b->capacity) nbytes = b->capacity;
    DeviceGuard guard(c->device);
    cudaStreamSynchronize(c->stream);
    return cudaMemcpy(dst, b->ptr, nbytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 3;
}
#endif

const char* DFT_B200_Version(void) { return "quantum_compute_dft_b200 0.1 (sm_100a)"; }

}  // extern "C"
