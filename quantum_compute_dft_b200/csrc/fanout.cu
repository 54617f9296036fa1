// fanout.cu -- single process, several GPUs, behind the UNMODIFIED ABI (SURVEY.md 8e "process model", 8b "additive
// exports (ii)").
//
// The reference's driver keeps every array on one device (dft.py:155-176) and calls DFT_ComputeXC from one
// process (dft.py:205-208), so the one-process-per-GPU path of comm.cu cannot serve it.  Here the solver that
// DFT_CreateSolver returned (the "primary" engine, on the caller's device) owns one child engine per GPU:
//
//   * the grid is dealt to the children in interleaved blocks of 1024 points (the deal of solver.shard_indices: with
//     AO screening a contiguous range -- one end of the molecule -- does not cost what another costs);
//   * every child keeps ITS shard of (Phi, grad Phi, w) resident on its own device.  The shards are pulled out of the
//     caller's arrays by the copy engines over NVLink -- one strided (2-D / 3-D peer) copy per plane and child, since
//     the blocks of a child are equidistant in the caller's array -- ONCE, and are reused by every later call as long
//     as the caller passes the same arrays (same pointers, sizes and a fingerprint of their contents: all weights and
//     2^17 samples of every plane).  That is the SCF pattern: dft.py uploads the AO arrays once (:155,:172) and only the
//     density matrix changes per iteration (:200).  C33H56N7O17P3S: 7/8 of 17.3 GB leaves device 0 once (~20 ms at
//     NVLink rates, the cost of ONE single-GPU build), then every iteration costs 1.14 MB per device;
//   * per call: D is copied to every child (nao^2 doubles), each child runs the ordinary single-GPU build on its own
//     stream (xc::run_build: the same kernels, the same plans and per-device adaptive state), and ONE kernel on the
//     primary device sums the children's [V_xc | E_xc] in a fixed order through peer loads over NVLink (staged peer
//     copies where the devices cannot address each other) straight into the caller's d_vxc.  No NCCL: all devices
//     belong to this process and the exchange is (nao^2 + 1) doubles per device;
//   * a child's share of a call is ~10 driver calls (40-50 us of host time), so children 1 .. n-1 each have a parked
//     worker thread that enqueues its device's work while the caller's thread does child 0's (C5 on 8 devices: 2.90 ->
//     2.76 ms per call).  The workers only ever touch their own child's state; hand-over is a sequence number
//     (release / acquire) with a short spin before they sleep on a condition variable.
//
// Selected by DFT_SetOption(solver, "devices", n) or, for a driver that knows nothing of options, by the environment
// variable DFT_B200_DEVICES=n|all read in DFT_CreateSolver.  Builds too small to pay for the hand-off
// (ngrid nao^2 < "devices_min_work", default 2e9: H2O, benzene) stay on the primary device.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <limits>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/dft_b200_ext.h"
#include "engine.h"

namespace {

constexpr int FAN_BLOCK = 1024;      // points per dealt block (even: shards of odd-nao planes stay 16-byte aligned)
constexpr int FAN_MAX_DEVICES = 16;
constexpr int FP_SAMPLES = 1 << 17;  // fingerprint samples per AO plane

struct FanChild {
    CublasHandleWrapper* eng = nullptr;
    int dev = 0;
    DeviceBuffer ao, grad, w, dm, out;   // this child's shard of the inputs, its copy of D, its [V_xc | E_xc]
    cudaEvent_t done = nullptr;
    int nreal = 0, nshard = 0;           // points dealt to the child; the same padded to an even count (zero-weight twin)
    bool have_counters = false;
    bool job_failed = false;             // outcome of this child's share of the current call
    std::thread worker;                  // children 1 .. n-1: enqueue their device's work in parallel with the caller's thread
    std::atomic<unsigned long long> done_seq{0};
};

// one DFT_ComputeXC call as the workers see it
struct FanJob {
    int xc_type = 0, ngrid = 0, nao = 0;
    const double *dm = nullptr, *ao = nullptr, *grad = nullptr, *w = nullptr;
    bool stale = false;
    int primary_device = 0;
};

struct FanOut {
    std::vector<FanChild> kids;
    bool peer_loads = true;      // the primary device can address every child's result buffer
    bool cache = true;           // option "ao_cache"
    double min_work = 2e9;       // option "devices_min_work": ngrid * nao^2 below which a build stays on the primary device
    // identity of the arrays the resident shards were cut from
    bool valid = false;
    const double *ao = nullptr, *grad = nullptr, *w = nullptr;
    int ngrid = 0, nao = 0, xc_type = -1;
    unsigned long long fingerprint = 0;
    DeviceBuffer stage;          // primary device: children's results when peer loads are not possible
    DeviceBuffer fp;             // primary device: fingerprint accumulator
    int scatters = 0;            // how many times the shards were (re)built
    bool last_call_fanned = false;
    // worker threads: a child's share of a call is ~10 driver calls (copy of D, five launches, counter read-back, event),
    // 40-50 us of host time; issued by one thread the eighth device would start 0.3 ms after the first
    bool threads = true;         // option "fan_threads"
    FanJob job;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<unsigned long long> seq{0};
    std::atomic<bool> quit{false};
};

struct FanSources {
    const double* p[FAN_MAX_DEVICES];
    int n;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// Order-independent fingerprint (a wrapping sum of mixed (index, bits) pairs) of all weights and FP_SAMPLES evenly
// spaced elements of each of the `nplanes` AO planes.  Cheap by construction: ~11 MB of weights + 0.5 M sectors.
__global__ void fingerprint_kernel(const double* __restrict__ ao, const double* __restrict__ grad,
                                   const double* __restrict__ w, size_t plane, int ngrid, int nplanes,
                                   unsigned long long* __restrict__ out) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (size_t i = tid; i < (size_t)ngrid; i += nthr)
        acc += mix64((unsigned long long)__double_as_longlong(w[i]) ^ (i * 0x9e3779b97f4a7c15ull));
    const size_t ns = plane < (size_t)FP_SAMPLES ? plane : (size_t)FP_SAMPLES;
    const size_t step = ns ? plane / ns : 0;
    for (int pl = 0; pl < nplanes; ++pl) {
        const double* src = pl == 0 ? ao : grad + (size_t)(pl - 1) * plane;
        for (size_t i = tid; i < ns; i += nthr) {
            const size_t idx = i * step;
            acc += mix64((unsigned long long)__double_as_longlong(src[idx]) ^ ((idx + plane * (pl + 1)) * 0xd6e8feb86659fd93ull));
        }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// V_xc = sum over the children of their partial matrices, E_xc likewise, in child order (bit-reproducible).  The
// sources live on the children's devices: peer loads over NVLink (or staged copies on this device).
__global__ void fan_reduce_kernel(FanSources s, double* __restrict__ vxc, double* __restrict__ e_out, size_t n2) {
    const size_t nthr = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n2; i += nthr) {
        double acc = 0.0;
        for (int c = 0; c < s.n; ++c) acc += s.p[c][i];
        if (i < n2) vxc[i] = acc;
        else e_out[0] = acc;
    }
}

// points of child r in a grid of `ngrid` points dealt to n children in blocks of FAN_BLOCK
void deal(int ngrid, int n, int r, int* nfull, int* tail) {
    const int nblk = (ngrid + FAN_BLOCK - 1) / FAN_BLOCK;
    const int rem = ngrid % FAN_BLOCK;            // points of a partial last block (0: the last block is full)
    int mine = r < nblk ? (nblk - r + n - 1) / n : 0;
    const bool owns_partial = rem != 0 && nblk > 0 && (nblk - 1) % n == r;
    *tail = owns_partial ? rem : 0;
    *nfull = owns_partial ? mine - 1 : mine;
}

// dst (contiguous) <- rows [block b of child r, b = 0 .. nfull) of a row-major (ngrid, width) array + the tail rows
// of the partial last block.  The child's blocks are equidistant in the source: ONE strided copy on the copy engines.
bool scatter_plane(CublasHandleWrapper* eng, int dst_dev, int src_dev, double* dst, const double* src, int width, int n,
                   int r, int nfull, int tail, int ngrid, cudaStream_t st) {
    const size_t chunk = (size_t)FAN_BLOCK * width * sizeof(double);
    const double* first = src + (size_t)r * FAN_BLOCK * width;
    cudaError_t e = cudaSuccess;
    if (nfull > 0) {
        const size_t spitch = chunk * n;
        int max_pitch = 0;
        cudaDeviceGetAttribute(&max_pitch, cudaDevAttrMaxPitch, dst_dev);
        if (nfull == 1 || n == 1) {
            e = cudaMemcpyAsync(dst, first, chunk * nfull, cudaMemcpyDefault, st);
        } else if (spitch <= (size_t)max_pitch && dst_dev == src_dev) {
            e = cudaMemcpy2DAsync(dst, chunk, first, spitch, chunk, nfull, cudaMemcpyDeviceToDevice, st);
        } else if (spitch <= (size_t)max_pitch) {
            cudaMemcpy3DPeerParms p;
            memset(&p, 0, sizeof(p));
            p.srcPtr = make_cudaPitchedPtr(const_cast<double*>(first), spitch, chunk, nfull);
            p.dstPtr = make_cudaPitchedPtr(dst, chunk, chunk, nfull);
            p.srcDevice = src_dev;
            p.dstDevice = dst_dev;
            p.extent = make_cudaExtent(chunk, nfull, 1);
            e = cudaMemcpy3DPeerAsync(&p, st);
        } else {
            for (int b = 0; b < nfull && e == cudaSuccess; ++b)
                e = cudaMemcpyAsync(reinterpret_cast<char*>(dst) + chunk * b, reinterpret_cast<const char*>(first) + spitch * b,
                                    chunk, cudaMemcpyDefault, st);
        }
    }
    if (e == cudaSuccess && tail > 0) {
        const int nblk = (ngrid + FAN_BLOCK - 1) / FAN_BLOCK;
        e = cudaMemcpyAsync(dst + (size_t)nfull * FAN_BLOCK * width, src + (size_t)(nblk - 1) * FAN_BLOCK * width,
                            (size_t)tail * width * sizeof(double), cudaMemcpyDefault, st);
    }
    DFT_CUDA_CHECK(eng, e);
    return e == cudaSuccess;
}

// One child's share of a call, on the child's device and stream: (shards,) D, the ordinary single-GPU build, the
// read-back of its counters, the event the primary stream waits for.  Touches only this child's state.
void enqueue_child(FanOut* f, int c) {
    const FanJob& j = f->job;
    FanChild& k = f->kids[c];
    const int n = (int)f->kids.size();
    const size_t n2 = (size_t)j.nao * j.nao, plane = (size_t)j.ngrid * j.nao;
    DeviceGuard child(k.dev);
    CublasHandleWrapper* e = k.eng;
    e->failed = false;
    k.job_failed = true;
    k.have_counters = false;
    cudaStream_t st = e->stream;
    if (j.stale) {
        int nfull = 0, tail = 0;
        deal(j.ngrid, n, c, &nfull, &tail);
        k.nreal = nfull * FAN_BLOCK + tail;
        k.nshard = k.nreal + (k.nreal & 1);
        const size_t rows = (size_t)(k.nshard > 0 ? k.nshard : 1);
        double* ao = (double*)k.ao.ensure(rows * j.nao * sizeof(double), &e->failed);
        double* gr = j.xc_type ? (double*)k.grad.ensure(3 * rows * j.nao * sizeof(double), &e->failed) : nullptr;
        double* w = (double*)k.w.ensure(rows * sizeof(double), &e->failed);
        if (e->failed) return;
        bool ok = scatter_plane(e, k.dev, j.primary_device, ao, j.ao, j.nao, n, c, nfull, tail, j.ngrid, st);
        for (int pl = 0; pl < 3 && j.xc_type && ok; ++pl)
            ok = scatter_plane(e, k.dev, j.primary_device, gr + (size_t)pl * k.nshard * j.nao, j.grad + (size_t)pl * plane,
                               j.nao, n, c, nfull, tail, j.ngrid, st);
        ok = ok && scatter_plane(e, k.dev, j.primary_device, w, j.w, 1, n, c, nfull, tail, j.ngrid, st);
        if (ok && k.nshard > k.nreal) {   // zero-weight twin of the last point: an even number of rows per plane
            const size_t row = (size_t)j.nao * sizeof(double);
            cudaMemcpyAsync(ao + (size_t)k.nreal * j.nao, ao + (size_t)(k.nreal - 1) * j.nao, row, cudaMemcpyDeviceToDevice, st);
            for (int pl = 0; pl < 3 && j.xc_type; ++pl) {
                double* g = gr + (size_t)pl * k.nshard * j.nao;
                cudaMemcpyAsync(g + (size_t)k.nreal * j.nao, g + (size_t)(k.nreal - 1) * j.nao, row, cudaMemcpyDeviceToDevice, st);
            }
            cudaMemsetAsync(w + k.nreal, 0, sizeof(double), st);
        }
        if (!ok) return;
    }
    double* dm = (double*)k.dm.ensure(n2 * sizeof(double), &e->failed);
    double* out = (double*)k.out.ensure((n2 + 1) * sizeof(double), &e->failed);
    if (!dm || !out) return;
    DFT_CUDA_CHECK(e, cudaMemcpyAsync(dm, j.dm, n2 * sizeof(double), cudaMemcpyDefault, st));
    const double r = xc::run_build(e, j.xc_type, k.nshard, j.nao, dm, (const double*)k.ao.ptr, (const double*)k.grad.ptr,
                                   (const double*)k.w.ptr, out, out + n2);
    k.have_counters = k.nshard > 0 && !std::isnan(r) && xc::enqueue_counter_readback(e);
    DFT_CUDA_CHECK(e, cudaEventRecord(k.done, st));
    k.job_failed = std::isnan(r) || e->failed;
}

// Worker of child c: spins briefly for the next call (an SCF loop calls every few ms), then sleeps on the condition
// variable.  `seq` is published with release semantics after the job is written; `done_seq` likewise after the
// child's state is.
void worker_main(FanOut* f, int c) {
    cudaSetDevice(f->kids[c].dev);
    unsigned long long seen = 0;
    for (;;) {
        unsigned long long s = f->seq.load(std::memory_order_acquire);
        if (s == seen && !f->quit.load(std::memory_order_acquire)) {
            const auto t0 = std::chrono::steady_clock::now();
            while ((s = f->seq.load(std::memory_order_acquire)) == seen && !f->quit.load(std::memory_order_acquire)) {
                if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) {
                    std::unique_lock<std::mutex> lk(f->mu);
                    f->cv.wait(lk, [&] { return f->seq.load(std::memory_order_acquire) != seen || f->quit.load(std::memory_order_acquire); });
                }
            }
        }
        if (f->quit.load(std::memory_order_acquire)) return;
        seen = s;
        enqueue_child(f, c);
        f->kids[c].done_seq.store(s, std::memory_order_release);
    }
}

void stop_workers(FanOut* f) {
    {
        std::lock_guard<std::mutex> lk(f->mu);
        f->quit.store(true, std::memory_order_release);
    }
    f->cv.notify_all();
    for (auto& k : f->kids)
        if (k.worker.joinable()) k.worker.join();
}

void release_child(FanChild& k) {
    DeviceGuard guard(k.dev);
    if (k.eng && k.eng->stream) cudaStreamSynchronize(k.eng->stream);
    k.ao.release(); k.grad.release(); k.w.release(); k.dm.release(); k.out.release();
    if (k.done) cudaEventDestroy(k.done);
    k.done = nullptr;
    delete k.eng;
    k.eng = nullptr;
}

}  // namespace

namespace xc {

void fanout_destroy(CublasHandleWrapper* ctx) {
    if (!ctx || !ctx->fan) return;
    FanOut* f = static_cast<FanOut*>(ctx->fan);
    stop_workers(f);
    for (auto& k : f->kids) release_child(k);
    {
        DeviceGuard guard(ctx->device);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        f->stage.release();
        f->fp.release();
    }
    delete f;
    ctx->fan = nullptr;
}

// n children on n devices: the primary's own device first, then the others in index order.  `allow_virtual`: more
// children than devices, dealt round-robin (several children per GPU: the whole mechanism on a one-GPU box -- tests).
int fanout_configure(CublasHandleWrapper* ctx, int ndev, bool allow_virtual) {
    if (!ctx || ctx->is_fan_child) return 1;
    if (ctx->nranks > 1) {
        fprintf(stderr, "[dft_b200] \"devices\": this solver already belongs to a multi-process communicator\n");
        return 3;
    }
    double min_work = 2e9;
    bool cache = true;
    if (ctx->fan) {   // keep the fan-out's own options across a reconfiguration
        min_work = static_cast<FanOut*>(ctx->fan)->min_work;
        cache = static_cast<FanOut*>(ctx->fan)->cache;
    }
    int have = 0;
    if (ndev > 1) {   // (validate before the existing fan-out is touched)
        if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) return 3;
        if (ndev > FAN_MAX_DEVICES) return 3;
        if (ndev > have && !allow_virtual) {
            fprintf(stderr, "[dft_b200] \"devices\" %d: only %d CUDA device(s) visible\n", ndev, have);
            return 3;
        }
    }
    fanout_destroy(ctx);
    if (ndev <= 1) return 0;
    FanOut* f = new FanOut();
    f->min_work = min_work;
    f->cache = cache;
    f->kids = std::vector<FanChild>(ndev);
    bool ok = true;
    for (int c = 0; c < ndev && ok; ++c) {
        FanChild& k = f->kids[c];
        k.dev = (ctx->device + c) % have;
        DeviceGuard guard(k.dev);
        k.eng = new CublasHandleWrapper();   // (takes the current device: stream, events, pinned block)
        k.eng->is_fan_child = true;
        ok = !k.eng->failed && cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming) == cudaSuccess;
        for (const auto& kv : ctx->option_log) set_option(k.eng, kv.first.c_str(), kv.second);
        if (ok && k.dev != ctx->device) {
            // both directions: the child's copy engines pull from the caller's arrays, the primary's reduction
            // kernel loads from the child's result
            int can_pc = 0, can_cp = 0;
            cudaDeviceCanAccessPeer(&can_cp, k.dev, ctx->device);
            cudaDeviceCanAccessPeer(&can_pc, ctx->device, k.dev);
            if (can_cp) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(ctx->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can_cp = 0;
                cudaGetLastError();
            }
            if (can_pc) {
                DeviceGuard primary(ctx->device);
                const cudaError_t e = cudaDeviceEnablePeerAccess(k.dev, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can_pc = 0;
                cudaGetLastError();
            }
            if (!can_pc) f->peer_loads = false;   // (copies still work without peer access: staged by the driver)
        }
    }
    ctx->fan = f;
    if (ok && f->threads)
        for (int c = 1; c < ndev; ++c) f->kids[c].worker = std::thread(worker_main, f, c);
    if (!ok) {
        fprintf(stderr, "[dft_b200] \"devices\" %d: could not create the per-device engines\n", ndev);
        fanout_destroy(ctx);
        return 4;
    }
    return 0;
}

bool fanout_wants(CublasHandleWrapper* ctx, int ngrid, int nao) {
    if (!ctx || !ctx->fan || ctx->is_fan_child || ctx->nranks > 1) return false;
    FanOut* f = static_cast<FanOut*>(ctx->fan);
    const int n = (int)f->kids.size();
    const bool go = n > 1 && nao > 0 && ngrid >= 2 * n * FAN_BLOCK && (double)ngrid * nao * nao >= f->min_work;
    if (!go) f->last_call_fanned = false;
    return go;
}

void fanout_invalidate(CublasHandleWrapper* ctx) {
    if (ctx && ctx->fan) static_cast<FanOut*>(ctx->fan)->valid = false;
}

int fanout_set_option(CublasHandleWrapper* ctx, const char* key, double value) {
    if (!strcmp(key, "devices")) return fanout_configure(ctx, (int)value, false);
    if (!strcmp(key, "virtual_devices")) return fanout_configure(ctx, (int)value, true);
    if (!strcmp(key, "ao_invalidate")) { fanout_invalidate(ctx); return 0; }
    if (!strcmp(key, "fan_threads")) {   // 0: the caller's thread enqueues every device's work itself (measurement)
        if (!ctx->fan) return 3;
        static_cast<FanOut*>(ctx->fan)->threads = value != 0.0;   // (idle workers stay parked)
        return 0;
    }
    if (!strcmp(key, "devices_min_work") || !strcmp(key, "ao_cache")) {
        if (!ctx->fan) return 3;   // (set "devices" first)
        FanOut* f = static_cast<FanOut*>(ctx->fan);
        if (key[0] == 'd') f->min_work = value;
        else { f->cache = value != 0.0; f->valid = false; }
        return 0;
    }
    return 2;
}

void fanout_forward_option(CublasHandleWrapper* ctx, const char* key, double value) {
    if (!ctx || !ctx->fan) return;
    for (auto& k : static_cast<FanOut*>(ctx->fan)->kids) {
        DeviceGuard guard(k.dev);
        set_option(k.eng, key, value);
    }
}

double fanout_stat(CublasHandleWrapper* ctx, const char* key, bool* known) {
    *known = true;
    FanOut* f = ctx ? static_cast<FanOut*>(ctx->fan) : nullptr;
    if (!strcmp(key, "devices")) return f ? (double)f->kids.size() : 1.0;
    if (!strcmp(key, "fan_active")) return f && f->last_call_fanned ? 1.0 : 0.0;
    if (!strcmp(key, "fan_scatters")) return f ? f->scatters : 0.0;
    if (!strcmp(key, "fan_peer_loads")) return f && f->peer_loads ? 1.0 : 0.0;
    if (!strcmp(key, "fan_resident_bytes")) {
        double b = 0.0;
        if (f) for (auto& k : f->kids) b += (double)(k.ao.capacity + k.grad.capacity + k.w.capacity);
        return b;
    }
    if (f && f->last_call_fanned && strstr(key, "_ms") && strcmp(key, "ao_ms") && strcmp(key, "total_ms")) {
        // kernels of the slowest device (every child records its own events when "timing" is on)
        float worst = 0.f;
        for (auto& k : f->kids) {
            resolve_times(k.eng);
            const float v = !strcmp(key, "density_ms") ? k.eng->stats.density_ms
                          : !strcmp(key, "vxc_ms") ? k.eng->stats.vxc_ms : k.eng->stats.reduce_ms;
            worst = v > worst ? v : worst;
        }
        return worst;
    }
    *known = false;
    return 0.0;
}

double run_fanout(CublasHandleWrapper* ctx, int xc_type, int ngrid, int nao, const double* d_dm, const double* d_ao,
                  const double* d_ao_grad, const double* d_w, double* d_vxc, double* d_exc_out) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    FanOut* f = static_cast<FanOut*>(ctx->fan);
    const int n = (int)f->kids.size();
    DeviceGuard guard(ctx->device);
    ctx->failed = false;
    ctx->times_pending = false;
    f->last_call_fanned = true;
    cudaStream_t s0 = ctx->stream;
    const size_t n2 = (size_t)nao * nao, plane = (size_t)ngrid * nao;
    const int nplanes = xc_type ? 4 : 1;
    double* packed = (double*)ctx->result.ensure(sizeof(double) * (n2 + 2), &ctx->failed);
    unsigned long long* d_fp = (unsigned long long*)f->fp.ensure(sizeof(unsigned long long), &ctx->failed);
    if (!packed || !d_fp) return nan;
    if (ctx->timing) cudaEventRecord(ctx->ev[0], s0);

    // 1. are the resident shards still cut from these arrays?  The engine stream is a blocking stream, so the wait
    //    below also puts everything the caller enqueued before this call (d_dm.set, dft.py:200) ahead of the
    //    children's copies on the other devices.
    if (ctx->num_sms <= 0) cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, ctx->device);
    cudaMemsetAsync(d_fp, 0, sizeof(unsigned long long), s0);
    fingerprint_kernel<<<ctx->num_sms * 2, 256, 0, s0>>>(d_ao, d_ao_grad, d_w, plane, ngrid, nplanes, d_fp);
    unsigned long long* h_fp = reinterpret_cast<unsigned long long*>(ctx->h_scalar + 12);
    DFT_CUDA_CHECK(ctx, cudaMemcpyAsync(h_fp, d_fp, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s0));
    DFT_CUDA_CHECK(ctx, cudaStreamSynchronize(s0));
    if (ctx->failed) return nan;
    const bool stale = !f->cache || !f->valid || f->ao != d_ao || f->grad != d_ao_grad || f->w != d_w || f->ngrid != ngrid ||
                       f->nao != nao || f->xc_type != xc_type || f->fingerprint != *h_fp;

    // 2. every child: (shards), D, the ordinary single-GPU build on its own device and stream -- child 0 from this
    //    thread, the others from their worker threads
    FanJob& j = f->job;
    j.xc_type = xc_type; j.ngrid = ngrid; j.nao = nao;
    j.dm = d_dm; j.ao = d_ao; j.grad = d_ao_grad; j.w = d_w;
    j.stale = stale;
    j.primary_device = ctx->device;
    bool failed = false;
    if (f->threads) {
        unsigned long long s;
        {
            std::lock_guard<std::mutex> lk(f->mu);
            s = f->seq.load(std::memory_order_relaxed) + 1;
            f->seq.store(s, std::memory_order_release);
        }
        f->cv.notify_all();
        enqueue_child(f, 0);
        for (int c = 1; c < n; ++c)
            while (f->kids[c].done_seq.load(std::memory_order_acquire) != s) std::this_thread::yield();
    } else {
        for (int c = 0; c < n; ++c) enqueue_child(f, c);
    }
    for (int c = 0; c < n; ++c) failed = failed || f->kids[c].job_failed;
    if (stale) {
        f->valid = !failed;
        f->ao = d_ao; f->grad = d_ao_grad; f->w = d_w; f->ngrid = ngrid; f->nao = nao; f->xc_type = xc_type;
        f->fingerprint = *h_fp;
        f->scatters += 1;
        if (getenv("DFT_B200_VERBOSE"))
            fprintf(stderr, "[dft_b200] fan-out: shards of %d x %d (%s) cut for %d devices (cut #%d)\n", ngrid, nao,
                    xc_type == 0 ? "LDA" : xc_type == 1 ? "GGA" : "B3LYP", n, f->scatters);
    }

    // 3. one reduction on the primary device, straight into the caller's array
    FanSources src;
    src.n = n;
    double* stage = nullptr;
    if (!f->peer_loads) stage = (double*)f->stage.ensure((size_t)n * (n2 + 1) * sizeof(double), &ctx->failed);
    for (int c = 0; c < n && !failed && !ctx->failed; ++c) {
        FanChild& k = f->kids[c];
        DFT_CUDA_CHECK(ctx, cudaStreamWaitEvent(s0, k.done, 0));
        if (k.dev != ctx->device && !f->peer_loads) {
            double* dst = stage + (size_t)c * (n2 + 1);
            DFT_CUDA_CHECK(ctx, cudaMemcpyPeerAsync(dst, ctx->device, k.out.ptr, k.dev, (n2 + 1) * sizeof(double), s0));
            src.p[c] = dst;
        } else {
            src.p[c] = (const double*)k.out.ptr;
        }
    }
    if (!failed && !ctx->failed) {
        const int threads = 256;
        size_t blocks = (n2 + 1 + threads - 1) / threads;
        if (blocks > (size_t)ctx->num_sms * 8) blocks = (size_t)ctx->num_sms * 8;
        fan_reduce_kernel<<<(unsigned)blocks, threads, 0, s0>>>(src, d_vxc, packed + n2, n2);
        if (d_exc_out) cudaMemcpyAsync(d_exc_out, packed + n2, sizeof(double), cudaMemcpyDeviceToDevice, s0);
        DFT_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_scalar, packed + n2, sizeof(double), cudaMemcpyDeviceToHost, s0));
    }
    if (ctx->timing) cudaEventRecord(ctx->ev[4], s0);
    DFT_CUDA_CHECK(ctx, cudaStreamSynchronize(s0));
    DFT_CUDA_CHECK(ctx, cudaGetLastError());
    if (failed || ctx->failed) {   // let every child drain before the caller sees the failure
        for (auto& k : f->kids) { DeviceGuard child(k.dev); cudaStreamSynchronize(k.eng->stream); }
        ctx->failed = true;
        f->valid = false;
        return nan;
    }

    // 4. per-device adaptive state (the V kernel's fragment deal, the skipping instance) and the statistics
    XcStats agg;
    agg.plans_built = 0;
    double skip = 0.0, vskip = 0.0;
    for (auto& k : f->kids) {
        DeviceGuard child(k.dev);
        if (k.have_counters) apply_counters(k.eng);
        k.eng->times_pending = k.eng->timing && k.nshard > 0;
        agg.launches += k.eng->stats.launches;
        agg.plans_built += k.eng->stats.plans_built;
        agg.density_units += k.eng->stats.density_units;
        agg.density_groups += k.eng->stats.density_groups;
        agg.dyn_units += k.eng->stats.dyn_units;
        skip += k.eng->stats.skip_fraction / n;
        vskip += k.eng->stats.vxc_skip_fraction / n;
    }
    agg.path = f->kids[0].eng->stats.path;
    agg.v_tiles_m = f->kids[0].eng->stats.v_tiles_m;
    agg.launches += 2;   // fingerprint + reduction
    agg.skip_fraction = skip;
    agg.vxc_skip_fraction = vskip;
    agg.ao_ms = ctx->stats.ao_ms;
    ctx->stats = agg;
    if (ctx->timing && cudaEventSynchronize(ctx->ev[4]) == cudaSuccess)
        cudaEventElapsedTime(&ctx->stats.total_ms, ctx->ev[0], ctx->ev[4]);
    return ctx->h_scalar[0];
}

}  // namespace xc

extern "C" {
// Points of shard `shard` when `ngrid` points are dealt to `nshards` shards in interleaved blocks of 1024 (the deal the
// fan-out uses; the same as solver.shard_indices of the Python harness).  Pure host arithmetic: -1 on bad arguments.
int DFT_ShardPoints(int ngrid, int nshards, int shard) {
    if (ngrid < 0 || nshards < 1 || shard < 0 || shard >= nshards) return -1;
    int nfull = 0, tail = 0;
    deal(ngrid, nshards, shard, &nfull, &tail);
    return nfull * FAN_BLOCK + tail;
}
}
