// comm.cu -- grid-sharded multi-GPU: one process per GPU, NCCL all-reduce of [V_xc | E_xc].
//
// The reference has no multi-device code at all.  Grid points are independent,
// so every rank integrates its own slice of (Phi, grad Phi, w) and the only exchange is one
// ncclAllReduce(sum, double) of nao*nao + 1 values over NVLink (<= 1.14 MB at nao = 377).
// NCCL is opened with dlopen on first use so that the single-GPU drop-in has no NCCL dependency
// (and inherits whichever libnccl.so.2 the host process already loaded, e.g. torch's).
#include <dlfcn.h>

#include <cstring>

#include "../../include/dft_b200_ext.h"
#include "engine.h"

namespace {

struct NcclId { char internal[128]; };  // layout of ncclUniqueId
typedef int (*fn_get_id)(NcclId*);
typedef int (*fn_init_rank)(void**, int, NcclId, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
typedef const char* (*fn_errstr)(int);
typedef int (*fn_group)(void);

struct NcclApi {
    void* handle = nullptr;
    fn_get_id get_id = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
    fn_group group_start = nullptr, group_end = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.handle || api.ok) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        fprintf(stderr, "[dft_b200] cannot dlopen libnccl.so.2: %s\n", dlerror());
        return api;
    }
    api.get_id = (fn_get_id)dlsym(api.handle, "ncclGetUniqueId");
    api.init_rank = (fn_init_rank)dlsym(api.handle, "ncclCommInitRank");
    api.allreduce = (fn_allreduce)dlsym(api.handle, "ncclAllReduce");
    api.destroy = (fn_destroy)dlsym(api.handle, "ncclCommDestroy");
    api.errstr = (fn_errstr)dlsym(api.handle, "ncclGetErrorString");
    api.group_start = (fn_group)dlsym(api.handle, "ncclGroupStart");
    api.group_end = (fn_group)dlsym(api.handle, "ncclGroupEnd");
    api.ok = api.get_id && api.init_rank && api.allreduce && api.destroy && api.group_start && api.group_end;
    return api;
}

constexpr int kNcclDouble = 8;  // ncclFloat64
constexpr int kNcclSum = 0;

}  // namespace

namespace xc {
// One grouped NCCL operation (a single launch): V_xc = sum over ranks of d_packed[0 .. n2), written OUT OF PLACE
// straight into the caller's array (no copy-back), and [E_xc | failed ranks] = d_packed[n2 .. n2 + 2) in place.
int allreduce_result(CublasHandleWrapper* ctx, double* d_packed, double* d_vxc, size_t n2) {
    if (!ctx || ctx->nranks <= 1 || !ctx->nccl_comm) return 0;
    NcclApi& api = nccl();
    if (!api.ok) return 1;
    int rc = api.group_start();
    if (rc == 0) rc = api.allreduce(d_packed, d_vxc, n2, kNcclDouble, kNcclSum, ctx->nccl_comm, ctx->stream);
    if (rc == 0) rc = api.allreduce(d_packed + n2, d_packed + n2, 2, kNcclDouble, kNcclSum, ctx->nccl_comm, ctx->stream);
    const int rc_end = api.group_end();
    if (rc == 0) rc = rc_end;
    if (rc != 0) {
        fprintf(stderr, "[dft_b200] ncclAllReduce failed: %s\n", api.errstr ? api.errstr(rc) : "?");
        return 2;
    }
    return 0;
}

void comm_destroy(CublasHandleWrapper* ctx) {
    if (!ctx) return;
    if (ctx->nccl_comm) {
        NcclApi& api = nccl();
        if (api.ok) api.destroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->rank = 0;
    ctx->nranks = 1;
}
}  // namespace xc

extern "C" {

int DFT_CommGetUniqueId(void* out_id_128_bytes) {
    if (!out_id_128_bytes) return 1;
    NcclApi& api = nccl();
    if (!api.ok) return 2;
    NcclId id;
    if (api.get_id(&id) != 0) return 3;
    memcpy(out_id_128_bytes, &id, sizeof(id));
    return 0;
}

int DFT_CommInit(XCSolver* solver, int rank, int nranks, const void* id_128_bytes) {
    if (!solver || !id_128_bytes || nranks < 1 || rank < 0 || rank >= nranks) return 1;
    CublasHandleWrapper* ctx = solver->context();
    DeviceGuard guard(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    xc::comm_destroy(ctx);   // (a second DFT_CommInit replaces the communicator instead of leaking it)
    if (nranks == 1) return 0;
    NcclApi& api = nccl();
    if (!api.ok) return 2;
    NcclId id;
    memcpy(&id, id_128_bytes, sizeof(id));
    void* comm = nullptr;
    int rc = api.init_rank(&comm, nranks, id, rank);
    if (rc != 0) {
        fprintf(stderr, "[dft_b200] ncclCommInitRank failed: %s\n", api.errstr ? api.errstr(rc) : "?");
        return 3;
    }
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return 0;
}

int DFT_CommDestroy(XCSolver* solver) {
    if (!solver) return 1;
    CublasHandleWrapper* ctx = solver->context();
    DeviceGuard guard(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    xc::comm_destroy(ctx);
    return 0;
}
}
