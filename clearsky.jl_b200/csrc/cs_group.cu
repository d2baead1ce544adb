// cs_group.cu -- single-process multi-GPU plumbing: one context per device + one NCCL all-reduce.
//
// The path shards by contiguous wavenumber slices (every nu is independent through K2..K6); the only cross-device
// step is the sum of the 2*np spectrally integrated fluxes.  NCCL is loaded with dlopen so that the library has no
// link-time dependency on it (a host process such as PyTorch may bring its own libnccl.so.2).
#include "cs_internal.cuh"
#include <dlfcn.h>

namespace {

typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
constexpr int NCCL_DOUBLE = 8;   // ncclFloat64
constexpr int NCCL_SUM = 0;

int32_t load_nccl(NcclApi& api)
{
    if (api.lib) return CS_OK;
    const char* names[] = {getenv("CLEARSKY_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    CS_REQUIRE(api.lib, CS_ERR_CUDA, "cannot dlopen libnccl.so.2 (set CLEARSKY_B200_NCCL): %s", dlerror());
#define CS_SYM(field, sym)                                                             \
    *(void**)(&api.field) = dlsym(api.lib, sym);                                       \
    CS_REQUIRE(api.field, CS_ERR_CUDA, "symbol %s missing from NCCL", sym)
    CS_SYM(CommInitAll, "ncclCommInitAll");
    CS_SYM(CommDestroy, "ncclCommDestroy");
    CS_SYM(GroupStart, "ncclGroupStart");
    CS_SYM(GroupEnd, "ncclGroupEnd");
    CS_SYM(AllReduce, "ncclAllReduce");
    CS_SYM(GetErrorString, "ncclGetErrorString");
#undef CS_SYM
    return CS_OK;
}

}  // namespace

struct cs_group {
    std::vector<cs_ctx*> ctx;
    std::vector<int> dev;
    std::vector<DevBuf> buf;
    std::vector<ncclComm_t> comm;
    NcclApi nccl;
    std::mutex mtx;
    bool peer_ok = false;        // every device of the group can map every other one's memory (NVLink / NVSwitch peer access)
    bool rcm_nccl = false;       // CS_GROUP_RCM_NCCL: keep the NCCL all-reduce in cs_group_rcm_step
};

extern "C" int32_t cs_group_create(int32_t ndev, const int32_t* devices, cs_group** out)
{
    CS_REQUIRE(out && devices && ndev >= 1, CS_ERR_ARG, "bad group arguments");
    *out = nullptr;
    cs_group* g = new cs_group();
    for (int i = 0; i < ndev; i++) {
        cs_ctx* c = nullptr;
        int32_t rc = cs_ctx_create(devices[i], &c);
        if (rc) {
            cs_group_free(g);
            return rc;
        }
        g->ctx.push_back(c);
        g->dev.push_back(devices[i]);
    }
    g->buf.resize((size_t)ndev);
    for (int i = 0; i < ndev; i++) g->buf[(size_t)i].st = g->ctx[(size_t)i]->stream;
    if (ndev > 1) {
        int32_t rc = load_nccl(g->nccl);
        if (rc) { cs_group_free(g); return rc; }
        g->comm.resize((size_t)ndev, nullptr);
        ncclResult_t r = g->nccl.CommInitAll(g->comm.data(), ndev, g->dev.data());
        if (r != 0) {
            cs_set_error("ncclCommInitAll: %s", g->nccl.GetErrorString(r));
            g->comm.clear();
            cs_group_free(g);
            return CS_ERR_CUDA;
        }
    }
    // peer access between all pairs (the fused RCM step stores its partial sums straight into the other devices' mailboxes)
    g->rcm_nccl = getenv("CS_GROUP_RCM_NCCL") != nullptr;
    g->peer_ok = ndev > 1;
    for (int i = 0; i < ndev && g->peer_ok; i++) {
        for (int j = 0; j < ndev && g->peer_ok; j++) {
            if (i == j) continue;
            int can = 0;
            if (g->dev[(size_t)i] == g->dev[(size_t)j] ||
                cudaDeviceCanAccessPeer(&can, g->dev[(size_t)i], g->dev[(size_t)j]) != cudaSuccess || !can) {
                g->peer_ok = g->dev[(size_t)i] == g->dev[(size_t)j] ? g->peer_ok : false;
                continue;
            }
            cudaSetDevice(g->dev[(size_t)i]);
            cudaError_t e = cudaDeviceEnablePeerAccess(g->dev[(size_t)j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) g->peer_ok = false;
            cudaGetLastError();
        }
    }
    *out = g;
    return CS_OK;
}

extern "C" int32_t cs_group_free(cs_group* g)
{
    if (!g) return CS_OK;
    for (size_t i = 0; i < g->comm.size(); i++)
        if (g->comm[i]) g->nccl.CommDestroy(g->comm[i]);
    for (size_t i = 0; i < g->ctx.size(); i++) {
        cudaSetDevice(g->dev[i]);
        if (i < g->buf.size()) g->buf[i].release();
        cs_ctx_free(g->ctx[i]);
    }
    delete g;
    return CS_OK;
}

extern "C" int32_t cs_group_size(cs_group* g, int32_t* n)
{
    CS_REQUIRE(g && n, CS_ERR_ARG, "null argument");
    *n = (int32_t)g->ctx.size();
    return CS_OK;
}

extern "C" int32_t cs_group_ctx(cs_group* g, int32_t i, cs_ctx** c)
{
    CS_REQUIRE(g && c && i >= 0 && i < (int32_t)g->ctx.size(), CS_ERR_ARG, "bad group member index");
    *c = g->ctx[(size_t)i];
    return CS_OK;
}

extern "C" int32_t cs_group_buffer(cs_group* g, int32_t i, int64_t count, double** d_ptr)
{
    CS_REQUIRE(g && d_ptr && count > 0 && i >= 0 && i < (int32_t)g->ctx.size(), CS_ERR_ARG, "bad group buffer request");
    std::lock_guard<std::mutex> lk(g->mtx);
    CS_CUDA(cudaSetDevice(g->dev[(size_t)i]));
    CS_TRY(g->buf[(size_t)i].reserve(sizeof(double) * (size_t)count));
    *d_ptr = g->buf[(size_t)i].as<double>();
    return CS_OK;
}

extern "C" int32_t cs_group_allreduce_sum(cs_group* g, int64_t count)
{
    CS_REQUIRE(g && count > 0, CS_ERR_ARG, "bad all-reduce arguments");
    const int n = (int)g->ctx.size();
    for (int i = 0; i < n; i++)
        CS_REQUIRE(g->buf[(size_t)i].cap >= sizeof(double) * (size_t)count, CS_ERR_ARG, "group buffer %d smaller than %lld doubles", i, (long long)count);
    if (n == 1) return cs_ctx_synchronize(g->ctx[0]);
    std::lock_guard<std::mutex> lk(g->mtx);
    ncclResult_t r = g->nccl.GroupStart();
    for (int i = 0; i < n && r == 0; i++) {
        double* p = g->buf[(size_t)i].as<double>();
        r = g->nccl.AllReduce(p, p, (size_t)count, NCCL_DOUBLE, NCCL_SUM, g->comm[(size_t)i], g->ctx[(size_t)i]->stream);
    }
    ncclResult_t r2 = g->nccl.GroupEnd();
    if (r == 0) r = r2;
    CS_REQUIRE(r == 0, CS_ERR_CUDA, "ncclAllReduce: %s", g->nccl.GetErrorString(r));
    for (int i = 0; i < n; i++) {
        CS_CUDA(cudaSetDevice(g->dev[(size_t)i]));
        CS_CUDA(cudaStreamSynchronize(g->ctx[(size_t)i]->stream));
    }
    return CS_OK;
}

extern "C" int32_t cs_group_read(cs_group* g, int32_t i, int64_t count, double* host)
{
    CS_REQUIRE(g && host && count > 0 && i >= 0 && i < (int32_t)g->ctx.size(), CS_ERR_ARG, "bad group read");
    CS_REQUIRE(g->buf[(size_t)i].cap >= sizeof(double) * (size_t)count, CS_ERR_ARG, "group buffer smaller than the read");
    CS_CUDA(cudaSetDevice(g->dev[(size_t)i]));
    cudaStream_t st = g->ctx[(size_t)i]->stream;
    CS_CUDA(cudaMemcpyAsync(host, g->buf[(size_t)i].p, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    return CS_OK;
}

// radiative-convective steps on a nu-sharded column: every device holds the same column state and its own slice of the
// spectrum.  With peer access between the devices a step is two kernels per device and no collective call: the flux
// kernel and a tail kernel that exchanges the 2*nrad partial sums through peer-memory mailboxes and updates the column
// (cs_rcm_enqueue_step_peer).  Otherwise (or with CS_GROUP_RCM_NCCL set): partial fluxes (2 kernels) -> ONE ncclAllReduce of
// 2*nrad doubles -> column update (1 kernel).  Everything is enqueued from this one host thread without any synchronisation
// until the last step; the devices then hold bit-identical temperatures (same summed fluxes, same arithmetic).
extern "C" int32_t cs_group_rcm_step(cs_group* g, cs_rcm* const* rcm, double dt, int64_t nsteps)
{
    CS_REQUIRE(g && rcm && nsteps >= 0, CS_ERR_ARG, "bad arguments");
    const int n = (int)g->ctx.size();
    int64_t nrad = 0;
    for (int i = 0; i < n; i++) {
        CS_REQUIRE(rcm[i], CS_ERR_ARG, "null RCM handle for device %d", i);
        cs_ctx* c = nullptr;
        int64_t nr = 0;
        CS_TRY(cs_rcm_ctx(rcm[i], &c));
        CS_REQUIRE(c == g->ctx[(size_t)i], CS_ERR_ARG, "RCM handle %d does not live on group member %d", i, i);
        CS_TRY(cs_rcm_info(rcm[i], nullptr, &nr, nullptr));
        CS_REQUIRE(i == 0 || nr == nrad, CS_ERR_ARG, "RCM handles disagree on the number of radiative levels");
        nrad = nr;
    }
    std::lock_guard<std::mutex> lk(g->mtx);
    std::vector<double*> buf((size_t)n);
    for (int i = 0; i < n; i++) {
        CS_CUDA(cudaSetDevice(g->dev[(size_t)i]));
        CS_TRY(g->buf[(size_t)i].reserve(sizeof(double) * 2 * (size_t)nrad));
        buf[(size_t)i] = g->buf[(size_t)i].as<double>();
    }
    if (n > 1 && g->peer_ok && !g->rcm_nccl) {
        // fused form: no collective call -- each device's tail kernel stores its partial sums into every mailbox and sums
        // what arrives in its own (cs_rt.cu: rcm_tail_kernel).  Mailboxes are created and connected on first use.
        int64_t done = 0;
        int32_t late = 0;
        if (cs_rcm_peer_status(rcm[0], &done, &late) != CS_OK) {
            std::vector<void*> box((size_t)n, nullptr);
            for (int i = 0; i < n; i++) CS_TRY(cs_rcm_peer_mailbox(rcm[i], n, &box[(size_t)i], nullptr));
            for (int i = 0; i < n; i++) CS_TRY(cs_rcm_peer_connect(rcm[i], i, n, box.data()));
        }
        for (int64_t k = 0; k < nsteps; k++)
            for (int i = 0; i < n; i++) CS_TRY(cs_rcm_enqueue_step_peer(rcm[i], dt));
        for (int i = 0; i < n; i++) {
            CS_CUDA(cudaSetDevice(g->dev[(size_t)i]));
            CS_CUDA(cudaStreamSynchronize(g->ctx[(size_t)i]->stream));
        }
        for (int i = 0; i < n; i++) {
            CS_TRY(cs_rcm_peer_status(rcm[i], &done, &late));
            CS_REQUIRE(!late, CS_ERR_CUDA, "device %d: the partial fluxes of another device did not arrive within the time limit", i);
        }
        return CS_OK;
    }
    for (int64_t k = 0; k < nsteps; k++) {
        for (int i = 0; i < n; i++) CS_TRY(cs_rcm_enqueue_fluxes(rcm[i], buf[(size_t)i]));
        if (n > 1) {
            ncclResult_t r = g->nccl.GroupStart();
            for (int i = 0; i < n && r == 0; i++)
                r = g->nccl.AllReduce(buf[(size_t)i], buf[(size_t)i], 2 * (size_t)nrad, NCCL_DOUBLE, NCCL_SUM, g->comm[(size_t)i],
                                      g->ctx[(size_t)i]->stream);
            ncclResult_t r2 = g->nccl.GroupEnd();
            if (r == 0) r = r2;
            CS_REQUIRE(r == 0, CS_ERR_CUDA, "ncclAllReduce: %s", g->nccl.GetErrorString(r));
        }
        for (int i = 0; i < n; i++) CS_TRY(cs_rcm_enqueue_update(rcm[i], buf[(size_t)i], dt));
    }
    for (int i = 0; i < n; i++) {
        CS_CUDA(cudaSetDevice(g->dev[(size_t)i]));
        CS_CUDA(cudaStreamSynchronize(g->ctx[(size_t)i]->stream));
    }
    return CS_OK;
}
