// fd_mgpu.cu -- fd_mgpu_*: several GPUs of one box behind one handle, for a single-process caller such as the Houdini
// plugin (one SOP node = one set of member handles, reference SOP_FaceDeform.hpp:108-113).
//
// The evaluation loop of the reference (SOP_FaceDeform.cpp:404-439) has no cross-vertex dependence (the disabled
// UTparallelFor at SOP_FaceDeform.hpp:181-188 says the same), so the vertices are split into contiguous ranges, one
// per device; the control-point system is assembled, factored and solved on the first device only and the weights
// cross NVLink once per solve:
//   p2p   every other device builds its evaluation tables (FP32 rows, or the column-scaled FP16 hi/lo tiles of the
//         tensor path) by reading the root's FP64 weight block through peer loads -- the broadcast and the pack are
//         the same kernels, nothing is staged -- and pulls the FP64 block itself only if the FP64 evaluation will run;
//   nccl  ncclBroadcast of the weight block and the radii inside one group call (libnccl.so.2 is loaded at run time:
//         the library has no link-time dependency on NCCL), then the ordinary fd_model_commit_weights.
// Per device one fd_ctx with its own stream; cross-device order is carried by events, the host never waits between the
// root's solve and the other devices' table builds.  fd_mgpu_eval drives the devices from one host thread each, so
// pageable caller buffers still copy concurrently.
#include <dlfcn.h>
#include <string.h>

#include <new>
#include <thread>
#include <vector>

#include "fd_internal.h"

namespace {

constexpr int MAX_DEV = 16;

// the five NCCL entry points used, resolved with dlsym (signatures as in nccl.h 2.x)
typedef struct ncclComm* nccl_comm_t;
typedef int (*nccl_comm_init_all_fn)(nccl_comm_t*, int, const int*);
typedef int (*nccl_comm_destroy_fn)(nccl_comm_t);
typedef int (*nccl_group_fn)(void);
typedef int (*nccl_broadcast_fn)(const void*, void*, size_t, int /*ncclDataType_t*/, int, nccl_comm_t, cudaStream_t);
typedef const char* (*nccl_error_fn)(int);
constexpr int NCCL_FLOAT64 = 8; // ncclFloat64 / ncclDouble

struct Nccl {
    void* lib = nullptr;
    nccl_comm_init_all_fn init_all = nullptr;
    nccl_comm_destroy_fn destroy = nullptr;
    nccl_group_fn group_start = nullptr, group_end = nullptr;
    nccl_broadcast_fn broadcast = nullptr;
    nccl_error_fn error_string = nullptr;
    bool load()
    {
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        init_all = (nccl_comm_init_all_fn)dlsym(lib, "ncclCommInitAll");
        destroy = (nccl_comm_destroy_fn)dlsym(lib, "ncclCommDestroy");
        group_start = (nccl_group_fn)dlsym(lib, "ncclGroupStart");
        group_end = (nccl_group_fn)dlsym(lib, "ncclGroupEnd");
        broadcast = (nccl_broadcast_fn)dlsym(lib, "ncclBroadcast");
        error_string = (nccl_error_fn)dlsym(lib, "ncclGetErrorString");
        return init_all && destroy && group_start && group_end && broadcast && error_string;
    }
};

} // namespace

struct fd_mgpu {
    int ndev = 0;
    int dev[MAX_DEV] = {};
    fd_ctx* ctx[MAX_DEV] = {};
    fd_model* model[MAX_DEV] = {}; // [0]: the fitted model; others: receivers
    int transport = FD_MGPU_P2P;
    Nccl nccl;
    nccl_comm_t comm[MAX_DEV] = {};
    bool have_comm = false;
    cudaEvent_t ev_solved = nullptr;                 // on the root's stream: the weights are final
    cudaEvent_t ev_begin[MAX_DEV] = {}, ev_ready[MAX_DEV] = {};
    fd_params prm;
    std::vector<float> rest;
    int N = 0, F = 0;
    int64_t bcast_bytes = 0;
    float bcast_ms = -1.f;
    bool timed = false;
    char err[512] = {};
};

#define MG_ERR(g, ...) snprintf((g)->err, sizeof((g)->err), __VA_ARGS__)

extern "C" {

int fd_mgpu_create(fd_mgpu** out, const int* devices, int32_t ndev, int32_t transport)
{
    if (!out || ndev < 1 || ndev > MAX_DEV || transport < FD_MGPU_AUTO || transport > FD_MGPU_P2P) return FD_E_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return FD_E_CUDA; // no GPU: no fallback
    fd_mgpu* g = new (std::nothrow) fd_mgpu();
    if (!g) return FD_E_NOMEM;
    g->ndev = ndev;
    for (int i = 0; i < ndev; ++i) {
        g->dev[i] = devices ? devices[i] : i;
        for (int k = 0; k < i; ++k)
            if (g->dev[k] == g->dev[i]) { delete g; return FD_E_INVALID; }
        if (g->dev[i] < 0 || g->dev[i] >= count) { delete g; return FD_E_INVALID; }
    }
    int prev = -1;
    cudaGetDevice(&prev);
    int st = FD_OK;
    for (int i = 0; i < ndev && st == FD_OK; ++i) st = fd_ctx_create(&g->ctx[i], g->dev[i], nullptr);
    // peer access to the root's memory from every other device (and back, for symmetry of later root changes)
    bool peer_ok = true;
    for (int i = 1; i < ndev && st == FD_OK; ++i) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, g->dev[i], g->dev[0]);
        if (!can) { peer_ok = false; continue; }
        cudaSetDevice(g->dev[i]);
        cudaError_t e = cudaDeviceEnablePeerAccess(g->dev[0], 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
        if (e != cudaSuccess) { peer_ok = false; continue; }
        // the models live in the root's stream-ordered pool (cudaMallocAsync): pools grant peer access separately
        cudaMemPool_t pool;
        cudaMemAccessDesc desc;
        memset(&desc, 0, sizeof(desc));
        desc.location.type = cudaMemLocationTypeDevice;
        desc.location.id = g->dev[i];
        desc.flags = cudaMemAccessFlagsProtReadWrite;
        if (cudaDeviceGetDefaultMemPool(&pool, g->dev[0]) != cudaSuccess || cudaMemPoolSetAccess(pool, &desc, 1) != cudaSuccess) {
            cudaGetLastError();
            peer_ok = false;
        }
    }
    if (st == FD_OK) {
        if (transport == FD_MGPU_P2P && !peer_ok && ndev > 1) {
            st = FD_E_UNSUPPORTED;
        } else if (transport == FD_MGPU_NCCL || (transport == FD_MGPU_AUTO && !peer_ok && ndev > 1)) {
            g->transport = FD_MGPU_NCCL;
            if (!g->nccl.load()) {
                st = FD_E_UNSUPPORTED;
            } else if (ndev > 1) {
                const int r = g->nccl.init_all(g->comm, ndev, g->dev);
                if (r != 0) st = FD_E_CUDA; else g->have_comm = true;
            }
        } else {
            g->transport = FD_MGPU_P2P;
        }
    }
    if (st == FD_OK) {
        cudaSetDevice(g->dev[0]);
        if (cudaEventCreateWithFlags(&g->ev_solved, cudaEventDisableTiming) != cudaSuccess) st = FD_E_CUDA;
        for (int i = 0; i < ndev && st == FD_OK; ++i) {
            cudaSetDevice(g->dev[i]);
            if (cudaEventCreate(&g->ev_begin[i]) != cudaSuccess || cudaEventCreate(&g->ev_ready[i]) != cudaSuccess) st = FD_E_CUDA;
        }
    }
    if (prev >= 0) cudaSetDevice(prev);
    if (st != FD_OK) {
        fd_mgpu_destroy(g);
        return st;
    }
    *out = g;
    return FD_OK;
}

void fd_mgpu_destroy(fd_mgpu* g)
{
    if (!g) return;
    int prev = -1;
    cudaGetDevice(&prev);
    for (int i = 0; i < g->ndev; ++i) {
        if (g->model[i]) fd_model_destroy(g->model[i]);
        if (g->ctx[i]) {
            cudaSetDevice(g->dev[i]);
            cudaStreamSynchronize(g->ctx[i]->stream);
            if (g->ev_begin[i]) cudaEventDestroy(g->ev_begin[i]);
            if (g->ev_ready[i]) cudaEventDestroy(g->ev_ready[i]);
        }
    }
    if (g->have_comm)
        for (int i = 0; i < g->ndev; ++i)
            if (g->comm[i]) g->nccl.destroy(g->comm[i]);
    if (g->ev_solved) { cudaSetDevice(g->dev[0]); cudaEventDestroy(g->ev_solved); }
    for (int i = 0; i < g->ndev; ++i)
        if (g->ctx[i]) fd_ctx_destroy(g->ctx[i]);
    if (prev >= 0) cudaSetDevice(prev);
    delete g;
}

const char* fd_mgpu_last_error(const fd_mgpu* g) { return g ? g->err : "null handle"; }
fd_ctx* fd_mgpu_ctx(fd_mgpu* g, int32_t i) { return (g && i >= 0 && i < g->ndev) ? g->ctx[i] : nullptr; }

int fd_mgpu_range(const fd_mgpu* g, int32_t i, int64_t n_vtx, int64_t* begin, int64_t* end)
{
    if (!g || i < 0 || i >= g->ndev || n_vtx < 0 || !begin || !end) return FD_E_INVALID;
    // the partition of facedeform_b200/shard.py: contiguous, the first n_vtx % ndev ranges one vertex longer
    const int64_t base = n_vtx / g->ndev, rem = n_vtx % g->ndev;
    *begin = i * base + (i < rem ? i : rem);
    *end = *begin + base + (i < rem ? 1 : 0);
    return FD_OK;
}

int fd_mgpu_info(const fd_mgpu* g, int32_t* ndev, int32_t* transport, int64_t* bcast_bytes, float* bcast_ms)
{
    if (!g) return FD_E_INVALID;
    if (ndev) *ndev = g->ndev;
    if (transport) *transport = g->transport;
    if (bcast_bytes) *bcast_bytes = g->bcast_bytes;
    if (bcast_ms) {
        float worst = -1.f;
        if (g->timed) {
            int prev = -1;
            cudaGetDevice(&prev);
            for (int i = 1; i < g->ndev; ++i) {
                float ms = -1.f;
                cudaSetDevice(g->dev[i]);
                if (cudaEventSynchronize(g->ev_ready[i]) == cudaSuccess &&
                    cudaEventElapsedTime(&ms, g->ev_begin[i], g->ev_ready[i]) == cudaSuccess && ms > worst)
                    worst = ms;
            }
            if (prev >= 0) cudaSetDevice(prev);
            if (g->ndev == 1) worst = 0.f;
        }
        *bcast_ms = worst;
    }
    return FD_OK;
}

// rbfcreate .. rbfbuildmodel on the root device (SOP_FaceDeform.cpp:331-363)
int fd_mgpu_fit(fd_mgpu* g, const fd_params* params, const float* rest_ctrl, int32_t n_ctrl, fd_report* report)
{
    if (!g || !params || !rest_ctrl || n_ctrl < 1) return FD_E_INVALID;
    for (int i = 0; i < g->ndev; ++i)
        if (g->model[i]) { fd_model_destroy(g->model[i]); g->model[i] = nullptr; }
    g->F = 0;
    g->timed = false;
    const int st = fd_rbf_fit(g->ctx[0], params, rest_ctrl, n_ctrl, &g->model[0], report);
    if (st != FD_OK) { MG_ERR(g, "%s", fd_last_error(g->ctx[0])); return st; }
    g->prm = *params;
    g->N = n_ctrl;
    g->rest.assign(rest_ctrl, rest_ctrl + (size_t)n_ctrl * 3);
    return FD_OK;
}

// the root solves all 3F right-hand sides, then the weights reach the other devices (see the file header)
int fd_mgpu_solve(fd_mgpu* g, const float* deform_ctrl, int32_t n_ctrl, int32_t frames, fd_report* report)
{
    if (!g || !deform_ctrl) return FD_E_INVALID;
    if (!g->model[0]) { MG_ERR(g, "solve before fit"); return FD_E_STATE; }
    fd_model* root = g->model[0];
    // host-pointer solve without the report's synchronisation: the status is read after the broadcast was enqueued
    const size_t bytes = (size_t)frames * n_ctrl * 3 * sizeof(float);
    if (n_ctrl != g->N) { MG_ERR(g, "%s", fd_status_string(FD_E_MISMATCH_POINT)); return FD_E_MISMATCH_POINT; }
    if (frames < 1) { MG_ERR(g, "frames must be >= 1"); return FD_E_INVALID; }
    int prev = -1;
    cudaGetDevice(&prev);
    int st = FD_OK;
    {
        cudaSetDevice(g->dev[0]);
        void* d_def = nullptr;
        st = fd_stage(g->ctx[0], FD_STAGE_P, bytes, &d_def);
        if (st == FD_OK && cudaMemcpyAsync(d_def, deform_ctrl, bytes, cudaMemcpyHostToDevice, g->ctx[0]->stream) != cudaSuccess) st = FD_E_CUDA;
        if (st == FD_OK) st = fd_rbf_solve_dev(root, (const float*)d_def, n_ctrl, frames);
        if (st == FD_OK && cudaEventRecord(g->ev_solved, g->ctx[0]->stream) != cudaSuccess) st = FD_E_CUDA;
        if (st != FD_OK) MG_ERR(g, "%s", fd_last_error(g->ctx[0]));
    }
    // receivers (re-created when the frame count grows past their reservation)
    for (int i = 1; i < g->ndev && st == FD_OK; ++i) {
        if (g->model[i] && g->F != frames) { fd_model_destroy(g->model[i]); g->model[i] = nullptr; }
        if (!g->model[i]) {
            st = fd_model_create_receiver(g->ctx[i], &g->prm, g->rest.data(), g->N, frames, &g->model[i]);
            if (st != FD_OK) MG_ERR(g, "device %d: %s", g->dev[i], fd_last_error(g->ctx[i]));
        }
    }
    void* w_ptr = nullptr;
    void* r_ptr = nullptr;
    size_t w_bytes = 0, r_bytes = 0;
    if (st == FD_OK) st = fd_model_weights_dev(root, &w_ptr, &w_bytes);
    if (st == FD_OK) st = fd_model_radii_dev(root, &r_ptr, &r_bytes);
    if (st == FD_OK && g->ndev > 1) {
        for (int i = 1; i < g->ndev; ++i) { // order every receiver's stream after the root's solve
            cudaSetDevice(g->dev[i]);
            cudaStreamWaitEvent(g->ctx[i]->stream, g->ev_solved, 0);
            cudaEventRecord(g->ev_begin[i], g->ctx[i]->stream);
        }
        if (g->transport == FD_MGPU_NCCL) {
            void *wi = nullptr, *ri = nullptr;
            size_t wb = 0, rb = 0;
            int r = g->nccl.group_start();
            for (int i = 0; i < g->ndev && r == 0; ++i) {
                if (i == 0) { wi = w_ptr; ri = r_ptr; }
                else { fd_model_weights_dev(g->model[i], &wi, &wb); fd_model_radii_dev(g->model[i], &ri, &rb); }
                r = g->nccl.broadcast(w_ptr, wi, w_bytes / sizeof(double), NCCL_FLOAT64, 0, g->comm[i], g->ctx[i]->stream);
                if (r == 0) r = g->nccl.broadcast(r_ptr, ri, r_bytes / sizeof(double), NCCL_FLOAT64, 0, g->comm[i], g->ctx[i]->stream);
            }
            const int r2 = g->nccl.group_end();
            if (r == 0) r = r2;
            if (r != 0) { MG_ERR(g, "NCCL: %s", g->nccl.error_string(r)); st = FD_E_CUDA; }
            for (int i = 1; i < g->ndev && st == FD_OK; ++i) {
                st = fd_model_commit_weights(g->model[i]);
                if (st != FD_OK) MG_ERR(g, "device %d: %s", g->dev[i], fd_last_error(g->ctx[i]));
            }
        } else {
            for (int i = 1; i < g->ndev && st == FD_OK; ++i) {
                st = fd_model_commit_from_peer(g->model[i], (const double*)w_ptr, (const double*)r_ptr, g->dev[0]);
                if (st != FD_OK) MG_ERR(g, "device %d: %s", g->dev[i], fd_last_error(g->ctx[i]));
            }
        }
        for (int i = 1; i < g->ndev; ++i) {
            cudaSetDevice(g->dev[i]);
            cudaEventRecord(g->ev_ready[i], g->ctx[i]->stream);
        }
        g->bcast_bytes = (int64_t)(w_bytes + r_bytes) * (g->ndev - 1);
        g->timed = st == FD_OK;
    } else if (st == FD_OK) {
        g->bcast_bytes = 0;
        g->timed = true;
    }
    if (prev >= 0) cudaSetDevice(prev);
    if (st != FD_OK) return st;
    g->F = frames;
    // the root's status (terminationtype == 1, SOP_FaceDeform.cpp:365-368); the root's weight block must not be
    // overwritten by a later solve before the receivers have read it: wait for them here (they are microseconds behind)
    st = fd_model_report(root, report);
    if (st != FD_OK) { MG_ERR(g, "%s", fd_last_error(g->ctx[0])); return st; }
    for (int i = 1; i < g->ndev; ++i) {
        const int s2 = fd_ctx_synchronize(g->ctx[i]);
        if (s2 != FD_OK) { MG_ERR(g, "device %d: %s", g->dev[i], fd_last_error(g->ctx[i])); return s2; }
    }
    return FD_OK;
}

// the vertex loop, one contiguous range per device (SOP_FaceDeform.cpp:404-439)
int fd_mgpu_eval(fd_mgpu* g, const float* P, int64_t n_vtx, const float* dist2, const float* tangentu,
                 const float* tangentv, const float* normal, float* P_out, float* falloff_out)
{
    if (!g || n_vtx < 0 || (n_vtx > 0 && (!P || !P_out))) return FD_E_INVALID;
    if (!g->model[0] || g->F < 1) { MG_ERR(g, "eval before solve"); return FD_E_STATE; }
    if (n_vtx == 0) return FD_OK;
    int status[MAX_DEV] = {};
    const size_t pitch = (size_t)n_vtx * 3 * sizeof(float);
    auto work = [&](int i) {
        int64_t b = 0, e = 0;
        fd_mgpu_range(g, i, n_vtx, &b, &e);
        if (e <= b) { status[i] = FD_OK; return; }
        status[i] = fd_eval_host_strided(g->model[i], P + 3 * b, e - b, dist2 ? dist2 + b : nullptr,
                                         tangentu ? tangentu + 3 * b : nullptr, tangentv ? tangentv + 3 * b : nullptr,
                                         normal ? normal + 3 * b : nullptr, P_out + 3 * b, pitch,
                                         falloff_out ? falloff_out + b : nullptr);
    };
    std::vector<std::thread> threads;
    for (int i = 1; i < g->ndev; ++i) threads.emplace_back(work, i);
    work(0);
    for (auto& t : threads) t.join();
    for (int i = 0; i < g->ndev; ++i)
        if (status[i] != FD_OK) {
            MG_ERR(g, "device %d: %s", g->dev[i], fd_last_error(g->ctx[i]));
            return status[i];
        }
    return FD_OK;
}

} // extern "C"
