// multigpu.cu -- include/ismpc_b200_multigpu.h: the batched MPCSolver::solve on all GPUs of one box.
// One handle + stream + persistent host thread per device, contiguous shards, no per-tick communication, one NCCL
// all-gather of the result records at the end of a run.  Built into lib/libismpc_b200_mg.so on top
// of the C ABI of libismpc_b200.so (it uses nothing of the library's internals), the CUDA runtime and NCCL.
#include <cuda_runtime.h>
#include <nccl.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ismpc_b200_multigpu.h"

namespace {

struct DBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(&p, bytes) != cudaSuccess) return -1;
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Shard {
    int rank = 0, device = 0;
    ismpc_handle* h = nullptr;
    cudaStream_t stream = nullptr;
    int first = 0, count = 0;                 // the resident shard (ismpc_group_formc_scatter)
    bool pushes_pending = false;
    DBuf state, walk, inst, push, status, status_tmp, gs, gw, gst;
    ncclComm_t comm = nullptr;
};

__global__ void or_into_kernel(int n, int32_t* __restrict__ acc, const int32_t* __restrict__ add)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] |= add[i];
}

}  // namespace

struct ismpc_group {
    std::vector<Shard> shards;
    int max_batch = 0, gather_mode = ISMPC_GATHER_NCCL;
    int n_total = 0;
    std::string err;
    // persistent workers: worker r runs job(r) on device shards[r].device
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<long long> generation{0};
    std::atomic<int> done{0};
    bool quit = false;
    std::atomic<bool> quit_flag{false};
    int spin_us = 300;
    std::function<int(Shard&)> job;
    std::vector<int> rcs;

    void worker_main(int r)
    {
        cudaSetDevice(shards[(size_t)r].device);
        long long seen = 0;
        for (;;) {
            // poll for a while (a tick that follows the previous one closely starts within a microsecond), then sleep
            bool go = false;
            const auto t_spin = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us);
            while (std::chrono::steady_clock::now() < t_spin)
                if (quit_flag.load(std::memory_order_acquire) || generation.load(std::memory_order_acquire) != seen) { go = true; break; }
            if (!go) {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return quit || generation.load(std::memory_order_acquire) != seen; });
            }
            if (quit_flag.load(std::memory_order_acquire)) return;
            seen = generation.load(std::memory_order_acquire);
            rcs[(size_t)r] = job(shards[(size_t)r]);
            done.fetch_add(1, std::memory_order_release);
        }
    }
    // run `f` for every shard, each on its own host thread (shard 0 on the caller's thread); first non-zero code wins
    int run(std::function<int(Shard&)> f)
    {
        const int G = (int)shards.size();
        job = std::move(f);
        done.store(0, std::memory_order_release);
        if (G > 1) {
            { std::lock_guard<std::mutex> lk(mu); generation.fetch_add(1, std::memory_order_release); }
            cv.notify_all();
        }
        cudaSetDevice(shards[0].device);
        rcs[0] = job(shards[0]);
        while (done.load(std::memory_order_acquire) < G - 1) std::this_thread::yield();
        for (int r = 0; r < G; ++r) if (rcs[(size_t)r] != 0) return rcs[(size_t)r];
        return 0;
    }
    int fail(int code, const char* what, Shard* s = nullptr)
    {
        char buf[512];
        snprintf(buf, sizeof(buf), "%s: %s%s%s", what, ismpc_error_string(code), s && s->h ? " -- " : "",
                 s && s->h ? ismpc_last_cuda_error(s->h) : "");
        std::lock_guard<std::mutex> lk(mu);
        if (err.empty()) err = buf;
        return code;
    }
};

static void shard_of(int n_total, int G, int r, int* first, int* count)
{
    const int base = n_total / G, extra = n_total % G;
    *first = r * base + (r < extra ? r : extra);
    *count = base + (r < extra ? 1 : 0);
}

extern "C" int ismpc_group_create(ismpc_group** out, const int* devices, int n_devices, int max_batch_per_device, int gather_mode)
{
    if (!out || !devices || n_devices < 1 || n_devices > 64 || max_batch_per_device < 1) return ISMPC_ERR_ARG;
    if (gather_mode != ISMPC_GATHER_NCCL && gather_mode != ISMPC_GATHER_HOST) return ISMPC_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) return ISMPC_ERR_CUDA;
    for (int r = 0; r < n_devices; ++r) if (devices[r] < 0 || devices[r] >= count) return ISMPC_ERR_CUDA;
    ismpc_group* g = new (std::nothrow) ismpc_group();
    if (!g) return ISMPC_ERR_ALLOC;
    g->max_batch = max_batch_per_device; g->gather_mode = gather_mode;
    if (const char* v = getenv("ISMPC_GROUP_SPIN_US")) g->spin_us = atoi(v) > 0 ? atoi(v) : 0;
    g->shards.resize((size_t)n_devices); g->rcs.assign((size_t)n_devices, 0);
    int rc = ISMPC_OK;
    for (int r = 0; r < n_devices && rc == ISMPC_OK; ++r) {
        Shard& s = g->shards[(size_t)r];
        s.rank = r; s.device = devices[r];
        rc = ismpc_create(&s.h, s.device, max_batch_per_device);
        if (rc != ISMPC_OK) break;
        if (cudaSetDevice(s.device) != cudaSuccess || cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) rc = ISMPC_ERR_CUDA;
    }
    if (rc == ISMPC_OK && gather_mode == ISMPC_GATHER_NCCL && n_devices > 1) {
        std::vector<ncclComm_t> comms((size_t)n_devices);
        if (ncclCommInitAll(comms.data(), n_devices, devices) != ncclSuccess) rc = ISMPC_ERR_CUDA;
        else for (int r = 0; r < n_devices; ++r) g->shards[(size_t)r].comm = comms[(size_t)r];
    }
    if (rc != ISMPC_OK) { ismpc_group_destroy(g); return rc; }
    for (int r = 1; r < n_devices; ++r) g->workers.emplace_back(&ismpc_group::worker_main, g, r);
    *out = g;
    return ISMPC_OK;
}

extern "C" int ismpc_group_destroy(ismpc_group* g)
{
    if (!g) return ISMPC_ERR_ARG;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->quit = true;
        g->quit_flag.store(true, std::memory_order_release);
    }
    g->cv.notify_all();
    for (std::thread& w : g->workers) w.join();
    for (Shard& s : g->shards) {
        cudaSetDevice(s.device);
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.comm) ncclCommDestroy(s.comm);
        DBuf* all[] = {&s.state, &s.walk, &s.inst, &s.push, &s.status, &s.status_tmp, &s.gs, &s.gw, &s.gst};
        for (DBuf* b : all) b->release();
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.h) ismpc_destroy(s.h);
    }
    delete g;
    return ISMPC_OK;
}

extern "C" int ismpc_group_size(const ismpc_group* g) { return g ? (int)g->shards.size() : 0; }
extern "C" const char* ismpc_group_last_error(const ismpc_group* g) { return g ? g->err.c_str() : ""; }
extern "C" ismpc_handle* ismpc_group_handle(ismpc_group* g, int rank)
{
    return (g && rank >= 0 && rank < (int)g->shards.size()) ? g->shards[(size_t)rank].h : nullptr;
}
extern "C" int64_t ismpc_group_kernel_launches(const ismpc_group* g)
{
    int64_t t = 0;
    if (g) for (const Shard& s : g->shards) t += ismpc_kernel_launches(s.h);
    return t;
}
extern "C" int ismpc_group_shard(const ismpc_group* g, int n_total, int rank, int* first, int* count)
{
    if (!g || n_total < 0 || rank < 0 || rank >= (int)g->shards.size() || !first || !count) return ISMPC_ERR_ARG;
    shard_of(n_total, (int)g->shards.size(), rank, first, count);
    return ISMPC_OK;
}

extern "C" int ismpc_group_formc_configure(ismpc_group* g, const ismpc_formc_model_t* model, int S, int F_ds,
                                           const double* plan_xyzt, int plan_rows)
{
    if (!g || !model || !plan_xyzt || plan_rows <= 0) return ISMPC_ERR_ARG;
    g->err.clear();
    return g->run([=](Shard& s) -> int {
        int rc = ismpc_formc_set_model(s.h, model);
        if (rc != ISMPC_OK) return g->fail(rc, "ismpc_formc_set_model", &s);
        if (S + F_ds > 0 && (rc = ismpc_formc_prepare_gait(s.h, S, F_ds)) != ISMPC_OK) return g->fail(rc, "ismpc_formc_prepare_gait", &s);
        if ((rc = ismpc_formc_set_plan(s.h, plan_xyzt, plan_rows, ISMPC_MEM_HOST)) != ISMPC_OK) return g->fail(rc, "ismpc_formc_set_plan", &s);
        return ISMPC_OK;
    });
}

extern "C" int ismpc_group_formc_solve_batch(ismpc_group* g, int n_total, const ismpc_state_t* state, const ismpc_walk_t* walk,
                                             const ismpc_formc_inst_t* inst, ismpc_formc_out_t* out)
{
    if (!g || n_total < 0 || !state || !walk || !inst || !out) return ISMPC_ERR_ARG;
    const int G = (int)g->shards.size();
    if ((n_total + G - 1) / G > g->max_batch) return ISMPC_ERR_ARG;
    g->err.clear();
    return g->run([=](Shard& s) -> int {
        int first, count;
        shard_of(n_total, G, s.rank, &first, &count);
        if (count == 0) return ISMPC_OK;
        int rc = ismpc_formc_solve_batch(s.h, count, state + first, walk + first, inst + first, nullptr, 0, out + first,
                                         nullptr, nullptr, ISMPC_MEM_HOST_ASYNC, s.stream);
        if (rc != ISMPC_OK) return g->fail(rc, "ismpc_formc_solve_batch", &s);
        if ((rc = ismpc_wait(s.h, s.stream)) != ISMPC_OK) return g->fail(rc, "ismpc_wait", &s);
        return ISMPC_OK;
    });
}

extern "C" int ismpc_group_formc_set_instances(ismpc_group* g, int n_total, const ismpc_formc_inst_t* inst)
{
    if (!g || n_total < 0 || (n_total > 0 && !inst)) return ISMPC_ERR_ARG;
    const int G = (int)g->shards.size();
    if ((n_total + G - 1) / G > g->max_batch) return ISMPC_ERR_ARG;
    g->err.clear();
    return g->run([=](Shard& s) -> int {
        int first, count;
        shard_of(n_total, G, s.rank, &first, &count);
        int rc = ismpc_formc_set_instances(s.h, count > 0 ? inst + first : nullptr, count, ISMPC_MEM_HOST);
        if (rc != ISMPC_OK) return g->fail(rc, "ismpc_formc_set_instances", &s);
        return ISMPC_OK;
    });
}

extern "C" int ismpc_group_formc_solve_batch_packed(ismpc_group* g, int n_total, const ismpc_formc_tick_t* tick,
                                                    ismpc_formc_out_t* out)
{
    if (!g || n_total < 0 || !tick || !out) return ISMPC_ERR_ARG;
    const int G = (int)g->shards.size();
    if ((n_total + G - 1) / G > g->max_batch) return ISMPC_ERR_ARG;
    g->err.clear();
    return g->run([=](Shard& s) -> int {
        int first, count;
        shard_of(n_total, G, s.rank, &first, &count);
        if (count == 0) return ISMPC_OK;
        int rc = ismpc_formc_solve_batch_packed(s.h, count, tick + first, nullptr, nullptr, 0, out + first, nullptr, nullptr,
                                                ISMPC_MEM_HOST_ASYNC, s.stream);
        if (rc != ISMPC_OK) return g->fail(rc, "ismpc_formc_solve_batch_packed", &s);
        if ((rc = ismpc_wait(s.h, s.stream)) != ISMPC_OK) return g->fail(rc, "ismpc_wait", &s);
        return ISMPC_OK;
    });
}

extern "C" int ismpc_group_formc_scatter(ismpc_group* g, int n_total, const ismpc_state_t* state, const ismpc_walk_t* walk,
                                         const ismpc_formc_inst_t* inst, const ismpc_push_t* push)
{
    if (!g || n_total < 0 || !state || !walk || !inst) return ISMPC_ERR_ARG;
    const int G = (int)g->shards.size();
    const int m = (n_total + G - 1) / G;
    if (m > g->max_batch) return ISMPC_ERR_ARG;
    g->err.clear();
    g->n_total = n_total;
    return g->run([=](Shard& s) -> int {
        shard_of(n_total, G, s.rank, &s.first, &s.count);
        const size_t cap = (size_t)(m > 0 ? m : 1);
        if (s.state.ensure(cap * sizeof(ismpc_state_t)) || s.walk.ensure(cap * sizeof(ismpc_walk_t)) ||
            s.inst.ensure(cap * sizeof(ismpc_formc_inst_t)) || s.push.ensure(cap * sizeof(ismpc_push_t)) ||
            s.status.ensure(cap * sizeof(int32_t)) || s.status_tmp.ensure(cap * sizeof(int32_t)))
            return g->fail(ISMPC_ERR_ALLOC, "cudaMalloc (shard buffers)", &s);
        cudaError_t e = cudaMemsetAsync(s.status.p, 0, cap * sizeof(int32_t), s.stream);
        // (records past the shard's count pad the all-gather: zeroed so that nothing uninitialised travels)
        if (e == cudaSuccess) e = cudaMemsetAsync(s.state.p, 0, cap * sizeof(ismpc_state_t), s.stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(s.walk.p, 0, cap * sizeof(ismpc_walk_t), s.stream);
        if (s.count > 0) {
            if (e == cudaSuccess) e = cudaMemcpyAsync(s.state.p, state + s.first, (size_t)s.count * sizeof(ismpc_state_t), cudaMemcpyHostToDevice, s.stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(s.walk.p, walk + s.first, (size_t)s.count * sizeof(ismpc_walk_t), cudaMemcpyHostToDevice, s.stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(s.inst.p, inst + s.first, (size_t)s.count * sizeof(ismpc_formc_inst_t), cudaMemcpyHostToDevice, s.stream);
            if (e == cudaSuccess && push) e = cudaMemcpyAsync(s.push.p, push + s.first, (size_t)s.count * sizeof(ismpc_push_t), cudaMemcpyHostToDevice, s.stream);
        }
        s.pushes_pending = push != nullptr;
        if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);      // the host arrays may be reused after the call
        if (e != cudaSuccess) return g->fail(ISMPC_ERR_CUDA, cudaGetErrorString(e), &s);
        return ISMPC_OK;
    });
}

extern "C" int ismpc_group_formc_rollout(ismpc_group* g, int n_ticks)
{
    if (!g || n_ticks < 0) return ISMPC_ERR_ARG;
    if (n_ticks == 0) return ISMPC_OK;
    g->err.clear();
    return g->run([=](Shard& s) -> int {
        if (s.count == 0) return ISMPC_OK;
        int rc = ismpc_formc_rollout(s.h, s.count, n_ticks, (ismpc_state_t*)s.state.p, (ismpc_walk_t*)s.walk.p,
                                     (const ismpc_formc_inst_t*)s.inst.p, nullptr, 0,
                                     s.pushes_pending ? (const ismpc_push_t*)s.push.p : nullptr, nullptr,
                                     (int32_t*)s.status_tmp.p, ISMPC_MEM_DEVICE, s.stream);
        if (rc != ISMPC_OK) return g->fail(rc, "ismpc_formc_rollout", &s);
        s.pushes_pending = false;
        or_into_kernel<<<(s.count + 255) / 256, 256, 0, s.stream>>>(s.count, (int32_t*)s.status.p, (const int32_t*)s.status_tmp.p);
        if (cudaGetLastError() != cudaSuccess) return g->fail(ISMPC_ERR_CUDA, "or_into_kernel", &s);
        return ISMPC_OK;
    });
}

extern "C" int ismpc_group_wait(ismpc_group* g)
{
    if (!g) return ISMPC_ERR_ARG;
    return g->run([=](Shard& s) -> int {
        const cudaError_t e = cudaStreamSynchronize(s.stream);
        return e == cudaSuccess ? ISMPC_OK : g->fail(ISMPC_ERR_CUDA, cudaGetErrorString(e), &s);
    });
}

extern "C" int ismpc_group_formc_gather(ismpc_group* g, ismpc_state_t* state_out, ismpc_walk_t* walk_out, int32_t* status_out)
{
    if (!g) return ISMPC_ERR_ARG;
    const int G = (int)g->shards.size();
    const int n_total = g->n_total;
    const int m = (n_total + G - 1) / G;
    if (n_total == 0) return ISMPC_OK;
    g->err.clear();
    if (g->gather_mode == ISMPC_GATHER_HOST || G == 1) {
        // every shard's records straight to their place in the host arrays
        return g->run([=](Shard& s) -> int {
            cudaError_t e = cudaSuccess;
            if (s.count > 0) {
                if (state_out) e = cudaMemcpyAsync(state_out + s.first, s.state.p, (size_t)s.count * sizeof(ismpc_state_t), cudaMemcpyDeviceToHost, s.stream);
                if (e == cudaSuccess && walk_out) e = cudaMemcpyAsync(walk_out + s.first, s.walk.p, (size_t)s.count * sizeof(ismpc_walk_t), cudaMemcpyDeviceToHost, s.stream);
                if (e == cudaSuccess && status_out) e = cudaMemcpyAsync(status_out + s.first, s.status.p, (size_t)s.count * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
            return e == cudaSuccess ? ISMPC_OK : g->fail(ISMPC_ERR_CUDA, cudaGetErrorString(e), &s);
        });
    }
    // ---- the one collective of the path: all-gather of the (padded) shard records over NVLink / NVSwitch ----
    int rc = g->run([=](Shard& s) -> int {
        if (s.gs.ensure((size_t)G * m * sizeof(ismpc_state_t)) || s.gw.ensure((size_t)G * m * sizeof(ismpc_walk_t)) ||
            s.gst.ensure((size_t)G * m * sizeof(int32_t)))
            return g->fail(ISMPC_ERR_ALLOC, "cudaMalloc (gather buffers)", &s);
        return ISMPC_OK;
    });
    if (rc != ISMPC_OK) return rc;
    // single-process multi-GPU: the calls of all ranks are issued inside one NCCL group from this thread, each on its
    // device's stream, i.e. behind that device's rollout
    ncclResult_t nr = ncclGroupStart();
    for (int r = 0; r < G && nr == ncclSuccess; ++r) {
        Shard& s = g->shards[(size_t)r];
        nr = ncclAllGather(s.state.p, s.gs.p, (size_t)m * sizeof(ismpc_state_t), ncclChar, s.comm, s.stream);
        if (nr == ncclSuccess) nr = ncclAllGather(s.walk.p, s.gw.p, (size_t)m * sizeof(ismpc_walk_t), ncclChar, s.comm, s.stream);
        if (nr == ncclSuccess) nr = ncclAllGather(s.status.p, s.gst.p, (size_t)m * sizeof(int32_t), ncclChar, s.comm, s.stream);
    }
    const ncclResult_t ne = ncclGroupEnd();
    if (nr != ncclSuccess || ne != ncclSuccess) {
        g->err = std::string("ncclAllGather: ") + ncclGetErrorString(nr != ncclSuccess ? nr : ne);
        return ISMPC_ERR_CUDA;
    }
    // every device now holds all records; the host reads them from the first one, shard by shard (padding skipped)
    Shard& s0 = g->shards[0];
    cudaSetDevice(s0.device);
    cudaError_t e = cudaSuccess;
    for (int r = 0; r < G && e == cudaSuccess; ++r) {
        int first, count;
        shard_of(n_total, G, r, &first, &count);
        if (count == 0) continue;
        if (state_out) e = cudaMemcpyAsync(state_out + first, (const ismpc_state_t*)s0.gs.p + (size_t)r * m, (size_t)count * sizeof(ismpc_state_t), cudaMemcpyDeviceToHost, s0.stream);
        if (e == cudaSuccess && walk_out) e = cudaMemcpyAsync(walk_out + first, (const ismpc_walk_t*)s0.gw.p + (size_t)r * m, (size_t)count * sizeof(ismpc_walk_t), cudaMemcpyDeviceToHost, s0.stream);
        if (e == cudaSuccess && status_out) e = cudaMemcpyAsync(status_out + first, (const int32_t*)s0.gst.p + (size_t)r * m, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost, s0.stream);
    }
    if (e != cudaSuccess) { g->err = std::string("gather copy: ") + cudaGetErrorString(e); return ISMPC_ERR_CUDA; }
    return ismpc_group_wait(g);
}
