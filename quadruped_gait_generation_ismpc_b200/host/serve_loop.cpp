// serve_loop.cpp -- C entry points around ismpc_host::FormCPipeline so that bench.py and the tests can drive the C++
// serving loop in-process (g++ only: everything goes through include/ismpc_b200.h).  -> lib/libismpc_host.so
#include <cstring>
#include <exception>
#include <string>
#include <thread>
#include <vector>

#include "FormCPipeline.hpp"

using ismpc_host::FormCPipeline;

static thread_local std::string g_err;

extern "C" const char* ismpc_host_last_error(void) { return g_err.c_str(); }

extern "C" void* ismpc_host_pipeline_create(int device, int n, int depth, const ismpc_formc_model_t* model, int S, int F_ds,
                                            const double* plan_xyzt, int plan_rows)
{
    try {
        return new FormCPipeline(device, n, depth, *model, S, F_ds, plan_xyzt, plan_rows, /*own_staging=*/false);
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

extern "C" void ismpc_host_pipeline_destroy(void* p) { delete static_cast<FormCPipeline*>(p); }

extern "C" long long ismpc_host_pipeline_launches(void* p) { return static_cast<FormCPipeline*>(p)->kernel_launches(); }

// Steps k0 .. k0+steps-1 of the serving loop.  Step k takes its inputs from the caller's pinned block
// in_blocks[k % n_blocks] (= [state n | walk n | inst n]), waits for the step that used its slot before, reads that
// step's result (status word of record 0, summed into *checksum), and submits.  Ends with every slot waited for; the
// result records of the last `depth` steps are then in out_copy_opt (depth x n records, slot-major) if given.
extern "C" int ismpc_host_pipeline_run(void* pv, int k0, int steps, const void* const* in_blocks, int n_blocks,
                                       long long* checksum, ismpc_formc_out_t* out_copy_opt)
{
    FormCPipeline& p = *static_cast<FormCPipeline*>(pv);
    const size_t n = (size_t)p.n();
    long long sum = 0;
    try {
        for (int k = k0; k < k0 + steps; ++k) {
            const int s = p.acquire();
            if (k - k0 >= p.depth()) sum += p.out(s)[0].status;
            const char* b = static_cast<const char*>(in_blocks[k % n_blocks]);
            p.submit_from(s, reinterpret_cast<const ismpc_state_t*>(b),
                          reinterpret_cast<const ismpc_walk_t*>(b + n * sizeof(ismpc_state_t)),
                          reinterpret_cast<const ismpc_formc_inst_t*>(b + n * (sizeof(ismpc_state_t) + sizeof(ismpc_walk_t))));
        }
        p.wait_all();
        if (out_copy_opt)
            for (int s = 0; s < p.depth(); ++s) std::memcpy(out_copy_opt + (size_t)s * n, p.out(s), n * sizeof(ismpc_formc_out_t));
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
    if (checksum) *checksum = sum;
    return 0;
}

// The same loop on T host threads, one pipeline (its own handles, streams and result buffers) per thread: thread t takes
// steps k0+t, k0+t+T, ...  (handles are not shared between threads; distinct handles are independent, ismpc_b200.h).
// out_copy_opt: T x depth x n records.  Returns 0, or -1 with the first thread's error in ismpc_host_last_error().
extern "C" int ismpc_host_pipelines_run(void* const* pipes, int T, int k0, int steps, const void* const* in_blocks,
                                        int n_blocks, long long* checksum, ismpc_formc_out_t* out_copy_opt)
{
    std::vector<std::thread> th;
    std::vector<long long> sums((size_t)T, 0);
    std::vector<std::string> errs((size_t)T);
    auto work = [&](int t) {
        {
            FormCPipeline& p = *static_cast<FormCPipeline*>(pipes[t]);
            const size_t n = (size_t)p.n();
            try {
                int done = 0;
                for (int k = k0 + t; k < k0 + steps; k += T, ++done) {
                    const int s = p.acquire();
                    if (done >= p.depth()) sums[t] += p.out(s)[0].status;
                    const char* b = static_cast<const char*>(in_blocks[k % n_blocks]);
                    p.submit_from(s, reinterpret_cast<const ismpc_state_t*>(b),
                                  reinterpret_cast<const ismpc_walk_t*>(b + n * sizeof(ismpc_state_t)),
                                  reinterpret_cast<const ismpc_formc_inst_t*>(b + n * (sizeof(ismpc_state_t) + sizeof(ismpc_walk_t))));
                }
                p.wait_all();
                if (out_copy_opt)
                    for (int s = 0; s < p.depth(); ++s)
                        std::memcpy(out_copy_opt + ((size_t)t * p.depth() + s) * n, p.out(s), n * sizeof(ismpc_formc_out_t));
            } catch (const std::exception& e) {
                errs[t] = e.what();
            }
        }
    };
    if (T == 1) work(0);                                   // no thread for a single pipeline
    else {
        for (int t = 0; t < T; ++t) th.emplace_back(work, t);
        for (std::thread& x : th) x.join();
    }
    long long sum = 0;
    for (int t = 0; t < T; ++t) {
        if (!errs[t].empty()) { g_err = errs[t]; return -1; }
        sum += sums[t];
    }
    if (checksum) *checksum = sum;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// The same serving loop with PERSISTENT host threads (what a real caller has: its worker threads exist before the first
// tick).  A pool owns T pipelines; thread 0 is the caller's own thread, threads 1..T-1 are parked between runs -- they
// spin on a generation counter for a short while (a run that follows the previous one closely starts within a
// microsecond) and then sleep on a condition variable.  ismpc_host_pool_run times the K steps itself: the clock starts
// before the first submit and stops when every stream of every pipeline has been waited for and every result read.
// ---------------------------------------------------------------------------------------------------------------------
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>

namespace {

struct HostPool {
    std::vector<std::unique_ptr<FormCPipeline>> pipes;
    std::vector<std::thread> workers;
    std::vector<long long> sums;
    std::vector<std::string> errs;
    std::vector<double> wait_s, submit_s;      // per thread, last run: seconds inside ismpc_wait / inside ismpc_formc_solve_batch
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<long long> generation{0};
    std::atomic<int> done{0};
    bool quit = false;
    int spin_us = 300;               // how long a parked worker polls before it sleeps (ismpc_host_pool_set_spin_us)
    // the job
    int k0 = 0, steps = 0, n_blocks = 0;
    const void* const* in_blocks = nullptr;
    ismpc_formc_out_t* out_copy = nullptr;
    bool packed = false;             // in_blocks are packed tick records: n_blocks per slot, slot-major (slot = t * depth + s)

    void work(int t)
    {
        FormCPipeline& p = *pipes[(size_t)t];
        const int T = (int)pipes.size();
        const size_t n = (size_t)p.n();
        sums[(size_t)t] = 0; errs[(size_t)t].clear();
        try {
            int submitted = 0;
            double t_wait = 0.0, t_submit = 0.0;
            p.rewind();              // every run ends with wait_all: slot j % depth serves this thread's j-th step of the run
            auto now = [] { return std::chrono::steady_clock::now(); };
            for (int k = k0 + t; k < k0 + steps; k += T, ++submitted) {
                const auto a0 = now();
                const int s = p.acquire();
                const auto a1 = now();
                if (submitted >= p.depth()) sums[(size_t)t] += p.out(s)[0].status;      // the result of the step that used the slot
                if (packed) {
                    // every slot serves its own fleet (resident constants and plans); the fleet's tick records rotate
                    // over the n_blocks pinned blocks the caller prepared for it
                    const size_t slot = (size_t)t * (size_t)p.depth() + (size_t)s;
                    const int visit = (k - k0) / (T * p.depth());
                    p.submit_packed_from(s, static_cast<const ismpc_formc_tick_t*>(in_blocks[slot * (size_t)n_blocks + (size_t)(visit % n_blocks)]));
                    const auto a2p = now();
                    t_wait += std::chrono::duration<double>(a1 - a0).count();
                    t_submit += std::chrono::duration<double>(a2p - a1).count();
                    continue;
                }
                const char* b = static_cast<const char*>(in_blocks[k % n_blocks]);
                p.submit_from(s, reinterpret_cast<const ismpc_state_t*>(b),
                              reinterpret_cast<const ismpc_walk_t*>(b + n * sizeof(ismpc_state_t)),
                              reinterpret_cast<const ismpc_formc_inst_t*>(b + n * (sizeof(ismpc_state_t) + sizeof(ismpc_walk_t))));
                const auto a2 = now();
                t_wait += std::chrono::duration<double>(a1 - a0).count();
                t_submit += std::chrono::duration<double>(a2 - a1).count();
            }
            const auto a3 = now();
            p.wait_all();
            t_wait += std::chrono::duration<double>(now() - a3).count();
            wait_s[(size_t)t] = t_wait; submit_s[(size_t)t] = t_submit;
            for (int s = 0; s < p.depth(); ++s) sums[(size_t)t] += p.out(s)[0].status;   // ... and of the last `depth` steps
            if (out_copy)
                for (int s = 0; s < p.depth(); ++s)
                    std::memcpy(out_copy + ((size_t)t * p.depth() + s) * n, p.out(s), n * sizeof(ismpc_formc_out_t));
        } catch (const std::exception& e) {
            errs[(size_t)t] = e.what();
        }
    }

    void worker_main(int t)
    {
        long long seen = 0;
        for (;;) {
            // spin briefly, then sleep
            bool go = false;
            const auto t_spin = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us);
            while (std::chrono::steady_clock::now() < t_spin) {
                if (generation.load(std::memory_order_acquire) != seen) { go = true; break; }
            }
            if (!go) {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return quit || generation.load(std::memory_order_acquire) != seen; });
                if (quit) return;
            }
            if (quit) return;
            seen = generation.load(std::memory_order_acquire);
            work(t);
            done.fetch_add(1, std::memory_order_release);
        }
    }

    ~HostPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
            generation.fetch_add(1, std::memory_order_release);
        }
        cv.notify_all();
        for (std::thread& w : workers) w.join();
    }
};

}  // namespace

extern "C" void* ismpc_host_pool_create(int device, int n, int threads, int depth, const ismpc_formc_model_t* model, int S,
                                        int F_ds, const double* plan_xyzt, int plan_rows)
{
    try {
        if (threads < 1 || threads > 64) throw std::runtime_error("ismpc_host_pool_create: threads must be 1..64");
        std::unique_ptr<HostPool> hp(new HostPool());
        for (int t = 0; t < threads; ++t)
            hp->pipes.emplace_back(new FormCPipeline(device, n, depth, *model, S, F_ds, plan_xyzt, plan_rows, /*own_staging=*/false));
        hp->sums.assign((size_t)threads, 0); hp->errs.assign((size_t)threads, std::string());
        hp->wait_s.assign((size_t)threads, 0.0); hp->submit_s.assign((size_t)threads, 0.0);
        for (int t = 1; t < threads; ++t) hp->workers.emplace_back(&HostPool::worker_main, hp.get(), t);
        return hp.release();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

extern "C" void ismpc_host_pool_destroy(void* p) { delete static_cast<HostPool*>(p); }

// Low-latency serving: parked workers poll for `us` microseconds before they go to sleep (default 300; a control loop
// that ticks every few milliseconds sets this above its period so that a tick never pays a futex wake-up).
extern "C" void ismpc_host_pool_set_spin_us(void* p, int us) { if (p && us >= 0) static_cast<HostPool*>(p)->spin_us = us; }

extern "C" long long ismpc_host_pool_launches(void* pv)
{
    long long t = 0;
    for (auto& p : static_cast<HostPool*>(pv)->pipes) t += p->kernel_launches();
    return t;
}

// Where the host threads spent the last run: seconds blocked in ismpc_wait (the GPU side -- copies and kernels -- was the
// slower party) and seconds inside ismpc_formc_solve_batch (driver calls: the host was), summed over the pool's threads.
extern "C" void ismpc_host_pool_stats(void* pv, double* wait_s, double* submit_s)
{
    HostPool& hp = *static_cast<HostPool*>(pv);
    double w = 0.0, s = 0.0;
    for (size_t t = 0; t < hp.pipes.size(); ++t) { w += hp.wait_s[t]; s += hp.submit_s[t]; }
    if (wait_s) *wait_s = w;
    if (submit_s) *submit_s = s;
}

// Steps k0 .. k0+steps-1: thread t takes steps k0+t, k0+t+T, ...  out_copy_opt: T x depth x n records (the result records
// of every pipeline's slots after the run).  elapsed_s_opt: wall-clock seconds from before the first submit until the
// last result has been read.  Returns 0, or -1 with the first error in ismpc_host_last_error().
static int pool_run(void* pv, int k0, int steps, const void* const* in_blocks, int n_blocks, long long* checksum,
                    ismpc_formc_out_t* out_copy_opt, double* elapsed_s_opt, bool packed);

extern "C" int ismpc_host_pool_run(void* pv, int k0, int steps, const void* const* in_blocks, int n_blocks,
                                   long long* checksum, ismpc_formc_out_t* out_copy_opt, double* elapsed_s_opt)
{
    return pool_run(pv, k0, steps, in_blocks, n_blocks, checksum, out_copy_opt, elapsed_s_opt, false);
}

// Packed mode.  ismpc_host_pool_set_instances gives slot `s` of thread `t` its fleet's constants (n records); the plans
// of all fleets are in the table the pool was created with.  ismpc_host_pool_run_packed: tick_blocks holds
// blocks_per_slot pinned arrays of n ismpc_formc_tick_t for every slot, slot-major (slot = t * depth + s); the j-th tick
// a slot serves takes block j % blocks_per_slot.  Otherwise as ismpc_host_pool_run.
extern "C" int ismpc_host_pool_set_instances(void* pv, int t, int s, const ismpc_formc_inst_t* inst)
{
    HostPool& hp = *static_cast<HostPool*>(pv);
    try {
        if (t < 0 || t >= (int)hp.pipes.size() || s < 0 || s >= hp.pipes[(size_t)t]->depth()) throw std::runtime_error("bad slot");
        hp.pipes[(size_t)t]->set_instances(s, inst);
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
    return 0;
}

extern "C" int ismpc_host_pool_set_option(void* pv, const char* name, int value)
{
    HostPool& hp = *static_cast<HostPool*>(pv);
    try {
        for (auto& p : hp.pipes) p->set_option(name, value);
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
    return 0;
}

extern "C" int ismpc_host_pool_run_packed(void* pv, int k0, int steps, const void* const* tick_blocks, int blocks_per_slot,
                                          long long* checksum, ismpc_formc_out_t* out_copy_opt, double* elapsed_s_opt)
{
    return pool_run(pv, k0, steps, tick_blocks, blocks_per_slot, checksum, out_copy_opt, elapsed_s_opt, true);
}

static int pool_run(void* pv, int k0, int steps, const void* const* in_blocks, int n_blocks, long long* checksum,
                    ismpc_formc_out_t* out_copy_opt, double* elapsed_s_opt, bool packed)
{
    HostPool& hp = *static_cast<HostPool*>(pv);
    const int T = (int)hp.pipes.size();
    hp.packed = packed;
    hp.k0 = k0; hp.steps = steps; hp.in_blocks = in_blocks; hp.n_blocks = n_blocks; hp.out_copy = out_copy_opt;
    hp.done.store(0, std::memory_order_release);
    const auto t0 = std::chrono::steady_clock::now();
    if (T > 1) {
        { std::lock_guard<std::mutex> lk(hp.mu); hp.generation.fetch_add(1, std::memory_order_release); }
        hp.cv.notify_all();
    }
    hp.work(0);
    while (hp.done.load(std::memory_order_acquire) < T - 1) { /* the workers finish within microseconds of thread 0 */ }
    const auto t1 = std::chrono::steady_clock::now();
    if (elapsed_s_opt) *elapsed_s_opt = std::chrono::duration<double>(t1 - t0).count();
    long long sum = 0;
    for (int t = 0; t < T; ++t) {
        if (!hp.errs[(size_t)t].empty()) { g_err = hp.errs[(size_t)t]; return -1; }
        sum += hp.sums[(size_t)t];
    }
    if (checksum) *checksum = sum;
    return 0;
}
