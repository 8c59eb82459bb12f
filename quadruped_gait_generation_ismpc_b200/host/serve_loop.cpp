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
