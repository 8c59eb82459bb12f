// FormCPipeline.hpp -- C++ serving loop of the batched MPCSolver::solve over the C ABI (no CUDA headers needed).
//
// The reference calls `desired = solver->solve(desired, walkState, ftsp)` once per 10 ms tick from one thread
// (AMR_code_DART/Controller.cpp:346-348).  A host that steps many robots per tick keeps several batch ticks in flight:
// DEPTH handles, each with its own stream and pinned staging, used round-robin; tick k's copy-in overlaps tick k-1's
// kernel and tick k-2's copy back.  The footstep plans are constructor data (MPCSolver::MPCSolver(ftsp_and_timings),
// MPCSolver.cpp:5) and live in the handles; a tick moves the per-tick records only.
//
//     FormCPipeline p(device, n, depth, model, S, F_ds, plan, plan_rows);
//     for (;;) { int s = p.acquire();            // waits for the tick that used this slot, its records are in p.out(s)
//                consume(p.out(s)); fill(p.state(s), p.walk(s), p.inst(s));
//                p.submit(s); }
//
// `submit_from` takes the caller's own pinned [state | walk | inst] block instead of the slot's staging.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ismpc_b200.h"

namespace ismpc_host {

class FormCPipeline {
public:
    FormCPipeline(int device, int n, int depth, const ismpc_formc_model_t& model, int S, int F_ds,
                  const double* plan_xyzt, int plan_rows, bool own_staging = true)
        : n_(n), slots_((size_t)depth)
    {
        if (n <= 0 || depth <= 0) throw std::runtime_error("FormCPipeline: bad sizes");
        for (Slot& s : slots_) {
            check(ismpc_create(&s.h, device, n), nullptr, "ismpc_create");
            check(ismpc_formc_set_model(s.h, &model), s.h, "ismpc_formc_set_model");
            if (S + F_ds > 0) check(ismpc_formc_prepare_gait(s.h, S, F_ds), s.h, "ismpc_formc_prepare_gait");
            check(ismpc_formc_set_plan(s.h, plan_xyzt, plan_rows, ISMPC_MEM_HOST), s.h, "ismpc_formc_set_plan");
            s.stream = ismpc_handle_stream(s.h);
            if (!s.stream) throw std::runtime_error("FormCPipeline: ismpc_handle_stream failed");
            const size_t in_bytes = (size_t)n * (sizeof(ismpc_state_t) + sizeof(ismpc_walk_t) + sizeof(ismpc_formc_inst_t));
            if (own_staging) {
                s.in = static_cast<char*>(ismpc_host_alloc(in_bytes));
                if (!s.in) throw std::runtime_error("FormCPipeline: ismpc_host_alloc failed");
            }
            s.out = static_cast<ismpc_formc_out_t*>(ismpc_host_alloc((size_t)n * sizeof(ismpc_formc_out_t)));
            if (!s.out) throw std::runtime_error("FormCPipeline: ismpc_host_alloc failed");
        }
    }
    ~FormCPipeline() = default;      // every Slot releases what it owns (also when the constructor throws half-way)
    FormCPipeline(const FormCPipeline&) = delete;
    FormCPipeline& operator=(const FormCPipeline&) = delete;

    int depth() const { return (int)slots_.size(); }
    int n() const { return n_; }
    // the slot's staging: three arrays back to back, which the library moves in one copy
    ismpc_state_t* state(int s) { return reinterpret_cast<ismpc_state_t*>(slots_[s].in); }
    ismpc_walk_t* walk(int s) { return reinterpret_cast<ismpc_walk_t*>(slots_[s].in + (size_t)n_ * sizeof(ismpc_state_t)); }
    ismpc_formc_inst_t* inst(int s)
    {
        return reinterpret_cast<ismpc_formc_inst_t*>(slots_[s].in + (size_t)n_ * (sizeof(ismpc_state_t) + sizeof(ismpc_walk_t)));
    }
    const ismpc_formc_out_t* out(int s) const { return slots_[s].out; }

    // next slot in round-robin order, after the tick that used it last has landed in out(slot)
    int acquire()
    {
        const int s = next_;
        next_ = (next_ + 1) % depth();
        wait(s);
        return s;
    }
    // start the round robin over at slot 0 (call with no tick in flight, e.g. after wait_all)
    void rewind() { next_ = 0; }
    void wait(int s) { check(ismpc_wait(slots_[s].h, slots_[s].stream), slots_[s].h, "ismpc_wait"); }
    void wait_all() { for (int s = 0; s < depth(); ++s) wait(s); }
    void submit(int s) { submit_from(s, state(s), walk(s), inst(s)); }
    void submit_from(int s, const ismpc_state_t* st, const ismpc_walk_t* wk, const ismpc_formc_inst_t* in)
    {
        Slot& q = slots_[s];
        check(ismpc_formc_solve_batch(q.h, n_, st, wk, in, nullptr, 0, q.out, nullptr, nullptr, ISMPC_MEM_HOST_ASYNC, q.stream),
              q.h, "ismpc_formc_solve_batch");
    }
    // tuning knob on every handle of the pipeline (ismpc_set_option), e.g. "formc_variant" = 16: the throughput build of
    // the tick kernel, the right one when 8 or more ticks are in flight on a GPU
    void set_option(const char* name, int value)
    {
        for (Slot& q : slots_) check(ismpc_set_option(q.h, name, value), q.h, "ismpc_set_option");
    }
    // ---- packed mode: the per-instance constants are handed to the slot's handle once (they are globals of the
    // reference, parameters.cpp:9-45), a tick then moves one 128-byte record per instance in and one out; with pinned
    // buffers (the slot's own staging is) the library lets the kernel read and write them in place, and a tick is a
    // single kernel launch -- no copy engine involved (ismpc_b200.h: ismpc_formc_solve_batch_packed) ----
    void set_instances(int s, const ismpc_formc_inst_t* inst)
    {
        check(ismpc_formc_set_instances(slots_[s].h, inst, n_, ISMPC_MEM_HOST), slots_[s].h, "ismpc_formc_set_instances");
    }
    ismpc_formc_tick_t* tick(int s)          // the slot's pinned staging for packed records (allocated on first use)
    {
        Slot& q = slots_[s];
        if (!q.tick) {
            q.tick = static_cast<ismpc_formc_tick_t*>(ismpc_host_alloc((size_t)n_ * sizeof(ismpc_formc_tick_t)));
            if (!q.tick) throw std::runtime_error("FormCPipeline: ismpc_host_alloc failed");
        }
        return q.tick;
    }
    void submit_packed(int s) { submit_packed_from(s, tick(s)); }
    void submit_packed_from(int s, const ismpc_formc_tick_t* tk)
    {
        Slot& q = slots_[s];
        check(ismpc_formc_solve_batch_packed(q.h, n_, tk, nullptr, nullptr, 0, q.out, nullptr, nullptr, ISMPC_MEM_HOST_ASYNC, q.stream),
              q.h, "ismpc_formc_solve_batch_packed");
    }
    long long kernel_launches() const
    {
        long long t = 0;
        for (const Slot& s : slots_) t += ismpc_kernel_launches(s.h);
        return t;
    }

private:
    struct Slot {
        ismpc_handle* h = nullptr; void* stream = nullptr; char* in = nullptr; ismpc_formc_out_t* out = nullptr;
        ismpc_formc_tick_t* tick = nullptr;
        Slot() = default;
        Slot(const Slot&) = delete;
        Slot& operator=(const Slot&) = delete;
        ~Slot()
        {
            if (h) { if (stream) ismpc_wait(h, stream); ismpc_destroy(h); }
            ismpc_host_free(in); ismpc_host_free(out); ismpc_host_free(tick);
        }
    };
    static void check(int rc, ismpc_handle* h, const char* what)
    {
        if (rc != ISMPC_OK)
            throw std::runtime_error(std::string(what) + ": " + ismpc_error_string(rc) + " [" + (h ? ismpc_last_cuda_error(h) : "") + "]");
    }
    int n_, next_ = 0;
    std::vector<Slot> slots_;
};

}  // namespace ismpc_host
