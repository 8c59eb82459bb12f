// MPCSolverMultiGpu.hpp -- MPCSolverBatch over ALL GPUs of a box (include/ismpc_b200_multigpu.h), plain g++.
//
// The reference steps ONE robot from one thread (AMR_code_DART/Controller.cpp:346-348).  A host that steps many keeps
// the same call -- `solve(robots, walkStates, ftsp_and_timings)` -- and lets the group shard the robots over the GPUs:
// contiguous ranges, one handle + stream + host thread per device, no communication between the devices per tick
// (SURVEY section 8e).  Two ways to use it:
//
//   * tick by tick with host buffers (what Controller::update would do for n robots):
//         MPCSolverMultiGpu<State, WalkState, Eigen::MatrixXd> solver(n, ftsp_and_time, devices);
//         solver.solve(robots, walkStates, ftsp_and_time);          // comPos / comVel advanced in place
//   * closed loop with the state resident on the GPUs: scatter once, advance, gather once -- the gather is the one
//     collective of the path (ncclAllGather of the result records, then one copy to the host):
//         solver.scatter(robots, walkStates, pushes);  solver.rollout(1000);  solver.gather(robots, walkStates, status);
//
// Same templates on the caller's own State / WalkState / matrix types as host/MPCSolver.hpp.
#pragma once

#include <vector>

#include "../../include/ismpc_b200_multigpu.h"
#include "MPCSolver.hpp"

namespace ismpc_host {

template <class T>
class PinnedArray {          // host array in pinned memory (ismpc_host_alloc), so that the group's copies are asynchronous
public:
    PinnedArray() = default;
    explicit PinnedArray(size_t n) { resize(n); }
    ~PinnedArray() { ismpc_host_free(p_); }
    PinnedArray(const PinnedArray&) = delete;
    PinnedArray& operator=(const PinnedArray&) = delete;
    void resize(size_t n)
    {
        ismpc_host_free(p_); p_ = nullptr; n_ = 0;
        if (n == 0) return;
        p_ = static_cast<T*>(ismpc_host_alloc(n * sizeof(T)));
        if (!p_) throw std::runtime_error("ismpc_host_alloc failed");
        n_ = n;
        std::memset(p_, 0, n * sizeof(T));
    }
    T* data() { return p_; }
    const T* data() const { return p_; }
    T& operator[](size_t i) { return p_[i]; }
    const T& operator[](size_t i) const { return p_[i]; }
    size_t size() const { return n_; }
private:
    T* p_ = nullptr; size_t n_ = 0;
};

template <class StateT, class WalkStateT, class MatrixT>
class MPCSolverMultiGpu {
public:
    // devices: CUDA ordinals, one shard each (e.g. {0,1,...,7}); gather_mode: ISMPC_GATHER_NCCL or ISMPC_GATHER_HOST
    MPCSolverMultiGpu(int n_robots, const MatrixT& ftsp_and_timings, const std::vector<int>& devices,
                      const Parameters& p = Parameters(), int gather_mode = ISMPC_GATHER_NCCL)
        : n_(n_robots), par_(p)
    {
        if (n_robots <= 0 || devices.empty()) throw std::runtime_error("MPCSolverMultiGpu: need robots and devices");
        const int G = (int)devices.size();
        int rc = ismpc_group_create(&g_, devices.data(), G, (n_robots + G - 1) / G, gather_mode);
        if (rc != ISMPC_OK) throw std::runtime_error(std::string("ismpc_group_create: ") + ismpc_error_string(rc));
        try {
            configure(ftsp_and_timings);
            st_.resize((size_t)n_); wk_.resize((size_t)n_); inst_.resize((size_t)n_); out_.resize((size_t)n_); tick_.resize((size_t)n_);
            for (int i = 0; i < n_; ++i) {
                inst_[i].com_height = p.comTargetHeight; inst_[i].box_w = p.footConstraintSquareWidth;
                inst_[i].box_w_init = 2.0; inst_[i].S = p.S(); inst_[i].F_ds = p.F();
                inst_[i].plan_first_row = 0; inst_[i].n_steps = plan_rows_;
            }
        } catch (...) {
            ismpc_group_destroy(g_); g_ = nullptr;
            throw;
        }
    }
    ~MPCSolverMultiGpu() { if (g_) ismpc_group_destroy(g_); }
    MPCSolverMultiGpu(const MPCSolverMultiGpu&) = delete;
    MPCSolverMultiGpu& operator=(const MPCSolverMultiGpu&) = delete;

    int devices() const { return ismpc_group_size(g_); }
    // robots [first, first + count) live on device `rank`
    void shard(int rank, int& first, int& count) const { check(ismpc_group_shard(g_, n_, rank, &first, &count), "ismpc_group_shard"); }

    // One tick for all robots on all devices, host buffers: comPos / comVel advanced in place, every other member stays.
    void solve(std::vector<StateT>& robots, const std::vector<WalkStateT>& walk, const MatrixT& ftsp_and_timings)
    {
        pack(robots, walk, ftsp_and_timings);
        // per-instance constants resident per shard, the tick moves one packed 128-byte record per robot in and one out
        // (pinned arrays: read / written in place by each device's kernel, one launch per device)
        if (inst_dirty_) { check(ismpc_group_formc_set_instances(g_, n_, inst_.data()), "ismpc_group_formc_set_instances"); inst_dirty_ = false; }
        for (int i = 0; i < n_; ++i) { tick_[i].state = st_[i]; tick_[i].walk = wk_[i]; }
        check(ismpc_group_formc_solve_batch_packed(g_, n_, tick_.data(), out_.data()), "ismpc_group_formc_solve_batch_packed");
        for (int i = 0; i < n_; ++i)
            for (int c = 0; c < 3; ++c) { robots[i].comPos(c) = out_[i].next.com_pos[c]; robots[i].comVel(c) = out_[i].next.com_vel[c]; }
    }
    const ismpc_formc_out_t& result(int i) const { return out_[(size_t)i]; }

    // Closed loop, state resident per GPU.
    void scatter(const std::vector<StateT>& robots, const std::vector<WalkStateT>& walk, const std::vector<ismpc_push_t>* pushes = nullptr)
    {
        pack(robots, walk, plan_matrix_dummy_);
        check(ismpc_group_formc_scatter(g_, n_, st_.data(), wk_.data(), inst_.data(), pushes ? pushes->data() : nullptr), "ismpc_group_formc_scatter");
    }
    void rollout(int n_ticks) { check(ismpc_group_formc_rollout(g_, n_ticks), "ismpc_group_formc_rollout"); }
    void wait() { check(ismpc_group_wait(g_), "ismpc_group_wait"); }
    // the one collective: all-gather over the devices, then to the host
    void gather(std::vector<StateT>& robots, std::vector<WalkStateT>& walk, std::vector<int32_t>& status)
    {
        status_.resize((size_t)n_);
        check(ismpc_group_formc_gather(g_, st_.data(), wk_.data(), status_.data()), "ismpc_group_formc_gather");
        status.assign(status_.data(), status_.data() + n_);
        for (int i = 0; i < n_; ++i) {
            for (int c = 0; c < 3; ++c) { robots[i].comPos(c) = st_[i].com_pos[c]; robots[i].comVel(c) = st_[i].com_vel[c]; }
            walk[i].simulationTime = wk_[i].sim_time; walk[i].mpcIter = wk_[i].mpc_iter; walk[i].controlIter = wk_[i].control_iter;
            walk[i].footstepCounter = wk_[i].footstep_counter; walk[i].supportFoot = wk_[i].support_foot != 0;
        }
    }
    long long kernel_launches() const { return ismpc_group_kernel_launches(g_); }

private:
    void check(int rc, const char* what) const
    {
        if (rc != ISMPC_OK) throw std::runtime_error(std::string(what) + ": " + ismpc_error_string(rc) + " [" + ismpc_group_last_error(g_) + "]");
    }
    void configure(const MatrixT& f)
    {
        plan_rows_of(f, plan_);
        plan_rows_ = (int)f.rows();
        const ismpc_formc_model_t m = model_of(par_);
        check(ismpc_group_formc_configure(g_, &m, par_.S(), par_.F(), plan_.data(), plan_rows_), "ismpc_group_formc_configure");
    }
    template <class V>
    void pack(const V& robots, const std::vector<WalkStateT>& walk, const MatrixT& f)
    {
        if ((int)robots.size() != n_ || (int)walk.size() != n_) throw std::runtime_error("MPCSolverMultiGpu: batch size mismatch");
        if (&f != &plan_matrix_dummy_) {
            bool changed = (int)f.rows() != plan_rows_ || f.cols() < 4;
            for (int i = 0; i < plan_rows_ && !changed; ++i)
                for (int c = 0; c < 4; ++c) if (plan_[(size_t)i * 4 + c] != f(i, c)) { changed = true; break; }
            if (changed) { configure(f); for (int i = 0; i < n_; ++i) inst_[i].n_steps = plan_rows_; inst_dirty_ = true; }
        }
        for (int i = 0; i < n_; ++i) {
            for (int c = 0; c < 3; ++c) {
                st_[i].com_pos[c] = robots[i].comPos(c); st_[i].com_vel[c] = robots[i].comVel(c); st_[i].zmp_pos[c] = robots[i].zmpPos(c);
            }
            wk_[i].sim_time = walk[i].simulationTime; wk_[i].mpc_iter = walk[i].mpcIter; wk_[i].control_iter = walk[i].controlIter;
            wk_[i].footstep_counter = walk[i].footstepCounter; wk_[i].support_foot = walk[i].supportFoot ? 1 : 0;
        }
    }
    int n_;
    Parameters par_;
    ismpc_group* g_ = nullptr;
    int plan_rows_ = 0;
    std::vector<double> plan_;
    MatrixT plan_matrix_dummy_;
    PinnedArray<ismpc_state_t> st_;
    PinnedArray<ismpc_walk_t> wk_;
    PinnedArray<ismpc_formc_inst_t> inst_;
    PinnedArray<ismpc_formc_out_t> out_;
    PinnedArray<ismpc_formc_tick_t> tick_;
    bool inst_dirty_ = true;
    PinnedArray<int32_t> status_;
};

}  // namespace ismpc_host
