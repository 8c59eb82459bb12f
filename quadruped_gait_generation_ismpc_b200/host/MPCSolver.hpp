// MPCSolver.hpp -- C++ host-side mirror of the reference's gait-MPC class over the C ABI.
//
// Same names, argument meaning and call sequence as AMR_code_DART/MPCSolver.hpp:16-28:
//
//     MPCSolver(const Eigen::MatrixXd& ftsp_and_timings);
//     State solve(State current, WalkState walkState, const Eigen::MatrixXd& ftsp_and_timings);
//     int itr, fsCount, old_fsCount, adaptation_memo, ds_samples, ct;  double xz_dot, yz_dot;
//
// The classes are templates on the caller's OWN state types, so that AMR_code_DART/Controller.cpp:105-106
// (`solver = new MPCSolver(ftsp_and_time_ref)`) and :346-348 (`desired = solver->solve(desired, walkState,
// ftsp_and_time_ref)`) compile against them with the reference's `State` (21 Eigen::Vector3d members and the getRel*
// methods, AMR_code_DART/types.hpp:7-74) and `WalkState` (:76-81) unchanged:
//
//     #include "types.hpp"                                             // the reference's
//     using MPCSolver = ismpc_host::BasicMPCSolver<State, WalkState, Eigen::MatrixXd>;
//
// (host/dropin/MPCSolver.hpp is exactly that, as a replacement for AMR_code_DART/MPCSolver.hpp.)  What the templates
// need from the types: comPos / comVel / zmpPos members with operator()(int), the six WalkState fields, rows() / cols()
// / operator()(i, j) on the matrix.  solve() returns `current` with comPos and comVel replaced and EVERY other member
// carried through, as the reference does (`State next = current;` MPCSolver.cpp:210, `return next;` :500).
// Hosts without the reference's headers use the small stand-ins at the end of this file
// (ismpc_host::State / WalkState / MatrixXd; ismpc_host::MPCSolver is the template on those).
//
// Error behaviour: the reference drops solver return codes (utils.cpp:128) and exit(1)s on missing data files
// (MPCSolver.cpp:9-12).  These classes never exit: construction failures throw std::runtime_error (nothing leaks: the
// handle is held by an RAII guard), per-tick problems are left in `status` (ISMPC_ST_* bits) and `solve` returns the
// state the kernel produced (unchanged on ISMPC_ST_WINDOW), as the reference would return `next = current` on a tick
// it skips (MPCSolver.cpp:210,214).
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ismpc_b200.h"

namespace ismpc_host {

// AMR_code_DART/parameters.cpp:9-45 as runtime values (defaults = the reference's constants)
struct Parameters {
    double mpcTimeStep = 0.01, controlTimeStep = 0.01;
    double singleSupportDuration = 0.35, doubleSupportDuration = 0.1, predictionTime = 1.0;
    double comTargetHeight = 0.69, footConstraintSquareWidth = 0.09;
    double mass_hrp4 = 50.0, g = 9.81;
    double q_p = 1005000.0, q_v = 100.0, q_u = 0.01, fz_max = 10000.0;   // MPCSolver.cpp:159,253-255
    int N() const { return (int)(predictionTime / mpcTimeStep + 0.5); }
    int S() const { return (int)(singleSupportDuration / mpcTimeStep + 0.5); }
    int F() const { return (int)(doubleSupportDuration / mpcTimeStep + 0.5); }
};

// Owns an ismpc_handle: destroyed on every path out of a constructor that throws.
class HandleGuard {
public:
    HandleGuard() = default;
    explicit HandleGuard(ismpc_handle* h) : h_(h) {}
    ~HandleGuard() { reset(); }
    HandleGuard(const HandleGuard&) = delete;
    HandleGuard& operator=(const HandleGuard&) = delete;
    HandleGuard(HandleGuard&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    HandleGuard& operator=(HandleGuard&& o) noexcept { if (this != &o) { reset(); h_ = o.h_; o.h_ = nullptr; } return *this; }
    void reset(ismpc_handle* h = nullptr) { if (h_) ismpc_destroy(h_); h_ = h; }
    ismpc_handle* get() const { return h_; }
private:
    ismpc_handle* h_ = nullptr;
};

inline void check(int rc, ismpc_handle* h, const char* what)
{
    if (rc != ISMPC_OK)
        throw std::runtime_error(std::string(what) + ": " + ismpc_error_string(rc) + " [" + (h ? ismpc_last_cuda_error(h) : "") + "]");
}

inline ismpc_formc_model_t model_of(const Parameters& p)
{
    ismpc_formc_model_t m{};
    m.dt = p.mpcTimeStep; m.dtc = p.controlTimeStep; m.mass = p.mass_hrp4; m.g = p.g;
    m.q_p = p.q_p; m.q_v = p.q_v; m.q_u = p.q_u; m.fz_max = p.fz_max; m.N = p.N();
    return m;
}

// ftsp_and_timings (Controller.cpp:89-97: rows x (x, y, z, t), Eigen column-major or anything with operator()(i, j))
// -> the row-major plan table of the C ABI.
template <class MatrixT>
inline void plan_rows_of(const MatrixT& f, std::vector<double>& plan)
{
    if (f.cols() < 4) throw std::runtime_error("MPCSolver: ftsp_and_timings needs 4 columns (x, y, z, t)");
    const int rows = (int)f.rows();
    plan.resize((size_t)rows * 4);
    for (int i = 0; i < rows; ++i)
        for (int c = 0; c < 4; ++c) plan[(size_t)i * 4 + c] = f(i, c);
}

template <class StateT, class WalkStateT, class MatrixT>
class BasicMPCSolverBatch {
public:
    BasicMPCSolverBatch(int n_robots, const MatrixT& ftsp_and_timings, const Parameters& p = Parameters(), int device = 0)
        : n_(n_robots), par_(p)
    {
        if (n_robots <= 0) throw std::runtime_error("MPCSolver: n_robots must be positive");
        ismpc_handle* h = nullptr;
        check(ismpc_create(&h, device, n_robots), nullptr, "ismpc_create");
        h_.reset(h);
        const ismpc_formc_model_t m = model_of(p);
        check(ismpc_formc_set_model(h, &m), h, "ismpc_formc_set_model");      // the constructor's matrix work, on the device
        set_plan(ftsp_and_timings);
        inst_.resize(n_); st_.resize(n_); wk_.resize(n_); out_.resize(n_);
        for (int i = 0; i < n_; ++i) {
            inst_[i].com_height = p.comTargetHeight; inst_[i].box_w = p.footConstraintSquareWidth;
            inst_[i].box_w_init = 2.0; inst_[i].S = p.S(); inst_[i].F_ds = p.F();
            inst_[i].plan_first_row = 0; inst_[i].n_steps = plan_rows_;
        }
    }
    BasicMPCSolverBatch(const BasicMPCSolverBatch&) = delete;
    BasicMPCSolverBatch& operator=(const BasicMPCSolverBatch&) = delete;

    // One tick for all robots, in place: comPos / comVel of every robot are advanced, every other member stays.
    // The reference passes the plan with every call (Controller.cpp:346-348); it is uploaded again only when its
    // CONTENTS differ from the table the handle holds (a 40 x 4 compare per tick).
    void solve(std::vector<StateT>& robots, const std::vector<WalkStateT>& walk, const MatrixT& ftsp_and_timings)
    {
        if ((int)robots.size() != n_ || (int)walk.size() != n_) throw std::runtime_error("MPCSolver: batch size mismatch");
        if (plan_changed(ftsp_and_timings)) set_plan(ftsp_and_timings);
        for (int i = 0; i < n_; ++i) {
            for (int c = 0; c < 3; ++c) {
                st_[i].com_pos[c] = robots[i].comPos(c); st_[i].com_vel[c] = robots[i].comVel(c);
                st_[i].zmp_pos[c] = robots[i].zmpPos(c);
            }
            wk_[i].sim_time = walk[i].simulationTime; wk_[i].mpc_iter = walk[i].mpcIter;
            wk_[i].control_iter = walk[i].controlIter; wk_[i].footstep_counter = walk[i].footstepCounter;
            wk_[i].support_foot = walk[i].supportFoot ? 1 : 0;
            inst_[i].n_steps = plan_rows_;
        }
        check(ismpc_formc_solve_batch(h_.get(), n_, st_.data(), wk_.data(), inst_.data(), /*plan: resident*/ nullptr, 0,
                                      out_.data(), nullptr, nullptr, ISMPC_MEM_HOST, nullptr),
              h_.get(), "ismpc_formc_solve_batch");
        for (int i = 0; i < n_; ++i)
            for (int c = 0; c < 3; ++c) {
                robots[i].comPos(c) = out_[i].next.com_pos[c]; robots[i].comVel(c) = out_[i].next.com_vel[c];
            }
    }
    const ismpc_formc_out_t& result(int i) const { return out_[i]; }
    Parameters& parameters() { return par_; }
    int plan_uploads() const { return plan_uploads_; }

private:
    bool plan_changed(const MatrixT& f) const
    {
        if ((int)f.rows() != plan_rows_ || f.cols() < 4) return true;
        for (int i = 0; i < plan_rows_; ++i)
            for (int c = 0; c < 4; ++c)
                if (plan_[(size_t)i * 4 + c] != f(i, c)) return true;
        return false;
    }
    void set_plan(const MatrixT& f)
    {
        plan_rows_of(f, plan_);
        plan_rows_ = (int)f.rows();
        // like the reference's constructor argument, the plan lives with the solver: one upload, then every tick
        // moves only the state / walk-state records (ismpc_formc_set_plan)
        check(ismpc_formc_set_plan(h_.get(), plan_.data(), plan_rows_, ISMPC_MEM_HOST), h_.get(), "ismpc_formc_set_plan");
        ++plan_uploads_;
    }
    int n_;
    Parameters par_;
    HandleGuard h_;
    int plan_rows_ = 0, plan_uploads_ = 0;
    std::vector<double> plan_;
    std::vector<ismpc_formc_inst_t> inst_;
    std::vector<ismpc_state_t> st_;
    std::vector<ismpc_walk_t> wk_;
    std::vector<ismpc_formc_out_t> out_;
};

// Drop-in for the reference class (AMR_code_DART/MPCSolver.hpp:16-28).
template <class StateT, class WalkStateT, class MatrixT>
class BasicMPCSolver {
public:
    explicit BasicMPCSolver(const MatrixT& ftsp_and_timings) : batch_(1, ftsp_and_timings), robots_(1), walk_(1) {}
    BasicMPCSolver(const MatrixT& ftsp_and_timings, const Parameters& p, int device = 0)
        : batch_(1, ftsp_and_timings, p, device), robots_(1), walk_(1) {}
    ~BasicMPCSolver() = default;

    // Compute the next desired state starting from the current state (MPCSolver.cpp:204-501)
    StateT solve(StateT current, WalkStateT walkState, const MatrixT& ftsp_and_timings)
    {
        itr = walkState.mpcIter;                 // MPCSolver.cpp:206
        fsCount = walkState.footstepCounter;     // MPCSolver.cpp:207
        robots_[0] = current;                    // State next = current;  (MPCSolver.cpp:210) -- all members carried
        walk_[0] = walkState;
        batch_.solve(robots_, walk_, ftsp_and_timings);
        status = batch_.result(0).status;
        zmp_x_input = batch_.result(0).zmp_in[0];
        zmp_y_input = batch_.result(0).zmp_in[1];
        return robots_[0];                       // return next;  (MPCSolver.cpp:500)
    }

    // some stuff (public members of the reference class, MPCSolver.hpp:24-28)
    int itr = 0;
    int fsCount = 0, old_fsCount = 0, adaptation_memo = 0, ds_samples = 0, ct = 0;
    double xz_dot = 0.0, yz_dot = 0.0;
    // additions: what the reference only prints (MPCSolver.cpp:402-403,425) or drops (utils.cpp:128)
    double zmp_x_input = 0.0, zmp_y_input = 0.0;
    int status = 0;
    int plan_uploads() const { return batch_.plan_uploads(); }

private:
    BasicMPCSolverBatch<StateT, WalkStateT, MatrixT> batch_;
    std::vector<StateT> robots_;
    std::vector<WalkStateT> walk_;
};

// ---------------------------------------------------------------------------------------------------------------------
// Stand-ins for hosts that have neither Eigen nor the reference's types.hpp (tests/cpp/shim_driver.cpp, plain C++
// callers): the members solve() and its caller touch, same names.
// ---------------------------------------------------------------------------------------------------------------------
struct Vector3d {
    double v[3] = {0, 0, 0};
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
};
struct MatrixXd {   // column-major like Eigen's default
    int r = 0, c = 0;
    std::vector<double> d;
    MatrixXd() = default;
    MatrixXd(int rows_, int cols_) : r(rows_), c(cols_), d((size_t)rows_ * cols_, 0.0) {}
    static MatrixXd Zero(int rows_, int cols_) { return MatrixXd(rows_, cols_); }
    double& operator()(int i, int j) { return d[(size_t)j * r + i]; }
    double operator()(int i, int j) const { return d[(size_t)j * r + i]; }
    int rows() const { return r; }
    int cols() const { return c; }
};
struct State {          // subset of AMR_code_DART/types.hpp:7-29
    Vector3d comPos, comVel, comAcc, zmpPos;
    Vector3d leftBackFootPos, rightBackFootPos, leftFrontFootPos, rightFrontFootPos;
    Vector3d torsoOrient;
};
struct WalkState {      // AMR_code_DART/types.hpp:76-81
    bool supportFoot = true;
    double simulationTime = 0.0;
    int mpcIter = 0, controlIter = 0, footstepCounter = 0, indInitial = 0;
};
using MPCSolverBatch = BasicMPCSolverBatch<State, WalkState, MatrixXd>;
using MPCSolver = BasicMPCSolver<State, WalkState, MatrixXd>;

}  // namespace ismpc_host
