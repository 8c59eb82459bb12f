// MPCSolver.hpp -- C++ host-side mirror of the reference's gait-MPC class over the C ABI.
//
// Same names, argument meaning and call sequence as AMR_code_DART/MPCSolver.hpp:16-28:
//
//     MPCSolver(const Eigen::MatrixXd& ftsp_and_timings);
//     State solve(State current, WalkState walkState, const Eigen::MatrixXd& ftsp_and_timings);
//     int itr, fsCount, old_fsCount, adaptation_memo, ds_samples, ct;  double xz_dot, yz_dot;
//
// so that AMR_code_DART/Controller.cpp:105-106 (`solver = new MPCSolver(ftsp_and_time_ref)`) and
// :346-348 (`desired = solver->solve(desired, walkState, ftsp_and_time_ref)`) compile against it
// unchanged.  It is a batch-of-1 client of include/ismpc_b200.h; `MPCSolverBatch` is the same thing
// for n independent robots.  Where Eigen is installed the Eigen types are used; otherwise the tiny
// stand-ins below (enough for the two call sites) keep the header self-contained.
//
// Error behaviour: the reference drops solver return codes (utils.cpp:128) and exit(1)s on missing
// data files (MPCSolver.cpp:9-12).  This class never exits: construction failures throw
// std::runtime_error, per-tick problems are left in `status` (ISMPC_ST_* bits) and `solve` returns
// the state the kernel produced (unchanged on ISMPC_ST_WINDOW), exactly as the reference would return
// `next = current` on a tick it skips (MPCSolver.cpp:210,214).
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ismpc_b200.h"

#if defined(__has_include)
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#define ISMPC_HAVE_EIGEN 1
#endif
#endif

namespace ismpc_host {

#ifdef ISMPC_HAVE_EIGEN
using Vector3d = Eigen::Vector3d;
using MatrixXd = Eigen::MatrixXd;
#else
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
#endif

// AMR_code_DART/types.hpp:7-29 (the members solve() and its caller touch; same names)
struct State {
    Vector3d comPos, comVel, comAcc, zmpPos;
    Vector3d leftBackFootPos, rightBackFootPos, leftFrontFootPos, rightFrontFootPos;
    Vector3d torsoOrient;
};

// AMR_code_DART/types.hpp:76-81
struct WalkState {
    bool supportFoot = true;
    double simulationTime = 0.0;
    int mpcIter = 0, controlIter = 0, footstepCounter = 0, indInitial = 0;
};

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

class MPCSolverBatch {
public:
    MPCSolverBatch(int n_robots, const MatrixXd& ftsp_and_timings, const Parameters& p = Parameters(), int device = 0)
        : n_(n_robots), par_(p)
    {
        if (n_robots <= 0) throw std::runtime_error("MPCSolver: n_robots must be positive");
        int rc = ismpc_create(&h_, device, n_robots);
        if (rc != ISMPC_OK) throw std::runtime_error(std::string("ismpc_create: ") + ismpc_error_string(rc));
        ismpc_formc_model_t m{};
        m.dt = p.mpcTimeStep; m.dtc = p.controlTimeStep; m.mass = p.mass_hrp4; m.g = p.g;
        m.q_p = p.q_p; m.q_v = p.q_v; m.q_u = p.q_u; m.fz_max = p.fz_max; m.N = p.N();
        rc = ismpc_formc_set_model(h_, &m);      // the constructor's matrix work, on the device
        if (rc != ISMPC_OK) {
            ismpc_destroy(h_);
            throw std::runtime_error(std::string("ismpc_formc_set_model: ") + ismpc_error_string(rc));
        }
        set_plan(ftsp_and_timings);
        inst_.resize(n_); st_.resize(n_); wk_.resize(n_); out_.resize(n_);
        for (int i = 0; i < n_; ++i) {
            inst_[i].com_height = p.comTargetHeight; inst_[i].box_w = p.footConstraintSquareWidth;
            inst_[i].box_w_init = 2.0; inst_[i].S = p.S(); inst_[i].F_ds = p.F();
            inst_[i].plan_first_row = 0; inst_[i].n_steps = plan_rows_;
        }
    }
    ~MPCSolverBatch() { if (h_) ismpc_destroy(h_); }
    MPCSolverBatch(const MPCSolverBatch&) = delete;
    MPCSolverBatch& operator=(const MPCSolverBatch&) = delete;

    // One tick for all robots.  The plan argument is accepted for signature parity; like the reference
    // (which only reads it for the unused footstepPredicted, MPCSolver.cpp:440-441) the midpoint sequence
    // is the one fixed at construction unless the matrix changed shape.
    void solve(std::vector<State>& robots, const std::vector<WalkState>& walk, const MatrixXd& ftsp_and_timings)
    {
        if ((int)robots.size() != n_ || (int)walk.size() != n_) throw std::runtime_error("MPCSolver: batch size mismatch");
        if (ftsp_and_timings.rows() != plan_rows_) set_plan(ftsp_and_timings);
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
        int rc = ismpc_formc_solve_batch(h_, n_, st_.data(), wk_.data(), inst_.data(), /*plan: resident*/ nullptr, 0,
                                         out_.data(), nullptr, nullptr, ISMPC_MEM_HOST, nullptr);
        if (rc != ISMPC_OK)
            throw std::runtime_error(std::string("ismpc_formc_solve_batch: ") + ismpc_error_string(rc) + " [" +
                                     ismpc_last_cuda_error(h_) + "]");
        for (int i = 0; i < n_; ++i)
            for (int c = 0; c < 3; ++c) {
                robots[i].comPos(c) = out_[i].next.com_pos[c]; robots[i].comVel(c) = out_[i].next.com_vel[c];
            }
    }
    const ismpc_formc_out_t& result(int i) const { return out_[i]; }
    Parameters& parameters() { return par_; }

private:
    void set_plan(const MatrixXd& f)
    {
        if (f.cols() < 4) throw std::runtime_error("MPCSolver: ftsp_and_timings needs 4 columns (x, y, z, t)");
        plan_rows_ = (int)f.rows();
        plan_.resize((size_t)plan_rows_ * 4);
        for (int i = 0; i < plan_rows_; ++i)
            for (int c = 0; c < 4; ++c) plan_[(size_t)i * 4 + c] = f(i, c);
        // like the reference's constructor argument, the plan lives with the solver: one upload, then every tick
        // moves only the state / walk-state records (ismpc_formc_set_plan)
        int rc = ismpc_formc_set_plan(h_, plan_.data(), plan_rows_, ISMPC_MEM_HOST);
        if (rc != ISMPC_OK) throw std::runtime_error(std::string("ismpc_formc_set_plan: ") + ismpc_error_string(rc));
    }
    int n_;
    Parameters par_;
    ismpc_handle* h_ = nullptr;
    int plan_rows_ = 0;
    std::vector<double> plan_;
    std::vector<ismpc_formc_inst_t> inst_;
    std::vector<ismpc_state_t> st_;
    std::vector<ismpc_walk_t> wk_;
    std::vector<ismpc_formc_out_t> out_;
};

// Drop-in for the reference class (AMR_code_DART/MPCSolver.hpp:16-28).
class MPCSolver {
public:
    explicit MPCSolver(const MatrixXd& ftsp_and_timings) : batch_(1, ftsp_and_timings) {}
    ~MPCSolver() = default;

    // Compute the next desired state starting from the current state (MPCSolver.cpp:204-501)
    State solve(State current, WalkState walkState, const MatrixXd& ftsp_and_timings)
    {
        itr = walkState.mpcIter;                 // MPCSolver.cpp:206
        fsCount = walkState.footstepCounter;     // MPCSolver.cpp:207
        std::vector<State> r(1, current);
        std::vector<WalkState> w(1, walkState);
        batch_.solve(r, w, ftsp_and_timings);
        status = batch_.result(0).status;
        zmp_x_input = batch_.result(0).zmp_in[0];
        zmp_y_input = batch_.result(0).zmp_in[1];
        return r[0];
    }

    // some stuff (public members of the reference class, MPCSolver.hpp:24-28)
    int itr = 0;
    int fsCount = 0, old_fsCount = 0, adaptation_memo = 0, ds_samples = 0, ct = 0;
    double xz_dot = 0.0, yz_dot = 0.0;
    // additions: what the reference only prints (MPCSolver.cpp:402-403,425) or drops (utils.cpp:128)
    double zmp_x_input = 0.0, zmp_y_input = 0.0;
    int status = 0;

private:
    MPCSolverBatch batch_;
};

}  // namespace ismpc_host
