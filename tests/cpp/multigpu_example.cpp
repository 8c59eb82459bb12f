// MPCSolverMultiGpu (host/MPCSolverMultiGpu.hpp) from plain C++: n copies of the DART app's robot with slightly different
// starts, sharded over the devices given on the command line.
//   multigpu_example <n_robots> <ticks> <gather: nccl|host> <device> [<device> ...]
// Prints, per robot: id, CoM x y z after `ticks` closed-loop ticks run (a) tick by tick through host buffers on all
// devices and (b) resident on the devices (scatter / rollout / gather), and the accumulated status of (b).
// Exit code 2 with the message on stderr if the group cannot be created (no GPU: there is no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../quadruped_gait_generation_ismpc_b200/host/MPCSolverMultiGpu.hpp"

using namespace ismpc_host;

int main(int argc, char** argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s n_robots ticks nccl|host device...\n", argv[0]); return 1; }
    const int n = atoi(argv[1]), ticks = atoi(argv[2]);
    const int mode = strcmp(argv[3], "host") == 0 ? ISMPC_GATHER_HOST : ISMPC_GATHER_NCCL;
    std::vector<int> devices;
    for (int a = 4; a < argc; ++a) devices.push_back(atoi(argv[a]));
    const int S = 35, F = 10, N_footsteps = 40;
    MatrixXd ftsp_and_time = MatrixXd::Zero(N_footsteps, 4);          // Controller.cpp:89-97
    for (int i = 1; i < N_footsteps; i++) {
        ftsp_and_time(i, 0) = (i - 1) * 0.2;
        ftsp_and_time(i, 1) = ((i - 1) % 2 == 0 ? 1.0 : -1.0) * 0.08;
        ftsp_and_time(i, 3) = (double)(S + F) * i;
    }
    try {
        MPCSolverMultiGpu<State, WalkState, MatrixXd> solver(n, ftsp_and_time, devices, Parameters(), mode);
        std::vector<State> a((size_t)n), b;
        std::vector<WalkState> wa((size_t)n), wb;
        for (int i = 0; i < n; ++i) {
            a[i].comPos(0) = 1e-3 * (i % 17); a[i].comPos(1) = 5e-4 * (i % 5); a[i].comPos(2) = 0.69;
            wa[i].footstepCounter = 2;
        }
        b = a; wb = wa;
        // (a) tick by tick, host buffers, Controller bookkeeping on the host (Controller.cpp:503-504; the footstep
        //     counter is left alone, as the shipped Controller does: `&& false`, :297)
        for (int k = 0; k < ticks; ++k) {
            for (int i = 0; i < n; ++i) wa[i].simulationTime = k;
            solver.solve(a, wa, ftsp_and_time);
            for (int i = 0; i < n; ++i) { ++wa[i].controlIter; wa[i].mpcIter = (int)floor(wa[i].controlIter * 0.01 / 0.01); }
        }
        // (b) resident closed loop: the library's rollout enables the footstep switch; with footstepCounter = 2 and a
        //     run shorter than the third step's start (t = 90 - 1) no switch happens, so (a) and (b) see the same ticks
        std::vector<int32_t> status;
        solver.scatter(b, wb);
        solver.rollout(ticks);
        solver.gather(b, wb, status);
        for (int i = 0; i < n; ++i)
            printf("%d %.17g %.17g %.17g %.17g %.17g %.17g %d\n", i, a[i].comPos(0), a[i].comPos(1), a[i].comPos(2),
                   b[i].comPos(0), b[i].comPos(1), b[i].comPos(2), status[(size_t)i]);
        fprintf(stderr, "devices %d, kernel launches %lld\n", solver.devices(), solver.kernel_launches());
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 2;
    }
    return 0;
}
