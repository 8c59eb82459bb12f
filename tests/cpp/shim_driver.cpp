// Drives the MPCSolver shim exactly as AMR_code_DART/Controller.cpp does (:89-106 plan + construction,
// :346-348 per-tick call, :503-504 iteration counters) and prints the CoM trajectory for the Python test.
#include <cstdio>
#include <cstdlib>
#include "../../quadruped_gait_generation_ismpc_b200/host/MPCSolver.hpp"

using namespace ismpc_host;

int main(int argc, char** argv)
{
    const int ticks = argc > 1 ? atoi(argv[1]) : 20;
    const int S = 35, F = 10, N_footsteps = 40;
    MatrixXd ftsp_and_time = MatrixXd::Zero(N_footsteps, 4);          // Controller.cpp:89-97
    for (int i = 1; i < N_footsteps; i++) {
        ftsp_and_time(i, 0) = (i - 1) * 0.2;
        ftsp_and_time(i, 1) = ((i - 1) % 2 == 0 ? 1.0 : -1.0) * 0.08;
        ftsp_and_time(i, 2) = 0.0;
        ftsp_and_time(i, 3) = (double)(S + F) * i;
    }
    MPCSolver* solver = new MPCSolver(ftsp_and_time);                 // Controller.cpp:105-106
    State desired;
    desired.comPos(0) = 0.0; desired.comPos(1) = 0.0; desired.comPos(2) = 0.69;
    WalkState walkState;
    walkState.footstepCounter = 2;
    for (int k = 0; k < ticks; ++k) {
        walkState.simulationTime = k;                                 // Controller.cpp:310
        desired = solver->solve(desired, walkState, ftsp_and_time);   // Controller.cpp:346-348
        printf("%.17g %.17g %.17g %.17g %.17g %.17g %d\n", desired.comPos(0), desired.comPos(1), desired.comPos(2),
               desired.comVel(0), desired.comVel(1), desired.comVel(2), solver->status);
        ++walkState.controlIter;                                      // Controller.cpp:503
        walkState.mpcIter = (int)(walkState.controlIter * 0.01 / 0.01); // Controller.cpp:504 (floor, in doubles)
    }
    delete solver;
    return 0;
}
