// The reference's own call sites, compiled against host/dropin/MPCSolver.hpp with the reference's `State` / `WalkState`
// (types.hpp of the reference where /root/reference exists, a generated stand-in with the same 21 members and the
// getRel* methods elsewhere; utils.cpp and Eigen are stubs -- HPIPM, BLASFEO and Eigen are not installed here).
//   argv[1] = ticks.  Prints the CoM trajectory + status per tick, then "carried <0|1>" (every untouched member of State
//   survives solve()) and "uploads <n>" (the plan goes to the device again only when its contents change).
#include <cstdio>
#include <cstdlib>
#include "MPCSolver.hpp"        // host/dropin/MPCSolver.hpp, found first on the include path

int main(int argc, char** argv)
{
    const int ticks = argc > 1 ? atoi(argv[1]) : 20;
    int N_footsteps = 40;
    Eigen::MatrixXd ftsp_and_time = Eigen::MatrixXd::Zero(N_footsteps, 4);      // Controller.cpp:89-97
    for (int i = 1; i < N_footsteps; i++) {
        ftsp_and_time(i, 0) = (i - 1) * 0.2;
        ftsp_and_time(i, 1) = ((i - 1) % 2 == 0 ? 1.0 : -1.0) * 0.08;
        ftsp_and_time(i, 2) = 0.0;
        ftsp_and_time(i, 3) = (double)(S + F) * i;                              // S, F: parameters.cpp globals
    }
    const Eigen::MatrixXd& ftsp_and_time_ref = ftsp_and_time;
    MPCSolver* solver = new MPCSolver(ftsp_and_time_ref);                      // Controller.cpp:105-106
    State desired;                                                              // Controller.hpp: State desired;
    desired.comPos << 0.0, 0.0, comTargetHeight;
    // members solve() must not touch: give each a distinct value
    Eigen::Vector3d* others[] = {&desired.comAcc, &desired.leftBackFootPos, &desired.leftBackFootVel, &desired.leftBackFootAcc,
                                 &desired.rightBackFootPos, &desired.rightBackFootVel, &desired.rightBackFootAcc,
                                 &desired.leftFrontFootPos, &desired.leftFrontFootVel, &desired.leftFrontFootAcc,
                                 &desired.rightFrontFootPos, &desired.rightFrontFootVel, &desired.rightFrontFootAcc,
                                 &desired.torsoOrient, &desired.leftBackFootOrient, &desired.rightBackFootOrient,
                                 &desired.leftFrontFootOrient, &desired.rightFrontFootOrient, &desired.zmpPos};
    const int n_others = (int)(sizeof(others) / sizeof(others[0]));
    for (int k = 0; k < n_others; ++k) for (int c = 0; c < 3; ++c) (*others[k])(c) = 100.0 * (k + 1) + c;
    WalkState walkState;
    walkState.supportFoot = true; walkState.simulationTime = 0; walkState.mpcIter = 0; walkState.controlIter = 0;
    walkState.footstepCounter = 2; walkState.indInitial = 0;
    int carried = 1;
    for (int k = 0; k < ticks; ++k) {
        walkState.simulationTime = k;                                          // Controller.cpp:310
        if (k == ticks / 2) ftsp_and_time(N_footsteps - 1, 0) += 0.5;           // the plan's CONTENTS change once (same shape)
        desired = solver->solve(desired, walkState, ftsp_and_time_ref);        // Controller.cpp:346-348
        printf("%.17g %.17g %.17g %.17g %.17g %.17g %d\n", desired.comPos(0), desired.comPos(1), desired.comPos(2),
               desired.comVel(0), desired.comVel(1), desired.comVel(2), solver->status);
        for (int q = 0; q < n_others; ++q) for (int c = 0; c < 3; ++c) if ((*others[q])(c) != 100.0 * (q + 1) + c) carried = 0;
        ++walkState.controlIter;                                               // Controller.cpp:503
        walkState.mpcIter = (int)floor(walkState.controlIter * controlTimeStep / mpcTimeStep);   // Controller.cpp:504
    }
    Eigen::VectorXd pose = desired.getRelComPose(walkState.supportFoot);       // the methods of the reference's State exist
    printf("carried %d\nuploads %d\npose %d\n", carried, solver->plan_uploads(), pose.size());
    delete solver;
    return 0;
}
