// MPCSolver.hpp -- replacement for AMR_code_DART/MPCSolver.hpp: put this directory in front of the include path (or copy
// the file over the reference's) and drop MPCSolver.cpp from the target; Controller.hpp / Controller.cpp stay as they are.
//
//   Controller.hpp:69            MPCSolver* solver;
//   Controller.cpp:105-106       solver = new MPCSolver(ftsp_and_time_ref);
//   Controller.cpp:346-348       desired = solver->solve(desired, walkState, ftsp_and_time_ref);
//
// The class is the B200-backed ismpc_host::BasicMPCSolver instantiated on the reference's OWN types: `State` (21
// Eigen::Vector3d members + the getRel* methods, types.hpp:7-74 -- all of them carried through solve(), which replaces
// comPos and comVel as MPCSolver.cpp:419-422,275-276 do), `WalkState` (types.hpp:76-81) and Eigen::MatrixXd.  Like the
// reference's header (MPCSolver.hpp:5-7) it pulls in types.hpp, parameters.cpp and utils.cpp, whose globals
// Controller.cpp uses; the solver is configured from those globals, so editing parameters.cpp keeps working.
#pragma once

#include <Eigen/Core>
#include "types.hpp"
#include "parameters.cpp"
#include "utils.cpp"
#include "../MPCSolver.hpp"

namespace ismpc_host {
inline Parameters reference_parameters()      // AMR_code_DART/parameters.cpp:9-45
{
    Parameters p;
    p.mpcTimeStep = ::mpcTimeStep; p.controlTimeStep = ::controlTimeStep;
    p.singleSupportDuration = ::singleSupportDuration; p.doubleSupportDuration = ::doubleSupportDuration;
    p.predictionTime = ::predictionTime; p.comTargetHeight = ::comTargetHeight;
    p.footConstraintSquareWidth = ::footConstraintSquareWidth; p.mass_hrp4 = ::mass_hrp4; p.g = ::g;
    return p;
}
}  // namespace ismpc_host

class MPCSolver : public ismpc_host::BasicMPCSolver<State, WalkState, Eigen::MatrixXd> {
public:
    explicit MPCSolver(const Eigen::MatrixXd& ftsp_and_timings)
        : ismpc_host::BasicMPCSolver<State, WalkState, Eigen::MatrixXd>(ftsp_and_timings, ismpc_host::reference_parameters()) {}
};
