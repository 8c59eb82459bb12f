// qp_dense.cu -- generic dense QP batch (the solveQP seam, AMR_code_DART/utils.cpp:89-139).
#include "common.cuh"
#include "das.cuh"
#include "launch.h"

namespace ismpc {

size_t qp_dense_work_doubles(int n, int nV, int nC) { (void)n; (void)nV; (void)nC; return 1; }

int qp_dense_launch(int, int, int, const double*, const double*, const double*, const double*, const double*,
                    double*, double*, signed char*, int32_t*, int32_t*, double*, cudaStream_t)
{
    return (int)cudaErrorNotSupported;   // filled in by the dense solver milestone
}

}  // namespace ismpc
