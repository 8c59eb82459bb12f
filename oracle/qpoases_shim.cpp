/*
 * qpoases_shim.cpp -- drives the reference's vendored qpOASES 3.2 (compiled unmodified, in place,
 * from /root/reference/AMR_code_DART/qpOASES) with the exact call form of the reference's solveQP
 * wrapper (AMR_code_DART/utils.cpp:121-130): Options::setToMPC(), printLevel = PL_NONE, nWSR = 300,
 * a fresh QProblem(nV,nC) per call, init(H,g,A,0,0,lbA,ubA,nWSR,NULL,...), getPrimalSolution.
 * Differences from the wrapper, none of which change results: buffers are caller/heap owned (the
 * wrapper's VLAs overflow the stack near nV~800, utils.cpp:97-102); the return code, duals, working
 * set and iteration count are returned instead of dropped (utils.cpp:128).
 *
 * TEST INFRASTRUCTURE ONLY -- see ismpc_oracle.h.  Built only where /root/reference exists; the
 * resulting oracle/_ref/libismpc_oracle_ref.so travels to the GPU box as a prebuilt file.
 */
#include "qpOASES.hpp"
#include "ismpc_oracle.h"
#include <vector>

static thread_local int g_nwsr_cap = 300; /* utils.cpp:124 */

extern "C" void oracle_qpoases_set_nwsr(int cap) { g_nwsr_cap = cap > 0 ? cap : 300; }

extern "C" int oracle_qpoases_solve(int nV, int nC, const double* H, const double* g,
                                    const double* A, const double* lbA, const double* ubA,
                                    double* x, double* y, int* ws, int* nwsr)
{
    qpOASES::Options options;
    options.setToMPC();                                   /* utils.cpp:122 */
    options.printLevel = qpOASES::PL_NONE;                /* utils.cpp:123 */
    qpOASES::int_t nWSR = g_nwsr_cap;                     /* utils.cpp:124 */
    qpOASES::QProblem qp(nV, nC);                         /* utils.cpp:126 */
    qp.setOptions(options);                               /* utils.cpp:127 */
    qpOASES::returnValue rv =
        qp.init(H, g, A, 0, 0, lbA, ubA, nWSR, NULL, NULL, NULL, NULL, NULL, NULL); /* utils.cpp:128 */
    qp.getPrimalSolution(x);                              /* utils.cpp:130 */
    if (nwsr) *nwsr = (int)nWSR;
    if (y) {
        std::vector<double> yy(nV + nC);
        qp.getDualSolution(yy.data());
        for (int i = 0; i < nC; ++i) y[i] = yy[nV + i];
    }
    if (ws) {
        std::vector<double> w(nC);
        qp.getWorkingSetConstraints(w.data());            /* qpOASES/QProblem.cpp:809-829 */
        for (int i = 0; i < nC; ++i) ws[i] = (int)w[i];
    }
    return (int)rv;
}

extern "C" int oracle_have_qpoases(void) { return 1; }
