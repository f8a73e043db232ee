// rd_mex.cpp -- MATLAB gateway of the QPBO path: drop-in for cpp/rd_mex.cpp of the reference
// (same name, 7 or 8 inputs, 4 outputs, rd_mex.cpp:14-100), forwarding to sb_rd_solve.
//   [labelling, energy, lower_bound, num_unlabelled] =
//       rd_mex(U0, U1, E00, E01, E10, E11, connectivity0, options)
#include "sb_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    SB_MEX_ASSERT(nrhs == 7 || nrhs == 8);            // rd_mex.cpp:20
    SB_MEX_ASSERT(nlhs == 4);                         // rd_mex.cpp:21
    for (int i = 0; i < 6; i++) SB_MEX_ASSERT(mxGetClassID(prhs[i]) == mxDOUBLE_CLASS);
    SB_MEX_ASSERT(mxGetClassID(prhs[6]) == mxUINT32_CLASS);
    const mwSize N = mxGetM(prhs[0]), E = mxGetN(prhs[2]);
    // rd_mex.cpp:36-49
    SB_MEX_ASSERT(mxGetM(prhs[1]) == N && mxGetN(prhs[0]) == 1 && mxGetN(prhs[1]) == 1);
    SB_MEX_ASSERT(mxGetN(prhs[3]) == E && mxGetN(prhs[4]) == E && mxGetN(prhs[5]) == E && mxGetN(prhs[6]) == E);
    SB_MEX_ASSERT(mxGetM(prhs[2]) == 1 && mxGetM(prhs[3]) == 1 && mxGetM(prhs[4]) == 1);
    SB_MEX_ASSERT(mxGetM(prhs[6]) == 2);
    const bool improve = sb_mex_option_bool(nrhs - 7, prhs + 7, "improve", false);   // rd_mex.cpp:34

    plhs[0] = sb_mex_matrix(N, 1);
    plhs[1] = mxCreateDoubleScalar(0);
    plhs[2] = mxCreateDoubleScalar(0);
    plhs[3] = mxCreateDoubleScalar(0);
    sb_mex_check(sb_rd_solve((int64_t)N, (int64_t)E, mxGetPr(prhs[0]), mxGetPr(prhs[1]), mxGetPr(prhs[2]),
                             mxGetPr(prhs[3]), mxGetPr(prhs[4]), mxGetPr(prhs[5]), (const uint32_t *)mxGetData(prhs[6]),
                             improve ? 1 : 0, mxGetPr(plhs[0]), mxGetPr(plhs[1]), mxGetPr(plhs[2]), mxGetPr(plhs[3])));
}
