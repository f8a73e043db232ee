// trws_mex.cpp -- MATLAB gateway of the TRW-S path: drop-in for cpp/trws_mex.cpp of the
// reference (same name, same 8 inputs / 4 outputs, trws_mex.cpp:27-56,134-163), forwarding to
// sb_trws_solve of libstereo_b200.so.
//   [labelling, energy, lower_bound, iterations] =
//       trws_mex(kernel, unary, connectivity0, q, qprim, alphas, tol, options)
#include "sb_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    SB_MEX_ASSERT(nrhs == 8);                         // trws_mex.cpp:152
    SB_MEX_ASSERT(nlhs == 4);                         // trws_mex.cpp:55
    const mxArray *a_kernel = prhs[0], *a_unary = prhs[1], *a_conn = prhs[2], *a_q = prhs[3], *a_qp = prhs[4],
                  *a_alpha = prhs[5], *a_tol = prhs[6];
    SB_MEX_ASSERT(mxGetClassID(a_kernel) == mxINT32_CLASS);
    SB_MEX_ASSERT(mxGetClassID(a_unary) == mxDOUBLE_CLASS && mxGetClassID(a_q) == mxDOUBLE_CLASS &&
                  mxGetClassID(a_qp) == mxDOUBLE_CLASS && mxGetClassID(a_alpha) == mxDOUBLE_CLASS &&
                  mxGetClassID(a_tol) == mxDOUBLE_CLASS);
    SB_MEX_ASSERT(mxGetClassID(a_conn) == mxUINT32_CLASS);
    const int kernel = *(const int *)mxGetData(a_kernel);
    if (kernel != 1 && kernel != 2) mexErrMsgTxt("Unsupported kernel");   // trws_mex.cpp:162
    const mwSize L = mxGetM(a_unary), N = mxGetN(a_unary), E = mxGetN(a_conn);
    // trws_mex.cpp:42-52
    SB_MEX_ASSERT(mxGetM(a_conn) == 2);
    SB_MEX_ASSERT(mxGetN(a_q) == E && mxGetN(a_qp) == E);
    SB_MEX_ASSERT(mxGetM(a_q) == L && mxGetM(a_qp) == L);
    SB_MEX_ASSERT(mxGetNumberOfElements(a_alpha) == E);
    SB_MEX_ASSERT(mxGetNumberOfElements(a_tol) == 1);

    sb_trws_options opt;
    sb_trws_default_options(&opt);
    opt.maxiter = sb_mex_option_double(nrhs - 7, prhs + 7, "maxiter", 1000);       // trws_mex.cpp:39
    opt.max_relgap = sb_mex_option_double(nrhs - 7, prhs + 7, "max_relgap", 0);    // trws_mex.cpp:40

    plhs[0] = sb_mex_matrix(N, 1);
    plhs[1] = mxCreateDoubleScalar(0);
    plhs[2] = mxCreateDoubleScalar(0);
    plhs[3] = mxCreateDoubleScalar(0);
    sb_mex_check(sb_trws_solve(kernel, (int)L, (int64_t)N, (int64_t)E, mxGetPr(a_unary),
                               (const uint32_t *)mxGetData(a_conn), mxGetPr(a_q), mxGetPr(a_qp), mxGetPr(a_alpha),
                               *mxGetPr(a_tol), &opt, mxGetPr(plhs[0]), mxGetPr(plhs[1]), mxGetPr(plhs[2]),
                               mxGetPr(plhs[3]), NULL));
}
