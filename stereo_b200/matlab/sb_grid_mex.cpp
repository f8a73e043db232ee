// sb_grid_mex.cpp -- MATLAB gateway of the GRID-NATIVE TRW-S entry (include/stereo_b200.h, sb_trws_grid_*): what
// dispmap_super.simultaneous_fusion (dispmap_super.m:153-198) calls instead of trws(...) when the L x E arrays q / qprim
// it would have to build do not fit (BASELINE configs 4-5: 140-560 GB).  The method hands over what it HOLDS:
//
//   [labelling, energy, lower_bound, iterations] =
//       sb_grid_mex(kernel, sz, proposals, unary, weights, tol, dnorm, options)
//
//   kernel     int32 1 | 2                                  (trws_mex.cpp:27)
//   sz         [H W]
//   proposals  4 x N x L doubles: cat(3, proposal_cell{:}), the current assignment last (dispmap_super.m:158)
//   unary      N x L doubles: unary_cost of each proposal  (dispmap_super.m:164-168, transposed: one column per label)
//   weights    E doubles, smooth_weights                    (dispmap_super.m:186)
//   tol        scalar;  dnorm = [d_min d_step]
//   options    struct or name / value pairs: maxiter (1000), max_relgap (0)   (trws_mex.cpp:39-40)
//
// Outputs as trws_mex.cpp:134-147: labelling N x 1 (1-based label per pixel), energy, lower bound, iterations.
// The proposals go to the device one label at a time (sb_trws_grid_set_labels), so nothing of size L x E exists anywhere.
#include "sb_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    SB_MEX_ASSERT(nrhs >= 7);
    SB_MEX_ASSERT(nlhs == 4);
    const mxArray *a_kernel = prhs[0], *a_sz = prhs[1], *a_prop = prhs[2], *a_unary = prhs[3], *a_w = prhs[4], *a_tol = prhs[5],
                  *a_dn = prhs[6];
    SB_MEX_ASSERT(mxGetClassID(a_kernel) == mxINT32_CLASS);
    SB_MEX_ASSERT(mxGetClassID(a_sz) == mxDOUBLE_CLASS && mxGetNumberOfElements(a_sz) == 2);
    SB_MEX_ASSERT(mxGetClassID(a_prop) == mxDOUBLE_CLASS && mxGetClassID(a_unary) == mxDOUBLE_CLASS &&
                  mxGetClassID(a_w) == mxDOUBLE_CLASS && mxGetClassID(a_tol) == mxDOUBLE_CLASS && mxGetClassID(a_dn) == mxDOUBLE_CLASS);
    SB_MEX_ASSERT(mxGetNumberOfElements(a_tol) == 1 && mxGetNumberOfElements(a_dn) == 2);
    const int kernel = *(const int *)mxGetData(a_kernel);
    if (kernel != 1 && kernel != 2) mexErrMsgTxt("Unsupported kernel");   // trws_mex.cpp:162
    const int H = (int)mxGetPr(a_sz)[0], W = (int)mxGetPr(a_sz)[1];
    SB_MEX_ASSERT(H >= 1 && W >= 1);
    const size_t N = (size_t)H * W;
    const size_t E = 2 * ((size_t)(H - 1) * W + (size_t)H * (W - 1));
    const size_t np = mxGetNumberOfElements(a_prop);
    SB_MEX_ASSERT(np % (4 * N) == 0 && np > 0);
    const int L = (int)(np / (4 * N));
    SB_MEX_ASSERT(mxGetNumberOfElements(a_unary) == N * (size_t)L);
    SB_MEX_ASSERT(mxGetNumberOfElements(a_w) == E);

    sb_trws_options opt;
    sb_trws_default_options(&opt);
    opt.maxiter = sb_mex_option_double(nrhs - 7, prhs + 7, "maxiter", 1000);       // trws_mex.cpp:39
    opt.max_relgap = sb_mex_option_double(nrhs - 7, prhs + 7, "max_relgap", 0);    // trws_mex.cpp:40

    sb_trws_grid *g = NULL;
    sb_mex_check(sb_trws_grid_create(kernel, H, W, L, *mxGetPr(a_tol), &opt, 0, 1, &g));
    // everything below must release the solver before it reports an error
    int rc = SB_OK;
    const double *prop = mxGetPr(a_prop), *un = mxGetPr(a_unary), *dn = mxGetPr(a_dn);
    for (int l = 0; l < L && rc == SB_OK; l++)
        rc = sb_trws_grid_set_labels(g, l, 1, prop + (size_t)l * 4 * N, un + (size_t)l * N, dn[0], dn[1]);
    if (rc == SB_OK) rc = sb_trws_grid_set_weights(g, mxGetPr(a_w));
    if (rc == SB_OK) rc = sb_trws_grid_finalize(g);
    plhs[0] = sb_mex_matrix(N, 1);
    plhs[1] = mxCreateDoubleScalar(0);
    plhs[2] = mxCreateDoubleScalar(0);
    plhs[3] = mxCreateDoubleScalar(0);
    if (rc == SB_OK)
        rc = sb_trws_grid_minimize(g, opt.maxiter, opt.max_relgap, mxGetPr(plhs[1]), mxGetPr(plhs[2]), mxGetPr(plhs[3]), NULL);
    if (rc == SB_OK) rc = sb_trws_grid_get_labels(g, mxGetPr(plhs[0]));
    std::string err = rc == SB_OK ? std::string() : std::string(sb_last_error());
    sb_trws_grid_destroy(g);
    if (rc != SB_OK) mexErrMsgTxt(err.c_str());
}
