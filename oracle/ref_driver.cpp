// ref_driver.cpp -- C ABI around the UNMODIFIED reference mex gateways.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile together with the
// reference sources where they lie under /root/reference (never copied into
// this repo) into oracle/_ref/libref_trws.so and oracle/_ref/libref_rd.so.
// The driver fabricates the mxArray arguments MATLAB would pass
// (rd.m:21, trws.m:33 after the connectivity-1 shift) and calls the
// reference mexFunction (cpp/trws_mex.cpp:149, cpp/rd_mex.cpp:14).
//
// The entry points have the same argument lists as our product C ABI
// (include/stereo_b200.h: sb_trws_solve / sb_rd_solve) so the parity tests
// call both with the same buffers.
//
// Compile with -DSB_REF_TRWS or -DSB_REF_RD (one gateway per shared object:
// each reference gateway defines mexFunction).
#include <stdint.h>
#include <string.h>
#include <string>
#include <exception>
#include "mex.h"

static std::string g_last_error;

extern "C" const char *ref_last_error(void) { return g_last_error.c_str(); }

#ifdef SB_REF_TRWS
// trws_mex(kernel, unary, connectivity, q, qprim, alphas, tol, options)
// (cpp/trws_mex.cpp:27-163).  labels come back 1-based, as the gateway
// returns them (trws_mex.cpp:138).
extern "C" int ref_trws_solve(int kernel, int L, int64_t N, int64_t E,
                              const double *unary, const uint32_t *conn,
                              const double *q, const double *qprim,
                              const double *alphas, double tol,
                              double maxiter, double max_relgap,
                              double *labels, double *energy,
                              double *lower_bound, double *iterations)
{
    int32_t k = kernel;
    mwSize d11[2] = {1, 1};
    mwSize dLN[2] = {L, (mwSize)N};
    mwSize d2E[2] = {2, (mwSize)E};
    mwSize dLE[2] = {L, (mwSize)E};
    mwSize dE1[2] = {(mwSize)E, 1};
    mxArray *a_kernel = sb_mxWrap(&k, mxINT32_CLASS, 2, d11);
    mxArray *a_unary = sb_mxWrap((void *)unary, mxDOUBLE_CLASS, 2, dLN);
    mxArray *a_conn = sb_mxWrap((void *)conn, mxUINT32_CLASS, 2, d2E);
    mxArray *a_q = sb_mxWrap((void *)q, mxDOUBLE_CLASS, 2, dLE);
    mxArray *a_qp = sb_mxWrap((void *)qprim, mxDOUBLE_CLASS, 2, dLE);
    mxArray *a_al = sb_mxWrap((void *)alphas, mxDOUBLE_CLASS, 2, dE1);
    mxArray *a_tol = sb_mxWrap(&tol, mxDOUBLE_CLASS, 2, d11);
    mxArray *opt = sb_mxCreateStruct();
    sb_mxAddField(opt, "maxiter", mxCreateDoubleScalar(maxiter));
    sb_mxAddField(opt, "max_relgap", mxCreateDoubleScalar(max_relgap));
    const mxArray *prhs[8] = {a_kernel, a_unary, a_conn, a_q, a_qp, a_al, a_tol, opt};
    mxArray *plhs[4] = {0, 0, 0, 0};
    int rc = 0;
    try {
        mexFunction(4, plhs, 8, prhs);
        memcpy(labels, mxGetPr(plhs[0]), sizeof(double) * (size_t)N);
        *energy = mxGetScalar(plhs[1]);
        *lower_bound = mxGetScalar(plhs[2]);
        *iterations = mxGetScalar(plhs[3]);
    } catch (const std::exception &e) {
        g_last_error = e.what();
        rc = -1;
    }
    for (int i = 0; i < 4; i++) mxDestroyArray(plhs[i]);
    mxDestroyArray(opt);
    mxArray *tmp[7] = {a_kernel, a_unary, a_conn, a_q, a_qp, a_al, a_tol};
    for (int i = 0; i < 7; i++) mxDestroyArray(tmp[i]);
    return rc;
}
#endif

#ifdef SB_REF_RD
// rd_mex(U0,U1,E00,E01,E10,E11,connectivity,options)  (cpp/rd_mex.cpp:14-100)
extern "C" int ref_rd_solve(int64_t N, int64_t E,
                            const double *U0, const double *U1,
                            const double *E00, const double *E01,
                            const double *E10, const double *E11,
                            const uint32_t *conn, int improve,
                            double *labels, double *energy,
                            double *lower_bound, double *num_unlabelled)
{
    mwSize dN1[2] = {(mwSize)N, 1};
    mwSize d1E[2] = {1, (mwSize)E};
    mwSize d2E[2] = {2, (mwSize)E};
    mxArray *a[7];
    a[0] = sb_mxWrap((void *)U0, mxDOUBLE_CLASS, 2, dN1);
    a[1] = sb_mxWrap((void *)U1, mxDOUBLE_CLASS, 2, dN1);
    a[2] = sb_mxWrap((void *)E00, mxDOUBLE_CLASS, 2, d1E);
    a[3] = sb_mxWrap((void *)E01, mxDOUBLE_CLASS, 2, d1E);
    a[4] = sb_mxWrap((void *)E10, mxDOUBLE_CLASS, 2, d1E);
    a[5] = sb_mxWrap((void *)E11, mxDOUBLE_CLASS, 2, d1E);
    a[6] = sb_mxWrap((void *)conn, mxUINT32_CLASS, 2, d2E);
    mxArray *opt = sb_mxCreateStruct();
    sb_mxAddField(opt, "improve", mxCreateLogicalScalar(improve));
    const mxArray *prhs[8] = {a[0], a[1], a[2], a[3], a[4], a[5], a[6], opt};
    mxArray *plhs[4] = {0, 0, 0, 0};
    int rc = 0;
    try {
        mexFunction(4, plhs, 8, prhs);
        memcpy(labels, mxGetPr(plhs[0]), sizeof(double) * (size_t)N);
        *energy = mxGetScalar(plhs[1]);
        *lower_bound = mxGetScalar(plhs[2]);
        *num_unlabelled = mxGetScalar(plhs[3]);
    } catch (const std::exception &e) {
        g_last_error = e.what();
        rc = -1;
    }
    // NB rd_mex.cpp:77-80 hands the matrix<> outputs to plhs *before* filling
    // them and the matrix<> destructors do not free them (cppmatrix.h:262).
    for (int i = 0; i < 4; i++) mxDestroyArray(plhs[i]);
    mxDestroyArray(opt);
    for (int i = 0; i < 7; i++) mxDestroyArray(a[i]);
    return rc;
}
#endif

#ifdef SB_REF_INTERP2
// vgg_interp2(A, X, Y, 'linear', oobv)  (imrender/vgg/vgg_interp2.cxx:43-145,
// linear kernel :246-322).  A is h x w x c double column-major; X,Y are n
// doubles (1-based, X = column coordinate); out is n x c double.
extern "C" int ref_interp2_linear(const double *A, int h, int w, int c,
                                  const double *X, const double *Y, int64_t n,
                                  double oobv, double *out)
{
    mwSize dA[3] = {h, w, c};
    mwSize dX[2] = {(mwSize)n, 1};
    mxArray *a_A = sb_mxWrap((void *)A, mxDOUBLE_CLASS, 3, dA);
    mxArray *a_X = sb_mxWrap((void *)X, mxDOUBLE_CLASS, 2, dX);
    mxArray *a_Y = sb_mxWrap((void *)Y, mxDOUBLE_CLASS, 2, dX);
    mxArray *a_m = sb_mxCreateString("linear");
    mxArray *a_o = mxCreateDoubleScalar(oobv);
    const mxArray *prhs[5] = {a_A, a_X, a_Y, a_m, a_o};
    mxArray *plhs[1] = {0};
    int rc = 0;
    try {
        mexFunction(1, plhs, 5, prhs);
        memcpy(out, mxGetData(plhs[0]), sizeof(double) * (size_t)n * (size_t)c);
    } catch (const std::exception &e) {
        g_last_error = e.what();
        rc = -1;
    }
    mxDestroyArray(plhs[0]);
    mxDestroyArray(a_A); mxDestroyArray(a_X); mxDestroyArray(a_Y);
    mxDestroyArray(a_m); mxDestroyArray(a_o);
    return rc;
}
#endif
