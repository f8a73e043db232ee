// gw_builders_driver.cpp -- calls the product's builders gateway (stereo_b200/matlab/sb_builders_mex.cpp) the way MATLAB
// would: an operation name plus numeric arrays as fabricated mxArrays (oracle/mex_shim/mex.h), outputs handed back as
// (class, dims, data) views that stay alive until the next call.
//
// TEST INFRASTRUCTURE ONLY (built by `make -C oracle gateway` into oracle/_build/libgw_builders.so; used by
// tests/test_gateway_gpu.py through oracle/oracle.py).  Nothing in the product links it.
#include <stdint.h>
#include <string.h>
#include <string>
#include <exception>
#include "mex.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

extern "C" {

struct GwArray {
    int classid;      // mxClassID
    int ndim;
    int dims[4];
    void *data;
};

static std::string g_err;
static mxArray *g_out[8] = {0, 0, 0, 0, 0, 0, 0, 0};

const char *gw_builders_last_error(void) { return g_err.c_str(); }

// args: nargs arrays borrowed from the caller; outs: nlhs views filled in.  Returns 0, or -1 with gw_builders_last_error().
int gw_builders_call(const char *op, int nargs, const GwArray *args, int nlhs, GwArray *outs)
{
    for (int i = 0; i < 8; i++) { mxDestroyArray(g_out[i]); g_out[i] = 0; }
    if (nargs > 15 || nlhs > 8) { g_err = "gw_builders_call: too many arguments"; return -1; }
    const mxArray *prhs[16];
    mxArray *own[16];
    own[0] = sb_mxCreateString(op);
    prhs[0] = own[0];
    for (int i = 0; i < nargs; i++) {
        mwSize d[4];
        for (int k = 0; k < 4; k++) d[k] = args[i].dims[k];
        own[i + 1] = sb_mxWrap(args[i].data, (mxClassID)args[i].classid, args[i].ndim, d);
        prhs[i + 1] = own[i + 1];
    }
    mxArray *plhs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int rc = 0;
    try {
        mexFunction(nlhs, plhs, nargs + 1, prhs);
    } catch (const std::exception &e) {
        g_err = e.what();
        rc = -1;
    }
    for (int i = 0; i <= nargs; i++) mxDestroyArray(own[i]);
    for (int i = 0; i < 8; i++) g_out[i] = plhs[i];
    if (rc == 0)
        for (int i = 0; i < nlhs; i++) {
            if (!plhs[i]) { g_err = "gw_builders_call: the gateway left an output empty"; return -1; }
            outs[i].classid = (int)mxGetClassID(plhs[i]);
            outs[i].ndim = mxGetNumberOfDimensions(plhs[i]);
            for (int k = 0; k < 4; k++) outs[i].dims[k] = k < outs[i].ndim ? (int)mxGetDimensions(plhs[i])[k] : 1;
            outs[i].data = mxGetData(plhs[i]);
        }
    return rc;
}

} // extern "C"

#ifdef SB_GW_GRID
// sb_grid_mex(kernel, sz, proposals, unary, weights, tol, dnorm, options) (stereo_b200/matlab/sb_grid_mex.cpp) with the
// arguments MATLAB would pass: int32 kernel, 4 x N x L proposals, N x L unaries, an options struct.
extern "C" int gw_grid_solve(int kernel, int H, int W, int L, const double *proposals, const double *unary, const double *weights,
                             double tol, double d_min, double d_step, double maxiter, double max_relgap, double *labels,
                             double *energy, double *lower_bound, double *iterations)
{
    const int N = H * W, E = 2 * ((H - 1) * W + H * (W - 1));
    int32_t k = kernel;
    double sz[2] = {(double)H, (double)W}, dn[2] = {d_min, d_step};
    mwSize d11[2] = {1, 1}, d12[2] = {1, 2}, dP[3] = {4, N, L}, dU[2] = {N, L}, dW[2] = {1, E};
    mxArray *a[7] = {sb_mxWrap(&k, mxINT32_CLASS, 2, d11),          sb_mxWrap(sz, mxDOUBLE_CLASS, 2, d12),
                     sb_mxWrap((void *)proposals, mxDOUBLE_CLASS, 3, dP), sb_mxWrap((void *)unary, mxDOUBLE_CLASS, 2, dU),
                     sb_mxWrap((void *)weights, mxDOUBLE_CLASS, 2, dW),   sb_mxWrap(&tol, mxDOUBLE_CLASS, 2, d11),
                     sb_mxWrap(dn, mxDOUBLE_CLASS, 2, d12)};
    mxArray *opt = sb_mxCreateStruct();
    sb_mxAddField(opt, "maxiter", mxCreateDoubleScalar(maxiter));
    sb_mxAddField(opt, "max_relgap", mxCreateDoubleScalar(max_relgap));
    const mxArray *prhs[8] = {a[0], a[1], a[2], a[3], a[4], a[5], a[6], opt};
    mxArray *plhs[4] = {0, 0, 0, 0};
    int rc = 0;
    try {
        mexFunction(4, plhs, 8, prhs);
        memcpy(labels, mxGetPr(plhs[0]), sizeof(double) * (size_t)N);
        *energy = mxGetScalar(plhs[1]);
        *lower_bound = mxGetScalar(plhs[2]);
        *iterations = mxGetScalar(plhs[3]);
    } catch (const std::exception &e) {
        g_err = e.what();
        rc = -1;
    }
    for (int i = 0; i < 4; i++) mxDestroyArray(plhs[i]);
    mxDestroyArray(opt);
    for (int i = 0; i < 7; i++) mxDestroyArray(a[i]);
    return rc;
}
#endif
