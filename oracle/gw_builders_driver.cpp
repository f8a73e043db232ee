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
