// sb_mex_common.h -- shared helpers of the MATLAB gateways (trws_mex.cpp, rd_mex.cpp,
// sb_builders_mex.cpp).  Only mex.h and the C ABI of libstereo_b200 (include/stereo_b200.h)
// are used: the gateways validate, fetch pointers, call the library and create the outputs --
// the same division of labour as the reference gateways (cpp/trws_mex.cpp, cpp/rd_mex.cpp),
// whose solvers are replaced by the CUDA library.
#pragma once
#include <string.h>
#include <string>
#include "mex.h"
#include "stereo_b200.h"

#define SB_MEX_ASSERT(cond)                                                              \
    do {                                                                                 \
        if (!(cond)) mexErrMsgTxt("Assertion failed: " #cond);                           \
    } while (0)

static inline void sb_mex_check(int rc)
{
    if (rc != SB_OK) mexErrMsgTxt(sb_last_error());   // solver errors -> mexErrMsgTxt (rd_mex.cpp:10-12)
}

// Option lookup with the conventions of MexParams (cpp/utils/mexutils.h:52-112): the trailing
// arguments are either one struct or name / value pairs; a missing name yields the default.
static inline const mxArray *sb_mex_option(int n, const mxArray **args, const char *name)
{
    if (n == 1 && mxGetClassID(args[0]) == mxSTRUCT_CLASS) {
        const int nf = mxGetNumberOfFields(args[0]);
        for (int f = 0; f < nf; f++)
            if (strcmp(mxGetFieldNameByNumber(args[0], f), name) == 0) return mxGetFieldByNumber(args[0], 0, f);
        return NULL;
    }
    for (int i = 0; i + 1 < n; i += 2) {
        char buf[256];
        if (mxGetClassID(args[i]) != mxCHAR_CLASS) continue;
        if (mxGetString(args[i], buf, sizeof(buf)) == 0 && strcmp(buf, name) == 0) return args[i + 1];
    }
    return NULL;
}

static inline double sb_mex_option_double(int n, const mxArray **args, const char *name, double def)
{
    const mxArray *a = sb_mex_option(n, args, name);
    if (!a) return def;
    SB_MEX_ASSERT(mxGetClassID(a) == mxDOUBLE_CLASS);   // matrix<double> asserts the class (cppmatrix.h:126)
    return *mxGetPr(a);
}

static inline bool sb_mex_option_bool(int n, const mxArray **args, const char *name, bool def)
{
    const mxArray *a = sb_mex_option(n, args, name);
    if (!a) return def;
    SB_MEX_ASSERT(mxGetClassID(a) == mxLOGICAL_CLASS);
    return *(const mxLogical *)mxGetData(a) != 0;
}

static inline mxArray *sb_mex_matrix(mwSize m, mwSize n)
{
    mwSize dims[2] = {m, n};
    return mxCreateNumericArray(2, dims, mxDOUBLE_CLASS, mxREAL);
}
