/* mex.h -- stand-in for MATLAB's mex.h so the reference mex gateways compile
 * without MATLAB.  TEST INFRASTRUCTURE ONLY (see oracle/README.md): it is used
 *   (a) by oracle/Makefile to build the *reference* gateways
 *       (/root/reference/cpp/{trws_mex,rd_mex}.cpp, imrender/vgg/vgg_interp2.cxx)
 *       into oracle/_ref/ as the parity oracle, and
 *   (b) to compile-check our own gateways in stereo_b200/matlab/.
 * It implements just the subset of the mx and mex API those files touch
 * (list in SURVEY.md section 8(c)).  Header-only; every function is inline.
 * An mxArray here is a plain heap struct; data is column-major like MATLAB.
 */
#ifndef SB_ORACLE_MEX_SHIM_H
#define SB_ORACLE_MEX_SHIM_H

#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>

#ifdef __cplusplus
#include <stdexcept>
#include <string>
#endif

typedef int mwSize;
typedef int mwIndex;
typedef unsigned char mxLogical;
typedef unsigned short mxChar;

typedef enum {
    mxUNKNOWN_CLASS = 0,
    mxCELL_CLASS,
    mxSTRUCT_CLASS,
    mxLOGICAL_CLASS,
    mxCHAR_CLASS,
    mxVOID_CLASS,
    mxDOUBLE_CLASS,
    mxSINGLE_CLASS,
    mxINT8_CLASS,
    mxUINT8_CLASS,
    mxINT16_CLASS,
    mxUINT16_CLASS,
    mxINT32_CLASS,
    mxUINT32_CLASS,
    mxINT64_CLASS,
    mxUINT64_CLASS,
    mxFUNCTION_CLASS
} mxClassID;

typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

#define SB_MX_MAXDIM 4
#define SB_MX_MAXFIELDS 16

typedef struct mxArray_tag {
    mxClassID classid;
    int ndim;
    mwSize dims[SB_MX_MAXDIM];
    void *data;
    int owns_data;
    /* struct arrays (1x1 only) */
    int nfields;
    const char *field_names[SB_MX_MAXFIELDS];
    struct mxArray_tag *field_values[SB_MX_MAXFIELDS];
} mxArray;

static inline size_t sb_mx_elsize(mxClassID c)
{
    switch (c) {
    case mxDOUBLE_CLASS: case mxINT64_CLASS: case mxUINT64_CLASS: return 8;
    case mxSINGLE_CLASS: case mxINT32_CLASS: case mxUINT32_CLASS: return 4;
    case mxINT16_CLASS: case mxUINT16_CLASS: case mxCHAR_CLASS: return 2;
    case mxINT8_CLASS: case mxUINT8_CLASS: case mxLOGICAL_CLASS: return 1;
    default: return 0;
    }
}

static inline size_t mxGetNumberOfElements(const mxArray *a)
{
    size_t n = 1;
    for (int i = 0; i < a->ndim; i++) n *= (size_t)a->dims[i];
    return n;
}

static inline mxArray *mxCreateNumericArray(int ndim, const mwSize *dims, mxClassID c, mxComplexity)
{
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    a->classid = c;
    a->ndim = ndim < 2 ? 2 : ndim;
    a->dims[0] = a->dims[1] = 1;
    for (int i = 0; i < ndim && i < SB_MX_MAXDIM; i++) a->dims[i] = dims[i];
    size_t n = mxGetNumberOfElements(a);
    a->data = calloc(n ? n : 1, sb_mx_elsize(c));
    a->owns_data = 1;
    return a;
}

static inline mxArray *mxCreateNumericMatrix(int m, int n, mxClassID c, mxComplexity f)
{
    mwSize d[2] = { m, n };
    return mxCreateNumericArray(2, d, c, f);
}

static inline mxArray *mxCreateDoubleMatrix(int m, int n, mxComplexity f)
{
    return mxCreateNumericMatrix(m, n, mxDOUBLE_CLASS, f);
}

static inline mxArray *mxCreateDoubleScalar(double v)
{
    mxArray *a = mxCreateNumericMatrix(1, 1, mxDOUBLE_CLASS, mxREAL);
    *(double *)a->data = v;
    return a;
}

static inline mxArray *mxCreateLogicalScalar(int v)
{
    mxArray *a = mxCreateNumericMatrix(1, 1, mxLOGICAL_CLASS, mxREAL);
    *(mxLogical *)a->data = (mxLogical)(v != 0);
    return a;
}

/* Borrow caller memory without copying (what MATLAB does for prhs). */
static inline mxArray *sb_mxWrap(void *data, mxClassID c, int ndim, const mwSize *dims)
{
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    a->classid = c;
    a->ndim = ndim < 2 ? 2 : ndim;
    a->dims[0] = a->dims[1] = 1;
    for (int i = 0; i < ndim && i < SB_MX_MAXDIM; i++) a->dims[i] = dims[i];
    a->data = data;
    a->owns_data = 0;
    return a;
}

static inline mxArray *sb_mxCreateStruct(void)
{
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    a->classid = mxSTRUCT_CLASS;
    a->ndim = 2;
    a->dims[0] = a->dims[1] = 1;
    return a;
}

static inline void sb_mxAddField(mxArray *s, const char *name, mxArray *value)
{
    if (s->nfields < SB_MX_MAXFIELDS) {
        s->field_names[s->nfields] = name;
        s->field_values[s->nfields] = value;
        s->nfields++;
    }
}

static inline mxArray *sb_mxCreateString(const char *str)
{
    int n = (int)strlen(str);
    mxArray *a = mxCreateNumericMatrix(1, n, mxCHAR_CLASS, mxREAL);
    for (int i = 0; i < n; i++) ((mxChar *)a->data)[i] = (mxChar)(unsigned char)str[i];
    return a;
}

static inline void mxDestroyArray(mxArray *a)
{
    if (!a) return;
    for (int i = 0; i < a->nfields; i++) mxDestroyArray(a->field_values[i]);
    if (a->owns_data) free(a->data);
    free(a);
}

/* deep copy of a numeric array (mxDuplicateArray); structs are not needed by the gateways */
static inline mxArray *mxDuplicateArray(const mxArray *src)
{
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    a->classid = src->classid;
    a->ndim = src->ndim;
    size_t n = 1;
    for (int i = 0; i < SB_MX_MAXDIM; i++) a->dims[i] = src->dims[i];
    for (int i = 0; i < src->ndim; i++) n *= (size_t)src->dims[i];
    const size_t bytes = n * sb_mx_elsize(src->classid);
    a->data = malloc(bytes ? bytes : 1);
    if (bytes) memcpy(a->data, src->data, bytes);
    a->owns_data = 1;
    return a;
}
/* shrink the column count of a matrix in place (mxSetN keeps the allocation) */
static inline void mxSetN(mxArray *a, mwSize n) { a->ndim = 2; a->dims[1] = n; }

static inline void *mxMalloc(size_t n) { return malloc(n ? n : 1); }
static inline void *mxCalloc(size_t n, size_t s) { return calloc(n ? n : 1, s ? s : 1); }
static inline void mxFree(void *p) { free(p); }

static inline int mxIsSparse(const mxArray *) { return 0; }
static inline int mxIsComplex(const mxArray *) { return 0; }
static inline int mxIsDouble(const mxArray *a) { return a->classid == mxDOUBLE_CLASS; }
static inline int mxIsStruct(const mxArray *a) { return a->classid == mxSTRUCT_CLASS; }
static inline int mxIsLogical(const mxArray *a) { return a->classid == mxLOGICAL_CLASS; }
static inline int mxIsChar(const mxArray *a) { return a->classid == mxCHAR_CLASS; }
static inline mxClassID mxGetClassID(const mxArray *a) { return a->classid; }
static inline int mxGetNumberOfDimensions(const mxArray *a) { return a->ndim; }
static inline const mwSize *mxGetDimensions(const mxArray *a) { return a->dims; }
static inline size_t mxGetM(const mxArray *a) { return (size_t)a->dims[0]; }
static inline size_t mxGetN(const mxArray *a)
{
    size_t n = 1;
    for (int i = 1; i < a->ndim; i++) n *= (size_t)a->dims[i];
    return n;
}
static inline double *mxGetPr(const mxArray *a) { return (double *)a->data; }
static inline void *mxGetData(const mxArray *a) { return a->data; }
static inline size_t mxGetElementSize(const mxArray *a) { return sb_mx_elsize(a->classid); }
static inline int mxGetNumberOfFields(const mxArray *a) { return a->nfields; }
static inline const char *mxGetFieldNameByNumber(const mxArray *a, int i)
{
    return (i >= 0 && i < a->nfields) ? a->field_names[i] : NULL;
}
static inline mxArray *mxGetFieldByNumber(const mxArray *a, mwIndex, int i)
{
    return (i >= 0 && i < a->nfields) ? a->field_values[i] : NULL;
}
static inline mxArray *mxGetField(const mxArray *a, mwIndex, const char *name)
{
    for (int i = 0; i < a->nfields; i++)
        if (strcmp(a->field_names[i], name) == 0) return a->field_values[i];
    return NULL;
}

static inline double mxGetScalar(const mxArray *a)
{
    switch (a->classid) {
    case mxDOUBLE_CLASS: return *(double *)a->data;
    case mxSINGLE_CLASS: return *(float *)a->data;
    case mxINT32_CLASS: return *(int *)a->data;
    case mxUINT32_CLASS: return *(unsigned *)a->data;
    case mxLOGICAL_CLASS: case mxUINT8_CLASS: return *(unsigned char *)a->data;
    case mxINT8_CLASS: return *(signed char *)a->data;
    case mxINT16_CLASS: return *(short *)a->data;
    case mxUINT16_CLASS: case mxCHAR_CLASS: return *(unsigned short *)a->data;
    case mxINT64_CLASS: return (double)*(long long *)a->data;
    case mxUINT64_CLASS: return (double)*(unsigned long long *)a->data;
    default: return 0.0;
    }
}

static inline double mxGetNaN(void) { return NAN; }
static inline double mxGetInf(void) { return INFINITY; }

/* returns 0 on success, 1 on failure (MATLAB convention) */
static inline int mxGetString(const mxArray *a, char *buf, mwSize buflen)
{
    if (a->classid != mxCHAR_CLASS) return 1;
    size_t n = mxGetNumberOfElements(a);
    if ((size_t)buflen < n + 1) return 1;
    for (size_t i = 0; i < n; i++) buf[i] = (char)((mxChar *)a->data)[i];
    buf[n] = 0;
    return 0;
}

static inline int mxSetDimensions(mxArray *a, const mwSize *dims, int ndim)
{
    a->ndim = ndim < 2 ? 2 : ndim;
    a->dims[0] = a->dims[1] = 1;
    for (int i = 0; i < ndim && i < SB_MX_MAXDIM; i++) a->dims[i] = dims[i];
    return 0;
}

static inline void mxSetData(mxArray *a, void *p)
{
    if (a->owns_data) free(a->data);
    a->data = p;
    a->owns_data = 1;
}

#ifdef __cplusplus
struct sb_mex_error : public std::runtime_error {
    explicit sb_mex_error(const std::string &m) : std::runtime_error(m) {}
};
static inline void mexErrMsgTxt(const char *msg) { throw sb_mex_error(msg ? msg : "mex error"); }
static inline void mexErrMsgIdAndTxt(const char *, const char *msg, ...) { throw sb_mex_error(msg ? msg : "mex error"); }
#else
static inline void mexErrMsgTxt(const char *msg) { fprintf(stderr, "%s\n", msg); abort(); }
#endif
static inline void mexWarnMsgTxt(const char *msg) { fprintf(stderr, "warning: %s\n", msg); }
static inline int mexEvalString(const char *) { return 0; }
#define mexPrintf printf

#ifdef __cplusplus
extern "C" {
#endif
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#ifdef __cplusplus
}
#endif

#endif /* SB_ORACLE_MEX_SHIM_H */
