/* mock_mex.h - a small FUNCTIONAL stand-in for MATLAB's C MEX API, enough to compile AND EXECUTE
 * qmri_b200_mex.cpp in a container without MATLAB / Octave (tests/test_mex_gateway.py drives mexFunction through
 * it and compares with the ctypes path).  NOT a MATLAB header: written from the public API documentation; build
 * against the real <mex.h> on a MATLAB host.
 *
 * Two complex storage models, like the real API:
 *   default                      interleaved complex (mex -R2018a): mxGetData of a complex array = (re, im) pairs
 *   -DQMRI_MOCK_SPLIT_COMPLEX    separate real / imaginary planes (pre-R2018a MEX API, GNU Octave): mxGetPr / mxGetPi
 * MX_HAS_INTERLEAVED_COMPLEX is defined accordingly (the macro the real header provides).
 */
#ifndef MOCK_MEX_H
#define MOCK_MEX_H
#include <stddef.h>
#include <stdint.h>
#ifdef QMRI_MOCK_SPLIT_COMPLEX
#define MX_HAS_INTERLEAVED_COMPLEX 0
#else
#define MX_HAS_INTERLEAVED_COMPLEX 1
#endif
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxUNKNOWN_CLASS = 0, mxCELL_CLASS = 1, mxSTRUCT_CLASS = 2, mxLOGICAL_CLASS = 3, mxCHAR_CLASS = 4, mxDOUBLE_CLASS = 6,
               mxSINGLE_CLASS = 7, mxINT32_CLASS = 12, mxUINT64_CLASS = 15, mxFUNCTION_CLASS = 16 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);
int mexCallMATLAB(int nlhs, mxArray* plhs[], int nrhs, mxArray* prhs[], const char* name);
void mexLock(void);
int mexAtExit(void (*fn)(void));
char* mxArrayToString(const mxArray* a);
void mxFree(void* p);
void* mxGetData(const mxArray* a);
void* mxGetImagData(const mxArray* a); /* split-complex model only (NULL otherwise) */
double* mxGetPr(const mxArray* a);
double* mxGetPi(const mxArray* a);
double mxGetScalar(const mxArray* a);
int mxIsComplex(const mxArray* a);
int mxIsDouble(const mxArray* a);
int mxIsSingle(const mxArray* a);
int mxIsClass(const mxArray* a, const char* name);
mwSize mxGetNumberOfDimensions(const mxArray* a);
const mwSize* mxGetDimensions(const mxArray* a);
size_t mxGetNumberOfElements(const mxArray* a);
mxArray* mxGetField(const mxArray* s, mwSize idx, const char* name);
mxArray* mxGetCell(const mxArray* c, mwSize idx);
mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID cls, mxComplexity cplx);
mxArray* mxCreateDoubleScalar(double v);
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity cplx);
mxArray* mxCreateString(const char* s);
mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char** names);
void mxSetField(mxArray* s, mwSize idx, const char* name, mxArray* v);
mxArray* mxCreateCellMatrix(mwSize m, mwSize n);
void mxSetCell(mxArray* c, mwSize idx, mxArray* v);
void mxDestroyArray(mxArray* a);

/* ---- test-driver side (not part of the MEX API) -------------------------------------------------------------- */
/* a "function handle": feval(handle, in) calls fn(user, in, &out); returns 0 on success */
typedef int (*mock_feval_fn)(void* user, const mxArray* in, mxArray** out);
mxArray* mock_create_function_handle(mock_feval_fn fn, void* user);
/* runs mexFunction; 0 = returned normally, 1 = mexErrMsgIdAndTxt was raised (message via mock_last_error) */
int mock_call_mex(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
const char* mock_last_error(void);
const char* mock_last_error_id(void);
int mock_is_locked(void);
void mock_run_atexit(void);
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
