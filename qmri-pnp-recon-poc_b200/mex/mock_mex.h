/* mock_mex.h - just enough of MATLAB's C MEX API (interleaved-complex, -R2018a) to
 * compile-check qmri_b200_mex.cpp in a container without MATLAB.  NOT a MATLAB header:
 * declarations only, written from the public API documentation; build against the real
 * <mex.h> with `mex -R2018a` on a MATLAB host. */
#ifndef MOCK_MEX_H
#define MOCK_MEX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxDOUBLE_CLASS = 6, mxSINGLE_CLASS = 7, mxINT32_CLASS = 12, mxUINT64_CLASS = 15 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);
int mexCallMATLAB(int nlhs, mxArray* plhs[], int nrhs, mxArray* prhs[], const char* name);
void mexLock(void);
int mexAtExit(void (*fn)(void));
char* mxArrayToString(const mxArray* a);
void mxFree(void* p);
void* mxGetData(const mxArray* a);
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
void mxDestroyArray(mxArray* a);
#ifdef __cplusplus
}
#endif
#endif
