// mock_mex.cpp - functional implementation of mock_mex.h (test infrastructure; see the header).
#include "mock_mex.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

struct mxArray_tag {
    mxClassID cls = mxUNKNOWN_CLASS;
    bool cplx = false;
    std::vector<mwSize> dims;
    std::vector<unsigned char> re, im;  // interleaved model: everything in `re`
    std::vector<std::pair<std::string, mxArray*>> fields;
    std::vector<mxArray*> cells;
    std::string str;
    mock_feval_fn fn = nullptr;
    void* user = nullptr;
};

namespace {
struct MexError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
std::string g_err, g_err_id;
bool g_locked = false;
void (*g_atexit)(void) = nullptr;

size_t elem_size(mxClassID c) {
    switch (c) {
        case mxDOUBLE_CLASS: case mxUINT64_CLASS: return 8;
        case mxSINGLE_CLASS: case mxINT32_CLASS: return 4;
        case mxLOGICAL_CLASS: case mxCHAR_CLASS: return 1;
        default: return 0;
    }
}
size_t numel(const mxArray* a) {
    size_t n = 1;
    for (mwSize d : a->dims) n *= d;
    return n;
}
}  // namespace

extern "C" {

void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    g_err_id = id ? id : "";
    throw MexError(buf);  // MATLAB unwinds out of the MEX function the same way
}
int mexCallMATLAB(int nlhs, mxArray* plhs[], int nrhs, mxArray* prhs[], const char* name) {
    if (strcmp(name, "feval") != 0 || nrhs != 2 || nlhs != 1 || !prhs[0] || prhs[0]->cls != mxFUNCTION_CLASS || !prhs[0]->fn) return 1;
    return prhs[0]->fn(prhs[0]->user, prhs[1], &plhs[0]);
}
void mexLock(void) { g_locked = true; }
int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }
char* mxArrayToString(const mxArray* a) {
    if (!a || a->cls != mxCHAR_CLASS) return nullptr;
    char* s = (char*)malloc(a->str.size() + 1);
    memcpy(s, a->str.c_str(), a->str.size() + 1);
    return s;
}
void mxFree(void* p) { free(p); }
void* mxGetData(const mxArray* a) { return a ? (void*)a->re.data() : nullptr; }
void* mxGetImagData(const mxArray* a) { return (a && a->cplx && !MX_HAS_INTERLEAVED_COMPLEX) ? (void*)a->im.data() : nullptr; }
double* mxGetPr(const mxArray* a) { return (double*)mxGetData(a); }
double* mxGetPi(const mxArray* a) { return (double*)mxGetImagData(a); }
double mxGetScalar(const mxArray* a) {
    if (!a || a->re.empty()) return 0.0;
    switch (a->cls) {
        case mxDOUBLE_CLASS: return *(const double*)a->re.data();
        case mxSINGLE_CLASS: return *(const float*)a->re.data();
        case mxINT32_CLASS: return *(const int32_t*)a->re.data();
        case mxUINT64_CLASS: return (double)*(const uint64_t*)a->re.data();
        case mxLOGICAL_CLASS: return *(const unsigned char*)a->re.data();
        default: return 0.0;
    }
}
int mxIsComplex(const mxArray* a) { return a && a->cplx; }
int mxIsDouble(const mxArray* a) { return a && a->cls == mxDOUBLE_CLASS; }
int mxIsSingle(const mxArray* a) { return a && a->cls == mxSINGLE_CLASS; }
int mxIsClass(const mxArray* a, const char* name) {
    if (!a) return 0;
    if (!strcmp(name, "function_handle")) return a->cls == mxFUNCTION_CLASS;
    if (!strcmp(name, "double")) return a->cls == mxDOUBLE_CLASS;
    if (!strcmp(name, "single")) return a->cls == mxSINGLE_CLASS;
    if (!strcmp(name, "struct")) return a->cls == mxSTRUCT_CLASS;
    if (!strcmp(name, "cell")) return a->cls == mxCELL_CLASS;
    if (!strcmp(name, "char")) return a->cls == mxCHAR_CLASS;
    return 0;
}
mwSize mxGetNumberOfDimensions(const mxArray* a) { return a->dims.size(); }
const mwSize* mxGetDimensions(const mxArray* a) { return a->dims.data(); }
size_t mxGetNumberOfElements(const mxArray* a) { return numel(a); }
mxArray* mxGetField(const mxArray* s, mwSize idx, const char* name) {
    if (!s || s->cls != mxSTRUCT_CLASS || idx != 0) return nullptr;
    for (auto& f : s->fields)
        if (f.first == name) return f.second;
    return nullptr;
}
mxArray* mxGetCell(const mxArray* c, mwSize idx) { return (c && c->cls == mxCELL_CLASS && idx < c->cells.size()) ? c->cells[idx] : nullptr; }
mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID cls, mxComplexity cplx) {
    mxArray* a = new mxArray_tag();
    a->cls = cls;
    a->cplx = cplx == mxCOMPLEX;
    a->dims.assign(dims, dims + ndim);
    while (a->dims.size() < 2) a->dims.push_back(1);
    while (a->dims.size() > 2 && a->dims.back() == 1) a->dims.pop_back();  // MATLAB drops trailing singleton dimensions
    const size_t bytes = numel(a) * elem_size(cls);
#if MX_HAS_INTERLEAVED_COMPLEX
    a->re.assign(bytes * (a->cplx ? 2 : 1), 0);
#else
    a->re.assign(bytes, 0);
    if (a->cplx) a->im.assign(bytes, 0);
#endif
    return a;
}
mxArray* mxCreateDoubleScalar(double v) {
    mwSize d[2] = {1, 1};
    mxArray* a = mxCreateNumericArray(2, d, mxDOUBLE_CLASS, mxREAL);
    *(double*)a->re.data() = v;
    return a;
}
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity cplx) {
    mwSize d[2] = {m, n};
    return mxCreateNumericArray(2, d, cls, cplx);
}
mxArray* mxCreateString(const char* s) {
    mxArray* a = new mxArray_tag();
    a->cls = mxCHAR_CLASS;
    a->str = s;
    a->dims = {1, a->str.size()};
    return a;
}
mxArray* mxCreateStructMatrix(mwSize, mwSize, int nfields, const char** names) {
    mxArray* a = new mxArray_tag();
    a->cls = mxSTRUCT_CLASS;
    a->dims = {1, 1};
    for (int i = 0; i < nfields; ++i) a->fields.emplace_back(names[i], nullptr);
    return a;
}
void mxSetField(mxArray* s, mwSize, const char* name, mxArray* v) {
    for (auto& f : s->fields)
        if (f.first == name) { f.second = v; return; }
    s->fields.emplace_back(name, v);
}
mxArray* mxCreateCellMatrix(mwSize m, mwSize n) {
    mxArray* a = new mxArray_tag();
    a->cls = mxCELL_CLASS;
    a->dims = {m, n};
    a->cells.assign(m * n, nullptr);
    return a;
}
void mxSetCell(mxArray* c, mwSize idx, mxArray* v) { if (idx < c->cells.size()) c->cells[idx] = v; }
void mxDestroyArray(mxArray* a) {
    if (!a) return;
    for (auto& f : a->fields) mxDestroyArray(f.second);
    for (auto* c : a->cells) mxDestroyArray(c);
    delete a;
}

mxArray* mock_create_function_handle(mock_feval_fn fn, void* user) {
    mxArray* a = new mxArray_tag();
    a->cls = mxFUNCTION_CLASS;
    a->dims = {1, 1};
    a->fn = fn;
    a->user = user;
    return a;
}
int mock_call_mex(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    g_err.clear();
    g_err_id.clear();
    try {
        mexFunction(nlhs, plhs, nrhs, prhs);
    } catch (const MexError&) {
        return 1;
    }
    return 0;
}
const char* mock_last_error(void) { return g_err.c_str(); }
const char* mock_last_error_id(void) { return g_err_id.c_str(); }
int mock_is_locked(void) { return g_locked; }
void mock_run_atexit(void) { if (g_atexit) g_atexit(); }
}  // extern "C"
