// qmri_b200_mex.cpp - single MEX gateway between MATLAB and libqmri_b200.so (include/qmri.h).
//
//   out = qmri_b200_mex(command, args...)
//
// Build on a MATLAB host (interleaved complex API):
//   mex -R2018a qmri_b200_mex.cpp -I../../include -L../lib -lqmri_b200
// In this repository's build container there is no MATLAB; the file is compile-checked against
// mex/mock_mex.h by tests/test_mex_gateway.py (g++ -fsyntax-only -DQMRI_MOCK_MEX).
//
// Handles (ctx / op / net / dict) are returned to MATLAB as uint64 scalars and kept alive across calls
// (mexLock); everything is released by `qmri_b200_mex('shutdown')` or at mexAtExit.  prhs arrays are never
// written.  Errors from the library become MATLAB exceptions `qmri:<code>`.
#ifdef QMRI_MOCK_MEX
#include "mock_mex.h"
#else
#include <mex.h>
#endif
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "qmri.h"

namespace {

std::vector<qmri_op*> g_ops;
std::vector<qmri_net*> g_nets;
std::vector<qmri_dict*> g_dicts;
qmri_ctx* g_ctx = nullptr;
bool g_locked = false;

void shutdown() {
    for (auto* o : g_ops) qmri_op_destroy(o);
    for (auto* n : g_nets) qmri_unetres_destroy(n);
    for (auto* d : g_dicts) qmri_dict_destroy(d);
    g_ops.clear(); g_nets.clear(); g_dicts.clear();
    if (g_ctx) qmri_ctx_destroy(g_ctx);
    g_ctx = nullptr;
}

void check(int rc) {
    if (rc != QMRI_OK) mexErrMsgIdAndTxt("qmri:error", "libqmri_b200 error %d: %s", rc, qmri_last_error());
}

qmri_ctx* ctx() {
    if (!g_ctx) {
        check(qmri_ctx_create(&g_ctx, 0));
        if (!g_locked) { mexLock(); mexAtExit(shutdown); g_locked = true; }
    }
    return g_ctx;
}

int dtype_of(const mxArray* a) {
    const bool c = mxIsComplex(a) != 0;
    if (mxIsDouble(a)) return c ? QMRI_C128 : QMRI_F64;
    if (mxIsSingle(a)) return c ? QMRI_C64 : QMRI_F32;
    mexErrMsgIdAndTxt("qmri:type", "arrays must be single or double");
    return -1;
}

mxArray* handle_out(void* p) {
    mxArray* a = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
    *(uint64_t*)mxGetData(a) = (uint64_t)(uintptr_t)p;
    return a;
}
template <typename T>
T* handle_in(const mxArray* a) { return (T*)(uintptr_t)(*(const uint64_t*)mxGetData(a)); }

int slices_of(const mxArray* a, int base_dims) {
    return mxGetNumberOfDimensions(a) > (mwSize)base_dims ? (int)mxGetDimensions(a)[base_dims] : 1;
}

// param.net as a MATLAB function handle: D2H hop per iteration (slow path, same results)
struct FevalCtx { const mxArray* fn; };
int feval_denoiser(void* user, const float* v_in, float* v_out, int H, int W, int Cin, int Cout, int S, int space, void*) {
    if (space != QMRI_HOST) return 2;
    FevalCtx* fc = (FevalCtx*)user;
    for (int s = 0; s < S; ++s) {
        mwSize dims[3] = {(mwSize)H, (mwSize)W, (mwSize)Cin};
        mxArray* in = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);  // PnP_ADMM hands the net a double array
        double* d = (double*)mxGetData(in);
        const float* src = v_in + (size_t)s * H * W * Cin;
        for (size_t i = 0; i < (size_t)H * W * Cin; ++i) d[i] = src[i];
        mxArray* rhs[2] = {const_cast<mxArray*>(fc->fn), in};
        mxArray* lhs[1] = {nullptr};
        if (mexCallMATLAB(1, lhs, 2, rhs, "feval") != 0 || !lhs[0]) { mxDestroyArray(in); return 1; }
        if (mxGetNumberOfElements(lhs[0]) != (size_t)H * W * Cout) { mxDestroyArray(in); mxDestroyArray(lhs[0]); return 3; }
        float* dst = v_out + (size_t)s * H * W * Cout;
        if (mxIsDouble(lhs[0])) { const double* o = (const double*)mxGetData(lhs[0]); for (size_t i = 0; i < (size_t)H * W * Cout; ++i) dst[i] = (float)o[i]; }
        else { const float* o = (const float*)mxGetData(lhs[0]); for (size_t i = 0; i < (size_t)H * W * Cout; ++i) dst[i] = o[i]; }
        mxDestroyArray(in);
        mxDestroyArray(lhs[0]);
    }
    return 0;
}

}  // namespace

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    (void)nlhs;
    if (nrhs < 1) mexErrMsgIdAndTxt("qmri:usage", "qmri_b200_mex(command, ...)");
    char* c = mxArrayToString(prhs[0]);
    std::string cmd = c ? c : "";
    mxFree(c);

    if (cmd == "op_spiral" || cmd == "op_epi") {            // (N, M, S|percentage, V) -> op handle
        const int N = (int)mxGetScalar(prhs[1]), M = (int)mxGetScalar(prhs[2]);
        const mxArray* V = prhs[4];
        const int L = (int)mxGetDimensions(V)[0], C = (int)mxGetDimensions(V)[1];
        if (!mxIsDouble(V) || mxIsComplex(V)) mexErrMsgIdAndTxt("qmri:type", "V must be real double (real(dict.V))");
        qmri_op* op = nullptr;
        if (cmd == "op_spiral") check(qmri_op_spiral(ctx(), N, M, (int)mxGetScalar(prhs[3]), (const double*)mxGetData(V), L, C, &op));
        else check(qmri_op_epi(ctx(), N, M, mxGetScalar(prhs[3]), (const double*)mxGetData(V), L, C, &op));
        g_ops.push_back(op);
        plhs[0] = handle_out(op);
    } else if (cmd == "forward") {                           // (op, x[N M C (S)]) -> y[nmeas (S)]
        qmri_op* op = handle_in<qmri_op>(prhs[1]);
        const int S = slices_of(prhs[2], 3);
        mwSize dims[2] = {(mwSize)qmri_op_nmeas(op), (mwSize)S};
        plhs[0] = mxCreateNumericArray(2, dims, mxDOUBLE_CLASS, mxCOMPLEX);
        check(qmri_forward(op, mxGetData(prhs[2]), dtype_of(prhs[2]), S, mxGetData(plhs[0]), QMRI_C128));
    } else if (cmd == "adjoint") {                           // (op, y, [N M C]) -> x
        qmri_op* op = handle_in<qmri_op>(prhs[1]);
        const int S = slices_of(prhs[2], 1);
        const double* sz = (const double*)mxGetData(prhs[3]);
        mwSize dims[4] = {(mwSize)sz[0], (mwSize)sz[1], (mwSize)sz[2], (mwSize)S};
        plhs[0] = mxCreateNumericArray(4, dims, mxDOUBLE_CLASS, mxCOMPLEX);
        check(qmri_adjoint(op, mxGetData(prhs[2]), dtype_of(prhs[2]), S, mxGetData(plhs[0]), QMRI_C128));
    } else if (cmd == "net_load") {                          // (in_nc, cell{64} of single weight tensors) -> net handle
        std::vector<const float*> w(64);
        for (int i = 0; i < 64; ++i) {
            const mxArray* t = mxGetCell(prhs[2], i);
            if (!t || !mxIsSingle(t)) mexErrMsgIdAndTxt("qmri:type", "weights must be a cell of 64 single arrays (PyTorch memory order)");
            w[i] = (const float*)mxGetData(t);
        }
        qmri_net* net = nullptr;
        check(qmri_unetres_load(ctx(), (int)mxGetScalar(prhs[1]), w.data(), 64, &net));
        g_nets.push_back(net);
        plhs[0] = handle_out(net);
    } else if (cmd == "denoise") {                           // (net, A[H W Cin (S)]) -> [H W 10 (S)], class of A
        qmri_net* net = handle_in<qmri_net>(prhs[1]);
        const mwSize* d = mxGetDimensions(prhs[2]);
        const int S = slices_of(prhs[2], 3);
        mwSize od[4] = {d[0], d[1], 10, (mwSize)S};
        plhs[0] = mxCreateNumericArray(4, od, mxIsDouble(prhs[2]) ? mxDOUBLE_CLASS : mxSINGLE_CLASS, mxREAL);
        check(qmri_unetres_denoise(net, mxGetData(prhs[2]), dtype_of(prhs[2]), mxGetData(plhs[0]), dtype_of(plhs[0]), S, (int)d[0], (int)d[1]));
    } else if (cmd == "pnp_admm") {                          // (op, y, param struct) -> x ; param.net = net handle or function handle
        qmri_op* op = handle_in<qmri_op>(prhs[1]);
        const mxArray* prm = prhs[3];
        const mxArray* X0 = mxGetField(prm, 0, "X0");
        const mxArray* net = mxGetField(prm, 0, "net");
        if (!X0 || !net) mexErrMsgIdAndTxt("qmri:param", "param.X0 and param.net are required");
        qmri_admm_params p;
        memset(&p, 0, sizeof(p));
        p.iters = (int)mxGetScalar(mxGetField(prm, 0, "iter"));
        p.gamma = mxGetScalar(mxGetField(prm, 0, "gamma"));
        const mxArray* tol = mxGetField(prm, 0, "cg_tol");
        p.cg_tol = tol ? mxGetScalar(tol) : 1e-4;
        const mxArray* dt = mxGetField(prm, 0, "denoiser_type");
        char* dts = dt ? mxArrayToString(dt) : nullptr;
        p.multi_level = dts && strcmp(dts, "multi_level") == 0;
        mxFree(dts);
        std::vector<float> nm;
        if (p.multi_level) {
            const mxArray* m = mxGetField(prm, 0, "noise_map");
            if (!m) mexErrMsgIdAndTxt("qmri:param", "param.noise_map is required for multi_level");
            nm.resize(mxGetNumberOfElements(m));
            if (mxIsDouble(m)) { const double* s = (const double*)mxGetData(m); for (size_t i = 0; i < nm.size(); ++i) nm[i] = (float)s[i]; }
            else memcpy(nm.data(), mxGetData(m), nm.size() * 4);
            p.noise_map = nm.data();
        }
        FevalCtx fc = {net};
        if (mxIsClass(net, "function_handle")) { p.fn = feval_denoiser; p.user = &fc; p.fn_space = QMRI_HOST; }
        else p.net = handle_in<qmri_net>(net);
        const int S = slices_of(X0, 3);
        const mwSize* d = mxGetDimensions(X0);
        mwSize od[4] = {d[0], d[1], d[2], (mwSize)S};
        plhs[0] = mxCreateNumericArray(4, od, mxDOUBLE_CLASS, mxCOMPLEX);
        check(qmri_pnp_admm(op, mxGetData(prhs[2]), dtype_of(prhs[2]), mxGetData(X0), dtype_of(X0), S, &p, mxGetData(plhs[0]), QMRI_C128));
    } else if (cmd == "dict_load") {                         // (D single KxC, normD single, lut single KxQ) -> dict handle
        const mwSize* d = mxGetDimensions(prhs[1]);
        const int Q = (int)mxGetDimensions(prhs[3])[1];
        if (!mxIsSingle(prhs[1]) || !mxIsSingle(prhs[2]) || !mxIsSingle(prhs[3])) mexErrMsgIdAndTxt("qmri:type", "pass single(dict.D), single(dict.normD), single(dict.lut)");
        qmri_dict* dd = nullptr;
        check(qmri_dict_load(ctx(), (const float*)mxGetData(prhs[1]), (const float*)mxGetData(prhs[2]), (const float*)mxGetData(prhs[3]),
                             (int64_t)d[0], (int)d[1], Q, 0, (int64_t)d[0], &dd));
        g_dicts.push_back(dd);
        plhs[0] = handle_out(dd);
    } else if (cmd == "match") {                             // (dict, x[npix x C], Q) -> qmap, pd, mt, dm
        qmri_dict* dd = handle_in<qmri_dict>(prhs[1]);
        const int64_t npix = (int64_t)mxGetDimensions(prhs[2])[0];
        const int Q = (int)mxGetScalar(prhs[3]);
        plhs[0] = mxCreateNumericMatrix((mwSize)npix, (mwSize)Q, mxSINGLE_CLASS, mxREAL);
        plhs[1] = mxCreateNumericMatrix((mwSize)npix, 1, mxSINGLE_CLASS, mxCOMPLEX);
        plhs[2] = mxCreateNumericMatrix((mwSize)npix, 1, mxSINGLE_CLASS, mxREAL);
        plhs[3] = mxCreateNumericMatrix((mwSize)npix, 1, mxINT32_CLASS, mxREAL);
        check(qmri_match(dd, mxGetData(prhs[2]), dtype_of(prhs[2]), npix, (float*)mxGetData(plhs[0]), (float*)mxGetData(plhs[1]),
                         (float*)mxGetData(plhs[2]), (int32_t*)mxGetData(plhs[3])));
    } else if (cmd == "shutdown") {
        shutdown();
    } else {
        mexErrMsgIdAndTxt("qmri:usage", "unknown command '%s'", cmd.c_str());
    }
}
