"""ctypes driver of the functional MEX mock (qmri-pnp-recon-poc_b200/mex/mock_mex.{h,cpp}): numpy <-> mxArray, and
``call(command, args..., nlhs=1)`` = what MATLAB does for ``[out1, ...] = qmri_b200_mex(command, args...)``.
Test infrastructure only."""
import ctypes as C
import os

import numpy as np

from conftest import PKG

CLS = {np.dtype(np.float64): 6, np.dtype(np.float32): 7, np.dtype(np.int32): 12, np.dtype(np.uint64): 15,
       np.dtype(np.complex128): 6, np.dtype(np.complex64): 7}
NPTYPE = {6: np.float64, 7: np.float32, 12: np.int32, 15: np.uint64}
FEVAL = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p))


class MexError(RuntimeError):
    def __init__(self, ident, msg):
        super().__init__(f"{ident}: {msg}")
        self.ident = ident


class FunctionHandle:
    def __init__(self, fn):
        self.fn = fn


class MexMock:
    def __init__(self, split_complex=False):
        name = "libqmri_mex_mock_split.so" if split_complex else "libqmri_mex_mock.so"
        self.split = split_complex
        self.lib = L = C.CDLL(os.path.join(PKG, "lib", name))
        vp, sz = C.c_void_p, C.c_size_t
        L.mxCreateNumericArray.restype = vp
        L.mxCreateNumericArray.argtypes = [sz, C.POINTER(sz), C.c_int, C.c_int]
        L.mxGetData.restype = vp
        L.mxGetData.argtypes = [vp]
        L.mxGetImagData.restype = vp
        L.mxGetImagData.argtypes = [vp]
        L.mxCreateString.restype = vp
        L.mxCreateString.argtypes = [C.c_char_p]
        L.mxCreateStructMatrix.restype = vp
        L.mxCreateStructMatrix.argtypes = [sz, sz, C.c_int, vp]
        L.mxSetField.argtypes = [vp, sz, C.c_char_p, vp]
        L.mxCreateCellMatrix.restype = vp
        L.mxCreateCellMatrix.argtypes = [sz, sz]
        L.mxSetCell.argtypes = [vp, sz, vp]
        L.mxDestroyArray.argtypes = [vp]
        L.mxIsComplex.argtypes = [vp]
        L.mxIsDouble.argtypes = [vp]
        L.mxIsSingle.argtypes = [vp]
        L.mxGetNumberOfDimensions.restype = sz
        L.mxGetNumberOfDimensions.argtypes = [vp]
        L.mxGetDimensions.restype = C.POINTER(sz)
        L.mxGetDimensions.argtypes = [vp]
        L.mock_create_function_handle.restype = vp
        L.mock_create_function_handle.argtypes = [FEVAL, vp]
        L.mock_call_mex.argtypes = [C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp)]
        L.mock_last_error.restype = C.c_char_p
        L.mock_last_error_id.restype = C.c_char_p
        self._keep = []

    # ---- numpy -> mxArray ----
    def to_mx(self, v):
        L = self.lib
        if isinstance(v, str):
            return L.mxCreateString(v.encode())
        if isinstance(v, FunctionHandle):
            def cb(user, in_mx, out_pp, _fn=v.fn):
                try:
                    out_pp[0] = self.to_mx(np.asarray(_fn(self.from_mx(in_mx))))
                    return 0
                except Exception:
                    return 1
            c = FEVAL(cb)
            self._keep.append(c)
            return L.mock_create_function_handle(c, None)
        if isinstance(v, dict):
            s = L.mxCreateStructMatrix(1, 1, 0, None)
            for k, x in v.items():
                L.mxSetField(s, 0, k.encode(), self.to_mx(x))
            return s
        if isinstance(v, (list, tuple)):
            c = L.mxCreateCellMatrix(1, len(v))
            for i, x in enumerate(v):
                L.mxSetCell(c, i, self.to_mx(x))
            return c
        a = np.asarray(v)
        if a.dtype not in CLS:            # MATLAB numbers are double unless stated
            a = a.astype(np.complex128 if np.iscomplexobj(a) else np.float64)
        a = np.asfortranarray(a)
        dims = a.shape if a.ndim >= 2 else ((1, 1) if a.ndim == 0 else (a.shape[0], 1))
        arr = (C.c_size_t * len(dims))(*dims)
        cplx = np.iscomplexobj(a)
        m = L.mxCreateNumericArray(len(dims), arr, CLS[a.dtype], 1 if cplx else 0)
        n = a.size
        flat = a.reshape(-1, order="F")
        if not n:
            return m
        if cplx and self.split:
            rt = np.float64 if a.dtype == np.complex128 else np.float32
            re = np.ascontiguousarray(flat.real.astype(rt))
            im = np.ascontiguousarray(flat.imag.astype(rt))
            C.memmove(L.mxGetData(m), re.ctypes.data, re.nbytes)
            C.memmove(L.mxGetImagData(m), im.ctypes.data, im.nbytes)
        else:
            flat = np.ascontiguousarray(flat)
            C.memmove(L.mxGetData(m), flat.ctypes.data, flat.nbytes)
        return m

    # ---- mxArray -> numpy (numeric arrays only) ----
    def from_mx(self, m):
        L = self.lib
        nd = L.mxGetNumberOfDimensions(m)
        dp = L.mxGetDimensions(m)
        dims = tuple(int(dp[i]) for i in range(nd))
        n = int(np.prod(dims))
        cplx = bool(L.mxIsComplex(m))
        rt = np.dtype(NPTYPE[self._class_of(m, dims)])

        def grab(ptr, count):
            if not count:
                return np.zeros(0, rt)
            return np.frombuffer((C.c_char * (count * rt.itemsize)).from_address(ptr), dtype=rt).copy()
        if cplx and not self.split:
            a = grab(L.mxGetData(m), 2 * n).view(np.complex128 if rt == np.float64 else np.complex64)
        elif cplx:
            a = (grab(L.mxGetData(m), n) + 1j * grab(L.mxGetImagData(m), n)).astype(np.complex128 if rt == np.float64 else np.complex64)
        else:
            a = grab(L.mxGetData(m), n)
        return a.reshape(dims, order="F")

    def _class_of(self, m, dims):
        L = self.lib
        if L.mxIsDouble(m):
            return 6
        if L.mxIsSingle(m):
            return 7
        # uint64 handles (1 x 1) and int32 index vectors are the only other outputs of the gateway
        return 15 if all(d == 1 for d in dims) else 12

    def call(self, command, *args, nlhs=1):
        L = self.lib
        prhs = [self.to_mx(command)] + [self.to_mx(a) for a in args]
        rhs = (C.c_void_p * len(prhs))(*prhs)
        lhs = (C.c_void_p * max(nlhs, 4))()
        rc = L.mock_call_mex(nlhs, lhs, len(prhs), rhs)
        try:
            if rc != 0:
                raise MexError(L.mock_last_error_id().decode(), L.mock_last_error().decode())
            outs = [self.from_mx(lhs[i]) for i in range(nlhs) if lhs[i]]
        finally:
            for p in prhs:
                L.mxDestroyArray(p)
            for i in range(len(lhs)):
                if lhs[i]:
                    L.mxDestroyArray(lhs[i])
        if nlhs == 0:
            return None
        return outs[0] if nlhs == 1 else outs
