"""ctypes access to the parity oracles.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; nothing under stereo_b200/ does.

Two oracles sit behind the same function names:

* ``kind == "reference"``: oracle/_ref/libref_*.so, the UNMODIFIED reference
  C++ (cpp/trws_mex.cpp, cpp/rd_mex.cpp, imrender/vgg/vgg_interp2.cxx and the
  vendored TRW-S / QPBO sources) compiled by oracle/Makefile against the fake
  mex.h.  Built in the container that has /root/reference; the .so files
  travel to the GPU box.
* ``kind == "port"``: oracle/_build/libsb_oracle.so, our plain-C restatement
  (oracle/*.c); always buildable.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_int32, c_int64, c_uint32

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
PORT_DIR = os.path.join(HERE, "_build")

_dp = POINTER(c_double)
_up = POINTER(c_uint32)
_ip = POINTER(c_int32)


def _ptr(a, t):
    return a.ctypes.data_as(t)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build(ref: bool = True, port: bool = True) -> None:
    """(Re)build the oracle libraries (make is incremental)."""
    if port:
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_cache: dict = {}


def _load(path):
    if path not in _cache:
        _cache[path] = ctypes.CDLL(path)
    return _cache[path]


def have_ref(name: str = "trws") -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libref_{name}.so"))


def have_port() -> bool:
    return os.path.exists(os.path.join(PORT_DIR, "libsb_oracle.so"))


def _ref(name):
    return _load(os.path.join(REF_DIR, f"libref_{name}.so"))


def _gateway(name):
    """Our own mex gateway (stereo_b200/matlab/<name>_mex.cpp) behind the same fabricated-mxArray
    driver as the reference gateway -- the product called through the MATLAB ABI."""
    path = os.path.join(PORT_DIR, f"libgw_{name}.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", HERE, "gateway"])
    return _load(path)


def _port():
    if not have_port():
        build(ref=False, port=True)
    return _load(os.path.join(PORT_DIR, "libsb_oracle.so"))


# ----------------------------------------------------------------------------
# TRW-S
# ----------------------------------------------------------------------------
_TRWS_ARGS = [c_int, c_int, c_int64, c_int64, _dp, _up, _dp, _dp, _dp, c_double,
              c_double, c_double, _dp, _dp, _dp, _dp]


def trws_solve(kernel, unary, conn, q, qprim, alphas, tol, maxiter=1000, max_relgap=0.0,
               kind="reference"):
    """trws_mex ABI (cpp/trws_mex.cpp:27-163): unary LxN, conn 2xE (0-based),
    q/qprim LxE, alphas E -- all in MATLAB (column-major) layout, i.e. numpy
    arrays of shape (N,L), (E,2), (E,L), (E,).  Returns (labels 1-based int64[N],
    energy, lower_bound, iterations)."""
    unary = _f64(unary)
    q = _f64(q)
    qprim = _f64(qprim)
    alphas = _f64(alphas)
    conn = np.ascontiguousarray(conn, dtype=np.uint32)
    N, L = unary.shape
    E = conn.shape[0]
    assert conn.shape == (E, 2) and q.shape == (E, L) and qprim.shape == (E, L) and alphas.shape == (E,)
    labels = np.zeros(N, dtype=np.float64)
    out = (c_double * 3)()
    if kind == "reference":
        fn = _ref("trws").ref_trws_solve
    elif kind == "gateway":
        fn = _gateway("trws").ref_trws_solve
    else:
        fn = _port().port_trws_solve
    fn.argtypes = _TRWS_ARGS
    fn.restype = c_int
    rc = fn(int(kernel), L, N, E, _ptr(unary, _dp), _ptr(conn, _up), _ptr(q, _dp), _ptr(qprim, _dp),
            _ptr(alphas, _dp), float(tol), float(maxiter), float(max_relgap), _ptr(labels, _dp),
            ctypes.cast(ctypes.byref(out, 0), _dp), ctypes.cast(ctypes.byref(out, 8), _dp),
            ctypes.cast(ctypes.byref(out, 16), _dp))
    if rc != 0:
        msg = ""
        if kind in ("reference", "gateway"):
            lib_ = _ref("trws") if kind == "reference" else _gateway("trws")
            lib_.ref_last_error.restype = ctypes.c_char_p
            msg = lib_.ref_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"oracle trws_solve ({kind}) failed rc={rc}: {msg}")
    return labels.astype(np.int64), out[0], out[1], int(out[2])


def trws_ordering(H, W, kind="reference"):
    """Node ordering (m_ordering) of SetAutomaticOrdering (ordering.cpp:7-157) on
    the H x W grid; returned as an (H, W) int32 array."""
    out = np.zeros(H * W, dtype=np.int32)
    fn = _ref("trws").ref_trws_ordering if kind == "reference" else _port().port_trws_ordering
    fn.argtypes = [c_int, c_int, _ip]
    fn.restype = c_int
    if fn(H, W, _ptr(out, _ip)) != 0:
        raise RuntimeError("oracle ordering failed")
    return out.reshape(W, H).T.copy()


def trws_update_message(kernel, Di, msg, stored0, stored1, alpha, lam, gamma, dir_, swapped,
                        kind="reference"):
    """One Edge::UpdateMessage (typeStereoLinear.h:329-487 /
    typeStereoQuadratic.h:329-501).  Returns (new message, vMin)."""
    Di = _f64(Di)
    msg = _f64(msg).copy()
    s0 = _f64(stored0)
    s1 = _f64(stored1)
    L = Di.shape[0]
    o0 = np.argsort(s0, kind="stable").astype(np.int32)
    o1 = np.argsort(s1, kind="stable").astype(np.int32)
    vmin = c_double()
    fn = _ref("trws").ref_trws_update_message if kind == "reference" else _port().port_trws_update_message
    fn.argtypes = [c_int, c_int, _dp, _dp, _dp, _dp, _ip, _ip, c_double, c_double, c_double, c_int, c_int, _dp]
    fn.restype = c_int
    rc = fn(int(kernel), L, _ptr(Di, _dp), _ptr(msg, _dp), _ptr(s0, _dp), _ptr(s1, _dp), _ptr(o0, _ip),
            _ptr(o1, _ip), float(alpha), float(lam), float(gamma), int(dir_), int(swapped), ctypes.byref(vmin))
    if rc != 0:
        raise RuntimeError("oracle update_message failed")
    return msg, vmin.value


# ----------------------------------------------------------------------------
# QPBO / roof duality
# ----------------------------------------------------------------------------
def rd_solve(U0, U1, E00, E01, E10, E11, conn, improve=False, kind="reference"):
    """rd_mex ABI (cpp/rd_mex.cpp:14-100).  conn is (E,2) 0-based.  Returns
    (labels float64[N] in {0,1,<0}, energy, lower_bound, num_unlabelled)."""
    U0, U1, E00, E01, E10, E11 = (_f64(x).ravel() for x in (U0, U1, E00, E01, E10, E11))
    conn = np.ascontiguousarray(conn, dtype=np.uint32)
    N = U0.shape[0]
    E = conn.shape[0]
    labels = np.zeros(N, dtype=np.float64)
    out = (c_double * 3)()
    fn = _ref("rd").ref_rd_solve if kind == "reference" else _gateway("rd").ref_rd_solve
    fn.argtypes = [c_int64, c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _up, c_int, _dp, _dp, _dp, _dp]
    fn.restype = c_int
    rc = fn(N, E, _ptr(U0, _dp), _ptr(U1, _dp), _ptr(E00, _dp), _ptr(E01, _dp), _ptr(E10, _dp), _ptr(E11, _dp),
            _ptr(conn, _up), int(bool(improve)), _ptr(labels, _dp),
            ctypes.cast(ctypes.byref(out, 0), _dp), ctypes.cast(ctypes.byref(out, 8), _dp),
            ctypes.cast(ctypes.byref(out, 16), _dp))
    if rc != 0:
        raise RuntimeError(f"oracle rd_solve ({kind}) failed rc={rc}")
    return labels, out[0], out[1], out[2]


# ----------------------------------------------------------------------------
# vgg_interp2 (linear)
# ----------------------------------------------------------------------------
def interp2_linear(A, X, Y, oobv, kind="reference"):
    """vgg_interp2(A, X, Y, 'linear', oobv) (vgg_interp2.cxx:246-322).
    A: (h, w, c) array; X, Y: n 1-based coordinates.  Returns (n, c)."""
    A = np.asarray(A, dtype=np.float64)
    if A.ndim == 2:
        A = A[:, :, None]
    h, w, c = A.shape
    Af = np.asfortranarray(A)
    X = _f64(X).ravel()
    Y = _f64(Y).ravel()
    n = X.shape[0]
    out = np.zeros((c, n), dtype=np.float64)
    fn = _ref("interp2").ref_interp2_linear if kind == "reference" else _port().port_interp2_linear
    fn.argtypes = [_dp, c_int, c_int, c_int, _dp, _dp, c_int64, c_double, _dp]
    fn.restype = c_int
    rc = fn(Af.ctypes.data_as(_dp), h, w, c, _ptr(X, _dp), _ptr(Y, _dp), n, float(oobv), _ptr(out, _dp))
    if rc != 0:
        raise RuntimeError("oracle interp2 failed")
    return out.T.copy()


# ----------------------------------------------------------------------------- builders gateway
class _GwArray(ctypes.Structure):
    _fields_ = [("classid", ctypes.c_int), ("ndim", ctypes.c_int), ("dims", ctypes.c_int * 4), ("data", ctypes.c_void_p)]


# mxClassID values of oracle/mex_shim/mex.h
_MX = {np.dtype(np.float64): 6, np.dtype(np.int32): 12, np.dtype(np.uint32): 13}
_MX_BACK = {6: np.float64, 12: np.int32, 13: np.uint32}


def builders_gateway(op, *args, nlhs=1):
    """sb_builders_mex(op, args...) with `nlhs` outputs, through the product's MATLAB gateway compiled against the mex.h
    stand-in (oracle/gw_builders_driver.cpp).  Arguments: numbers (1 x 1 doubles) or NumPy arrays of float64 / int32 /
    uint32 with up to 4 dimensions, handed over in MATLAB (column-major) order.  Returns a list of NumPy arrays (copies),
    MATLAB shapes; raises RuntimeError with the mexErrMsgTxt text."""
    lib_ = _gateway("builders")
    lib_.gw_builders_call.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(_GwArray), ctypes.c_int, ctypes.POINTER(_GwArray)]
    lib_.gw_builders_last_error.restype = ctypes.c_char_p
    keep = []
    arr = (_GwArray * max(len(args), 1))()
    for i, a in enumerate(args):
        a = np.asarray(a)
        if a.dtype not in _MX:
            a = a.astype(np.float64)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        if a.ndim == 1:
            a = a.reshape(1, -1)
        a = np.asfortranarray(a)
        keep.append(a)
        arr[i].classid = _MX[a.dtype]
        arr[i].ndim = a.ndim
        for k in range(4):
            arr[i].dims[k] = a.shape[k] if k < a.ndim else 1
        arr[i].data = a.ctypes.data
    outs = (_GwArray * max(nlhs, 1))()
    rc = lib_.gw_builders_call(op.encode(), len(args), arr, nlhs, outs)
    if rc != 0:
        raise RuntimeError(lib_.gw_builders_last_error().decode("utf-8", "replace"))
    res = []
    for i in range(nlhs):
        o = outs[i]
        shape = tuple(o.dims[k] for k in range(o.ndim))
        n = int(np.prod(shape))
        dt = _MX_BACK[o.classid]
        buf = (ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(o.data) if n else b""
        res.append(np.frombuffer(buf, dtype=dt, count=n).reshape(shape, order="F").copy())
    return res


def grid_gateway_solve(kernel, H, W, proposals, unary, weights, tol, d_min=0.0, d_step=1.0, maxiter=1000, max_relgap=0.0):
    """sb_grid_mex(kernel, sz, proposals, unary, weights, tol, dnorm, options) through the product's grid-native MATLAB gateway
    (stereo_b200/matlab/sb_grid_mex.cpp behind the mex.h stand-in).  proposals: L x 4 x N, unary: L x N, weights: E.
    Returns (labels N, energy, lower_bound, iterations); raises RuntimeError with the mexErrMsgTxt text."""
    lib_ = _gateway("grid")
    proposals = np.asarray(proposals, dtype=np.float64)
    L = proposals.shape[0]
    N = int(H) * int(W)
    assert proposals.shape == (L, 4, N)
    pl = np.ascontiguousarray(proposals.transpose(0, 2, 1))          # label-major, each label a MATLAB 4 x N array
    un = np.ascontiguousarray(np.asarray(unary, dtype=np.float64).reshape(L, N))
    wt = np.ascontiguousarray(np.asarray(weights, dtype=np.float64).reshape(-1))
    labels = np.zeros(N, dtype=np.float64)
    e, lb, it = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    dp = ctypes.POINTER(ctypes.c_double)
    lib_.gw_grid_solve.argtypes = [ctypes.c_int] * 4 + [dp, dp, dp] + [ctypes.c_double] * 5 + [dp, dp, dp, dp]
    lib_.gw_builders_last_error.restype = ctypes.c_char_p
    rc = lib_.gw_grid_solve(int(kernel), int(H), int(W), L, pl.ctypes.data_as(dp), un.ctypes.data_as(dp), wt.ctypes.data_as(dp),
                            float(tol), float(d_min), float(d_step), float(maxiter), float(max_relgap), labels.ctypes.data_as(dp),
                            ctypes.byref(e), ctypes.byref(lb), ctypes.byref(it))
    if rc != 0:
        raise RuntimeError(lib_.gw_builders_last_error().decode("utf-8", "replace"))
    return labels, e.value, lb.value, it.value
