"""Device pipeline of the SIMCA hot path: thin Python over the C ABI (include/ocm_simca.h).

torch is used for device memory, streams, the random start block of the subspace iteration and (multi-GPU)
torch.distributed; the numerical steps are calls into libocm_simca_b200.so.  One library call is left on a fallback
path: `eigh_topk` (torch.linalg.eigh, cuSOLVER) for a covariance whose subspace iteration does not converge (flat
spectrum at n_components) or a fit that asks for `full_spectrum=True` WITH vectors.  Nothing here falls back to the CPU.

    moments -> covariance -> eigh -> (mean, loadings, eigenvalues)      fit   (SIMCA.py:62-75)
    fused score kernel -> T2, Q, accept                                 predict (SIMCA.py:120-154)
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import native
from .limits import RULES

_PRE = {"snv": 1, "center": 2, "autoscale": 3, "snv_center": 4}


def _torch():
    return native.require_cuda()


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    """torch's current stream of the current device as the `void* stream` of the C ABI.  The raw getter costs ~1 us;
    `torch.cuda.current_stream().cuda_stream` builds a Stream object per call (~19 us x ~160 calls in one CV grid)."""
    torch = _torch()
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:
        return C.c_void_p(raw(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


UPLOAD_BOUNCE_BYTES = 64 << 20      # numpy matrices above this go through ocm_upload_rows


def to_device(X, device=None, dtype=None):
    """numpy / torch (any device) -> 2-D row-contiguous CUDA tensor, fp32 or fp64."""
    torch = _torch()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if isinstance(X, torch.Tensor):
        t = X
    else:
        a = np.asarray(X)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        if (a.ndim == 2 and a.nbytes > UPLOAD_BOUNCE_BYTES and a.shape[1] > 0 and a.strides[1] == a.itemsize
                and a.strides[0] >= a.shape[1] * a.itemsize and dtype in (None, torch.from_numpy(a[:0]).dtype)):
            # large pageable matrix: parallel bounce fill + double-buffered copies (ocm_upload_rows), ~3x torch's .to()
            with torch.cuda.device(device):
                torch.cuda.current_stream().synchronize()      # the copies run on the library's streams: nothing of ours may still use a recycled block
                t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, device=device)
                native.check(native.lib().ocm_upload_rows(C.c_void_p(a.ctypes.data), a.shape[0], a.shape[1] * a.itemsize,
                                                          a.strides[0], _p(t), t.stride(0) * a.itemsize), "ocm_upload_rows")
            return t
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    if dtype is not None:
        t = t.to(dtype)
    t = t.to(device, non_blocking=False)
    if t.dim() != 2:
        raise ValueError("expected a 2-D array [n_samples, n_features]")
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def index_tensor(idx, device):
    torch = _torch()
    if idx is None:
        return None
    if isinstance(idx, torch.Tensor):
        return idx.to(device=device, dtype=torch.int64).contiguous()
    return torch.as_tensor(np.asarray(idx, dtype=np.int64), device=device)


# ------------------------------------------------------------------------------------------
# preprocessing
# ------------------------------------------------------------------------------------------
def preprocess(X, mode, mu=None, sigma=None, eps=1e-8, out=None):
    """SNV / centre / autoscale / SNV+centre on fp32 device rows (ocm_preprocess_f32)."""
    torch = _torch()
    X = to_device(X, dtype=torch.float32)
    if out is None:
        out = torch.empty_like(X)
    mu_t = None if mu is None else to_device(torch.as_tensor(mu).reshape(1, -1), X.device, torch.float32)
    sg_t = None if sigma is None else to_device(torch.as_tensor(sigma).reshape(1, -1), X.device, torch.float32)
    rc = native.lib().ocm_preprocess_f32(_p(X), X.shape[0], X.shape[1], X.stride(0), _PRE[mode], _p(mu_t),
                                         _p(sg_t), float(eps), _p(out), out.stride(0), _stream())
    native.check(rc, "ocm_preprocess_f32")
    return out


def snv(X, eps=1e-8):
    return preprocess(X, "snv", eps=eps)


_SG_CACHE = {}
_SG_DEVICE = {}      # (window, order, deriv, delta, device) -> (host taps, device edge matrices)


def _savgol_tables(window, polyorder, deriv, delta):
    """Interior taps and the two edge matrices of scipy's savgol_filter(mode='interp'), obtained by pushing unit
    impulses through scipy itself, so the device filter is the same linear map by construction."""
    key = (window, polyorder, deriv, float(delta))
    if key not in _SG_CACHE:
        from scipy.signal import savgol_filter
        h = window // 2
        L = 3 * window
        resp = savgol_filter(np.eye(L), window, polyorder, deriv=deriv, delta=delta, axis=1)   # row j: response to impulse at j
        c = L // 2
        g = np.array([resp[c + d - h, c] for d in range(window)])                 # y[c] = sum_d g[d] x[c + d - h]
        edge_l = resp[:window, :h].T.copy()                                       # y[i<h] from x[0:window]
        edge_r = resp[L - window:, L - h:].T.copy()                               # y[L-h+i] from x[L-window:L]
        _SG_CACHE[key] = (g.astype(np.float32), edge_l.astype(np.float32), edge_r.astype(np.float32))
    return _SG_CACHE[key]


def savgol(X, window_length=5, polyorder=2, deriv=1, delta=1.0, snv=False, snv_eps=1e-8, out=None):
    """[SNV +] Savitzky-Golay filter of fp32 device rows in one pass (ocm_savgol_f32)."""
    torch = _torch()
    X = to_device(X, dtype=torch.float32)
    if out is None:
        out = torch.empty_like(X)
    key = (int(window_length), int(polyorder), int(deriv), float(delta), str(X.device))
    if key not in _SG_DEVICE:
        g, el, er = _savgol_tables(int(window_length), int(polyorder), int(deriv), delta)
        _SG_DEVICE[key] = (g, torch.as_tensor(el, device=X.device).contiguous(), torch.as_tensor(er, device=X.device).contiguous())
    g, el_d, er_d = _SG_DEVICE[key]
    rc = native.lib().ocm_savgol_f32(_p(X), X.shape[0], X.shape[1], X.stride(0), int(bool(snv)), float(snv_eps),
                                     g.ctypes.data_as(C.c_void_p), int(window_length), _p(el_d), _p(er_d), _p(out),
                                     out.stride(0), _stream())
    native.check(rc, "ocm_savgol_f32")
    return out


def mahalanobis_outlier_mask(X, n_components=10, percentile=95.0):
    """PCA score-space Mahalanobis outlier filter (simca_nuts.py:127-147, utils/data_utils.py:65-80):
    fit PCA(n_components) on X, d_i = sqrt(t_i^T cov(T)^-1 t_i) (= sqrt of the Hotelling T2 the scoring kernel returns,
    because cov(T) = diag(lambda)), keep rows with d_i <= percentile(d).  Returns (keep mask [n] bool on the device,
    distances [n] fp32, threshold)."""
    torch = _torch()
    X = to_device(X)
    n, p = X.shape
    k = min(int(n_components), p, max(1, n - 1))
    pca = fit_pca(X, None, k)
    Sinv = torch.diag(torch.as_tensor(1.0 / pca.eigs_all[:k], device=X.device))
    T2, _, _ = score(X, pca.mean, pca.P, Sinv)
    d = torch.sqrt(T2.clamp_min(0).double())
    thr = torch.quantile(d, float(percentile) / 100.0, interpolation="linear")
    return d <= thr, d, float(thr.item())


# ------------------------------------------------------------------------------------------
# PCA fit: moments -> covariance -> eigh
# ------------------------------------------------------------------------------------------
def colstats(X, row_idx=None, shift=None, want_sq=False):
    torch = _torch()
    n = X.shape[0] if row_idx is None else row_idx.numel()
    p = X.shape[1]
    s = torch.zeros(p, dtype=torch.float64, device=X.device)
    ss = torch.zeros(p, dtype=torch.float64, device=X.device) if want_sq else None
    fn = native.lib().ocm_colstats_f32 if X.dtype == torch.float32 else native.lib().ocm_colstats_f64
    native.check(fn(_p(X), X.stride(0), _p(row_idx), n, p, _p(shift), _p(s), _p(ss), _stream()), "ocm_colstats")
    return s, ss, n


def colstats_minmax(X, row_idx=None):
    """One pass: column sums (fp64) and column extrema (fp32, rounded outwards) of the selected rows."""
    torch = _torch()
    n = X.shape[0] if row_idx is None else row_idx.numel()
    p = X.shape[1]
    s = torch.zeros(p, dtype=torch.float64, device=X.device)
    mn = torch.full((p,), float("inf"), dtype=torch.float32, device=X.device)
    mx = torch.full((p,), float("-inf"), dtype=torch.float32, device=X.device)
    fn = native.lib().ocm_colstats_minmax_f32 if X.dtype == torch.float32 else native.lib().ocm_colstats_minmax_f64
    native.check(fn(_p(X), X.stride(0), _p(row_idx), n, p, _p(s), _p(mn), _p(mx), _stream()), "ocm_colstats_minmax")
    return s, mn, mx, n


_GRAM_TC_WS = {}     # device index -> uint8 tensor (grown on demand)


def _gram_tc_workspace(n, p, device):
    torch = _torch()
    need = C.c_int64(0)
    native.check(native.lib().ocm_gram_tc_workspace_bytes(int(n), int(p), C.byref(need)), "ocm_gram_tc_workspace_bytes")
    key = device.index if device.index is not None else torch.cuda.current_device()
    ws = _GRAM_TC_WS.get(key)
    if ws is None or ws.numel() < need.value + 1024:
        _GRAM_TC_WS[key] = ws = torch.empty(need.value + 1024, dtype=torch.uint8, device=device)
    off = (-ws.data_ptr()) % 1024
    return ws, off, need.value


def gram_tc(X, row_idx, shift, colmin, colmax, products=13, out=None, colsum=None):
    """Upper triangle of the Gram of the quantised shifted rows on the integer tensor cores (ocm_gram_tc_*), plus the
    exact column sums of the same quantised rows.  Returns (G [p,p] f64, colsum [p] f64, flags [2] int64 on the device:
    clamped values, non-finite inputs)."""
    torch = _torch()
    n = X.shape[0] if row_idx is None else row_idx.numel()
    p = X.shape[1]
    if out is None:
        out = torch.zeros((p, p), dtype=torch.float64, device=X.device)
    if colsum is None:
        colsum = torch.zeros(p, dtype=torch.float64, device=X.device)
    flags = torch.zeros(2, dtype=torch.int64, device=X.device)
    ws, off, nbytes = _gram_tc_workspace(n, p, X.device)
    fn = native.lib().ocm_gram_tc_f32 if X.dtype == torch.float32 else native.lib().ocm_gram_tc_f64
    rc = fn(_p(X), X.stride(0), _p(row_idx), n, p, _p(shift), _p(colmin), _p(colmax), int(products),
            C.c_void_p(ws.data_ptr() + off), nbytes, _p(out), _p(colsum), _p(flags), _stream())
    native.check(rc, "ocm_gram_tc")
    return out, colsum, flags


def gram(X, row_idx=None, shift=None, out=None):
    """Upper triangle of sum_r (x_r - shift)(x_r - shift)^T, fp64 accumulation."""
    torch = _torch()
    n = X.shape[0] if row_idx is None else row_idx.numel()
    p = X.shape[1]
    if out is None:
        out = torch.zeros((p, p), dtype=torch.float64, device=X.device)
    fn = native.lib().ocm_gram_f64acc_f32 if X.dtype == torch.float32 else native.lib().ocm_gram_f64acc_f64
    native.check(fn(_p(X), X.stride(0), _p(row_idx), n, p, _p(shift), _p(out), _stream()), "ocm_gram_f64acc")
    return out


def cov_from_parts(Gparts, sparts, counts, weights, shift):
    """Combine partition moments into (C [p,p] full symmetric, mean [p]) -- ocm_cov_from_parts_f64."""
    torch = _torch()
    ng, p = sparts.shape
    Cm = torch.empty((p, p), dtype=torch.float64, device=sparts.device)
    mean = torch.empty(p, dtype=torch.float64, device=sparts.device)
    delta = torch.empty(p, dtype=torch.float64, device=sparts.device)
    cnt = (C.c_int64 * ng)(*[int(c) for c in counts])
    wts = (C.c_double * ng)(*[float(w) for w in weights])
    rc = native.lib().ocm_cov_from_parts_f64(_p(Gparts), _p(sparts), cnt, wts, ng, p, _p(shift), _p(Cm),
                                             _p(mean), _p(delta), _stream())
    native.check(rc, "ocm_cov_from_parts_f64")
    return Cm, mean


class PhaseTimer:
    """CUDA-event marks between the phases of a fit (bench.py's `fit` leg): mark(name) closes the phase `name`."""

    def __init__(self):
        torch = _torch()
        self._ev = [("", torch.cuda.Event(enable_timing=True))]
        self._ev[0][1].record()

    def mark(self, name):
        torch = _torch()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self._ev.append((name, e))

    def result(self):
        torch = _torch()
        torch.cuda.synchronize()
        out = {}
        for (_, a), (name, b) in zip(self._ev[:-1], self._ev[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out


GRAM_TC_MIN_ROWS = 16384     # below this the FP64 tensor-pipe (DMMA) kernel is used: CV folds, configs 1 / 2
GRAM_TC_PRODUCTS = 13        # digit products of the integer Gram: 16 exact / 13 (<= 2^-37) / 10 (<= 2^-28)


def covariance(X, row_idx=None, group=None, gram_path="auto", products=None, timer=None):
    """Sample covariance (ddof=1) and column mean of the selected rows.

    Column sums give a provisional mean that is used as the shift of the Gram accumulation, so G - n*d*d^T never
    cancels.  Bulk fits (>= GRAM_TC_MIN_ROWS rows, `gram_path` "auto" or "tc") take the integer tcgen05
    Gram (csrc/gram_tc.cu): one pass for sums / extrema, one pass that quantises and accumulates the exact column sums,
    then int8 MMAs; small fits take the FP64 tensor-pipe kernel ("dmma").  With torch.distributed initialised and
    `group` given, the partial moments of the row shards are all-reduced (NCCL over NVLink) before the combination."""
    torch = _torch()
    n_rows = X.shape[0] if row_idx is None else row_idx.numel()
    use_tc = gram_path == "tc" or (gram_path == "auto" and n_rows >= GRAM_TC_MIN_ROWS)
    if group is not None:
        import torch.distributed as dist
    mark = timer.mark if timer is not None else (lambda name: None)
    if use_tc:
        s0, mn, mx, n = colstats_minmax(X, row_idx)
    else:
        s0, _, n = colstats(X, row_idx)
    mark("colstats")
    n_tot = torch.tensor([float(n)], dtype=torch.float64, device=X.device)
    if group is not None:
        dist.all_reduce(s0, group=group)
        dist.all_reduce(n_tot, group=group)
        mark("allreduce")
    shift = s0 / n_tot
    if use_tc:
        G, s1, flags = gram_tc(X, row_idx, shift, mn, mx, products=products or GRAM_TC_PRODUCTS)
    else:
        G = gram(X, row_idx, shift)
        s1, _, _ = colstats(X, row_idx, shift)
        flags = None
    mark("gram")
    if group is not None:
        dist.all_reduce(G, group=group)
        dist.all_reduce(s1, group=group)
        if flags is not None:
            dist.all_reduce(flags, group=group)
        mark("allreduce")
    N = int(round(float(n_tot.item())))
    if flags is not None:
        clamped, nonfinite = (int(v) for v in flags.cpu().tolist())
        if clamped:
            raise native.NativeError(f"integer Gram: {clamped} values exceeded the column bounds (internal error)")
        if nonfinite:
            G.fill_(float("nan"))          # what an IEEE accumulation of the same rows would have produced
    Cm, mean = cov_from_parts(G.unsqueeze(0), s1.unsqueeze(0), [N], [1.0], shift)
    mark("covariance")
    return Cm, mean, N


def sign_normalise_(V):
    """sklearn svd_flip(u_based_decision=False): the largest-|entry| of each row positive."""
    torch = _torch()
    idx = V.abs().argmax(dim=-1, keepdim=True)           # works on [k, p] and on a batch [F, k, p] alike
    sgn = torch.sign(torch.gather(V, -1, idx))
    sgn = torch.where(sgn == 0, torch.ones_like(sgn), sgn)
    V.mul_(sgn)
    return V


def eigh_topk(Cm, k):
    """All eigenvalues (descending) and the k leading eigenvectors (rows) of a symmetric fp64 matrix, on the device
    (cuSOLVER through torch.linalg.eigh).  Only the full-spectrum path uses it: a fit that asks for `full_spectrum`,
    a subspace iteration that did not converge; the default fit and the CV folds use eig_topk_subspace."""
    torch = _torch()
    native.require_cuda()
    w, V = torch.linalg.eigh(Cm)
    w = torch.flip(w, dims=[-1])
    V = torch.flip(V, dims=[-1])[..., :k].transpose(-1, -2).contiguous()
    if V.dim() == 2:
        sign_normalise_(V)
    else:
        for b in range(V.shape[0]):
            sign_normalise_(V[b])
    return w, V


def eigh_small(H):
    """Batched eigen-decomposition of small symmetric fp64 matrices [B, m, m], m <= 64 (ocm_eigh_small_f64: parallel
    Jacobi, one CTA per matrix).  Returns (w [B, m] descending, U [B, m, m] eigenvectors as columns)."""
    torch = _torch()
    H = H.contiguous()
    B, m, _ = H.shape
    w = torch.empty((B, m), dtype=torch.float64, device=H.device)
    U = torch.empty((B, m, m), dtype=torch.float64, device=H.device)
    native.check(native.lib().ocm_eigh_small_f64(_p(H), m, B, _p(w), _p(U), _stream()), "ocm_eigh_small_f64")
    return w, U


def _symm_mul(Cs, Q):
    """Z = C Q for a batch of symmetric fp64 matrices (ocm_eig_symm_mul_f64: FP64 tensor pipe)."""
    torch = _torch()
    F, p, kb = Q.shape
    Z = torch.empty_like(Q)
    native.check(native.lib().ocm_eig_symm_mul_f64(_p(Cs), _p(Q), _p(Z), p, kb, F, _stream()), "ocm_eig_symm_mul_f64")
    return Z


def _tn_small(A, B):
    """H = A^T B, [F, ka, kb] (ocm_eig_tn_small_f64)."""
    torch = _torch()
    F, p, ka = A.shape
    kb = B.shape[2]
    H = torch.empty((F, ka, kb), dtype=torch.float64, device=A.device)
    native.check(native.lib().ocm_eig_tn_small_f64(_p(A), _p(B), _p(H), p, ka, kb, F, _stream()), "ocm_eig_tn_small_f64")
    return H


def _mul_small(Z, U):
    """Y = Z U, [F, p, kb] x [F, kb, ku] (ocm_eig_mul_small_f64)."""
    torch = _torch()
    F, p, kb = Z.shape
    ku = U.shape[2]
    Y = torch.empty((F, p, ku), dtype=torch.float64, device=Z.device)
    native.check(native.lib().ocm_eig_mul_small_f64(_p(Z), _p(U.contiguous()), _p(Y), p, kb, ku, F, _stream()), "ocm_eig_mul_small_f64")
    return Y


def _orthonormalise(Z, fail):
    """Batched tall-skinny orthonormalisation [F, p, kb] by Cholesky-QR applied twice, in the library's own kernels: two
    small Gram products (ocm_eig_tn_small_f64), two in-CTA kb x kb Cholesky factorisations and row-wise triangular solves
    (ocm_eig_chol_solve_f64).  CholQR2 is orthogonal to machine precision while cond(Z)^2 * eps < 1; a factorisation that
    meets a non-positive pivot (rank-deficient block) is recorded in `fail` [F] (int32, accumulated on the device, read by
    the caller at its next synchronisation) and that matrix goes to the full solver."""
    torch = _torch()
    F, p, kb = Z.shape
    Y = Z.contiguous().clone()
    L = torch.empty((F, kb, kb), dtype=torch.float64, device=Z.device)
    info = torch.empty(F, dtype=torch.int32, device=Z.device)
    for _ in range(2):
        G = _tn_small(Y, Y)
        native.check(native.lib().ocm_eig_chol_solve_f64(_p(G), _p(Y), _p(L), _p(info), p, kb, F, _stream()), "ocm_eig_chol_solve_f64")
        fail |= info != 0
    return Y


def _eig_cycle(Cs, Q, fail, res2, fail_ritz=None):
    """One cycle of the block subspace iteration: two plain power steps (multiply, orthonormalise), then a Rayleigh-Ritz
    step (multiply, project, small eigensolve, rotate, residuals) and the orthonormalised block the next cycle starts
    from.  Returns (w [F, kb], X [F, p, kb], Q_next); residual norms squared land in res2, rank losses in fail
    (fail_ritz: the flags as they stood when the Ritz values were formed)."""
    lib = native.lib()
    F, p, kb = Q.shape
    for _ in range(2):
        Q = _orthonormalise(_symm_mul(Cs, Q), fail)
    Z = _symm_mul(Cs, Q)
    H = _tn_small(Q, Z)
    w, U = eigh_small(H)                           # hand-written batched Jacobi (descending)
    ZU = _mul_small(Z, U)
    X = _mul_small(Q, U)
    native.check(lib.ocm_eig_ritz_residual_f64(_p(ZU), _p(X), _p(w), _p(res2), p, kb, F, _stream()), "ocm_eig_ritz_residual_f64")
    if fail_ritz is not None:
        fail_ritz.copy_(fail)
    return w, X, _orthonormalise(ZU, fail)


# Opt-in (OCM_EIG_GRAPH=1): the cycle (26 launches) captured once per (device, batch, p, block) into a CUDA graph over
# static buffers and replayed, the convergence test between cycles staying on the host.  Same kernels in the same order on
# the same data (results agree to the last few ulps: the small Gram products accumulate with fp64 atomics).  Measured on
# the CV grid's solve (11 covariances of 1050 x 1050, three cycles): 5.5 ms against 5.8 ms eager -- the solve is bound by
# its kernels (C Q products, the 32 x 32 Jacobi sweeps, the deflated traces), not by their launches -- while the capture
# costs 550 ms on first use and the static buffers hold a copy of the batch.  Hence off by default.
_EIG_GRAPHS = {}
EIG_GRAPH = os.environ.get("OCM_EIG_GRAPH", "0") == "1"
graph_replayed_launches = 0       # kernel launches issued through graph replays (native.launch_count() sees captures only)


def _eig_graph(F, p, kb, device):
    torch = _torch()
    key = (device.index if device.index is not None else torch.cuda.current_device(), F, p, kb)
    if key in _EIG_GRAPHS:
        return _EIG_GRAPHS[key]
    from types import SimpleNamespace
    g = None
    try:
        with torch.cuda.device(device):
            st = SimpleNamespace()
            st.Cs = torch.eye(p, dtype=torch.float64, device=device).repeat(F, 1, 1).contiguous()
            st.Q = torch.zeros((F, p, kb), dtype=torch.float64, device=device)
            st.Q[:, :kb, :] = torch.eye(kb, dtype=torch.float64, device=device)
            st.fail = torch.zeros(F, dtype=torch.bool, device=device)
            st.fail_ritz = torch.zeros(F, dtype=torch.bool, device=device)
            st.res2 = torch.empty((F, kb), dtype=torch.float64, device=device)
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):              # one eager cycle first: one-time attribute calls happen outside the capture
                _eig_cycle(st.Cs, st.Q, st.fail, st.res2, st.fail_ritz)
            torch.cuda.current_stream().wait_stream(side)
            l0 = native.launch_count()
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                st.w, st.X, Qn = _eig_cycle(st.Cs, st.Q, st.fail, st.res2, st.fail_ritz)
                st.Q.copy_(Qn)
            st.nodes = native.launch_count() - l0
            st.graph = cg
            g = st
    except Exception:                                  # no graphs here (capture refused): the eager loop runs instead
        torch.cuda.synchronize()
        g = None
    if len(_EIG_GRAPHS) >= 6:
        _EIG_GRAPHS.pop(next(iter(_EIG_GRAPHS)))
    _EIG_GRAPHS[key] = g
    return g


def eig_topk_subspace(Cs, k_need, tol=1e-11, max_iter=60, seed=0):
    """Leading eigenpairs of a BATCH of symmetric fp64 matrices by block subspace iteration with Rayleigh-Ritz,
    plus the power sums of the remaining spectrum from traces of the deflated matrices.

    This is what the CV grid needs per fold (loadings up to A_max, eigenvalues up to A_max, theta_1..3 of the rest for
    the Jackson-Mudholkar / chi2box / ci limits) without a full eigensolve per fold -- ten 1050 x 1050 fp64 `eigh`
    calls were 90 % of the grid's wall time.  Every numerical step is a kernel of this library (csrc/eig_sub.cu,
    csrc/eig_small.cu, csrc/moments.cu); torch only allocates and draws the random start block.  One batched product
    C Q per iteration, two plain power steps per Rayleigh-Ritz step (`_eig_cycle`); a matrix
    that has not converged after `max_iter` iterations (flat spectrum at A_max) or whose block lost rank is flagged so
    the caller can fall back to the full solver.

    Cs [F, p, p].  Returns (w [F, kb] descending Ritz values, X [F, p, kb] Ritz vectors, tails [F, 3] =
    sums of lambda, lambda^2, lambda^3 beyond the first k_need, converged [F] bool)."""
    global graph_replayed_launches
    torch = _torch()
    lib = native.lib()
    Cs = Cs.contiguous()
    F, p, _ = Cs.shape
    kb = min(p, max(k_need + 12, 32), 64)
    if k_need > kb:
        raise ValueError("at most 64 leading eigenpairs are supported by the subspace solver")
    g = torch.Generator(device=Cs.device)
    g.manual_seed(seed)
    fail = torch.zeros(F, dtype=torch.bool, device=Cs.device)
    Q = _orthonormalise(torch.randn((F, p, kb), generator=g, dtype=torch.float64, device=Cs.device), fail)
    conv = torch.zeros(F, dtype=torch.bool, device=Cs.device)
    res2 = torch.empty((F, kb), dtype=torch.float64, device=Cs.device)
    w = X = None
    cycles = max(1, (max_iter + 2) // 3)
    eg = _eig_graph(F, p, kb, Cs.device) if EIG_GRAPH and Cs.is_cuda else None
    if eg is not None:
        eg.Cs.copy_(Cs); eg.Q.copy_(Q); eg.fail.copy_(fail)
        for _ in range(cycles):
            eg.graph.replay()
            graph_replayed_launches += eg.nodes
            conv = (eg.res2.sqrt()[:, :k_need].max(dim=1).values <= tol * eg.w[:, 0]) & ~eg.fail_ritz
            if bool((conv | eg.fail_ritz).all()):
                break
        w, X, fail = eg.w.clone(), eg.X.clone(), eg.fail_ritz.clone()
    else:
        for _ in range(cycles):
            fail_ritz = torch.empty_like(fail)
            w, X, Qn = _eig_cycle(Cs, Q, fail, res2, fail_ritz)
            conv = (res2.sqrt()[:, :k_need].max(dim=1).values <= tol * w[:, 0]) & ~fail_ritz
            if bool((conv | fail_ritz).all()):
                fail = fail_ritz
                break
            Q = Qn
    conv = conv & ~fail
    # tail power sums from the explicitly deflated matrix (no cancellation): R = C - X_k diag(w_k) X_k^T,
    # tr R, |R|_F^2, tr R^3 = sum_ij (R^2)_ij R_ij with R^2 = R^T R from the FP64 tensor-pipe Gram kernel
    R = torch.empty_like(Cs)
    R2 = torch.empty_like(Cs)
    tails = torch.empty((F, 3), dtype=torch.float64, device=Cs.device)
    native.check(lib.ocm_eig_tail_sums_f64(_p(Cs), _p(X), _p(w), p, kb, int(k_need), F, _p(R), _p(R2), _p(tails), _stream()),
                 "ocm_eig_tail_sums_f64")
    return w, X, tails, conv


@dataclass
class PCAModel:
    """Device-resident PCA of one class: everything predict needs, plus what the reference
    exposes under `_model[cls]` (SIMCA.py:83-99)."""
    mean: "object"        # torch f64 [p]
    P: "object"           # torch f64 [A, p]  (rows sign-normalised)
    eigs_all: np.ndarray  # host f64 [min(n, p)]  explained variances, descending
    n_samples: int
    x_rms: float = 0.0    # sqrt(trace(cov) / p): size of one centred element (ocm_score_params.x_rms; 0 = unknown)

    def arrays32(self):
        torch = _torch()
        return self.mean.to(torch.float32), self.P.to(torch.float32).contiguous()


def _fold_covariances(X, row_sets):
    """Covariance + mean + count of every row subset.  When the subsets are 'everything but a held-out part' of a common
    pool (k-fold training sets), the Gram is DOWNDATED: one Gram + column sums of the pool, one small Gram per held-out
    part, combined with weights (+1, -1) by ocm_cov_from_parts_f64 around the pool mean -- instead of
    one full-size Gram per fold.  Anything else falls back to one covariance per subset."""
    torch = _torch()
    sets_h = [s.cpu().numpy() for s in row_sets]
    # membership through flag arrays (np.unique / setdiff1d on ten index sets cost 5 ms of a 23 ms grid)
    n_all = X.shape[0]
    in_pool = np.zeros(n_all, bool)
    distinct = True
    for sh in sets_h:
        cnt = np.bincount(sh, minlength=n_all) if sh.size else np.zeros(n_all, np.int64)
        distinct = distinct and (cnt.max(initial=0) <= 1)
        in_pool |= cnt > 0
    pool = np.flatnonzero(in_pool)
    held = []
    for sh in sets_h:
        m = in_pool.copy()
        m[sh] = False
        held.append(np.flatnonzero(m))
    downdate = len(sets_h) >= 2 and distinct and \
        all(0 < len(h) < len(s) for h, s in zip(held, sets_h)) and X.shape[0] < GRAM_TC_MIN_ROWS * 4
    if not downdate:
        out = [covariance(X, idx) for idx in row_sets]
        return [c for c, _, _ in out], [m for _, m, _ in out], [n for _, _, n in out], None
    dev = X.device
    pool_t = index_tensor(pool, dev)
    s0, _, n_pool = colstats(X, pool_t)
    shift = s0 / float(n_pool)
    G_pool = gram(X, pool_t, shift)
    s_pool, _, _ = colstats(X, pool_t, shift)
    covs, means, ns = [], [], []
    for h in held:
        h_t = index_tensor(h, dev)
        G_h = gram(X, h_t, shift)
        s_h, _, _ = colstats(X, h_t, shift)
        Cm, mean = cov_from_parts(torch.stack([G_pool, G_h]), torch.stack([s_pool, s_h]), [n_pool, len(h)], [1.0, -1.0], shift)
        covs.append(Cm); means.append(mean); ns.append(n_pool - len(h))
    # the pool itself (all rows that are in some training set): a k-fold refit on "everything" is this covariance
    Cp, mp = cov_from_parts(G_pool.unsqueeze(0), s_pool.unsqueeze(0), [n_pool], [1.0], shift)
    return covs, means, ns, dict(rows=pool, cov=Cp, mean=mp, n=int(n_pool))


def fit_pca_folds(X, row_sets, k) -> list:
    """PCA models of several row subsets of the same matrix (CV folds): downdated Grams (see _fold_covariances), ONE
    batched partial eigensolve for all of them (eig_topk_subspace), full solver only for a subset that did not converge.
    The returned models carry a `limits.Spectrum` (k leading eigenvalues + tail power sums) instead of the full
    eigenvalue array."""
    from .limits import Spectrum
    torch = _torch()
    covs, means, ns, pool = _fold_covariances(X, row_sets)
    Cs = torch.stack(covs + ([pool["cov"]] if pool is not None else []))      # the pool rides in the same batched solve
    w, V, tails, conv = eig_topk_subspace(Cs, k)
    w_h, tails_h, conv_h = w[:, :k].cpu().numpy(), tails.cpu().numpy(), conv.cpu().numpy()
    P_all = sign_normalise_(V[:, :, :k].transpose(1, 2).contiguous())      # [F(+1), k, p], one batched pass
    P32_all = P_all.to(torch.float32)                                      # what the scoring launches read: converted once, not per fold
    inv_all = 1.0 / w[:, :k].clamp_min(0.0)
    inv32_all = inv_all.to(torch.float32)
    models = []
    for f in range(len(row_sets)):
        if not conv_h[f]:
            models.append(fit_pca(X, row_sets[f], k))
            continue
        P = P_all[f]
        lam = np.maximum(w_h[f], 0.0)
        models.append(PCAModel(mean=means[f], P=P, eigs_all=Spectrum(lam, np.maximum(tails_h[f], 0.0)), n_samples=ns[f]))
        models[-1].inv_eig_device = inv_all[f]      # stays on the device for the multi-A scoring launch
        models[-1].P32, models[-1].inv_eig32 = P32_all[f], inv32_all[f]
    fit_pca_folds.last_pool = None
    if pool is not None and conv_h[-1]:
        from .limits import LazySpectrum
        f = len(row_sets)
        P = P_all[f]
        Cp, n_p, p_ = pool["cov"], pool["n"], X.shape[1]
        spec = LazySpectrum(np.maximum(w_h[f], 0.0), np.maximum(tails_h[f], 0.0), lambda: _full_spectrum(Cp, n_p, p_))
        x_rms = float(np.sqrt(max(float(torch.diagonal(Cp).sum()), 0.0) / p_))
        fit_pca_folds.last_pool = dict(rows=pool["rows"], X=X,
                                       model=PCAModel(mean=pool["mean"], P=P, eigs_all=spec, n_samples=n_p, x_rms=x_rms))
    return models


fit_pca_folds.last_pool = None      # PCA of the union of the row sets of the last call (k components), for a refit on it


def eigvals_sym(Cm):
    """All eigenvalues (descending, fp64 device tensor) of a symmetric matrix by the library's own Householder
    tridiagonalisation + Sturm bisection (ocm_eigvals_sym_f64, csrc/eig_full.cu)."""
    torch = _torch()
    p = Cm.shape[0]
    A = Cm.contiguous().clone()                       # overwritten by the reduction
    work = torch.empty(2 * p + 8, dtype=torch.float64, device=Cm.device)
    d = torch.empty(p, dtype=torch.float64, device=Cm.device)
    e = torch.empty(p, dtype=torch.float64, device=Cm.device)
    w = torch.empty(p, dtype=torch.float64, device=Cm.device)
    native.check(native.lib().ocm_eigvals_sym_f64(_p(A), p, _p(work), _p(d), _p(e), _p(w), _stream()), "ocm_eigvals_sym_f64")
    return w


def _full_spectrum(Cm, n, p):
    """All eigenvalues of the covariance, descending, as the reference keeps them (min(n, p) entries)."""
    evals = eigvals_sym(Cm)
    m = min(n, p)
    eigs = evals[:m].clamp_min(0.0).cpu().numpy()
    if n <= p:
        eigs[n - 1:] = 0.0     # rank n-1: the reference's trailing values are ~1e-30 noise
    return eigs


def fit_pca(X, row_idx=None, k=2, group=None, full_spectrum=False, gram_path="auto", products=None, timer=None) -> PCAModel:
    """PCA of the selected rows: mean, k leading loadings (sign-normalised rows) and the eigenvalue spectrum.

    Default: block subspace iteration for the k leading pairs + traces of the deflated covariance for the tail sums
    (everything the limits need); `eigs_all` is then a `limits.LazySpectrum` that runs the full eigenvalue solve only
    if the whole array is asked for.  `full_spectrum=True` (or a subspace iteration that did not converge: flat
    spectrum at k) takes the full symmetric eigensolve straight away."""
    from .limits import LazySpectrum
    Cm, mean, n = covariance(X, row_idx, group, gram_path=gram_path, products=products, timer=timer)
    p = X.shape[1]
    x_rms = float(np.sqrt(max(float(_torch().diagonal(Cm).sum()), 0.0) / p))
    if not full_spectrum and k < p:
        w, V, tails, conv = eig_topk_subspace(Cm.unsqueeze(0), k, tol=1e-13, max_iter=90)
        if timer is not None:
            timer.mark("eigensolve")
        if bool(conv[0]):
            P = V[0, :, :k].transpose(0, 1).contiguous()
            sign_normalise_(P)
            lam = np.maximum(w[0, :k].cpu().numpy(), 0.0)
            spec = LazySpectrum(lam, np.maximum(tails[0].cpu().numpy(), 0.0), lambda: _full_spectrum(Cm, n, p))
            return PCAModel(mean=mean, P=P, eigs_all=spec, n_samples=n, x_rms=x_rms)
    evals, evecs = eigh_topk(Cm, k)
    m = min(n, p)
    eigs = evals[:m].clamp_min(0.0).cpu().numpy()
    if n <= p:
        eigs[n - 1:] = 0.0     # rank n-1: the reference's trailing values are ~1e-30 noise
    return PCAModel(mean=mean, P=evecs, eigs_all=eigs, n_samples=n, x_rms=x_rms)


# ------------------------------------------------------------------------------------------
# fused scoring
# ------------------------------------------------------------------------------------------
def score(X, mean, P, Sinv, rule="alt", ct=1.0, cq=1.0, dlim=np.sqrt(2), snv=False, snv_eps=1e-8,
          want_scores=False, force_simt=False, x_rms=0.0):
    """Fused T2 / Q / accept for the rows of X against one model.

    x_rms (PCAModel.x_rms of the model's fit; 0 = unknown) lets the tcgen05 kernel use its faster fp16 operand mode, see
    ocm_score_params in include/ocm_simca.h.

    X: CUDA tensor [n,p] fp32 or fp64.  mean [p], P [A,p], Sinv [A,A] any float dtype on the
    device.  Returns (T2, Q, accept[, scores]) as device tensors (fp32, fp32, uint8[, fp32])."""
    torch = _torch()
    lib = native.lib()
    n, p = X.shape
    A = P.shape[0]
    dev = X.device
    T2 = torch.empty(n, dtype=torch.float32, device=dev)
    Q = torch.empty(n, dtype=torch.float32, device=dev)
    acc = torch.empty(n, dtype=torch.uint8, device=dev)
    sc = torch.empty((n, A), dtype=torch.float32, device=dev) if want_scores else None
    prm = native.ScoreParams(A=A, rule=RULES[rule] if isinstance(rule, str) else int(rule), snv=int(bool(snv)),
                             snv_eps=float(snv_eps), ct=float(ct), cq=float(cq), dlim=float(dlim), x_rms=float(x_rms or 0.0))
    P32 = P.to(torch.float32).contiguous()
    S32 = Sinv.to(torch.float32).contiguous()
    if X.dtype == torch.float64:
        mu = mean.to(torch.float64).contiguous()
        rc = lib.ocm_simca_score_f64in(_p(X), n, p, X.stride(0), _p(mu), _p(P32), _p(S32), C.byref(prm),
                                       _p(T2), _p(Q), _p(acc), _p(sc), _stream())
    else:
        mu = mean.to(torch.float32).contiguous()
        fn = lib.ocm_simca_score_simt_f32 if force_simt else lib.ocm_simca_score_f32
        rc = fn(_p(X), n, p, X.stride(0), _p(mu), _p(P32), _p(S32), C.byref(prm), _p(T2), _p(Q), _p(acc),
                _p(sc), _stream())
    native.check(rc, "ocm_simca_score")
    return (T2, Q, acc, sc) if want_scores else (T2, Q, acc)


def score_savgol(X, mean, P, Sinv, window_length=5, polyorder=2, deriv=1, delta=1.0, pre_snv=True, snv_eps=1e-8,
                 rule="alt", ct=1.0, cq=1.0, dlim=np.sqrt(2), x_rms=0.0):
    """Raw fp32 device rows -> (T2, Q, accept) against a model fitted on [SNV +] Savitzky-Golay filtered spectra
    (ocm_simca_score_savgol_f32): the filter runs chunk by chunk into a bounded scratch, the filtered matrix is never
    materialised.  Same numbers as `score(savgol(X, ..., snv=pre_snv), ...)`."""
    torch = _torch()
    X = to_device(X, dtype=torch.float32)
    n, p = X.shape
    A = P.shape[0]
    dev = X.device
    key = (int(window_length), int(polyorder), int(deriv), float(delta), str(dev))
    if key not in _SG_DEVICE:
        g, el, er = _savgol_tables(int(window_length), int(polyorder), int(deriv), delta)
        _SG_DEVICE[key] = (g, torch.as_tensor(el, device=dev).contiguous(), torch.as_tensor(er, device=dev).contiguous())
    g, el_d, er_d = _SG_DEVICE[key]
    T2 = torch.empty(n, dtype=torch.float32, device=dev)
    Q = torch.empty(n, dtype=torch.float32, device=dev)
    acc = torch.empty(n, dtype=torch.uint8, device=dev)
    prm = native.ScoreParams(A=A, rule=RULES[rule] if isinstance(rule, str) else int(rule), snv=0, snv_eps=float(snv_eps),
                             ct=float(ct), cq=float(cq), dlim=float(dlim), x_rms=float(x_rms or 0.0))
    mu = mean.to(device=dev, dtype=torch.float32).contiguous()
    P32 = P.to(device=dev, dtype=torch.float32).contiguous()
    S32 = Sinv.to(device=dev, dtype=torch.float32).contiguous()
    rc = native.lib().ocm_simca_score_savgol_f32(_p(X), n, p, X.stride(0), int(bool(pre_snv)), float(snv_eps),
                                                 g.ctypes.data_as(C.c_void_p), int(window_length), _p(el_d), _p(er_d),
                                                 _p(mu), _p(P32), _p(S32), C.byref(prm), _p(T2), _p(Q), _p(acc), _stream())
    native.check(rc, "ocm_simca_score_savgol_f32")
    return T2, Q, acc


HOST_STREAM_BYTES = 256 << 20      # host matrices above this are streamed through the host-buffer call, not uploaded whole


def host_streamable(X):
    """True for a host matrix that `score_host` takes as it is: 2-D numpy fp32 / fp64, unit column stride, above
    HOST_STREAM_BYTES (smaller inputs are uploaded whole: one copy, one launch)."""
    return (isinstance(X, np.ndarray) and X.ndim == 2 and X.dtype in (np.float32, np.float64) and X.shape[0] > 0
            and X.strides[1] == X.itemsize and X.strides[0] >= X.shape[1] * X.itemsize and X.strides[0] % X.itemsize == 0
            and X.nbytes > HOST_STREAM_BYTES)


def score_host(X, mean, P, Sinv, rule="alt", ct=1.0, cq=1.0, dlim=np.sqrt(2), snv=False, snv_eps=1e-8, x_rms=0.0,
               chunk_rows=0):
    """Fused T2 / Q / accept for the rows of a HOST matrix (numpy fp32 or fp64, pageable or pinned) against one model,
    streamed chunk by chunk over the link with the copies overlapping the kernel (ocm_simca_score_host_f32 / _f64in): the
    matrix never has to fit in device memory.  Returns numpy (T2 fp32, Q fp32, accept uint8)."""
    torch = native.require_cuda()
    lib = native.lib()
    n, p = X.shape
    A = P.shape[0]
    f64 = X.dtype == np.float64
    mu_h = np.ascontiguousarray(mean.detach().cpu().numpy() if isinstance(mean, torch.Tensor) else mean,
                                dtype=np.float64 if f64 else np.float32)
    P_h = np.ascontiguousarray(P.detach().cpu().numpy() if isinstance(P, torch.Tensor) else P, dtype=np.float32)
    S_h = np.ascontiguousarray(Sinv.detach().cpu().numpy() if isinstance(Sinv, torch.Tensor) else Sinv, dtype=np.float32)
    T2 = np.empty(n, np.float32); Q = np.empty(n, np.float32); acc = np.empty(n, np.uint8)
    prm = native.ScoreParams(A=A, rule=RULES[rule] if isinstance(rule, str) else int(rule), snv=int(bool(snv)),
                             snv_eps=float(snv_eps), ct=float(ct), cq=float(cq), dlim=float(dlim), x_rms=float(x_rms or 0.0))
    hp = lambda a: C.c_void_p(a.ctypes.data)
    fn = lib.ocm_simca_score_host_f64in if f64 else lib.ocm_simca_score_host_f32
    rc = fn(hp(X), n, p, X.strides[0] // X.itemsize, hp(mu_h), hp(P_h), hp(S_h), C.byref(prm), int(chunk_rows),
            hp(T2), hp(Q), hp(acc))
    native.check(rc, "ocm_simca_score_host")
    return T2, Q, acc


def score_multiA(X, mean, P, inv_eig, rule="alt", ct=None, cq=None, dlim=None, snv=False, snv_eps=1e-8):
    """All nested models A = 1..Amax at once.  Returns T2, Q [n,Amax] fp32 and accept [n,Amax] uint8
    (accept is None when no limits are given)."""
    torch = _torch()
    lib = native.lib()
    n, p = X.shape
    Amax = P.shape[0]
    dev = X.device
    T2 = torch.empty((n, Amax), dtype=torch.float32, device=dev)
    Q = torch.empty((n, Amax), dtype=torch.float32, device=dev)
    have_lim = ct is not None
    acc = torch.empty((n, Amax), dtype=torch.uint8, device=dev) if have_lim else None
    f32 = lambda v: None if v is None else torch.as_tensor(np.asarray(v, dtype=np.float32), device=dev) \
        if not isinstance(v, torch.Tensor) else v.to(device=dev, dtype=torch.float32).contiguous()
    ct_t, cq_t, dl_t = f32(ct), f32(cq), f32(dlim)
    P32 = P.to(torch.float32).contiguous()
    ie = inv_eig.to(torch.float32).contiguous()
    r = RULES[rule] if isinstance(rule, str) else int(rule)
    if X.dtype == torch.float64:
        mu = mean.to(torch.float64).contiguous()
        fn = lib.ocm_simca_score_multiA_f64in
    else:
        mu = mean.to(torch.float32).contiguous()
        fn = lib.ocm_simca_score_multiA_f32
    rc = fn(_p(X), n, p, X.stride(0), _p(mu), _p(P32), _p(ie), Amax, r, int(bool(snv)), float(snv_eps),
            _p(ct_t), _p(cq_t), _p(dl_t), _p(T2), _p(Q), _p(acc), _stream())
    native.check(rc, "ocm_simca_score_multiA")
    return T2, Q, acc


def conformity_counts(accept, is_target):
    """[m,4] int64 {TP,TN,FP,FN} per decision column (ocm_conformity_counts)."""
    torch = _torch()
    if accept.dim() == 1:
        accept = accept.unsqueeze(1)
    accept = accept.contiguous()
    n, m = accept.shape
    counts = torch.zeros((m, 4), dtype=torch.int64, device=accept.device)
    tgt = is_target.to(device=accept.device, dtype=torch.uint8).contiguous()
    native.check(native.lib().ocm_conformity_counts(_p(accept), _p(tgt), n, m, _p(counts), _stream()),
                 "ocm_conformity_counts")
    return counts


ACTS = {"none": 0, "elu": 1, "gelu": 2}


def vae_conv_in(x, spec_mean, spec_std, w, bias, act="elu"):
    """First ConvVAE1D encoder block in one kernel (ocm_vae_conv_in_bf16): standardise fp32 rows x [n, L], Conv1d(1 -> C,
    K taps, stride 1, 'same' zero padding) with w [C, K] / bias [C] (fp32, BatchNorm already folded in), activation;
    returns bf16 [n, L, C] (channels-last)."""
    torch = _torch()
    n, L = x.shape
    Cc, K = w.shape
    out = torch.empty((n, L, Cc), dtype=torch.bfloat16, device=x.device)
    rc = native.lib().ocm_vae_conv_in_bf16(_p(x), n, L, x.stride(0), _p(spec_mean), _p(spec_std), _p(w), _p(bias), Cc, K,
                                           ACTS[act], _p(out), _stream())
    native.check(rc, "ocm_vae_conv_in_bf16")
    return out


def bias_act_(y, bias, act="elu"):
    """In place y[..., c] <- act(y[..., c] + bias[c]) on a dense bf16 tensor whose LAST (memory) axis is the channel
    (ocm_bias_act_bf16)."""
    Cc = y.shape[-1]
    rc = native.lib().ocm_bias_act_bf16(_p(y), y.numel() // Cc, Cc, _p(bias), ACTS[act], _stream())
    native.check(rc, "ocm_bias_act_bf16")
    return y


def vae_tail_q(y, bias, act, w_out, b_out, x, spec_mean, spec_std, want_xrec=False):
    """Decoder tail + reconstruction Q in one kernel (ocm_vae_tail_q_bf16): y [n, Ly, C] bf16 (raw output of the last
    transposed convolution, channels-last), x [n, L] fp32.  Returns q [n] fp32 (and x_rec [n, L] bf16 when asked)."""
    torch = _torch()
    n, Ly, Cc = y.shape
    L = x.shape[1]
    q = torch.empty(n, dtype=torch.float32, device=x.device)
    xr = torch.empty((n, L), dtype=torch.bfloat16, device=x.device) if want_xrec else None
    rc = native.lib().ocm_vae_tail_q_bf16(_p(y), n, Ly, Cc, _p(bias), ACTS[act], _p(w_out), float(b_out), _p(x), x.stride(0), L,
                                          _p(spec_mean), _p(spec_std), _p(q), _p(xr), _stream())
    native.check(rc, "ocm_vae_tail_q_bf16")
    return (q, xr) if want_xrec else q


def latent_score(z=None, m=None, Sinv=None, x=None, xrec=None):
    """VAE-SIMCA epilogue: d2 = (z-m)^T Sinv (z-m), q = |x - xrec|^2 (ocm_latent_score_f32)."""
    torch = _torch()
    ref = z if z is not None else x
    n, dev = ref.shape[0], ref.device
    d2 = torch.empty(n, dtype=torch.float32, device=dev) if z is not None else None
    q = torch.empty(n, dtype=torch.float32, device=dev) if x is not None else None
    is_bf16 = 0
    if z is not None:
        z = z.to(torch.float32).contiguous()
        m = m.to(device=dev, dtype=torch.float32).contiguous()
        Sinv = Sinv.to(device=dev, dtype=torch.float32).contiguous()
    if x is not None:
        x = x.to(torch.float32).contiguous()
        if xrec.dtype == torch.bfloat16:
            is_bf16 = 1
            xrec = xrec.contiguous()
        else:
            xrec = xrec.to(torch.float32).contiguous()
    rc = native.lib().ocm_latent_score_f32(
        _p(z), n, 0 if z is None else z.shape[1], 0 if z is None else z.stride(0), _p(m), _p(Sinv),
        _p(x), _p(xrec), is_bf16, 0 if x is None else x.shape[1], 0 if x is None else x.stride(0),
        0 if x is None else xrec.stride(0), _p(d2), _p(q), _stream())
    native.check(rc, "ocm_latent_score_f32")
    return d2, q
