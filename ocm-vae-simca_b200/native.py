"""ctypes binding of libocm_simca_b200.so (the C ABI in include/ocm_simca.h).

There is deliberately NO fallback: if the shared library has not been built, or a
compute call is made without a CUDA device, this module raises.  Build with
`python -c "import __graft_entry__ as g; g.build()"` (or `native.build()`).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.environ.get("OCM_SIMCA_LIB") or os.path.join(PKG_DIR, "libocm_simca_b200.so")   # override: A/B builds
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]
LINK_LIBS = ["-lcuda"]

_lib = None


class NativeError(RuntimeError):
    pass


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(s) > t for s in deps)


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a into the in-tree shared library."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise NativeError("nvcc not found; cannot build libocm_simca_b200.so")
    objs = []
    build_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        deps_newer = (not os.path.exists(obj)) or force or any(
            os.path.getmtime(d) > os.path.getmtime(obj)
            for d in [src] + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h")))
        if deps_newer:
            cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                   "-Xcompiler", "-fPIC", "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, pr in procs:
        out, _ = pr.communicate()
        if verbose and out:
            print(out.decode())
        if pr.returncode != 0:
            raise NativeError(f"nvcc failed on {src}:\n{out.decode()}")
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + LINK_LIBS
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise NativeError("link failed:\n" + r.stdout.decode())
    global _lib
    _lib = None
    return LIB_PATH


class ScoreParams(C.Structure):
    _fields_ = [("A", C.c_int32), ("rule", C.c_int32), ("snv", C.c_int32), ("snv_eps", C.c_float),
                ("ct", C.c_float), ("cq", C.c_float), ("dlim", C.c_float), ("x_rms", C.c_float)]


_vp, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float

# name -> (restype, argtypes): mirrors include/ocm_simca.h one to one
SIGNATURES = {
    "ocm_abi_version": (C.c_int, []),
    "ocm_last_error": (C.c_char_p, []),
    "ocm_launch_count": (C.c_uint64, []),
    "ocm_preprocess_f32": (C.c_int, [_vp, _i64, _i32, _i64, _i32, _vp, _vp, _f32, _vp, _i64, _vp]),
    "ocm_savgol_f32": (C.c_int, [_vp, _i64, _i32, _i64, _i32, _f32, _vp, _i32, _vp, _vp, _vp, _i64, _vp]),
    "ocm_colstats_f32": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "ocm_colstats_f64": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "ocm_colstats_minmax_f32": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "ocm_colstats_minmax_f64": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "ocm_gram_tc_workspace_bytes": (C.c_int, [_i64, _i32, C.POINTER(_i64)]),
    "ocm_gram_tc_f32": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp]),
    "ocm_gram_tc_f64": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp]),
    "ocm_gram_f64acc_f32": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp]),
    "ocm_gram_f64acc_f64": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp]),
    "ocm_cov_from_parts_f64": (C.c_int, [_vp, _vp, C.POINTER(_i64), C.POINTER(C.c_double), _i32, _i32,
                                         _vp, _vp, _vp, _vp, _vp]),
    "ocm_eig_symm_mul_f64": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "ocm_eig_tn_small_f64": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ocm_eig_chol_solve_f64": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "ocm_eig_mul_small_f64": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ocm_eig_ritz_residual_f64": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "ocm_eig_tail_sums_f64": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "ocm_eigvals_sym_f64": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "ocm_eigh_small_f64": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "ocm_simca_score_f32": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, C.POINTER(ScoreParams),
                                      _vp, _vp, _vp, _vp, _vp]),
    "ocm_simca_score_simt_f32": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, C.POINTER(ScoreParams),
                                           _vp, _vp, _vp, _vp, _vp]),
    "ocm_simca_score_rowteam_supported": (C.c_int, [_i64, _i32, _i64, _i32, _i32, _vp]),
    "ocm_simca_score_rowteam_f32": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, C.POINTER(ScoreParams),
                                              _vp, _vp, _vp, _vp]),
    "ocm_simca_score_tc_f32": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, C.POINTER(ScoreParams),
                                         _vp, _vp, _vp, _vp]),
    "ocm_simca_score_tc2_supported": (C.c_int, [_i64, _i32, _i64, _i32, _i32, _vp]),
    "ocm_simca_score_tc2_f32": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, C.POINTER(ScoreParams),
                                          _vp, _vp, _vp, _i32, _vp]),
    "ocm_simca_score_tc_debug_words": (C.c_int, [_vp]),
    "ocm_simca_score_tc_status": (C.c_int, []),
    "ocm_simca_score_tc_inject_stall": (C.c_int, [_i32]),
    "ocm_simca_host_release": (C.c_int, []),
    "ocm_simca_score_tc_supported": (C.c_int, [_i64, _i32, _i64, _i32, _i32, _vp]),
    "ocm_simca_score_f64in": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, C.POINTER(ScoreParams),
                                        _vp, _vp, _vp, _vp, _vp]),
    "ocm_simca_score_simt_f64in": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, C.POINTER(ScoreParams),
                                             _vp, _vp, _vp, _vp, _vp]),
    "ocm_simca_score_snv_staged_supported": (C.c_int, [_i64, _i32, _i64, _i32, _vp, _vp]),
    "ocm_simca_score_snv_staged_f32": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, C.POINTER(ScoreParams),
                                                 _vp, _vp, _vp, _vp]),
    "ocm_simca_score_savgol_f32": (C.c_int, [_vp, _i64, _i32, _i64, _i32, _f32, _vp, _i32, _vp, _vp, _vp, _vp, _vp,
                                             C.POINTER(ScoreParams), _vp, _vp, _vp, _vp]),
    "ocm_simca_score_f64in_bulk_supported": (C.c_int, [_i64, _i32, _i64, _i32, _vp, _vp]),
    "ocm_simca_score_multiA_f32": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _f32,
                                             _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ocm_simca_score_multiA_f64in": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _f32,
                                               _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ocm_conformity_counts": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "ocm_latent_score_f32": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _i64,
                                       _vp, _vp, _vp]),
    "ocm_simca_score_host_f32": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, C.POINTER(ScoreParams), _i64,
                                           _vp, _vp, _vp]),
    "ocm_vae_conv_in_bf16": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "ocm_bias_act_bf16": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _vp]),
    "ocm_vae_tail_q_bf16": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _f32, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "ocm_upload_rows": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _i64]),
    "ocm_simca_score_host_f64in": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, C.POINTER(ScoreParams), _i64,
                                             _vp, _vp, _vp]),
}


def lib():
    """Load (once) and return the shared library; raises NativeError when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing -- the CUDA library has not been built "
                "(run __graft_entry__.build()).  There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError as e:
                raise NativeError(f"{LIB_PATH} does not export {name}; rebuild it") from e
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ocm_last_error().decode(errors="replace")
        raise NativeError(f"{what or 'ocm call'} failed with code {rc}: {msg}")


def launch_count():
    return int(lib().ocm_launch_count())


_torch_mod = None


def require_cuda():
    """torch, once a CUDA device has been seen (the check is cached: it is on the path of every kernel call)."""
    global _torch_mod
    if _torch_mod is None:
        import torch
        if not torch.cuda.is_available():
            raise NativeError("no CUDA device: the SIMCA hot path has no CPU fallback")
        _torch_mod = torch
    return _torch_mod
