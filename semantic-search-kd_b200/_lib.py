"""ctypes loader (and in-tree builder) of libb200search.so -- the C ABI in include/b200search.h.

The library is compiled in-tree with nvcc for sm_100a only; the built .so is git-ignored but
travels to the GPU box with the repo snapshot.  There is no CPU fallback: if the library cannot
be loaded, or no B200 is present, every compute call raises.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_ROOT = _PKG.parent
_CSRC = _PKG / "csrc"
_SO = _PKG / "libb200search.so"
_SOURCES = ["b2s_api.cu", "select.cuh", "scan_topk.cuh", "merge_topk.cuh", "util_kernels.cuh",
            "gemm_topk_tc.cuh", "gemm_topk_host.inl", "ptx.cuh", "ance_filter.cuh", "exchange.cuh", "rescore.cuh"]

B2S_OK = 0
B2S_ERR_INVALID = -1
B2S_ERR_CUDA = -2
B2S_ERR_NOMEM = -3
B2S_ERR_UNSUPPORTED = -4
B2S_ERR_NO_DEVICE = -5
METRIC_INNER_PRODUCT = 0
METRIC_COSINE = 1
DTYPE_F32 = 0
DTYPE_BF16 = 1
PATH_AUTO, PATH_SCAN, PATH_TENSOR = 0, 1, 2
SEARCH_STABLE_QUERIES = 1

# every symbol include/b200search.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "b2s_version", "b2s_last_error", "b2s_device_count", "b2s_create", "b2s_destroy", "b2s_reserve",
    "b2s_add_f32", "b2s_add_bf16", "b2s_ntotal", "b2s_dim", "b2s_reset", "b2s_set_id_offset",
    "b2s_set_option", "b2s_get_option", "b2s_search", "b2s_search_device", "b2s_merge_device",
    "b2s_similarity", "b2s_read_rows_f32", "b2s_rows_device", "b2s_last_stats", "b2s_read_timings",
    "b2s_packed_bytes", "b2s_merge_packed_device", "b2s_score_rows_device", "b2s_ance_filter_device",
    "b2s_exchange_create", "b2s_exchange_local", "b2s_exchange_connect", "b2s_exchange_status",
    "b2s_search_sharded_device", "b2s_search_sharded", "b2s_maxsim_device", "b2s_add_prepared", "b2s_read_trace",
]


class Stats(ctypes.Structure):
    _fields_ = [("path", ctypes.c_int32), ("kernel_launches", ctypes.c_int32),
                ("seeded", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("corpus_bytes", ctypes.c_int64), ("passes", ctypes.c_int64),
                ("dominant_ms", ctypes.c_float), ("total_ms", ctypes.c_float)]


def nvcc_command(out: Path = _SO, extra=()):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-o", str(out)]
    if not (_CSRC / "gemm_topk_tc.cuh").exists():
        cmd.append("-DB2S_NO_TENSOR_PATH")
    cmd += list(extra)
    cmd.append(str(_CSRC / "b2s_api.cu"))
    return cmd


def is_stale() -> bool:
    if not _SO.exists():
        return True
    t = _SO.stat().st_mtime
    hdr = _ROOT / "include" / "b200search.h"
    srcs = [_CSRC / s for s in _SOURCES if (_CSRC / s).exists()] + [hdr]
    return any(s.stat().st_mtime > t for s in srcs)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/ into libb200search.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if force or is_stale():
        cmd = nvcc_command(extra=["-Xptxas", "-v"] if verbose else [])
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        if verbose:
            import re
            for m in re.finditer(r"Compiling entry function '(\w+)'.*?Used (\d+) registers[^\n]*", res.stderr, re.S):
                print(f"{m.group(2):>4} regs  {m.group(1)[:90]}")
            for ln in res.stderr.splitlines():
                if "spill" in ln and "0 bytes spill stores, 0 bytes spill loads" not in ln:
                    print(ln)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    """Load the library (building it if nvcc is available and the .so is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _SO.exists():
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            build()
        else:
            raise RuntimeError(f"{_SO} is missing and nvcc is not available; the CUDA extension is "
                               "required (no CPU fallback)")
    L = ctypes.CDLL(str(_SO))
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    L.b2s_version.restype = i32
    L.b2s_last_error.restype = ctypes.c_char_p
    L.b2s_device_count.restype = i32
    L.b2s_create.argtypes = [i32, i32, i32, ctypes.POINTER(vp)]
    L.b2s_destroy.argtypes = [vp]
    L.b2s_reserve.argtypes = [vp, i64]
    L.b2s_add_f32.argtypes = [vp, vp, i64, i32]
    L.b2s_add_bf16.argtypes = [vp, vp, i64, i32]
    L.b2s_ntotal.argtypes = [vp]
    L.b2s_ntotal.restype = i64
    L.b2s_dim.argtypes = [vp]
    L.b2s_reset.argtypes = [vp]
    L.b2s_set_id_offset.argtypes = [vp, i64]
    L.b2s_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    L.b2s_get_option.argtypes = [vp, ctypes.c_char_p]
    L.b2s_get_option.restype = i64
    L.b2s_search.argtypes = [vp, vp, i64, i32, vp, vp]
    L.b2s_search_device.argtypes = [vp, vp, i32, i64, i32, vp, vp, vp, ctypes.c_uint]
    L.b2s_merge_device.argtypes = [i32, vp, vp, i32, i64, i32, vp, vp, vp]
    L.b2s_similarity.argtypes = [i32, vp, i64, vp, i64, i32, vp]
    L.b2s_read_rows_f32.argtypes = [vp, i64, i64, vp]
    L.b2s_rows_device.argtypes = [vp]
    L.b2s_rows_device.restype = vp
    L.b2s_last_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    L.b2s_read_timings.argtypes = [vp, vp, vp, i32]
    L.b2s_read_timings.restype = i32
    L.b2s_packed_bytes.argtypes = [i64, i32]
    L.b2s_packed_bytes.restype = i64
    L.b2s_merge_packed_device.argtypes = [i32, vp, i32, i64, i32, vp, vp, vp]
    L.b2s_merge_packed_device.restype = i32
    L.b2s_score_rows_device.argtypes = [vp, vp, i32, i32, i64, vp, i32, vp, vp]
    L.b2s_score_rows_device.restype = i32
    L.b2s_ance_filter_device.argtypes = [i32, vp, vp, i64, i32, vp, vp, i32, ctypes.c_float, i32, vp, vp, vp, vp]
    L.b2s_ance_filter_device.restype = i32
    L.b2s_exchange_create.argtypes = [vp, i32, i32, i64, i32, vp]
    L.b2s_exchange_create.restype = i32
    L.b2s_exchange_local.argtypes = [vp]
    L.b2s_exchange_local.restype = vp
    L.b2s_exchange_connect.argtypes = [vp, vp, i32]
    L.b2s_exchange_connect.restype = i32
    L.b2s_exchange_status.argtypes = [vp]
    L.b2s_exchange_status.restype = i32
    L.b2s_search_sharded_device.argtypes = [vp, vp, i32, i64, i32, vp, vp, vp, i32, ctypes.c_uint]
    L.b2s_search_sharded_device.restype = i32
    L.b2s_search_sharded.argtypes = [vp, vp, i64, i32, vp, vp]
    L.b2s_search_sharded.restype = i32
    L.b2s_maxsim_device.argtypes = [i32, vp, vp, i64, i32, vp, i64, i32, vp, vp, vp, vp]
    L.b2s_maxsim_device.restype = i32
    L.b2s_add_prepared.argtypes = [vp, vp, i32, i64, i32]
    L.b2s_add_prepared.restype = i32
    L.b2s_read_trace.argtypes = [vp, vp, i32]
    L.b2s_read_trace.restype = i32
    for name in ("b2s_create", "b2s_destroy", "b2s_reserve", "b2s_add_f32", "b2s_add_bf16", "b2s_dim",
                 "b2s_reset", "b2s_set_id_offset", "b2s_set_option", "b2s_search", "b2s_search_device",
                 "b2s_merge_device", "b2s_similarity", "b2s_read_rows_f32", "b2s_last_stats"):
        getattr(L, name).restype = i32
    _lib = L
    return L


def last_error() -> str:
    return lib().b2s_last_error().decode("utf-8", "replace")
