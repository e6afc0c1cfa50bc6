"""ctypes binding of the C ABI declared in `include/mccube_b200.h`.

This is the ONLY way the Python layer reaches the GPU kernels, and there is no CPU
fallback: if `libmccube_b200.so` is missing or a call fails, an exception is raised.
torch is used for device memory and streams only (`tensor.data_ptr()` in, results in
pre-allocated tensors).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np
import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmccube_b200.so"

BOX_MAX_DIMS = 8


class MccubeB200Error(RuntimeError):
    pass


class Rng(C.Structure):
    _fields_ = [
        ("key", C.c_uint32 * 2),
        ("key_dev", C.c_void_p),
        ("t", C.c_float),
        ("t_dev", C.c_void_p),
        ("leaf", C.c_uint32),
        ("n_leaves", C.c_uint32),
        ("partitionable", C.c_int32),
        ("fold_time", C.c_int32),
    ]


class Drift(C.Structure):
    _fields_ = [("kind", C.c_int32), ("f", C.c_void_p), ("mean", C.c_void_p), ("inv_var", C.c_void_p)]


class Cubature(C.Structure):
    _fields_ = [("k", C.c_int32), ("points", C.c_void_p), ("scale", C.c_float), ("lam", C.c_void_p)]


_SIGNATURES = {
    # name: (restype, argtypes)
    "mccb200_version": (C.c_int, []),
    "mccb200_last_error": (C.c_char_p, []),
    "mccb200_sm_count": (C.c_int, []),
    "mccb200_launch_count": (C.c_ulonglong, []),
    "mccb200_profile_enable": (None, [C.c_int]),
    "mccb200_profile_read": (C.c_int, [C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]),
    "mccb200_mc_indices_workspace_bytes": (C.c_size_t, [C.c_uint64, C.c_uint64, C.c_int]),
    "mccb200_mc_indices": (C.c_int, [C.POINTER(Rng), C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_size_t, C.c_void_p]),
    "mccb200_weighted_workspace_bytes": (C.c_size_t, [C.c_uint64]),
    "mccb200_gumbel_keys": (C.c_int, [C.POINTER(Rng), C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p,
                                      C.c_size_t, C.c_void_p]),
    "mccb200_weighted_choice_replace": (C.c_int, [C.POINTER(Rng), C.c_void_p, C.c_int, C.c_uint64, C.c_uint64,
                                                  C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mccb200_sort_pairs_workspace_bytes": (C.c_size_t, [C.c_uint64]),
    "mccb200_sort_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    "mccb200_expand_select": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(Drift), C.c_float,
                                        C.POINTER(Cubature), C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                                        C.c_uint64, C.c_void_p, C.c_void_p]),
    "mccb200_expand_select_substeps": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Drift), C.c_int,
                                                 C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_void_p,
                                                 C.c_uint64, C.c_void_p, C.c_void_p]),
    "mccb200_expand_scatter_peers": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Drift), C.c_float,
                                               C.POINTER(Cubature), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                               C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mccb200_expand_scatter_slots": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Drift), C.c_float,
                                               C.POINTER(Cubature), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                               C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mccb200_shard_bucket_ids_workspace_bytes": (C.c_size_t, [C.c_int]),
    "mccb200_shard_bucket_ids": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mccb200_normalised_copy": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mccb200_expand_select_wsum": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Drift), C.c_float,
                                             C.POINTER(Cubature), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                             C.c_void_p]),
    "mccb200_kd_partition_levels": (C.c_int, [C.c_uint64, C.c_uint64]),
    "mccb200_kd_partition_workspace_bytes": (C.c_size_t, [C.c_uint64, C.c_int, C.c_int]),
    "mccb200_kd_partition": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_void_p]),
    "mccb200_expand_select_draw_workspace_bytes": (C.c_size_t, []),
    "mccb200_expand_select_draw_supported": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "mccb200_expand_select_draw": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Drift), C.c_float,
                                             C.POINTER(Cubature), C.POINTER(Rng), C.c_uint64, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_size_t, C.c_void_p]),
    "mccb200_expand_select_strided": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Drift), C.c_float,
                                               C.POINTER(Cubature), C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mccb200_shard_bucket_ids_regions": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                                  C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                                  C.c_void_p]),
    "mccb200_peer_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mccb200_set_expand_residency": (None, [C.c_int]),
    "mccb200_set_side_residency": (None, [C.c_int]),
    "mccb200_unzip_pairs": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mccb200_compact_slots": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mccb200_enable_peer_access": (C.c_int, [C.c_int]),
    "mccb200_mc_indices_slice": (C.c_int, [C.POINTER(Rng), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                           C.c_void_p, C.c_size_t, C.c_void_p]),
    "mccb200_expand_gather_peers": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.POINTER(Drift), C.c_float,
                                              C.POINTER(Cubature), C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                              C.c_void_p]),
    "mccb200_mc_indices_slice64": (C.c_int, [C.POINTER(Rng), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                             C.c_void_p, C.c_size_t, C.c_void_p]),
    "mccb200_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "mccb200_peer_free": (C.c_int, [C.c_void_p]),
    "mccb200_peer_export": (C.c_int, [C.c_void_p, C.c_char_p]),
    "mccb200_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "mccb200_peer_close": (C.c_int, [C.c_void_p]),
    "mccb200_expand_all": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(Drift), C.c_float,
                                     C.POINTER(Cubature), C.c_void_p, C.c_void_p]),
    "mccb200_gather_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mccb200_normalise_weights": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]),
    "mccb200_stratified_keys": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "mccb200_box_bounds_workspace_bytes": (C.c_size_t, []),
    "mccb200_box_bounds": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mccb200_box_keys": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "mccb200_box_child_bounds": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(Drift), C.c_float,
                                           C.POINTER(Cubature), C.POINTER(C.c_int32), C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mccb200_box_child_minmax": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(Drift), C.c_float,
                                           C.POINTER(Cubature), C.POINTER(C.c_int32), C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mccb200_box_child_keys": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(Drift), C.c_float,
                                         C.POINTER(Cubature), C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "mccb200_box_prk_step_workspace_bytes": (C.c_size_t, [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "mccb200_box_prk_step": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Drift), C.c_float,
                                       C.POINTER(Cubature), C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.c_void_p]),
    "mccb200_box_prk_tables_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "mccb200_box_prk_tables": (C.c_int, [C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                         C.c_void_p]),
    "mccb200_box_prk_selection_bytes": (C.c_size_t, [C.c_uint64, C.c_uint64]),
    "mccb200_box_prk_selection": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mccb200_box_prk_count": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Drift), C.c_float,
                                        C.POINTER(Cubature), C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_void_p,
                                        C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mccb200_box_prk_rank_gather": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Drift), C.c_float,
                                              C.POINTER(Cubature), C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_uint64,
                                              C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_size_t, C.c_void_p]),
    "mccb200_box_prk_rank_global": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Drift), C.c_float,
                                              C.POINTER(Cubature), C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_uint64,
                                              C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_size_t, C.c_void_p]),
    "mccb200_box_prk_key_start_offset": (C.c_size_t, [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                                      C.POINTER(C.c_uint32)]),
    "mccb200_moments_sum": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "mccb200_moments_m2": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "mccb200_shard_owner_keys": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p]),
    "mccb200_shard_send_ids": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                         C.c_void_p]),
    "mccb200_scatter_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p,
                                       C.c_void_p]),
    "mccb200_gaussian_full_drift": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]),
    "mccb200_gaussian_full_drift_tc_supported": (C.c_int, [C.c_uint64, C.c_int, C.c_int]),
    "mccb200_gaussian_full_drift_tc_prepare_bytes": (C.c_size_t, [C.c_int]),
    "mccb200_gaussian_full_drift_tc_prepare": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mccb200_gaussian_full_drift_tc": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the C-ABI library (no fallback: raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    global LIB_PATH
    if os.environ.get("MCCB200_LIB"):          # experiments: a library built with other -D flags
        LIB_PATH = Path(os.environ["MCCB200_LIB"])
    if not LIB_PATH.exists():
        raise MccubeB200Error(
            f"{LIB_PATH} is missing: build it with `python -m mccube_b200._build` "
            "(there is no CPU fallback for the CUDA path)"
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def launch_count() -> int:
    return int(load().mccb200_launch_count())


def profile_enable(on: bool):
    load().mccb200_profile_enable(int(on))


def profile_read(max_entries: int = 4096):
    """[(phase, ms), ...] in launch order since the last profile_enable(True)."""
    names = (C.c_char_p * max_entries)()
    ms = (C.c_float * max_entries)()
    n = load().mccb200_profile_read(names, ms, max_entries)
    return [(names[i].decode(), float(ms[i])) for i in range(n)]


def exported_symbols():
    return list(_SIGNATURES)


def _check(rc: int, what: str):
    if rc != 0:
        msg = load().mccb200_last_error().decode(errors="replace")
        raise MccubeB200Error(f"{what} failed (status {rc}): {msg}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _require_cuda(*tensors):
    """Every tensor must live on the CURRENT CUDA device: the wrappers launch on its current stream and the
    workspace cache is keyed by it (use `with torch.cuda.device(t.device):` around calls for another GPU)."""
    cur = None
    for t in tensors:
        if t is None or isinstance(t, _DevicePtr):
            continue
        if not t.is_cuda:
            raise MccubeB200Error("mccube_b200 kernels need CUDA tensors (no CPU fallback)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise MccubeB200Error(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: "
                                  "wrap the call in `with torch.cuda.device(tensor.device)`")


def _i32c(t, what="index"):
    """Index / permutation / slot vectors are read as 32-bit words (uint32 bit patterns)."""
    if t is not None and (t.dtype not in (torch.int32, torch.uint32) or not t.is_contiguous()):
        raise MccubeB200Error(f"{what}: expected a contiguous int32 tensor, got {t.dtype} contiguous={t.is_contiguous()}")
    return t


def _f32w(t, n=None):
    """Optional weight vector: 1-D float32, any stride (the kernels read 4-byte words `stride` apart)."""
    if t is not None:
        if t.dtype != torch.float32 or t.dim() != 1:
            raise MccubeB200Error(f"weights: expected a 1-D float32 tensor, got {t.dtype} with {t.dim()} dims")
        if n is not None and t.numel() != n:
            raise MccubeB200Error(f"weights: expected {n} entries, got {t.numel()}")
    return t


def _f32c(t):
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise MccubeB200Error(f"expected contiguous float32 tensor, got {t.dtype} contiguous={t.is_contiguous()}")
    return t


# ------------------------------------------------------------------ descriptors
def make_rng(key, t, *, leaf=0, n_leaves=1, partitionable=False, fold_time=True) -> Rng:
    r = Rng()
    r.key[0], r.key[1] = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    r.key_dev = None
    r.t = float(np.float32(t))
    r.t_dev = None
    r.leaf, r.n_leaves = leaf, n_leaves
    r.partitionable = int(bool(partitionable))
    r.fold_time = int(bool(fold_time))
    return r


def make_drift(kind=0, f=None, mean=None, inv_var=None) -> Drift:
    d = Drift()
    d.kind = kind
    d.f, d.mean, d.inv_var = _ptr(f), _ptr(mean), _ptr(inv_var)
    d._keep = (f, mean, inv_var)
    return d


def make_cubature(k, scale=0.0, points=None, lam=None) -> Cubature:
    c = Cubature()
    c.k = int(k)
    c.scale = float(np.float32(scale))
    c.points, c.lam = _ptr(points), _ptr(lam)
    c._keep = (points, lam)
    return c


def _dims_array(dims):
    arr = (C.c_int32 * len(dims))(*[int(x) for x in dims])
    return arr


# ------------------------------------------------------------------ workspace cache
_ws_cache: dict = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream): the C ABI never allocates, and two
    streams must never share scratch."""
    key = (torch.device(device).index or 0, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        _ws_cache[key] = buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
    return buf


def release_workspace():
    _ws_cache.clear()


# ------------------------------------------------------------------ calls
def mc_indices(rng: Rng, n_inputs: int, count: int, replace: bool, device, out=None) -> torch.Tensor:
    lib = load()
    out = torch.empty(count, dtype=torch.int32, device=device) if out is None else out
    _require_cuda(out)
    nbytes = lib.mccb200_mc_indices_workspace_bytes(n_inputs, count, int(replace))
    ws = workspace(nbytes, device)
    _check(lib.mccb200_mc_indices(C.byref(rng), n_inputs, count, int(replace), out.data_ptr(), ws.data_ptr(),
                                  ws.numel(), _stream()), "mc_indices")
    return out


def mc_indices_slice(rng: Rng, n_inputs: int, count: int, first: int, size: int, device, x64: bool = False) -> torch.Tensor:
    """Entries [first, first + size) of the with_replacement=True index vector of `count` draws.
    x64: `jax_enable_x64` semantics (int64 randint, 64-bit draws), int64 result, n_inputs may exceed 2^32."""
    lib = load()
    if x64:
        out = torch.empty(size, dtype=torch.int64, device=device)
        ws = workspace(4096, device)
        _check(lib.mccb200_mc_indices_slice64(C.byref(rng), n_inputs, count, first, size, out.data_ptr(), ws.data_ptr(),
                                              ws.numel(), _stream()), "mc_indices_slice64")
        return out
    out = torch.empty(size, dtype=torch.int32, device=device)
    nbytes = lib.mccb200_mc_indices_workspace_bytes(n_inputs, count, 1)
    ws = workspace(nbytes, device)
    _check(lib.mccb200_mc_indices_slice(C.byref(rng), n_inputs, count, first, size, out.data_ptr(), ws.data_ptr(),
                                        ws.numel(), _stream()), "mc_indices_slice")
    return out


def sort_pairs(keys: torch.Tensor, values, key_bits: int = 32, want_keys: bool = False):
    """Stable argsort / sort_key_val of uint32 keys (int32 tensors reinterpreted)."""
    lib = load()
    _require_cuda(keys, values)
    n = keys.numel()
    vals_out = torch.empty(n, dtype=torch.int32, device=keys.device)
    keys_out = torch.empty(n, dtype=torch.int32, device=keys.device) if want_keys else None
    ws = workspace(lib.mccb200_sort_pairs_workspace_bytes(n), keys.device)
    _check(lib.mccb200_sort_pairs(keys.data_ptr(), _ptr(values), n, key_bits, _ptr(keys_out), vals_out.data_ptr(),
                                  ws.data_ptr(), ws.numel(), _stream()), "sort_pairs")
    return (keys_out, vals_out) if want_keys else vals_out


def expand_select(y0, d, weighted, drift: Drift, dt, cub: Cubature, idx, perm=None, out_per_part=1, in_per_part=1,
                  n_out=None, out=None):
    lib = load()
    _require_cuda(y0, idx, perm)
    _f32c(y0)
    _i32c(idx, "expand_select idx"); _i32c(perm, "expand_select perm")
    n, ld = y0.shape
    n_out = idx.numel() if n_out is None else n_out
    out = torch.empty((n_out, ld), dtype=torch.float32, device=y0.device) if out is None else out
    _check(lib.mccb200_expand_select(y0.data_ptr(), n, d, int(weighted), C.byref(drift), float(np.float32(dt)),
                                     C.byref(cub), idx.data_ptr(), _ptr(perm), out_per_part, in_per_part, n_out,
                                     out.data_ptr(), _stream()), "expand_select")
    return out


def kd_partition_levels(n: int, leaf_size: int) -> int:
    return int(load().mccb200_kd_partition_levels(int(n), int(leaf_size)))


def kd_partition(y, d: int, levels: int) -> torch.Tensor:
    """Row indices of `y` [n, ld] ordered leaf by leaf of a `levels`-deep median-split tree on the first d columns."""
    lib = load()
    _require_cuda(y)
    _f32c(y)
    n, ld = y.shape
    out = torch.empty(n, dtype=torch.int32, device=y.device)
    ws = workspace(lib.mccb200_kd_partition_workspace_bytes(n, d, levels), y.device)
    _check(lib.mccb200_kd_partition(y.data_ptr(), n, d, ld, levels, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
           "kd_partition")
    return out


def expand_select_draw_supported(d: int, ld: int, weighted: bool) -> bool:
    return bool(load().mccb200_expand_select_draw_supported(int(d), int(ld), int(weighted)))


def expand_select_draw(y0, d, drift: Drift, dt, cub: Cubature, rng: Rng, n_out: int, out=None, idx_out=None):
    """`MonteCarloKernel(with_replacement=True)` + expand in one kernel: the child ids are drawn inside the gather
    (no index vector in HBM).  `idx_out` (int32 [n_out], optional) receives the drawn ids."""
    lib = load()
    _require_cuda(y0)
    _f32c(y0)
    n, ld = y0.shape
    out = torch.empty((n_out, ld), dtype=torch.float32, device=y0.device) if out is None else out
    if idx_out is not None:
        _require_cuda(idx_out)
        _i32c(idx_out, "expand_select_draw idx_out")
    ws = workspace(lib.mccb200_expand_select_draw_workspace_bytes(), y0.device)
    _check(lib.mccb200_expand_select_draw(y0.data_ptr(), n, d, C.byref(drift), float(np.float32(dt)), C.byref(cub),
                                          C.byref(rng), n_out, out.data_ptr(),
                                          idx_out.data_ptr() if idx_out is not None else None, ws.data_ptr(), ws.numel(),
                                          _stream()), "expand_select_draw")
    return out


def expand_select_pairs(y0, d, drift: Drift, dt, cub: Cubature, pairs, n_out=None, out=None, first=0):
    """expand_select reading the child ids from column 1 of packed (slot, id) pairs [n, 2] int32, rows
    [first, first + n_out)."""
    lib = load()
    _require_cuda(y0, pairs)
    _f32c(y0)
    _i32c(pairs, "expand_select_pairs")
    n, ld = y0.shape
    n_out = pairs.shape[0] - first if n_out is None else n_out
    out = torch.empty((n_out, ld), dtype=torch.float32, device=y0.device) if out is None else out
    _check(lib.mccb200_expand_select_strided(y0.data_ptr(), n, d, C.byref(drift), float(np.float32(dt)), C.byref(cub),
                                             pairs.data_ptr() + 8 * first + 4, 2, n_out, out.data_ptr(), _stream()),
           "expand_select_strided")
    return out


def expand_select_substeps(y0, d, drift: Drift, dts, scales, k, idx):
    """Fused n_substeps > 1 step (Hadamard, drift kind 0 / 2): only the selected children of the k^s
    expansion are computed."""
    lib = load()
    _require_cuda(y0, idx)
    _f32c(y0)
    s = len(dts)
    arr_dt = (C.c_float * s)(*[float(np.float32(x)) for x in dts])
    arr_sc = (C.c_float * s)(*[float(np.float32(x)) for x in scales])
    out = torch.empty((idx.numel(), y0.shape[1]), dtype=torch.float32, device=y0.device)
    _check(lib.mccb200_expand_select_substeps(y0.data_ptr(), y0.shape[0], d, C.byref(drift), s, arr_dt, arr_sc, int(k),
                                              idx.data_ptr(), idx.numel(), out.data_ptr(), _stream()),
           "expand_select_substeps")
    return out


class _DevicePtr:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view it."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 3}


def peer_alloc(shape, device) -> torch.Tensor:
    """float32 tensor over a plain cudaMalloc buffer (IPC-exportable, outside torch's caching allocator)."""
    lib = load()
    nbytes = 4 * int(np.prod(shape))
    ptr = C.c_void_p()
    with torch.cuda.device(device):
        _check(lib.mccb200_peer_alloc(nbytes, C.byref(ptr)), "peer_alloc")
        return torch.as_tensor(_DevicePtr(ptr.value, shape), device=device)


def peer_export(t: torch.Tensor) -> bytes:
    buf = C.create_string_buffer(64)
    _check(load().mccb200_peer_export(t.data_ptr(), buf), "peer_export")
    return buf.raw


def peer_open(handle: bytes, device) -> int:
    """Maps a peer's buffer for kernels of `device`; returns the device pointer."""
    ptr = C.c_void_p()
    with torch.cuda.device(device):
        _check(load().mccb200_peer_open(handle, C.byref(ptr)), "peer_open")
    return int(ptr.value)


def peer_close(ptr: int):
    _check(load().mccb200_peer_close(C.c_void_p(ptr)), "peer_close")


def peer_free(t: torch.Tensor):
    _check(load().mccb200_peer_free(C.c_void_p(t.data_ptr())), "peer_free")


def expand_gather_peers(peer_ptrs, n_ranks, rows_per_rank, d, drift: Drift, dt, cub: Cubature, idx, out):
    """Pull form: out[q] = child idx[q] (global child ids), parents read from the shards `peer_ptrs`."""
    lib = load()
    _require_cuda(peer_ptrs, idx, out)
    wide = idx.dtype == torch.int64
    _check(lib.mccb200_expand_gather_peers(peer_ptrs.data_ptr(), int(n_ranks), int(rows_per_rank), d, C.byref(drift),
                                           float(np.float32(dt)), C.byref(cub), None if wide else idx.data_ptr(),
                                           idx.data_ptr() if wide else None, idx.numel(), out.data_ptr(), _stream()),
           "expand_gather_peers")
    return out


def enable_peer_access(peer_device: int):
    _check(load().mccb200_enable_peer_access(int(peer_device)), "enable_peer_access")


def expand_scatter_peers(y0, d, drift: Drift, dt, cub: Cubature, idx, order, counts, rank, n_ranks, rows_per_rank,
                         peer_ptrs):
    """Fused expand + exchange: the selected children of this rank's parents go straight into the
    shards `peer_ptrs` (int64 device tensor of n_ranks base pointers) over peer memory."""
    lib = load()
    _require_cuda(y0, idx, order, counts, peer_ptrs)
    _f32c(y0)
    _check(lib.mccb200_expand_scatter_peers(y0.data_ptr(), y0.shape[0], d, C.byref(drift), float(np.float32(dt)),
                                            C.byref(cub), idx.data_ptr(), order.data_ptr(), counts.data_ptr(), int(rank),
                                            int(n_ranks), int(rows_per_rank), peer_ptrs.data_ptr(), _stream()),
           "expand_scatter_peers")


def expand_all(y0, d, weighted, drift: Drift, dt, cub: Cubature):
    lib = load()
    _require_cuda(y0)
    _f32c(y0)
    n, ld = y0.shape
    out = torch.empty((n * cub.k, ld), dtype=torch.float32, device=y0.device)
    _check(lib.mccb200_expand_all(y0.data_ptr(), n, d, int(weighted), C.byref(drift), float(np.float32(dt)),
                                  C.byref(cub), out.data_ptr(), _stream()), "expand_all")
    return out


def gather_rows(p, idx):
    lib = load()
    _require_cuda(p, idx)
    _f32c(p)
    _i32c(idx, "gather_rows idx")
    out = torch.empty((idx.numel(), p.shape[1]), dtype=torch.float32, device=p.device)
    _check(lib.mccb200_gather_rows(p.data_ptr(), p.shape[1], idx.data_ptr(), idx.numel(), out.data_ptr(), _stream()),
           "gather_rows")
    return out


def normalise_weights_(y):
    lib = load()
    _require_cuda(y)
    _f32c(y)
    ws = torch.empty(1, dtype=torch.float64, device=y.device)
    _check(lib.mccb200_normalise_weights(y.data_ptr(), y.shape[0], y.shape[1], ws.data_ptr(), _stream()),
           "normalise_weights")
    return y


def normalised_copy(y, total=None):
    """A copy of `y` [n, ld] with the weight column divided by its sum (`total`: device float64 [1] holding that sum
    when the producer already has it)."""
    lib = load()
    _require_cuda(y)
    _f32c(y)
    out = torch.empty_like(y)
    ws = torch.empty(1, dtype=torch.float64, device=y.device)
    _check(lib.mccb200_normalised_copy(y.data_ptr(), y.shape[0], y.shape[1], _ptr(total), out.data_ptr(), ws.data_ptr(),
                                       _stream()), "normalised_copy")
    return out


def expand_select_wsum(y0, d, drift: Drift, dt, cub: Cubature, idx):
    """Weighted expand_select that also returns the sum of the child weights (device float64 [1])."""
    lib = load()
    _require_cuda(y0, idx)
    _f32c(y0)
    _i32c(idx, "expand_select_wsum idx")
    n, ld = y0.shape
    out = torch.empty((idx.numel(), ld), dtype=torch.float32, device=y0.device)
    wsum = torch.empty(1, dtype=torch.float64, device=y0.device)
    _check(lib.mccb200_expand_select_wsum(y0.data_ptr(), n, d, C.byref(drift), float(np.float32(dt)), C.byref(cub),
                                          idx.data_ptr(), idx.numel(), out.data_ptr(), wsum.data_ptr(), _stream()),
           "expand_select_wsum")
    return out, wsum


def stratified_keys(p, d, com):
    lib = load()
    _require_cuda(p, com)
    _f32c(p)
    keys = torch.empty(p.shape[0], dtype=torch.int32, device=p.device)
    _check(lib.mccb200_stratified_keys(p.data_ptr(), p.shape[0], d, p.shape[1], com.data_ptr(), keys.data_ptr(),
                                       _stream()), "stratified_keys")
    return keys


def box_bounds(p, dims, bits):
    lib = load()
    _require_cuda(p)
    _f32c(p)
    bounds = torch.empty(2 * len(dims), dtype=torch.float32, device=p.device)
    ws = workspace(lib.mccb200_box_bounds_workspace_bytes(), p.device)
    _check(lib.mccb200_box_bounds(p.data_ptr(), p.shape[0], p.shape[1], _dims_array(dims), len(dims), bits,
                                  bounds.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "box_bounds")
    return bounds


def box_keys(p, dims, bits, bounds):
    lib = load()
    _require_cuda(p, bounds)
    _f32c(p)
    keys = torch.empty(p.shape[0], dtype=torch.int32, device=p.device)
    _check(lib.mccb200_box_keys(p.data_ptr(), p.shape[0], p.shape[1], _dims_array(dims), len(dims), bits,
                                bounds.data_ptr(), keys.data_ptr(), _stream()), "box_keys")
    return keys


def box_child_bounds(y0, d, drift, dt, cub, dims, bits, keycols=None):
    lib = load()
    _require_cuda(y0, keycols)
    _f32c(y0)
    bounds = torch.empty(2 * len(dims), dtype=torch.float32, device=y0.device)
    ws = workspace(lib.mccb200_box_bounds_workspace_bytes(), y0.device)
    _check(lib.mccb200_box_child_bounds(y0.data_ptr(), y0.shape[0], d, y0.shape[1], C.byref(drift),
                                        float(np.float32(dt)), C.byref(cub), _dims_array(dims), len(dims), bits,
                                        _ptr(keycols), bounds.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
           "box_child_bounds")
    return bounds


def box_child_keys(y0, d, drift, dt, cub, dims, bits, bounds):
    lib = load()
    _require_cuda(y0, bounds)
    _f32c(y0)
    keys = torch.empty(y0.shape[0] * cub.k, dtype=torch.int32, device=y0.device)
    _check(lib.mccb200_box_child_keys(y0.data_ptr(), y0.shape[0], d, y0.shape[1], C.byref(drift),
                                      float(np.float32(dt)), C.byref(cub), _dims_array(dims), len(dims), bits,
                                      bounds.data_ptr(), keys.data_ptr(), _stream()), "box_child_keys")
    return keys


def box_prk_supported(n, d, cub, dims, bits, in_per_part) -> bool:
    """Whether the fused counting-sort kernel covers this shape (Hadamard, n_dims*bits <= 12,
    k <= 1024); the workspace query returns 0 otherwise."""
    if cub.points is not None:
        return False
    return load().mccb200_box_prk_step_workspace_bytes(n, cub.k, len(dims), bits, in_per_part) > 0


def box_prk_step(y0, d, drift, dt, cub, dims, bits, bounds, idx_local, in_per_part, out=None):
    lib = load()
    _require_cuda(y0, bounds, idx_local)
    _f32c(y0)
    n = y0.shape[0]
    out = torch.empty_like(y0) if out is None else out
    nbytes = lib.mccb200_box_prk_step_workspace_bytes(n, cub.k, len(dims), bits, in_per_part)
    ws = workspace(nbytes, y0.device)
    _check(lib.mccb200_box_prk_step(y0.data_ptr(), n, d, C.byref(drift), float(np.float32(dt)), C.byref(cub),
                                    _dims_array(dims), len(dims), bits, bounds.data_ptr(), idx_local.data_ptr(),
                                    idx_local.numel(), in_per_part, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                    _stream()), "box_prk_step")
    return out


_prk_tables_cache: dict = {}


def box_prk_tables(k, dims, bits, device):
    """Class tables of the fused Box step: a function of (k, dims, bits) only, built once per
    device and kept (the cubature and the partitioner do not change during a solve)."""
    lib = load()
    key = (torch.device(device).index or 0, int(k), tuple(dims), int(bits))
    hit = _prk_tables_cache.get(key)
    if hit is not None:
        return hit
    nbytes = lib.mccb200_box_prk_tables_bytes(int(k), len(dims), int(bits))
    if nbytes == 0:
        raise MccubeB200Error(f"box_prk_tables: unsupported shape k={k} dims={dims} bits={bits}")
    buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    _check(lib.mccb200_box_prk_tables(int(k), _dims_array(dims), len(dims), int(bits), buf.data_ptr(), nbytes,
                                      _stream()), "box_prk_tables")
    _prk_tables_cache[key] = buf
    return buf


def box_prk_selection(idx_local, in_per_part):
    """Prefix-count tables of the shared local index vector (runs on the current stream)."""
    lib = load()
    _require_cuda(idx_local)
    nbytes = lib.mccb200_box_prk_selection_bytes(int(in_per_part), idx_local.numel())
    if nbytes == 0:
        raise MccubeB200Error("box_prk_selection: partition sizes out of range")
    sel = torch.empty(nbytes, dtype=torch.uint8, device=idx_local.device)
    _check(lib.mccb200_box_prk_selection(idx_local.data_ptr(), idx_local.numel(), int(in_per_part), sel.data_ptr(),
                                         nbytes, _stream()), "box_prk_selection")
    return sel


def box_prk_count(y0, d, drift, dt, cub, dims, bits, bounds, in_per_part, tables, keycols=None):
    """Counting pass + scan; returns the workspace that box_prk_rank_gather consumes."""
    lib = load()
    _require_cuda(y0, bounds, tables)
    _f32c(y0)
    n = y0.shape[0]
    nbytes = lib.mccb200_box_prk_step_workspace_bytes(n, cub.k, len(dims), bits, in_per_part)
    ws = workspace(nbytes, y0.device)
    _check(lib.mccb200_box_prk_count(y0.data_ptr(), n, d, C.byref(drift), float(np.float32(dt)), C.byref(cub),
                                     _dims_array(dims), len(dims), bits, bounds.data_ptr(), in_per_part,
                                     tables.data_ptr(), _ptr(keycols), ws.data_ptr(), ws.numel(), _stream()),
           "box_prk_count")
    return ws


def box_child_minmax(y0, d, drift, dt, cub, dims, bits, keycols=None):
    """[min..., max...] of every key dim over the implicit children (2 * len(dims) float32)."""
    lib = load()
    _require_cuda(y0, keycols)
    _f32c(y0)
    out = torch.empty(2 * len(dims), dtype=torch.float32, device=y0.device)
    ws = workspace(lib.mccb200_box_bounds_workspace_bytes(), y0.device)
    _check(lib.mccb200_box_child_minmax(y0.data_ptr(), y0.shape[0], d, y0.shape[1], C.byref(drift),
                                        float(np.float32(dt)), C.byref(cub), _dims_array(dims), len(dims), bits,
                                        _ptr(keycols), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
           "box_child_minmax")
    return out


def box_prk_local_key_starts(ws, n, k, n_dims, bits, in_per_part, out_per_part):
    """View (int32 bit patterns of uint32) of the local sorted position of every key's first child inside the
    workspace `box_prk_count` returned."""
    lib = load()
    n_keys = C.c_uint32(0)
    off = lib.mccb200_box_prk_key_start_offset(n, k, n_dims, bits, in_per_part, out_per_part, C.byref(n_keys))
    if off == C.c_size_t(-1).value:
        raise MccubeB200Error("box_prk_local_key_starts: unsupported shape")
    return ws[off:off + 4 * n_keys.value].view(torch.int32)


def box_prk_rank_global(y0, d, drift, dt, cub, dims, bits, out_per_part, in_per_part, tables, sel, ws, key_start,
                        idx_slots):
    """Ranking pass with global key starts: fills idx_slots[global slot] = local child id for this shard's hits."""
    lib = load()
    _require_cuda(y0, tables, sel, ws, key_start, idx_slots)
    _check(lib.mccb200_box_prk_rank_global(y0.data_ptr(), y0.shape[0], d, C.byref(drift), float(np.float32(dt)),
                                           C.byref(cub), _dims_array(dims), len(dims), bits, out_per_part, in_per_part,
                                           tables.data_ptr(), sel.data_ptr(), key_start.data_ptr(), idx_slots.data_ptr(),
                                           ws.data_ptr(), ws.numel(), _stream()), "box_prk_rank_global")
    return idx_slots


def compact_slots(idx_slots, slot_list, cursor):
    """slot_list[: cursor] = the slots s with idx_slots[s] != -1 (any order); cursor: 1-element int32 view, zeroed."""
    lib = load()
    _require_cuda(idx_slots, slot_list, cursor)
    _check(lib.mccb200_compact_slots(idx_slots.data_ptr(), idx_slots.numel(), slot_list.data_ptr(), slot_list.numel(),
                                     cursor.data_ptr(), _stream()), "compact_slots")


def expand_scatter_slots(y0, d, drift: Drift, dt, cub: Cubature, idx_slots, slot_list, counts, rank, n_ranks,
                         rows_per_rank, peer_ptrs):
    """Fused expand + exchange driven by a slot list: child idx_slots[s] of this rank's parents goes to row
    s % rows_per_rank of shard s // rows_per_rank."""
    lib = load()
    _require_cuda(y0, idx_slots, slot_list, counts, peer_ptrs)
    _f32c(y0)
    _check(lib.mccb200_expand_scatter_slots(y0.data_ptr(), y0.shape[0], d, C.byref(drift), float(np.float32(dt)),
                                            C.byref(cub), idx_slots.data_ptr(), slot_list.data_ptr(), counts.data_ptr(),
                                            int(rank), int(n_ranks), int(rows_per_rank), peer_ptrs.data_ptr(), _stream()),
           "expand_scatter_slots")


def box_prk_rank_gather(y0, d, drift, dt, cub, dims, bits, out_per_part, in_per_part, tables, sel, ws, out=None,
                        keycols_out=None):
    """`keycols_out` [n_out, 4] float32 (optional): receives y1[:, dims[0]:dims[0]+4] packed, for the
    next step's box_child_bounds / box_prk_count."""
    lib = load()
    _require_cuda(y0, tables, sel, ws)
    n = y0.shape[0]
    n_out = n * cub.k // in_per_part * out_per_part
    out = torch.empty((n_out, y0.shape[1]), dtype=torch.float32, device=y0.device) if out is None else out
    _check(lib.mccb200_box_prk_rank_gather(y0.data_ptr(), n, d, C.byref(drift), float(np.float32(dt)), C.byref(cub),
                                           _dims_array(dims), len(dims), bits, out_per_part, in_per_part,
                                           tables.data_ptr(), sel.data_ptr(), out.data_ptr(), _ptr(keycols_out),
                                           ws.data_ptr(), ws.numel(), _stream()), "box_prk_rank_gather")
    return out


def column_means(p, d, weights=None):
    """(mean[d] float64, weight_total) -- center_of_mass."""
    lib = load()
    _require_cuda(p, weights)
    _f32c(p)
    n, ld = p.shape
    _f32w(weights, n)
    sums = torch.zeros(d + 1, dtype=torch.float64, device=p.device)
    ws = 1 if weights is None else weights.stride(0)
    _check(lib.mccb200_moments_sum(p.data_ptr(), n, d, ld, _ptr(weights), ws, sums.data_ptr(), _stream()),
           "moments_sum")
    return sums[:d] / sums[d], sums[d]


def second_moment(p, d, centre, weights=None):
    """sum_r w_r (p_r - centre)(p_r - centre)^T as float64 [d, d] (accumulated on the GPU)."""
    lib = load()
    _require_cuda(p, weights, centre)
    _f32c(p); _f32c(centre)
    n, ld = p.shape
    _f32w(weights, n)
    w_stride = 1 if weights is None else weights.stride(0)
    m2 = torch.zeros((d, d), dtype=torch.float64, device=p.device)
    _check(lib.mccb200_moments_m2(p.data_ptr(), n, d, ld, _ptr(weights), w_stride, centre.data_ptr(), m2.data_ptr(),
                                  _stream()), "moments_m2")
    return m2


def moments(p, d, weights=None):
    """(sum[d], weight_total, m2[d, d]) in float64 about the weighted mean."""
    lib = load()
    _require_cuda(p, weights)
    _f32c(p)
    n, ld = p.shape
    _f32w(weights, n)
    w_stride = 1 if weights is None else weights.stride(0)
    sums = torch.zeros(d + 1, dtype=torch.float64, device=p.device)
    _check(lib.mccb200_moments_sum(p.data_ptr(), n, d, ld, _ptr(weights), w_stride, sums.data_ptr(), _stream()),
           "moments_sum")
    mean64 = sums[:d] / sums[d]
    centre = mean64.to(torch.float32).contiguous()
    m2 = torch.zeros((d, d), dtype=torch.float64, device=p.device)
    _check(lib.mccb200_moments_m2(p.data_ptr(), n, d, ld, _ptr(weights), w_stride, centre.data_ptr(), m2.data_ptr(),
                                  _stream()), "moments_m2")
    # exact shift back from the fp32 centre to the fp64 mean: sum w (x-c)(x-c)^T - W (m-c)(m-c)^T
    delta = mean64 - centre.to(torch.float64)
    m2 = m2 - sums[d] * torch.outer(delta, delta)
    return mean64, sums[d], m2


def gaussian_full_drift(y, d, mean, prec):
    lib = load()
    _require_cuda(y, mean, prec)
    _f32c(y)
    f = torch.empty((y.shape[0], d), dtype=torch.float32, device=y.device)
    _check(lib.mccb200_gaussian_full_drift(y.data_ptr(), y.shape[0], d, y.shape[1], mean.data_ptr(), prec.data_ptr(),
                                           f.data_ptr(), _stream()), "gaussian_full_drift")
    return f


def gaussian_full_drift_tc_supported(n, d, ld) -> bool:
    return bool(load().mccb200_gaussian_full_drift_tc_supported(int(n), int(d), int(ld)))


def gaussian_full_drift_tc_prepare(prec):
    """3xTF32 split of the precision matrix, [2, d, d] (hi / lo of prec^T); done once per target."""
    lib = load()
    _require_cuda(prec)
    d = prec.shape[0]
    split = torch.empty((2, d, d), dtype=torch.float32, device=prec.device)
    _check(lib.mccb200_gaussian_full_drift_tc_prepare(prec.contiguous().data_ptr(), d, split.data_ptr(), _stream()),
           "gaussian_full_drift_tc_prepare")
    return split


def gaussian_full_drift_tc(y, d, mean, prec_split):
    """-(y - mean) @ prec on the tensor cores (tcgen05, 3xTF32, fp32 accumulation)."""
    lib = load()
    _require_cuda(y, mean, prec_split)
    _f32c(y)
    f = torch.empty((y.shape[0], d), dtype=torch.float32, device=y.device)
    _check(lib.mccb200_gaussian_full_drift_tc(y.data_ptr(), y.shape[0], d, y.shape[1], mean.data_ptr(),
                                              prec_split.data_ptr(), f.data_ptr(), _stream()), "gaussian_full_drift_tc")
    return f


def shard_owner_keys(idx, n_loc, k, n_ranks):
    """(keys[n_tot] = owner rank of every output slot, counts[n_ranks, n_ranks] = slots per (owner, dest))."""
    lib = load()
    _require_cuda(idx)
    keys = torch.empty(idx.numel(), dtype=torch.int32, device=idx.device)
    counts = torch.empty((n_ranks, n_ranks), dtype=torch.int32, device=idx.device)
    _check(lib.mccb200_shard_owner_keys(idx.data_ptr(), idx.numel(), n_loc, k, n_ranks, keys.data_ptr(),
                                        counts.data_ptr(), _stream()), "shard_owner_keys")
    return keys, counts


def shard_bucket_ids(idx, first_slot, n_loc, k, n_ranks, packed: bool = False):
    """Send buffers of the id exchange from this rank's slice `idx` (int32 bit patterns or int64) of the global index
    vector: (slots [n] int32, ids [n] int32 = owner-local child ids, counts [n_ranks] int32), both grouped by owner
    rank and ascending in slot inside a group; `packed` returns (pairs [n, 2] int32, counts) for one all-to-all."""
    lib = load()
    _require_cuda(idx)
    if idx.dtype not in (torch.int32, torch.uint32, torch.int64) or not idx.is_contiguous():
        raise MccubeB200Error(f"shard_bucket_ids: expected a contiguous int32 / int64 index slice, got {idx.dtype}")
    n = idx.numel()
    slots = torch.empty((n, 2) if packed else (n,), dtype=torch.int32, device=idx.device)
    ids = None if packed else torch.empty(n, dtype=torch.int32, device=idx.device)
    counts = torch.empty(n_ranks, dtype=torch.int32, device=idx.device)
    nbytes = lib.mccb200_shard_bucket_ids_workspace_bytes(int(n_ranks))
    if nbytes == 0:
        raise MccubeB200Error(f"shard_bucket_ids: unsupported rank count {n_ranks}")
    ws = workspace(nbytes, idx.device)
    _check(lib.mccb200_shard_bucket_ids(idx.data_ptr(), int(idx.dtype == torch.int64), n, int(first_slot), int(n_loc),
                                        int(k), int(n_ranks), slots.data_ptr(), _ptr(ids), counts.data_ptr(),
                                        ws.data_ptr(), ws.numel(), _stream()), "shard_bucket_ids")
    return (slots, counts) if packed else (slots, ids, counts)


def shard_bucket_regions(idx, first_slot, n_loc, k, n_ranks, region_cap, out_pairs, out_counts):
    """`shard_bucket_ids(packed=True)` with owner g's group written at out_pairs[g * region_cap :] (caller-provided
    buffers: [n_ranks * region_cap, 2] int32 and [n_ranks] int32)."""
    lib = load()
    _require_cuda(idx, out_pairs, out_counts)
    if idx.dtype not in (torch.int32, torch.uint32, torch.int64) or not idx.is_contiguous():
        raise MccubeB200Error(f"shard_bucket_regions: expected a contiguous int32 / int64 index slice, got {idx.dtype}")
    _i32c(out_pairs, "shard_bucket_regions pairs"); _i32c(out_counts, "shard_bucket_regions counts")
    if out_pairs.numel() < 2 * n_ranks * region_cap or out_counts.numel() < n_ranks:
        raise MccubeB200Error("shard_bucket_regions: output buffers too small")
    nbytes = lib.mccb200_shard_bucket_ids_workspace_bytes(int(n_ranks))
    ws = workspace(nbytes, idx.device)
    _check(lib.mccb200_shard_bucket_ids_regions(idx.data_ptr(), int(idx.dtype == torch.int64), idx.numel(), int(first_slot),
                                                int(n_loc), int(k), int(n_ranks), int(region_cap), out_pairs.data_ptr(),
                                                None, out_counts.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
           "shard_bucket_ids_regions")


def peer_copy(dst_ptr: int, src: torch.Tensor, nbytes: int):
    """cudaMemcpyAsync D2D of the first `nbytes` bytes of `src` to the (possibly peer-mapped) address `dst_ptr`."""
    _require_cuda(src)
    _check(load().mccb200_peer_copy(int(dst_ptr), src.data_ptr(), int(nbytes), _stream()), "peer_copy")


class expand_residency:
    """`with expand_residency(3): ...` -- the gather kernels launched inside leave a quarter of every SM free."""

    def __init__(self, ctas_per_sm: int):
        self.ctas = int(ctas_per_sm)

    def __enter__(self):
        load().mccb200_set_expand_residency(self.ctas)

    def __exit__(self, *exc):
        load().mccb200_set_expand_residency(0)


class side_residency:
    """`with side_residency(1): ...` -- the index-draw and id-bucketing kernels launched inside use one CTA per SM, so
    that they fit beside a gather running under `expand_residency(3)`."""

    def __init__(self, ctas_per_sm: int):
        self.ctas = int(ctas_per_sm)

    def __enter__(self):
        load().mccb200_set_side_residency(self.ctas)

    def __exit__(self, *exc):
        load().mccb200_set_side_residency(0)


def unzip_slots_(pairs, first: int, count: int, out):
    """out[:count] = pairs[first : first + count, 0]  (the slot column of packed (slot, id) pairs)."""
    lib = load()
    _require_cuda(pairs, out)
    _i32c(pairs, "unzip_slots"); _i32c(out, "unzip_slots out")
    _check(lib.mccb200_unzip_pairs(pairs.data_ptr() + 8 * int(first), int(count), out.data_ptr(), None, _stream()),
           "unzip_pairs")
    return out


def unzip_pairs(pairs):
    """(slots [n], ids [n]) from packed pairs [n, 2] int32."""
    lib = load()
    _require_cuda(pairs)
    _i32c(pairs, "unzip_pairs")
    n = pairs.shape[0]
    slots = torch.empty(n, dtype=torch.int32, device=pairs.device)
    ids = torch.empty(n, dtype=torch.int32, device=pairs.device)
    _check(lib.mccb200_unzip_pairs(pairs.data_ptr(), n, slots.data_ptr(), ids.data_ptr(), _stream()), "unzip_pairs")
    return slots, ids


def shard_send_ids(idx, order, start, count, child_offset):
    lib = load()
    _require_cuda(idx, order)
    out = torch.empty(count, dtype=torch.int32, device=idx.device)
    _check(lib.mccb200_shard_send_ids(idx.data_ptr(), order.data_ptr(), start, count, child_offset, out.data_ptr(),
                                      _stream()), "shard_send_ids")
    return out


def scatter_rows_(out, rows, slots, slot_offset=0):
    """out[slots[r] - slot_offset, :] = rows[r, :]"""
    lib = load()
    _require_cuda(out, rows, slots)
    _f32c(rows); _f32c(out)
    _i32c(slots, "scatter_rows slots")
    _check(lib.mccb200_scatter_rows(rows.data_ptr(), rows.shape[1], slots.data_ptr(), slot_offset, rows.shape[0],
                                    out.data_ptr(), _stream()), "scatter_rows")
    return out


def weighted_indices(rng: Rng, p: torch.Tensor, count: int, replace: bool) -> torch.Tensor:
    """Index vector of jr.choice(key, a, (count,), replace, p) for a probability vector p
    (float32, possibly a strided column view)."""
    lib = load()
    _require_cuda(p)
    if p.dtype != torch.float32 or p.dim() != 1:
        raise MccubeB200Error("weighted_indices: p must be a 1-D float32 tensor")
    n = p.numel()
    ws = workspace(lib.mccb200_weighted_workspace_bytes(n), p.device)
    if replace:
        out = torch.empty(count, dtype=torch.int32, device=p.device)
        _check(lib.mccb200_weighted_choice_replace(C.byref(rng), p.data_ptr(), p.stride(0), n, count, out.data_ptr(),
                                                   ws.data_ptr(), ws.numel(), _stream()), "weighted_choice_replace")
        return out
    keys = torch.empty(n, dtype=torch.int32, device=p.device)
    _check(lib.mccb200_gumbel_keys(C.byref(rng), p.data_ptr(), p.stride(0), n, keys.data_ptr(), ws.data_ptr(),
                                   ws.numel(), _stream()), "gumbel_keys")
    return sort_pairs(keys, None, 32)[:count].contiguous()
