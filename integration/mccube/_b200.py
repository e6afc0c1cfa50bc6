"""mccube/_b200.py -- what a maintainer of tttc3/MCCube adds to run `MCCSolver.step` on the B200 kernels.

Drop this file into the reference's `mccube/` package and add to `MCCSolver.step`
(/root/reference/mccube/_solvers.py:115, before the existing body):

    from . import _b200
    if _b200.can_fuse(self, terms, y0):
        return _b200.fused_step(self, terms, t0, t1, y0, args, solver_state)

Everything below is `jax.ffi.ffi_call` on the handlers of libmccube_b200_ffi.so
(mccube_b200/csrc/xla_ffi_shim.cc, one per C-ABI entry point of include/mccube_b200.h).

STATUS: UNVERIFIED UNDER REAL JAX -- jax / diffrax / equinox are not installable in the image this was written
in.  The argument lists are kept in lock step with the shim by tests/test_ffi_shim.py (target
names) and the shim itself is compiled against the C ABI there; the same C ABI is what the tested PyTorch host
layer (mccube_b200/_lib.py) calls.
"""
from __future__ import annotations

import ctypes

import numpy as np

try:  # pragma: no cover - not importable where this repo is developed
    import jax
    import jax.numpy as jnp
except ImportError:  # keeps the file importable for the text checks of tests/test_ffi_shim.py
    jax = jnp = None

_TARGETS = (
    "mccb200_mc_indices_ffi", "mccb200_expand_select_ffi", "mccb200_expand_select_draw_ffi", "mccb200_kd_partition_ffi", "mccb200_expand_all_ffi", "mccb200_box_child_bounds_ffi",
    "mccb200_box_prk_tables_ffi", "mccb200_box_prk_selection_ffi", "mccb200_box_prk_count_ffi",
    "mccb200_box_prk_rank_gather_ffi", "mccb200_box_prk_step_ffi", "mccb200_gaussian_full_drift_tc_prepare_ffi",
    "mccb200_gaussian_full_drift_tc_ffi", "mccb200_gaussian_full_drift_ffi", "mccb200_moments_sum_ffi",
    "mccb200_moments_m2_ffi", "mccb200_stratified_keys_ffi", "mccb200_sort_pairs_ffi", "mccb200_gather_rows_ffi",
    "mccb200_normalise_weights_ffi",
)
_core = None


def _load():
    """Registers the FFI targets once; `_core` answers the host-side size queries."""
    global _core
    if _core is not None:
        return _core
    ffi_lib = ctypes.CDLL("libmccube_b200_ffi.so")
    for name in _TARGETS:
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(ffi_lib, name)), platform="CUDA")
    _core = ctypes.CDLL("libmccube_b200.so")
    for fn in ("mccb200_mc_indices_workspace_bytes", "mccb200_box_prk_step_workspace_bytes",
               "mccb200_box_prk_tables_bytes", "mccb200_box_prk_selection_bytes", "mccb200_box_bounds_workspace_bytes",
               "mccb200_sort_pairs_workspace_bytes", "mccb200_gaussian_full_drift_tc_prepare_bytes",
               "mccb200_expand_select_draw_workspace_bytes", "mccb200_kd_partition_workspace_bytes"):
        getattr(_core, fn).restype = ctypes.c_size_t
    return _core


def _u8(nbytes):
    return jax.ShapeDtypeStruct((int(nbytes),), jnp.uint8)


def _f32(*shape):
    return jax.ShapeDtypeStruct(tuple(int(s) for s in shape), jnp.float32)


_DUMMY = None


def _drift_args(drift_kind, f=None, mean=None, inv_var=None):
    """(f, mean, inv_var) with 1-element dummies for the buffers a drift kind does not use."""
    global _DUMMY
    if _DUMMY is None:
        _DUMMY = jnp.zeros((1,), jnp.float32)
    return (f if drift_kind == 1 else _DUMMY, mean if drift_kind == 2 else _DUMMY, inv_var if drift_kind == 2 else _DUMMY)


# ---------------------------------------------------------------------------------------------- A8
def mc_indices(key, t, n_inputs, count, replace, leaf=0, n_leaves=1):
    """Index form of `jr.choice(key_at(t), p, (count,), replace)` (mccube/_kernels/random.py:56-65)."""
    core = _load()
    ws = core.mccb200_mc_indices_workspace_bytes(ctypes.c_uint64(n_inputs), ctypes.c_uint64(count), int(replace))
    idx, _ = jax.ffi.ffi_call("mccb200_mc_indices_ffi", (jax.ShapeDtypeStruct((count,), jnp.uint32), _u8(ws)))(
        jax.random.key_data(key).astype(jnp.uint32), jnp.asarray(t, jnp.float32).reshape(1),
        n_inputs=np.int64(n_inputs), replace=np.int64(replace), leaf=np.int64(leaf), n_leaves=np.int64(n_leaves),
        partitionable=np.int64(bool(jax.config.jax_threefry_partitionable)))
    return idx


# ---------------------------------------------------------------------------------------------- A1-A5
def expand_select(y0, idx, k, dt, scale, drift_kind=0, weighted=False, f=None, mean=None, inv_var=None):
    """y1[r] = y0[i] + (dt * f(y0[i]) + noise[j]), (i, j) = divmod(idx[r], k) (mccube/_solvers.py:125-144)."""
    _load()
    return jax.ffi.ffi_call("mccb200_expand_select_ffi", _f32(idx.shape[0], y0.shape[1]))(
        y0, *_drift_args(drift_kind, f, mean, inv_var), idx, drift_kind=np.int64(drift_kind),
        weighted=np.int64(weighted), k=np.int64(k), dt=np.float32(dt), scale=np.float32(scale))


def expand_select_draw(y0, key, t, n_out, k, dt, scale, drift_kind=0, f=None, mean=None, inv_var=None):
    """MonteCarloKernel(with_replacement=True) fused with the expansion: the child ids
    jr.choice(step_key(key, t), n * k, (n_out,), replace=True) are drawn inside the gather kernel (random.py:45-66)."""
    core = _load()
    ws = core.mccb200_expand_select_draw_workspace_bytes()
    y1, _ = jax.ffi.ffi_call("mccb200_expand_select_draw_ffi", (_f32(n_out, y0.shape[1]), _u8(ws)))(
        y0, *_drift_args(drift_kind, f, mean, inv_var), jax.random.key_data(key).astype(jnp.uint32),
        jnp.asarray(t, jnp.float32).reshape(1), drift_kind=np.int64(drift_kind), k=np.int64(k), dt=np.float32(dt),
        scale=np.float32(scale), partitionable=np.int64(bool(jax.config.jax_threefry_partitionable)))
    return y1


def kd_partition(y, leaf_size, weighted=False):
    """Drop-in for the `jax.pure_callback` into sklearn of BinaryTreePartitioningKernel (mccube/_kernels/tree.py:44-58):
    idx[n // leaf_size, leaf_size] of a median-split tree built on the device (same node point sets as sklearn's)."""
    core = _load()
    n, ld = y.shape
    d = ld - (1 if weighted else 0)
    levels = core.mccb200_kd_partition_levels(ctypes.c_uint64(n), ctypes.c_uint64(leaf_size))
    ws = core.mccb200_kd_partition_workspace_bytes(ctypes.c_uint64(n), d, levels)
    idx, _ = jax.ffi.ffi_call("mccb200_kd_partition_ffi", (jax.ShapeDtypeStruct((n,), jnp.uint32), _u8(ws)))(
        y, d=np.int64(d), levels=np.int64(levels))
    return idx.reshape(-1, leaf_size)


def expand_all(y0, k, dt, scale, drift_kind=0, weighted=False, f=None, mean=None, inv_var=None):
    _load()
    return jax.ffi.ffi_call("mccb200_expand_all_ffi", _f32(y0.shape[0] * k, y0.shape[1]))(
        y0, *_drift_args(drift_kind, f, mean, inv_var), drift_kind=np.int64(drift_kind), weighted=np.int64(weighted),
        k=np.int64(k), dt=np.float32(dt), scale=np.float32(scale))


def gather_rows(p, idx):
    _load()
    return jax.ffi.ffi_call("mccb200_gather_rows_ffi", _f32(idx.shape[0], p.shape[1]))(p, idx)


def normalise_weights(y):
    _load()
    out, _ = jax.ffi.ffi_call("mccb200_normalise_weights_ffi", (_f32(*y.shape), jax.ShapeDtypeStruct((1,), jnp.float64)),
                              input_output_aliases={0: 0})(y)
    return out


# ---------------------------------------------------------------------------------------------- Box + PRK
def box_prk_tables(k, dims, bits):
    core = _load()
    nbytes = core.mccb200_box_prk_tables_bytes(int(k), len(dims), int(bits))
    return jax.ffi.ffi_call("mccb200_box_prk_tables_ffi", _u8(nbytes))(
        k=np.int64(k), bits=np.int64(bits), dims=np.asarray(dims, np.int32))


def box_child_bounds(y0, k, dt, scale, dims, bits, drift_kind=0, f=None, mean=None, inv_var=None, keycols=None):
    core = _load()
    ws = core.mccb200_box_bounds_workspace_bytes()
    bounds, _ = jax.ffi.ffi_call("mccb200_box_child_bounds_ffi", (_f32(2 * len(dims)), _u8(ws)))(
        y0, *_drift_args(drift_kind, f, mean, inv_var), keycols if keycols is not None else _DUMMY,
        drift_kind=np.int64(drift_kind), k=np.int64(k), dt=np.float32(dt), scale=np.float32(scale), bits=np.int64(bits),
        has_keycols=np.int64(keycols is not None), dims=np.asarray(dims, np.int32))
    return bounds


def box_prk_selection(idx_local, in_per_part):
    core = _load()
    nbytes = core.mccb200_box_prk_selection_bytes(ctypes.c_uint64(in_per_part), ctypes.c_uint64(idx_local.shape[0]))
    return jax.ffi.ffi_call("mccb200_box_prk_selection_ffi", _u8(nbytes))(idx_local, in_per_part=np.int64(in_per_part))


def box_prk_count(y0, bounds, tables, k, dt, scale, dims, bits, in_per_part, drift_kind=0, f=None, mean=None,
                  inv_var=None, keycols=None):
    core = _load()
    ws = core.mccb200_box_prk_step_workspace_bytes(ctypes.c_uint64(y0.shape[0]), int(k), len(dims), int(bits),
                                                   ctypes.c_uint64(in_per_part))
    return jax.ffi.ffi_call("mccb200_box_prk_count_ffi", _u8(ws))(
        y0, *_drift_args(drift_kind, f, mean, inv_var), bounds, tables, keycols if keycols is not None else _DUMMY,
        drift_kind=np.int64(drift_kind), k=np.int64(k), dt=np.float32(dt), scale=np.float32(scale), bits=np.int64(bits),
        in_per_part=np.int64(in_per_part), has_keycols=np.int64(keycols is not None), dims=np.asarray(dims, np.int32))


def box_prk_rank_gather(y0, tables, sel, workspace, k, dt, scale, dims, bits, out_per_part, in_per_part, drift_kind=0,
                        f=None, mean=None, inv_var=None, want_keycols=True):
    """Returns (y1, keycols_out): the recombined state and its packed key columns for the next step."""
    _load()
    n_out = y0.shape[0] * k // in_per_part * out_per_part
    y1, keycols_out, _ = jax.ffi.ffi_call(
        "mccb200_box_prk_rank_gather_ffi", (_f32(n_out, y0.shape[1]), _f32(n_out, 4), _u8(workspace.shape[0])),
        input_output_aliases={6: 2})(
        y0, *_drift_args(drift_kind, f, mean, inv_var), tables, sel, workspace, drift_kind=np.int64(drift_kind),
        k=np.int64(k), dt=np.float32(dt), scale=np.float32(scale), bits=np.int64(bits),
        out_per_part=np.int64(out_per_part), in_per_part=np.int64(in_per_part), want_keycols=np.int64(want_keycols),
        dims=np.asarray(dims, np.int32))
    return y1, keycols_out


def box_prk_step(y0, bounds, idx_local, k, dt, scale, dims, bits, in_per_part, drift_kind=0, f=None, mean=None,
                 inv_var=None):
    """Tables + selection + count + rank/gather on one stream (mccube/_kernels/base.py:152-177)."""
    core = _load()
    ws = core.mccb200_box_prk_step_workspace_bytes(ctypes.c_uint64(y0.shape[0]), int(k), len(dims), int(bits),
                                                   ctypes.c_uint64(in_per_part))
    n_out = y0.shape[0] * k // in_per_part * idx_local.shape[0]
    y1, _ = jax.ffi.ffi_call("mccb200_box_prk_step_ffi", (_f32(n_out, y0.shape[1]), _u8(ws)))(
        y0, *_drift_args(drift_kind, f, mean, inv_var), bounds, idx_local, drift_kind=np.int64(drift_kind),
        k=np.int64(k), dt=np.float32(dt), scale=np.float32(scale), bits=np.int64(bits),
        in_per_part=np.int64(in_per_part), dims=np.asarray(dims, np.int32))
    return y1


# ---------------------------------------------------------------------------------------------- drift, metrics
def gaussian_full_drift_prepare(prec):
    core = _load()
    nfloat = core.mccb200_gaussian_full_drift_tc_prepare_bytes(int(prec.shape[0])) // 4
    return jax.ffi.ffi_call("mccb200_gaussian_full_drift_tc_prepare_ffi", _f32(nfloat))(prec)


def gaussian_full_drift(y, mean, prec_split):
    """f = -(y - mean) @ prec on the tensor cores (tcgen05 kind::tf32, 3xTF32): config 4's drift."""
    _load()
    return jax.ffi.ffi_call("mccb200_gaussian_full_drift_tc_ffi", _f32(*y.shape))(y, mean, prec_split)


def gaussian_full_drift_simt(y, mean, prec):
    _load()
    return jax.ffi.ffi_call("mccb200_gaussian_full_drift_ffi", _f32(*y.shape))(y, mean, prec)


def moments_sum(p, d, w=None):
    """[sum_r w_r p_r (d entries), sum_r w_r] in float64 (mccube/_metrics.py:79-96)."""
    _load()
    return jax.ffi.ffi_call("mccb200_moments_sum_ffi", jax.ShapeDtypeStruct((d + 1,), jnp.float64))(
        p, w if w is not None else _DUMMY, d=np.int64(d), weighted=np.int64(w is not None))


def moments_m2(p, d, centre, w=None):
    _load()
    return jax.ffi.ffi_call("mccb200_moments_m2_ffi", jax.ShapeDtypeStruct((d, d), jnp.float64))(
        p, w if w is not None else _DUMMY, centre, d=np.int64(d), weighted=np.int64(w is not None))


def stratified_keys(p, d, com):
    _load()
    return jax.ffi.ffi_call("mccb200_stratified_keys_ffi", jax.ShapeDtypeStruct((p.shape[0],), jnp.uint32))(
        p, com, d=np.int64(d))


def sort_pairs(keys, values=None, key_bits=32):
    """Stable `lax.sort_key_val` on bits [0, key_bits); values=None sorts an iota (argsort)."""
    core = _load()
    ws = core.mccb200_sort_pairs_workspace_bytes(ctypes.c_uint64(keys.shape[0]))
    u32 = jax.ShapeDtypeStruct(keys.shape, jnp.uint32)
    k_out, v_out, _ = jax.ffi.ffi_call("mccb200_sort_pairs_ffi", (u32, u32, _u8(ws)))(
        keys, values if values is not None else keys, key_bits=np.int64(key_bits), has_values=np.int64(values is not None))
    return k_out, v_out


# ---------------------------------------------------------------------------------------------- the step
def can_fuse(solver, terms, y0) -> bool:
    """Euler + LocalLinearCubaturePath(Hadamard) + MonteCarloKernel (optionally under a Box partitioner, once the
    reference gains one), float32 single-array state on a CUDA device, n_substeps == 1, unweighted."""
    import diffrax

    from ._formulae import Hadamard
    from ._kernels import MonteCarloKernel
    from ._path import LocalLinearCubaturePath

    ctrl = terms.cde.control
    return (type(solver.solver) is diffrax.Euler and solver.n_substeps == 1 and not solver.weighted
            and isinstance(ctrl, LocalLinearCubaturePath) and isinstance(ctrl.gaussian_cubature, Hadamard)
            and type(solver.recombination_kernel) is MonteCarloKernel
            and isinstance(y0, jax.Array) and y0.dtype == jnp.float32 and y0.ndim == 2
            and list(y0.devices())[0].platform == "gpu")


def fused_step(solver, terms, t0, t1, y0, args, solver_state):
    """`MCCSolver.step` without the [n*k, d] expansion: same 5-tuple as the reference returns."""
    import diffrax

    kern = solver.recombination_kernel
    cub = terms.cde.control.gaussian_cubature
    k = cub.point_count
    n = y0.shape[0]
    # quirk Q1 (mccube/_solvers.py:127-134): with n_substeps == 1 the inner step runs on [t1, t1 + (t1 - t0)]
    dt = (t1 - t0).astype(jnp.float32)
    g = terms.cde.vector_field(t1, y0, args)
    scale = (g * jnp.sqrt(dt)).astype(jnp.float32)
    f = terms.ode.vector_field(t1, y0, args).astype(jnp.float32)          # evaluated on the parents only
    idx = mc_indices(kern.key, t0, n * k, kern.recombination_count, kern.with_replacement)   # quirk Q2: outer t0
    y1 = expand_select(y0, idx, k, dt, scale, drift_kind=1, f=f)
    return y1, None, dict(y0=y0, y1=y1), solver_state, diffrax.RESULTS.successful
