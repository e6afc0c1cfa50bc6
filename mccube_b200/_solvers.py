"""MCC solver (mirror of /root/reference/mccube/_solvers.py:33-155) on the CUDA hot path.

`MCCSolver.step` keeps the reference's protocol -- `step(terms, t0, t1, y0, args,
solver_state, made_jump) -> (y1, y_error, dense_info, solver_state, RESULTS)` -- and its
quirks, but the `[n*k, d]` expansion of `_solvers.py:136` is never
built when the recombination kernel is one the library knows how to fuse:

  MonteCarloKernel                                   -> mc_indices + expand_select
  PartitioningRecombinationKernel(BoxPartitionKernel,
                                  MonteCarloKernel)  -> box_child_bounds + box_prk_step
                                                        (or child keys + sort + expand_select)
  anything else (user kernels, Stratified, BinaryTree, n_substeps > 1, weighted)
                                                     -> expand_all on the GPU, then the
                                                        kernel protocol on the children.

There is no CPU fallback anywhere: unsupported term structures raise.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch

from . import _lib
from ._drift import GaussianDiagDrift
from ._kernels import (
    AbstractRecombinationKernel,
    BoxPartitionKernel,
    MonteCarloKernel,
    PartitioningRecombinationKernel,
)
from ._path import LocalLinearCubaturePath
from ._formulae import Hadamard
from ._term import MCCTerm
from ._utils import nop, pack_particles, unpack_particles
from .diffrax_compat import (
    AbstractSolver,
    AbstractWrappedSolver,
    Euler,
    LocalLinearInterpolation,
    RESULTS,
)


_side_streams: dict = {}


def _side_stream(device) -> "torch.cuda.Stream":
    key = torch.device(device).index or 0
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


class MCCSolver(AbstractWrappedSolver):
    """`MCCSolver(solver, recombination_kernel, n_substeps=1, weighted=False)`
    (_solvers.py:87-90).  `time_dtype` is the dtype diffrax carries `t` in (float32 unless
    jax_enable_x64): the sub-step interval and the cubature scale are computed in it."""

    term_structure = MCCTerm
    interpolation_cls = LocalLinearInterpolation

    def __init__(self, solver: AbstractSolver, recombination_kernel: AbstractRecombinationKernel,
                 n_substeps: int = 1, weighted: bool = False, *, time_dtype=np.float32, fused: bool = True):
        self.solver = solver
        self.recombination_kernel = recombination_kernel
        self.n_substeps = n_substeps
        self.weighted = weighted
        self.time_dtype = np.dtype(time_dtype)
        self.fused = fused
        self.fuse_draw = True   # with_replacement=True: draw the child ids inside the gather kernel when the shape allows
        self.last_path = None  # which implementation the last step took (for tests / bench)
        self._point_tables: dict = {}
        self.__check_init__()

    def __check_init__(self):
        """_solvers.py:96-103 (same messages)."""
        if not isinstance(self.solver, Euler) or type(self.solver) is not Euler:
            warnings.warn(
                f"""Support is only provided for the diffrax.Euler solver at present;
                got {self.solver}. Expect undefined behaviour!"""
            )
        if self.n_substeps < 1:
            raise ValueError(f"n_substeps must be at least one; got {self.n_substeps}")

    def init(self, terms, t0, t1, y0, args):
        return self.solver.init(terms, t0, t1, y0, args)

    # ------------------------------------------------------------------ step pieces
    def substep_times(self, t0, t1):
        """(_t0, _t1) of every inner step, with quirk Q1 (`_t0 = _t1` runs first,
        _solvers.py:127-134), in the time dtype."""
        ty = self.time_dtype.type
        t0, t1 = ty(t0), ty(t1)
        h = ty(ty(t1 - t0) / ty(self.n_substeps))
        _t0, _t1 = t0, ty(t0 + h)
        out = []
        for _ in range(self.n_substeps):
            _t0 = _t1
            _t1 = ty(_t0 + h)
            out.append((_t0, _t1))
        return out

    def _describe(self, terms, y, t_a, t_b, args):
        """Build the (drift, cubature, dt) descriptors of one inner Euler step, or raise."""
        if not isinstance(terms, MCCTerm):
            raise NotImplementedError(f"MCCSolver needs an MCCTerm (term_structure); got {type(terms).__name__}")
        control = terms.cde.control
        if not isinstance(control, LocalLinearCubaturePath):
            raise NotImplementedError("MCCSolver's CUDA path needs a LocalLinearCubaturePath control")
        ty = self.time_dtype.type
        dt = np.float32(ty(t_b) - ty(t_a))                      # ODETerm.contr
        coeff = np.float32(control.coefficient(t_a, t_b, self.time_dtype))
        g = terms.cde.vector_field(t_a, y, args)
        if isinstance(g, torch.Tensor):
            if g.numel() != 1:
                raise NotImplementedError("state-dependent diffusion is not supported by the CUDA path")
            g = g.item()
        g = np.float32(g)
        cubature = control.gaussian_cubature
        k = cubature.point_count
        d = y.shape[1]
        lam = None
        if isinstance(cubature, Hadamard):
            cub = _lib.make_cubature(k, scale=np.float32(g * coeff))
        else:
            # scaled point table, uploaded once per (formula, g, coeff): a constant-step solve re-uses it
            # every step, and a step that uploads nothing can be captured in a CUDA graph
            ck = (id(cubature), float(g), float(coeff), str(y.device))
            hit = self._point_tables.get(ck)
            if hit is None:
                pts = (g * (coeff * cubature.stacked_points.astype(np.float32))).astype(np.float32)
                hit = (torch.tensor(pts, dtype=torch.float32, device=y.device),
                       torch.tensor(cubature.stacked_weights.astype(np.float32), device=y.device), cubature)
                if len(self._point_tables) >= 256:   # distinct float32 step lengths of one solve: a handful
                    self._point_tables.pop(next(iter(self._point_tables)))
                self._point_tables[ck] = hit
            cub = _lib.make_cubature(k, points=hit[0], lam=hit[1])
        f = terms.ode.vector_field
        if isinstance(f, GaussianDiagDrift) and self.fused:
            drift = _lib.make_drift(2, mean=f.mean, inv_var=f.inv_var)
        else:
            fv = f(t_a, y, args)
            if fv is None:
                drift = _lib.make_drift(0)
            else:
                fv = torch.as_tensor(fv, dtype=torch.float32, device=y.device).contiguous()
                if fv.shape != (y.shape[0], d):
                    raise ValueError(f"drift must return [n, d]={tuple(y.shape)}; got {tuple(fv.shape)}")
                drift = _lib.make_drift(1, f=fv)
        return drift, cub, dt, k, d

    # ------------------------------------------------------------------ the step
    def step(self, terms, t0, t1, y0, args, solver_state, made_jump):
        del made_jump
        if isinstance(y0, (dict, list, tuple)):
            raise NotImplementedError("MCCSolver.step: PyTree states are only supported by the kernels, not the solver")
        if not y0.is_cuda:
            raise _lib.MccubeB200Error("MCCSolver.step needs a CUDA tensor state (no CPU fallback)")
        y0c = y0.contiguous()
        kernel = self.recombination_kernel
        times = self.substep_times(t0, t1)

        fusable = (self.fused and self.n_substeps == 1 and not self.weighted)
        y1_hat = None
        weight_total = None
        # weighted states with the default weighting_function=nop sample uniformly (quirk Q6): the fused
        # kernel carries the weight column along (child weight = w_parent * lambda, k-major quirk Q3)
        fusable_mc = (self.fused and self.n_substeps == 1 and type(kernel) is MonteCarloKernel
                      and isinstance(kernel.recombination_count, int)
                      and (not self.weighted or kernel.weighting_function is nop))
        if fusable_mc:
            yv = y0c[:, :-1] if self.weighted else y0c
            drift, cub, dt, k, d = self._describe(terms, yv, *times[0], args)
            n = y0c.shape[0]
            draw_in_kernel = (self.fuse_draw and kernel.with_replacement and not getattr(kernel, "x64", False)
                              and n * k < 2 ** 32 - 1
                              and _lib.expand_select_draw_supported(d, y0c.shape[1], self.weighted))
            if draw_in_kernel:
                # with_replacement=True draws are counter based: the gather kernel draws its own child ids
                rng = _lib.make_rng(kernel.key, t0, partitionable=kernel.threefry_partitionable)
                y1_hat = _lib.expand_select_draw(y0c, d, drift, dt, cub, rng, kernel.recombination_count)
                self.last_path = "fused:expand_select_draw"
            else:
                idx = kernel.indices(t0, n * k, kernel.recombination_count, y0c.device)
                if self.weighted:   # the gather also sums the child weights it writes: no pass of its own for them
                    y1_hat, weight_total = _lib.expand_select_wsum(y0c, d, drift, dt, cub, idx)
                else:
                    y1_hat = _lib.expand_select(y0c, d, self.weighted, drift, dt, cub, idx)
                self.last_path = "fused:mc_indices+expand_select" + ("(weighted)" if self.weighted else "")
        elif self._substeps_fusable(terms, kernel, y0c):
            y1_hat = self._substeps_step(terms, t0, times, y0c, args, kernel)
        elif (fusable and type(kernel) is PartitioningRecombinationKernel
              and type(kernel.partitioning_kernel) is BoxPartitionKernel
              and type(kernel.recombination_kernel) is MonteCarloKernel
              and isinstance(kernel.partitioning_kernel.partition_count, int)):
            y1_hat, solver_state = self._box_prk_step(terms, t0, times[0], y0c, args, kernel, solver_state)
        if y1_hat is None:
            y1_hat = self._generic_step(terms, t0, times, y0c, args, kernel)
        # renormalise the weights after recombination (_solvers.py:146)
        if self.weighted:
            y1_res = _lib.normalised_copy(y1_hat, weight_total)   # clone + weights / sum(weights) in one pass
        else:
            y1_res = y1_hat
        dense_info = dict(y0=y0, y1=y1_hat)
        return y1_res, None, dense_info, solver_state, RESULTS.successful

    MAX_FUSED_SUBSTEPS = 4

    def _substeps_fusable(self, terms, kernel, y0c) -> bool:
        """n_substeps in 2..4 with the Hadamard formula, no drift or the analytic diagonal-Gaussian drift
        (an arbitrary callable is only evaluated on arrays, i.e. at materialised intermediate states),
        unweighted, global MonteCarloKernel, and at most 2^32 children."""
        if not (self.fused and 1 < self.n_substeps <= self.MAX_FUSED_SUBSTEPS and not self.weighted):
            return False
        if type(kernel) is not MonteCarloKernel or not isinstance(kernel.recombination_count, int):
            return False
        if not isinstance(terms, MCCTerm) or not isinstance(terms.cde.control, LocalLinearCubaturePath):
            return False
        cubature = terms.cde.control.gaussian_cubature
        if not isinstance(cubature, Hadamard) or not isinstance(terms.ode.vector_field, GaussianDiagDrift):
            return False
        return y0c.shape[0] * cubature.point_count ** self.n_substeps <= 2 ** 32

    def _substeps_step(self, terms, t0, times, y0c, args, kernel):
        """All inner Euler steps in one kernel: only the selected children of the k^s expansion are computed."""
        dts, scales = [], []
        for (a, b) in times:
            drift, cub, dt, k, d = self._describe(terms, y0c, a, b, args)
            dts.append(dt)
            scales.append(cub.scale)
        n = y0c.shape[0]
        idx = kernel.indices(t0, n * k ** self.n_substeps, kernel.recombination_count, y0c.device)
        self.last_path = f"fused:mc_indices+expand_select(substeps={self.n_substeps})"
        return _lib.expand_select_substeps(y0c, d, drift, dts, scales, k, idx)

    @staticmethod
    def _carried_keycols(solver_state, y0c, dims):
        """The packed key columns the previous step's gather wrote for exactly this state tensor
        (same object, not modified since), or None.  Carried in `solver_state`, which diffrax hands
        from one step to the next (the reference's MCCSolver carries None there)."""
        if not isinstance(solver_state, dict):
            return None
        kc = solver_state.get("mccb200_keycols")
        if kc is None or solver_state.get("y1") is not y0c or solver_state.get("version") != y0c._version \
                or solver_state.get("dims") != tuple(dims):
            return None
        return kc

    def _box_prk_step(self, terms, t0, times0, y0c, args, kernel, solver_state=None):
        part: BoxPartitionKernel = kernel.partitioning_kernel
        mc: MonteCarloKernel = kernel.recombination_kernel
        drift, cub, dt, k, d = self._describe(terms, y0c, *times0, args)
        n = y0c.shape[0]
        m = part.partition_count
        total = n * k
        if total % m != 0:
            raise ValueError(f"partition_count {m} does not divide the {total} expanded particles")
        in_per_part = total // m
        out_per_part = mc.recombination_count
        dims = part.resolve_dims(d)
        use_kernel = (getattr(self, "use_box_prk_kernel", True)
                      and _lib.box_prk_supported(n, d, cub, dims, part.bits, in_per_part))
        # The shared local index vector depends only on (key, t0): it is drawn, and its selection
        # tables are built, on a side stream while the bounds / counting passes run on the main one.
        main = torch.cuda.current_stream()
        side = _side_stream(y0c.device)
        tables = _lib.box_prk_tables(k, dims, part.bits, y0c.device) if use_kernel else None
        # (no side.wait_stream(main): the draw reads nothing the main stream produces, so it may
        # also overlap the previous step's gather)
        if torch.cuda.is_current_stream_capturing():
            side.wait_stream(main)       # fork the side stream into the capture
        with torch.cuda.stream(side):
            idx_local = mc.indices(t0, in_per_part, out_per_part, y0c.device)
            sel = _lib.box_prk_selection(idx_local, in_per_part) if use_kernel else None
        # 4 consecutive, 16-byte aligned key columns: the gather also emits them packed for the next step
        vec4 = (use_kernel and len(dims) == 4 and dims[0] % 4 == 0 and tuple(dims) == tuple(range(dims[0], dims[0] + 4))
                and d % 4 == 0 and getattr(self, "carry_key_columns", True))
        keycols = self._carried_keycols(solver_state, y0c, dims) if vec4 else None
        bounds = part.fixed_bounds(dims, y0c.device)
        if bounds is None:
            bounds = _lib.box_child_bounds(y0c, d, drift, dt, cub, dims, part.bits, keycols=keycols)
        if use_kernel:
            ws = _lib.box_prk_count(y0c, d, drift, dt, cub, dims, part.bits, bounds, in_per_part, tables, keycols=keycols)
            main.wait_stream(side)
            idx_local.record_stream(main)
            sel.record_stream(main)
            n_out = m * out_per_part
            keycols_out = torch.empty((n_out, 4), dtype=torch.float32, device=y0c.device) if vec4 else None
            out = _lib.box_prk_rank_gather(y0c, d, drift, dt, cub, dims, part.bits, out_per_part, in_per_part,
                                           tables, sel, ws, keycols_out=keycols_out)
            self.last_path = "fused:box_prk_step" + ("+keycols" if keycols is not None else "")
            state = solver_state
            if vec4:
                state = dict(solver_state) if isinstance(solver_state, dict) else {}
                state.update(mccb200_keycols=keycols_out, y1=out, version=out._version, dims=tuple(dims))
            return out, state
        main.wait_stream(side)
        idx_local.record_stream(main)
        keys = _lib.box_child_keys(y0c, d, drift, dt, cub, dims, part.bits, bounds)
        perm = _lib.sort_pairs(keys, None, len(dims) * part.bits)
        out = _lib.expand_select(y0c, d, False, drift, dt, cub, idx_local, perm=perm, out_per_part=out_per_part,
                                 in_per_part=in_per_part, n_out=m * out_per_part)
        self.last_path = "fused:box_child_keys+sort+expand_select"
        return out, solver_state

    def _generic_step(self, terms, t0, times, y0c, args, kernel):
        """The reference's data flow with the expansion materialised on the GPU
        (_solvers.py:132-144): needed for arbitrary kernels, n_substeps > 1 and weights."""
        y, w = unpack_particles(y0c, self.weighted)
        y = y.contiguous()
        for (a, b) in times:
            drift, cub, dt, k, d = self._describe(terms, y, a, b, args)
            y = _lib.expand_all(y, d, False, drift, dt, cub)
            if w is not None:
                lam = (torch.tensor(terms.cde.control.weights, dtype=torch.float32, device=y.device))
                w = (w[None, :] * lam[:, None]).reshape(-1)       # quirk Q3: k-major (_solvers.py:138-141)
        y1_packed = pack_particles(y, w) if w is not None else y
        self.last_path = "generic:expand_all+kernel"
        return kernel(t0, y1_packed.contiguous(), args, self.weighted)

    def func(self, terms, t0, y0, args):
        """_solvers.py:150-155."""
        return self.recombination_kernel(t0, self.solver.func(terms, t0, y0, args), args, self.weighted)
