"""Stand-ins for the slice of diffrax the MCC hot path touches.

diffrax / jax / equinox are not installed in this image, so the solver
protocol the reference plugs into is restated here with the same names and argument
meaning: `ODETerm`, `WeaklyDiagonalControlTerm`, `Euler`, `SaveAt`, `diffeqsolve`
(ConstantStepSize only), `RESULTS`, `LocalLinearInterpolation`.  Where real diffrax
exists the `MCCSolver` / kernels are driven by it instead.

Time arithmetic follows diffrax exactly (float32 unless `time_dtype=np.float64`):
`ConstantStepSize.init -> t0 + dt0`, `adapt_step_size -> (t1, t1 + dt0)`, and the
`_clip_to_end` tolerance of diffrax/_integrate.py; the RNG fold-in word of each step
depends on these bits (quirk Q2).
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Any, Callable, Optional, Sequence

import numpy as np
import torch


class RESULTS(enum.IntEnum):
    successful = 0
    max_steps_reached = 1


class AbstractTerm:
    def vf(self, t, y, args):
        raise NotImplementedError

    def contr(self, t0, t1):
        raise NotImplementedError

    def prod(self, vf, control):
        raise NotImplementedError

    def vf_prod(self, t, y, args, control):
        return self.prod(self.vf(t, y, args), control)


class ODETerm(AbstractTerm):
    """dy = f(t, y, args) dt; `prod(vf, control) = control * vf`."""

    def __init__(self, vector_field: Callable):
        self.vector_field = vector_field

    def vf(self, t, y, args):
        return self.vector_field(t, y, args)

    def contr(self, t0, t1):
        return t1 - t0

    def prod(self, vf, control):
        from ._tree import tree_map

        return tree_map(lambda v: control * v, vf)


class WeaklyDiagonalControlTerm(AbstractTerm):
    """dy = g(t, y, args) * dx(t) elementwise; `prod(vf, control) = vf * control`."""

    def __init__(self, vector_field: Callable, control):
        self.vector_field = vector_field
        self.control = control

    def vf(self, t, y, args):
        return self.vector_field(t, y, args)

    def contr(self, t0, t1):
        return self.control.evaluate(t0, t1)

    def prod(self, vf, control):
        from ._tree import tree_map

        if isinstance(control, (dict, list, tuple)):
            return tree_map(lambda v, c: v * c, vf, control)
        return vf * control


class AbstractSolver:
    term_structure: Any = AbstractTerm
    interpolation_cls: Any = None

    def init(self, terms, t0, t1, y0, args):
        return None

    def func(self, terms, t0, y0, args):
        return terms.vf(t0, y0, args)


class AbstractWrappedSolver(AbstractSolver):
    solver: AbstractSolver


class Euler(AbstractSolver):
    """y1 = y0 + terms.vf_prod(t0, y0, args, terms.contr(t0, t1))."""

    def step(self, terms, t0, t1, y0, args, solver_state, made_jump):
        from ._tree import tree_map

        del solver_state, made_jump
        control = terms.contr(t0, t1)
        y1 = tree_map(lambda a, b: a + b, y0, terms.vf_prod(t0, y0, args, control))
        return y1, None, dict(y0=y0, y1=y1), None, RESULTS.successful


class EulerHeun(Euler):
    """Present only so `MCCSolver.__check_init__`'s warning path can be exercised
    (/root/reference/tests/test_solvers.py:55-64)."""


class LocalLinearInterpolation:
    def __init__(self, *, t0, t1, y0, y1):
        self.t0, self.t1, self.y0, self.y1 = t0, t1, y0, y1

    def evaluate(self, t0, t1=None, left=True):
        from ._tree import tree_map

        del left
        if t1 is not None:
            return tree_map(lambda a, b: a - b, self.evaluate(t1), self.evaluate(t0))
        denom = float(self.t1) - float(self.t0)
        coeff = 0.0 if denom == 0.0 else (float(t0) - float(self.t0)) / denom
        return tree_map(lambda a, b: a + coeff * (b - a), self.y0, self.y1)


Euler.interpolation_cls = LocalLinearInterpolation        # as diffrax.Euler (defined above the interpolation class)


@dataclass
class SaveAt:
    t0: bool = False
    t1: bool = False
    ts: Optional[Sequence[float]] = None
    steps: bool = False
    dense: bool = False


@dataclass
class Solution:
    t0: float
    t1: float
    ts: Any
    ys: Any
    result: RESULTS
    stats: dict = field(default_factory=dict)
    interpolation: Any = None

    def evaluate(self, t):
        if self.interpolation is None:
            raise ValueError("Dense solution has not been saved; pass SaveAt(dense=True).")
        for interp in self.interpolation:
            if float(interp.t0) <= float(t) <= float(interp.t1):
                return interp.evaluate(t)
        return self.interpolation[-1].evaluate(t)


def _clip_to_end(tprev, tnext, t1, dtype):
    tol = 1e-10 if np.dtype(dtype) == np.float64 else 1e-6
    return dtype(t1) if tnext > t1 - dtype(tol) else tnext


def step_times(t0, t1, dt0, time_dtype=np.float32, max_steps=4096):
    """(tprev, tnext) pairs of a ConstantStepSize solve, in `time_dtype` arithmetic."""
    dtype = np.dtype(time_dtype).type
    t0, t1, dt0 = dtype(t0), dtype(t1), dtype(dt0)
    tprev, tnext = t0, _clip_to_end(t0, dtype(t0 + dt0), t1, dtype)
    out = []
    while tprev < t1 and len(out) < max_steps:
        out.append((tprev, tnext))
        tprev, tnext = tnext, dtype(tnext + dt0)
        tprev = min(tprev, t1)
        tnext = _clip_to_end(tprev, tnext, t1, dtype)
    return out


def _stack(ys):
    from ._tree import tree_map

    if not ys:
        return None
    return tree_map(lambda *leaves: torch.stack(leaves), ys[0], *ys[1:])


def diffeqsolve(terms, solver, t0, t1, dt0, y0, args=None, *, saveat: SaveAt = SaveAt(t1=True),
                max_steps: int = 4096, time_dtype=np.float32, progress: Callable | None = None) -> Solution:
    """Constant-step integration loop with diffrax's call protocol:
    `solver.step(terms, tprev, tnext, y, args, solver_state, made_jump)` each step."""
    times = step_times(t0, t1, dt0, time_dtype, max_steps)
    state = solver.init(terms, times[0][0] if times else t0, times[0][1] if times else t1, y0, args)
    y = y0
    ts_out, ys_out, interps = [], [], []
    want_ts = None if saveat.ts is None else [np.dtype(time_dtype).type(t) for t in saveat.ts]
    ts_cursor = 0
    if saveat.t0:
        ts_out.append(np.dtype(time_dtype).type(t0)); ys_out.append(y0)
    if want_ts is not None:
        while ts_cursor < len(want_ts) and want_ts[ts_cursor] <= np.dtype(time_dtype).type(t0):
            ts_out.append(want_ts[ts_cursor]); ys_out.append(y0); ts_cursor += 1
    result = RESULTS.successful
    for i, (a, b) in enumerate(times):
        y_new, _, dense_info, state, result = solver.step(terms, a, b, y, args, state, False)
        if want_ts is not None or saveat.dense:
            interp = solver.interpolation_cls(t0=a, t1=b, **dense_info)
            if saveat.dense:
                interps.append(interp)
            while want_ts is not None and ts_cursor < len(want_ts) and want_ts[ts_cursor] <= b:
                ts_out.append(want_ts[ts_cursor]); ys_out.append(interp.evaluate(want_ts[ts_cursor])); ts_cursor += 1
        if saveat.steps:
            ts_out.append(b); ys_out.append(y_new)
        y = y_new
        if progress is not None:
            progress(i, b, y)
    if len(times) >= max_steps and times and times[-1][1] < np.dtype(time_dtype).type(t1):
        result = RESULTS.max_steps_reached
    if saveat.t1 and not saveat.steps:
        ts_out.append(np.dtype(time_dtype).type(t1)); ys_out.append(y)
    return Solution(t0=t0, t1=t1, ts=np.array(ts_out), ys=_stack(ys_out), result=result,
                    stats=dict(num_steps=len(times)), interpolation=interps or None)


class CapturedSolve:
    """A whole constant-step solve captured ONCE as a CUDA graph and replayed per call: the counterpart
    of `jax.jit(diffeqsolve)` (the reference runs the loop inside one XLA executable, README.md:82-84).
    Every step's scalars (t0 bits folded into the key, dt, cubature scale) are baked into the graph's
    kernel arguments, so a replay launches the same kernels on the same addresses with no Python and no
    per-kernel launch latency -- what makes the small configurations (README: n=512, 512 steps) fast.

        solve = capture_solve(terms, solver, t0, t1, dt0, y0, args, saveat=...)
        sol = solve(y0)            # any y0 of the captured shape / dtype / device
    """

    def __init__(self, terms, solver, t0, t1, dt0, y0, args=None, *, saveat: SaveAt = SaveAt(t1=True),
                 max_steps: int = 4096, time_dtype=np.float32):
        if not (isinstance(y0, torch.Tensor) and y0.is_cuda):
            raise ValueError("capture_solve needs a CUDA tensor state")
        if saveat.dense:
            raise ValueError("capture_solve: SaveAt(dense=True) keeps every step alive; use ts= or steps=")
        self._y0 = y0.detach().clone()
        run = lambda: diffeqsolve(terms, solver, t0, t1, dt0, self._y0, args, saveat=saveat,
                                  max_steps=max_steps, time_dtype=time_dtype)
        # one eager solve first: lazy one-time work (function attributes, cached tables and point uploads,
        # the caching allocator's pool) must not happen inside the capture
        warm = torch.cuda.Stream(device=y0.device)
        warm.wait_stream(torch.cuda.current_stream(y0.device))
        with torch.cuda.stream(warm):
            run()
        torch.cuda.current_stream(y0.device).wait_stream(warm)
        self.graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self.graph):
                self._sol = run()
        except RuntimeError as e:
            # a step that synchronises with the host or copies from pageable memory (the generic path with a
            # host-side partitioner, a drift / diffusion callable that calls .item() or .cpu(), ...) cannot be recorded
            torch.cuda.synchronize(y0.device)
            raise ValueError("capture_solve: this solve cannot be recorded as a CUDA graph (a step synchronises with "
                             f"the host or uploads from pageable memory); use diffeqsolve instead [{e}]") from e

    def __call__(self, y0) -> Solution:
        from ._tree import tree_map

        if y0.shape != self._y0.shape or y0.dtype != self._y0.dtype or y0.device != self._y0.device:
            raise ValueError(f"captured for {tuple(self._y0.shape)} {self._y0.dtype} on {self._y0.device}")
        self._y0.copy_(y0)
        self.graph.replay()
        s = self._sol
        # the graph's buffers are overwritten by the next replay: hand out copies
        return Solution(t0=s.t0, t1=s.t1, ts=s.ts.copy(), ys=tree_map(lambda a: a.clone(), s.ys),
                        result=s.result, stats=dict(s.stats, cuda_graph=True))


def capture_solve(*a, **kw) -> CapturedSolve:
    return CapturedSolve(*a, **kw)
