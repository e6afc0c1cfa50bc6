"""Generates tests/golden/*.npz from the oracle (run from the repo root:
`python tests/golden/make_golden.py`).

The reference (JAX/diffrax/equinox) cannot be imported in this image, so the
vectors come from oracle/ -- whose whole RNG/step chain is pinned to the reference's own
published quickstart number (tests/test_oracle_pin.py).  Fixtures are small on purpose.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import jax_random as jr  # noqa: E402
from oracle import mccube_ref as ref  # noqa: E402

OUT = Path(__file__).resolve().parent


def indices():
    key = jr.prng_key(42)
    cases = {}
    for part in (False, True):
        for replace in (False, True):
            for (P, n, t) in [(16, 4, 0.0), (16, 4, 0.05), (1000, 37, 1.25), (16384, 512, 0.0), (16384, 512, 0.05),
                              (65537, 1024, 3.5)]:
                k = jr.mc_step_key(key, t, partitionable=part)
                idx = jr.choice_indices(k, P, n, replace, partitionable=part)
                cases[f"P{P}_n{n}_t{t}_r{int(replace)}_p{int(part)}"] = idx.astype(np.int64)
    np.savez_compressed(OUT / "mc_indices.npz", **cases)


def readme_trajectory():
    """README.md:60-89 (config 1) in float32: n=512, d=10, dt0=0.05, 512 epochs, y0=ones."""
    n, d = 512, 10
    key = jr.prng_key(42)
    y0 = np.ones((n, d), np.float32)
    t0, dt0 = 0.0, 0.05
    t1 = t0 + dt0 * 512
    mean = (2 * np.ones(d)).astype(np.float32)
    inv_var = (np.ones(d) / 3.0).astype(np.float32)
    M, lam = ref.hadamard_points(d), ref.hadamard_weights(d)
    out = {}
    for replace in (False, True):
        kern = ref.MonteCarloKernel(n, with_replacement=replace, key=key)
        times = ref.step_times(t0, t1, dt0, np.float32)
        save = {0, 1, 63, len(times) - 1}
        yf, saved = ref.diffeqsolve(y0, t0, t1, dt0, lambda y: ref.gaussian_diag_drift(y, mean, inv_var), M, lam,
                                    np.float32(np.sqrt(2.0)), kern, time_dtype=np.float32, save_steps=save)
        m = yf.astype(np.float64).mean(0)
        c = np.cov(yf.astype(np.float64), rowvar=False)
        w2 = ref.gaussian_wasserstein_metric((2 * np.ones(d), m), (3 * np.eye(d), c))
        tag = f"r{int(replace)}"
        out[f"n_steps_{tag}"] = np.array(len(times))
        out[f"w2_{tag}"] = np.array(w2.real)
        for s, y in saved.items():
            out[f"y_step{s}_{tag}"] = y
    np.savez_compressed(OUT / "readme_trajectory.npz", **out)


def box_and_stratified():
    rng = np.random.default_rng(7)
    n, d = 96, 6
    y0 = rng.normal(size=(n, d)).astype(np.float32)
    mean = np.linspace(-1, 1, d).astype(np.float32)
    inv_var = (1.0 / np.linspace(1, 3, d)).astype(np.float32)
    M, lam = ref.hadamard_points(d), ref.hadamard_weights(d)
    drift = lambda y: ref.gaussian_diag_drift(y, mean, inv_var)  # noqa: E731
    t0, t1 = np.float32(0.15), np.float32(0.2)
    children = ref.expand(y0, t0, t1, drift, M, lam, np.float32(np.sqrt(2.0)))
    box = ref.BoxPartitionKernel(32, dims=(0, 2, 5), bits=2)
    strat = ref.StratifiedPartitioningKernel(32)
    mc = ref.MonteCarloKernel(n // 32, key=jr.prng_key(42))
    np.savez_compressed(
        OUT / "partition_step.npz", y0=y0, mean=mean, inv_var=inv_var, t0=t0, t1=t1, children=children,
        box_idx=box.indices(children), strat_idx=strat(0, children, None) if False else ref.stratified_indices(children, 32)[0],
        box_step=ref.mcc_step(y0, t0, t1, drift, M, lam, np.float32(np.sqrt(2.0)),
                              ref.PartitioningRecombinationKernel(box, mc)),
        strat_step=ref.mcc_step(y0, t0, t1, drift, M, lam, np.float32(np.sqrt(2.0)),
                                ref.PartitioningRecombinationKernel(strat, mc)),
    )


if __name__ == "__main__":
    indices()
    readme_trajectory()
    box_and_stratified()
    print("golden vectors written to", OUT)
