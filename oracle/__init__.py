"""CPU oracle for the MCCube hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a NumPy restatement of the reference algorithm
(tttc3/MCCube v0.0.3, `/root/reference/mccube`) and of the third-party
arithmetic it leans on (`jax.random`, `diffrax.Euler`, `ConstantStepSize`),
none of which is installed in this image.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it. Nothing under `mccube_b200/` imports it;
the product path fails loudly when the CUDA library is missing.

Parity pin status ("Oracle pinning"):
  * threefry2x32: pinned to the three Random123 known-answer vectors.
  * jax.random layering (non-partitionable mode, the reference's era): pinned to
    JAX's published `split(PRNGKey(0))`, `split(PRNGKey(42))`,
    `random_bits(PRNGKey(1701))`, `uniform/normal(PRNGKey(0))` values.
  * whole chain (fold_in(t bits) -> split -> 2-round sort shuffle -> expansion
    order -> Euler update -> W2): pinned to the reference's own published
    quickstart numbers (docs/quickstart.md:184-208: 3.3503528099696664), see
    tests/test_oracle_pin.py.
  * partitionable mode (JAX >= 0.5 default): structural checks only, NO external
    golden => "parity unpinned" for that mode.
  * x64 index draw (`randint64`, jax_enable_x64 semantics, used by the sharded pull
    path for more than 2^32 children): structural checks only => "parity unpinned".
  * BoxPartitionKernel: does not exist in the reference => oracle is the
    definition; "parity unpinned" by construction.
"""
