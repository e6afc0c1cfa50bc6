"""NumPy restatement of the `jax.random` chain used by MCCube (TEST INFRASTRUCTURE).

The reference never vendors this arithmetic: it calls `jr.fold_in`, `jr.split`,
`jr.choice` (`/root/reference/mccube/_kernels/random.py:56-58,65`) from an
un-pinned `jax>=0.4.23` (`/root/reference/pyproject.toml:31`).  What follows is the
published algorithm of `jax/_src/prng.py` (threefry2x32, `threefry_split`,
`threefry_fold_in`, `threefry_random_bits`) and `jax/_src/random.py`
(`_shuffle`, `randint`, `uniform`, `gumbel`, `normal`, `choice`) restated with
uint32 wrap-around arithmetic.  Both values of the `jax_threefry_partitionable`
flag are implemented (`partitionable=False` was the default of the JAX the
reference was written against; `True` is the default from JAX 0.5).

Keys are plain `(k0, k1)` tuples of Python ints / uint32.
"""
from __future__ import annotations

import math

import numpy as np

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = U32(0x1BD11BDA)


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(key, x0, x1):
    """Threefry-2x32, 20 rounds (Salmon et al., Random123). Vectorised over x0/x1."""
    with np.errstate(over="ignore"):
        k0, k1 = U32(key[0]), U32(key[1])
        ks = (k0, k1, k0 ^ k1 ^ _PARITY)
        x0 = np.asarray(x0, dtype=U32) + ks[0]
        x1 = np.asarray(x1, dtype=U32) + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + U32(i + 1)
    return x0, x1


def _hash_original(key, counts):
    """`threefry_2x32(keypair, count)` of jax/_src/prng.py: pad to even, split in halves."""
    counts = np.asarray(counts, dtype=U32).ravel()
    n = counts.size
    if n % 2:
        counts = np.concatenate([counts, np.zeros(1, U32)])
    h = counts.size // 2
    o0, o1 = threefry2x32(key, counts[:h], counts[h:])
    return np.concatenate([o0, o1])[:n]


def prng_key(seed: int):
    """`jax.random.PRNGKey(seed)` / `jr.key(seed)` for the threefry impl."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return (U32(seed >> 32), U32(seed & 0xFFFFFFFF))


def fold_in(key, data):
    """`jr.fold_in(key, data)`; `data` is reduced to uint32 (int32 bit pattern wraps)."""
    o0, o1 = threefry2x32(key, U32(0), U32(int(data) & 0xFFFFFFFF))
    return (U32(o0), U32(o1))


def split(key, num: int = 2, *, partitionable: bool = False):
    """`jr.split(key, num)` -> list of `num` keys."""
    if partitionable:
        o0, o1 = threefry2x32(key, np.zeros(num, U32), np.arange(num, dtype=U32))
        return [(U32(a), U32(b)) for a, b in zip(o0, o1)]
    out = _hash_original(key, np.arange(2 * num, dtype=U32)).reshape(num, 2)
    return [(U32(a), U32(b)) for a, b in out]


def random_bits(key, bit_width: int, size: int, *, partitionable: bool = False):
    """`_random_bits(key, bit_width, (size,))` (flat C-order for n-d shapes)."""
    assert bit_width in (32, 64)
    size = int(size)
    if partitionable:
        idx = np.arange(size, dtype=np.uint64)
        hi = (idx >> np.uint64(32)).astype(U32)
        lo = (idx & np.uint64(0xFFFFFFFF)).astype(U32)
        o0, o1 = threefry2x32(key, hi, lo)
        if bit_width == 32:
            return o0 ^ o1
        return (o0.astype(np.uint64) << np.uint64(32)) | o1.astype(np.uint64)
    max_count = (bit_width * size + 31) // 32
    assert max_count < 0xFFFFFFFF, "multi-block random_bits never exercised by the reference"
    bits = _hash_original(key, np.arange(max_count, dtype=U32))
    if bit_width == 32:
        return bits
    hi, lo = bits[:size], bits[size:]
    return (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)


def randint(key, size: int, minval: int, maxval: int, *, partitionable: bool = False):
    """`jr.randint(key, (size,), minval, maxval)` for int32."""
    k1, k2 = split(key, 2, partitionable=partitionable)
    hi = random_bits(k1, 32, size, partitionable=partitionable)
    lo = random_bits(k2, 32, size, partitionable=partitionable)
    span = U32(maxval - minval) if maxval > minval else U32(1)
    with np.errstate(over="ignore"):
        mult = U32(1 << 16) % span
        mult = U32(mult * mult) % span
        off = ((hi % span) * mult + (lo % span)) % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32)


def randint64(key, size: int, minval: int, maxval: int, *, partitionable: bool = False):
    """`jr.randint(key, (size,), minval, maxval)` with `jax_enable_x64` (int64): the same recipe on 64-bit
    words -- two 64-bit draws, multiplier = ((2^32 % span)^2 mod 2^64) % span, all arithmetic wrapping in
    uint64 (jax/_src/random.py `_randint`).  UNPINNED: restated from the public source, no JAX here."""
    U64 = np.uint64
    k1, k2 = split(key, 2, partitionable=partitionable)
    hi = random_bits(k1, 64, size, partitionable=partitionable)
    lo = random_bits(k2, 64, size, partitionable=partitionable)
    span = U64(maxval - minval) if maxval > minval else U64(1)
    with np.errstate(over="ignore"):
        mult = U64(1 << 32) % span
        mult = U64(mult * mult) % span
        off = ((hi % span) * mult + (lo % span)) % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int64)


def uniform(key, size: int, dtype=np.float32, minval=0.0, maxval=1.0, *, partitionable=False):
    """`jr.uniform(key, (size,), dtype, minval, maxval)`."""
    dtype = np.dtype(dtype)
    nbits = dtype.itemsize * 8
    nmant = np.finfo(dtype).nmant
    bits = random_bits(key, nbits, size, partitionable=partitionable)
    utype = np.uint32 if nbits == 32 else np.uint64
    one_bits = np.array(1.0, dtype).view(utype)
    fbits = (bits >> utype(nbits - nmant)) | one_bits
    floats = fbits.view(dtype) - dtype.type(1.0)
    minval = dtype.type(minval)
    maxval = dtype.type(maxval)
    return np.maximum(minval, (floats * (maxval - minval) + minval).astype(dtype))


def gumbel(key, size: int, dtype=np.float32, *, partitionable=False):
    dtype = np.dtype(dtype)
    u = uniform(key, size, dtype, np.finfo(dtype).tiny, 1.0, partitionable=partitionable)
    return (-np.log(-np.log(u))).astype(dtype)


def normal(key, size: int, dtype=np.float32, *, partitionable=False):
    """`jr.normal`: sqrt(2) * erfinv(uniform(nextafter(-1, 0), 1))."""
    from scipy.special import erfinv

    dtype = np.dtype(dtype)
    lo = np.nextafter(dtype.type(-1.0), dtype.type(0.0))
    u = uniform(key, size, dtype, lo, 1.0, partitionable=partitionable)
    return (dtype.type(np.sqrt(2)) * erfinv(u).astype(dtype)).astype(dtype)


def shuffle_rounds(size: int) -> int:
    """Number of sort rounds of `jax._src.random._shuffle`."""
    return int(np.ceil(3 * np.log(max(1, size)) / np.log(np.iinfo(np.uint32).max)))


def shuffle_perm(key, size: int, *, partitionable: bool = False):
    """`_shuffle(key, arange(size))`: `rounds` stable sorts by fresh 32-bit keys."""
    x = np.arange(size, dtype=np.int64)
    for _ in range(shuffle_rounds(size)):
        key, sub = split(key, 2, partitionable=partitionable)
        sk = random_bits(sub, 32, size, partitionable=partitionable)
        x = x[np.argsort(sk, kind="stable")]
    return x


def choice_indices(key, n_inputs: int, n_draws: int, replace: bool, p=None, *,
                   partitionable: bool = False, x64: bool = False):
    """Index vector that `jr.choice(key, a, (n_draws,), replace, p)` gathers with.

    The reference never exposes this vector (`random.py:65` returns `a[ind]`); it is
    what the CUDA path must reproduce bit for bit on the integer branches.
    """
    if p is None:
        if replace:
            if x64:   # jax_enable_x64: randint's default dtype is int64
                return randint64(key, n_draws, 0, n_inputs, partitionable=partitionable)
            return randint(key, n_draws, 0, n_inputs, partitionable=partitionable).astype(np.int64)
        return shuffle_perm(key, n_inputs, partitionable=partitionable)[:n_draws]
    p = np.asarray(p)
    if replace:
        c = np.cumsum(p)
        u = uniform(key, n_draws, c.dtype, partitionable=partitionable)
        r = c[-1] * (c.dtype.type(1) - u)
        return np.searchsorted(c, r).astype(np.int64)
    g = -gumbel(key, n_inputs, p.dtype, partitionable=partitionable) - np.log(p)
    return np.argsort(g, kind="stable")[:n_draws].astype(np.int64)


def float_time_bits(t, dtype=np.float32) -> int:
    """`force_bitcast_convert_type(t, int32)` of diffrax._misc: cast to float32, bitcast."""
    del dtype
    return int(np.array(t, dtype=np.float32).view(np.uint32))


def mc_step_key(base_key, t, *, leaf: int = 0, n_leaves: int = 1, partitionable: bool = False):
    """Key used by MonteCarloKernel at time `t` (`_kernels/random.py:56-58`)."""
    k = fold_in(base_key, float_time_bits(t))
    return split(k, n_leaves, partitionable=partitionable)[leaf]
