/*
 * mccube_b200 -- C ABI of the B200-native MCCube hot path (sm_100a).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / jax types.  Every
 * entry point only ENQUEUES work on the caller's CUDA stream and returns; nothing
 * synchronises, nothing allocates (scratch is caller-provided through the
 * *_workspace_bytes() queries), results stay on the device.  That is the contract an
 * XLA-FFI custom call needs (INTEGRATION.md shows the jax.ffi binding for each one)
 * and it makes every call CUDA-graph capturable.
 *
 * The reference (tttc3/MCCube v0.0.3) is pure Python on JAX; the "FFI for this path"
 * is the set of XLA ops its step lowers to.  Each function below names the reference
 * lines it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - all data pointers are DEVICE pointers unless the name ends in _host;
 *   - particle state is row-major float32 [n, ld] with ld = d (+1 when weighted: the
 *     weight is the LAST column, mccube/_utils.py:47-51,82-84);
 *   - child id c of the expansion is c = i*k + j  (parent i, cubature point j;
 *     mccube/_solvers.py:136); ids are uint32 (n*k <= 2^32);
 *   - return value: 0 on success, negative mccb200_status on error;
 *     mccb200_last_error() returns a thread-local message.
 */
#ifndef MCCUBE_B200_H_
#define MCCUBE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCCB200_VERSION 100
#define MCCB200_API __attribute__((visibility("default")))

typedef enum {
  MCCB200_OK = 0,
  MCCB200_ERR_INVALID_ARGUMENT = -1,
  MCCB200_ERR_WORKSPACE_TOO_SMALL = -2,
  MCCB200_ERR_CUDA = -3,
  MCCB200_ERR_UNSUPPORTED = -4
} mccb200_status;

MCCB200_API int mccb200_version(void);
MCCB200_API const char* mccb200_last_error(void);
/* Number of SMs of the current device (grid sizing is done inside the library). */
MCCB200_API int mccb200_sm_count(void);
/* Number of kernels (and device-side copies) this library has enqueued so far in this
 * process: the evidence behind bench.py's gpu_launches. */
MCCB200_API unsigned long long mccb200_launch_count(void);
/* Optional phase timing for benchmarks: when enabled, every kernel group of the library
 * records a CUDA event on the launching stream; profile_read synchronises on the last one
 * and returns (phase name, milliseconds) pairs in launch order.  Off by default. */
MCCB200_API void mccb200_profile_enable(int on);
MCCB200_API int mccb200_profile_read(const char** names, float* ms, int max_entries);

/* ------------------------------------------------------------------ RNG description
 * The key MonteCarloKernel uses at time t (mccube/_kernels/random.py:56-58):
 *     _t   = bitcast_int32(float32(t))            diffrax._misc.force_bitcast_convert_type
 *     key  = jr.fold_in(self.key, _t)
 *     key  = split_by_tree(key, particles)[leaf]  == jr.split(key, n_leaves)[leaf]
 * key_dev / t_dev let the words come from device memory (what an XLA custom call gets
 * inside diffrax's while_loop); when NULL the by-value fields are used. */
typedef struct {
  uint32_t key[2];          /* jr.key(seed) words: {seed >> 32, seed & 0xffffffff}      */
  const uint32_t* key_dev;  /* optional: 2 words on the device                          */
  float t;                  /* time the kernel is called with (outer t0, _solvers.py:144) */
  const float* t_dev;       /* optional: float32 t on the device                        */
  uint32_t leaf;            /* position of the array in the particles PyTree            */
  uint32_t n_leaves;        /* number of PyTree leaves (1 for a single array)           */
  int32_t partitionable;    /* jax_threefry_partitionable flag (0 = JAX < 0.5 default)  */
  int32_t fold_time;        /* 1: apply the fold_in/split chain above; 0: use key as is */
} mccb200_rng;

/* ------------------------------------------------------------------ A8: resampling indices
 * jr.choice(key, p[n_inputs, .], (count,), replace, p=None)  (random.py:65) as an INDEX
 * vector:  replace=0 -> _shuffle(key, arange(n_inputs))[:count]  (`rounds` stable
 * sort_key_val rounds over fresh 32-bit threefry keys); replace=1 -> randint(key, 0, n_inputs).
 * Bit-exact with jax.random for both threefry modes.  n_inputs < 2^32 - 1. */
MCCB200_API size_t mccb200_mc_indices_workspace_bytes(uint64_t n_inputs, uint64_t count, int replace);
MCCB200_API int mccb200_mc_indices(const mccb200_rng* rng, uint64_t n_inputs, uint64_t count, int replace,
                       uint32_t* idx_out, void* workspace, size_t workspace_bytes, void* stream);

/* Weighted branches of jr.choice (random.py:60-65 with a real `weighting_function`):
 *   gumbel_keys: keys_out[i] = order-preserving uint32 of g_i = -gumbel(key)_i - log(p_i); the
 *     caller sorts them (sort_pairs, 32 bits) and keeps the first `count`  (Gumbel top-k).
 *   weighted_choice_replace: c = cumsum(p) (fp64 accumulate), r = c[-1] * (1 - uniform(key)),
 *     idx = searchsorted(c, r).
 * p may be strided (the weight column of packed particles).  Random bits are jax.random's; the
 * float arithmetic is not promised bit-identical to XLA. */
MCCB200_API size_t mccb200_weighted_workspace_bytes(uint64_t n_inputs);
MCCB200_API int mccb200_gumbel_keys(const mccb200_rng* rng, const float* p, int p_stride, uint64_t n_inputs,
                                    uint32_t* keys_out, void* workspace, size_t workspace_bytes, void* stream);
MCCB200_API int mccb200_weighted_choice_replace(const mccb200_rng* rng, const float* p, int p_stride,
                                                uint64_t n_inputs, uint64_t count, uint32_t* idx_out,
                                                void* workspace, size_t workspace_bytes, void* stream);

/* Stable LSD radix sort of (uint32 key, uint32 value) pairs on bits [0, key_bits)
 * == lax.sort_key_val(keys, values, is_stable=True) (jax _shuffle) / jnp.argsort
 * (stratified.py:43).  values_in == NULL sorts an implicit iota (argsort).
 * Sorted pairs land in keys_out / values_out (keys_out may be NULL). */
MCCB200_API size_t mccb200_sort_pairs_workspace_bytes(uint64_t count);
MCCB200_API int mccb200_sort_pairs(const uint32_t* keys_in, const uint32_t* values_in, uint64_t count, int key_bits,
                       uint32_t* keys_out, uint32_t* values_out, void* workspace, size_t workspace_bytes,
                       void* stream);

/* ------------------------------------------------------------------ drift description (A6)
 * kind 0: no drift.  kind 1: f[n, d] evaluated by the caller (any user callable,
 * README.md:76; evaluated on PARENTS only, mccube/_term.py:58-63).  kind 2: fused
 * analytic drift of a diagonal Gaussian target, f = (mean - y) * inv_var. */
typedef struct {
  int32_t kind;
  const float* f;        /* kind 1: [n, d]  */
  const float* mean;     /* kind 2: [d]     */
  const float* inv_var;  /* kind 2: [d]     */
} mccb200_drift;

/* Cubature description (A4/A5).  points == NULL selects the Hadamard formula
 * (mccube/_formulae.py:211-227): sign(M[j,c]) = (-1)^popcount((j mod k/2) & c), negated
 * for j >= k/2, generated in-kernel; noise = +-scale with scale = g*sqrt(dt)
 * (mccube/_path.py:60-71, _term.py:70-77).  Otherwise points is the [k, d] table
 * g * (coeff * M) already scaled by the caller. */
typedef struct {
  int32_t k;             /* point count */
  const float* points;   /* optional [k, d] scaled table */
  float scale;           /* Hadamard: fl(g * fl(sqrt(dt))) */
  const float* lam;      /* [k] cubature weights (weighted state only), may be NULL = 1/k */
} mccb200_cubature;

/* ------------------------------------------------------------------ A1-A5 + gather
 * Fused expand+select: for every output row r the selected child id is
 *     c = perm ? perm[(r / out_per_part) * in_per_part + idx[r % out_per_part]] : idx[r]
 * (the second form is PartitioningRecombinationKernel's shared local index vector,
 * mccube/_kernels/base.py:159-177, quirk Q4), and
 *     y1[r, :] = y0[i, :] + (dt * f(y0[i, :]) + noise[j, :]),   i = c / k, j = c % k
 * in exactly that fp32 op order, no FMA contraction (diffrax Euler.step through
 * mccube/_term.py:77).  The n*k*d expansion of _solvers.py:136 never exists.
 * weighted != 0: ld = d + 1 and the child weight is w[c % n] * lam[c / n]
 * (_solvers.py:138-141 lays weights out k-major: quirk Q3), unnormalised. */
MCCB200_API int mccb200_expand_select(const float* y0, uint64_t n, int d, int weighted, const mccb200_drift* drift,
                          float dt, const mccb200_cubature* cub, const uint32_t* idx, const uint32_t* perm,
                          uint64_t out_per_part, uint64_t in_per_part, uint64_t n_out, float* y1,
                          void* stream);

/* The same with n_substeps inner Euler steps fused (MCCSolver(..., n_substeps=s), mccube/_solvers.py:127-136):
 * child c = ((i*k + j_1)*k + j_2)..., y <- y + (dts[s] * f(y) + scales[s] * M[j_s, :]) for s = 0..n_substeps-1,
 * the drift re-evaluated at every intermediate state.  Hadamard cubature, drift kinds 0 / 2 (an array drift
 * is only known at the parents), unweighted, n * k^n_substeps <= 2^32.  dts / scales are HOST arrays. */
/* The same with child ids `idx_stride` 32-bit words apart (idx_stride = 2: the id column of packed (slot, id) pairs
 * received from the id exchange, mccb200_shard_bucket_ids); unweighted, no permutation. */
MCCB200_API int mccb200_expand_select_strided(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                              const mccb200_cubature* cub, const uint32_t* idx, int idx_stride,
                                              uint64_t n_out, float* y1, void* stream);
/* y_out = y_in with the weight column (last of ld) divided by its sum: the renormalised copy MCCSolver.step returns
 * next to the raw y1 of dense_info (mccube/_solvers.py:146-155), in one streaming pass.  total: device double with the
 * sum of the weights when a producer supplies it (mccb200_expand_select_wsum), NULL = summed here (workspace: one
 * double). */
MCCB200_API int mccb200_normalised_copy(const float* y_in, uint64_t n, int ld, const double* total, float* y_out,
                                        double* workspace, void* stream);
/* mccb200_expand_select (weighted = 1, no permutation) that also accumulates the sum of the child weights it writes
 * into *wsum (device double, zeroed by the call). */
MCCB200_API int mccb200_expand_select_wsum(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                           const mccb200_cubature* cub, const uint32_t* idx, uint64_t n_out, float* y1,
                                           double* wsum, void* stream);
/* Median-split (KD) partitioner on the device: replaces the host callback of mccube/_kernels/tree.py:44-58 (sklearn
 * KDTree / BallTree through jax.pure_callback).  idx_out[n] lists the rows of y [n, ld] (first d columns = coordinates)
 * leaf by leaf: every level splits every node at its median ((end - start) / 2, as sklearn's _recursive_build) along
 * the dimension of largest max - min.  The point SETS of the nodes equal sklearn's at the same depth; the order inside
 * a leaf is sorted along the last split dimension (sklearn's is std::nth_element's, platform dependent).
 * mccb200_kd_partition_levels(n, leaf_size) = floor(log2(n / leaf_size)): leaves of leaf_size .. 2 * leaf_size - 1
 * points, one level deeper than sklearn's tree when n / leaf_size is a power of two. */
MCCB200_API int mccb200_kd_partition_levels(uint64_t n, uint64_t leaf_size);
MCCB200_API size_t mccb200_kd_partition_workspace_bytes(uint64_t n, int d, int levels);
MCCB200_API int mccb200_kd_partition(const float* y, uint64_t n, int d, int ld, int levels, uint32_t* idx_out,
                                     void* workspace, size_t workspace_bytes, void* stream);
/* MonteCarloKernel(with_replacement=True) and the expand step in ONE kernel (replaces mccube/_kernels/random.py:45-66
 * followed by the gather of mccube/_solvers.py:117-142): output row r is child
 * jr.choice(key_at(t), n * k, (n_out,), replace=True)[r], the 32-bit randint drawn inside the kernel, so the index
 * vector never exists in HBM (idx_out, optional, receives it).  Unweighted states with d / 4 a power of two in
 * 4 .. 32 (ask mccb200_expand_select_draw_supported); n * k < 2^32 - 1. */
MCCB200_API size_t mccb200_expand_select_draw_workspace_bytes(void);
MCCB200_API int mccb200_expand_select_draw_supported(int d, int ld, int weighted);
MCCB200_API int mccb200_expand_select_draw(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                           const mccb200_cubature* cub, const mccb200_rng* rng, uint64_t n_out,
                                           float* y1, uint32_t* idx_out, void* workspace, size_t workspace_bytes,
                                           void* stream);
/* Residency of the gather kernels launched by THIS thread from now on: at most `ctas_per_sm` CTAs per SM (1..8; 0
 * restores the default, which fills the SM).  A gather that leaves a quarter of every SM free lets the index draw and
 * the bucketing of the next step run under it (mccube_b200/_distributed.py::step_owner_ids). */
MCCB200_API void mccb200_set_expand_residency(int ctas_per_sm);
/* The counterpart for the kernels meant to run UNDER such a gather (index draw with replacement, id bucketing): at
 * most `ctas_per_sm` CTAs per SM (0 = default grids). */
MCCB200_API void mccb200_set_side_residency(int ctas_per_sm);
MCCB200_API int mccb200_expand_select_substeps(const float* y0, uint64_t n, int d, const mccb200_drift* drift,
                                               int n_substeps, const float* dts_host, const float* scales_host,
                                               int k, const uint32_t* idx, uint64_t n_out, float* y1, void* stream);

/* Materialise the full [n*k, ld] expansion (what _solvers.py:136 builds).  Only for the
 * generic kernel protocol (user-defined recombination kernels) and for parity tests. */
MCCB200_API int mccb200_expand_all(const float* y0, uint64_t n, int d, int weighted, const mccb200_drift* drift,
                       float dt, const mccb200_cubature* cub, float* y1, void* stream);

/* out[r, :] = in[idx[r], :]  (jnp.take / p[indices]: stratified.py:45, tree.py:58, random.py:65). */
MCCB200_API int mccb200_gather_rows(const float* in, int ld, const uint32_t* idx, uint64_t n_out, float* out, void* stream);

/* Weight renormalisation of pack_particles(*unpack_particles(.)) (_solvers.py:146,
 * _utils.py:47-51): last column divided by its sum.  workspace: 1 double. */
MCCB200_API int mccb200_normalise_weights(float* y, uint64_t n, int ld, double* workspace, void* stream);

/* ------------------------------------------------------------------ A10: partitioning keys
 * StratifiedPartitioningKernel (mccube/_kernels/stratified.py:36-45): key = ||p - com||_2
 * as an order-preserving uint32 (fp32 sequential sum of squares, sqrt).  com[d] on device. */
MCCB200_API int mccb200_stratified_keys(const float* p, uint64_t n, int d, int ld, const float* com, uint32_t* keys,
                            void* stream);

/* BoxPartitionKernel (NEW: absent from the reference).  n_dims key
 * dimensions dims[i] (host array), `bits` bits per dimension:
 *     cell = clamp(floor((x - lo) * inv_h), 0, 2^bits - 1);  key = lexicographic cell id.
 * bounds = device [2 * n_dims]: lo[0..), inv_h[0..).  mccb200_box_bounds fills it with
 * the bounding box of the rows (lo = min, inv_h = 2^bits / (max - min)). */
#define MCCB200_BOX_MAX_DIMS 8
#define MCCB200_MAX_SUBSTEPS 4
MCCB200_API int mccb200_box_bounds(const float* p, uint64_t n, int ld, const int32_t* dims_host, int n_dims, int bits,
                       float* bounds, void* workspace, size_t workspace_bytes, void* stream);
MCCB200_API size_t mccb200_box_bounds_workspace_bytes(void);
MCCB200_API int mccb200_box_keys(const float* p, uint64_t n, int ld, const int32_t* dims_host, int n_dims, int bits,
                     const float* bounds, uint32_t* keys, void* stream);

/* Same two operations on the IMPLICIT children of y0 (never materialised): bounds and
 * keys of child c = i*k + j computed from parent i in-kernel. */
MCCB200_API int mccb200_box_child_bounds(const float* y0, uint64_t n, int d, int ld, const mccb200_drift* drift, float dt,
                             const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                             const float* keycols /* optional, see box_prk_rank_gather */, float* bounds,
                             void* workspace, size_t workspace_bytes, void* stream);
MCCB200_API int mccb200_box_child_keys(const float* y0, uint64_t n, int d, int ld, const mccb200_drift* drift, float dt,
                           const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                           const float* bounds, uint32_t* keys, void* stream);
/* minmax[q] = min, minmax[n_dims + q] = max of key dim q over the implicit children (what a partition over several
 * shards reduces across devices before forming bounds = (min, 2^bits / (max - min)) in float32). */
MCCB200_API int mccb200_box_child_minmax(const float* y0, uint64_t n, int d, int ld, const mccb200_drift* drift, float dt,
                             const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                             const float* keycols /* optional */, float* minmax, void* workspace,
                             size_t workspace_bytes, void* stream);

/* Fused BoxPartitionKernel + PartitioningRecombinationKernel + MonteCarloKernel step on
 * the implicit children (Hadamard cubature only): counting-sort ranks of all n*k
 * children by box key (stable in child id) without storing keys or ids, then selection of
 * local positions idx_local[0..out_per_part) in every partition of in_per_part children,
 * then the fused Euler update of the selected children.  Result == expand_all ->
 * box_keys -> sort_pairs -> expand_select(perm) bit for bit.  n_dims*bits <= 12.
 * Replaces PartitioningRecombinationKernel.__call__ (mccube/_kernels/base.py:152-177) applied to
 * the expansion of MCCSolver.step (mccube/_solvers.py:132-144). */
MCCB200_API size_t mccb200_box_prk_step_workspace_bytes(uint64_t n, int k, int n_dims, int bits, uint64_t in_per_part);
MCCB200_API int mccb200_box_prk_step(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                         const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                         const float* bounds, const uint32_t* idx_local, uint64_t out_per_part,
                         uint64_t in_per_part, float* y1, void* workspace, size_t workspace_bytes,
                         void* stream);

/* The same step as four calls, so that a caller can (a) keep the class tables, which depend on
 * (k, dims, bits) only, across steps and (b) draw idx_local and build its selection tables on a
 * second stream while the counting pass runs:
 *     tables    (once)          -> `tables`  (mccb200_box_prk_tables_bytes)
 *     selection (idx_local)     -> `sel`     (mccb200_box_prk_selection_bytes)
 *     count     (y0, bounds, tables)         -> workspace (same size as for box_prk_step)
 *     rank_gather (tables, sel, workspace of count) -> y1
 * box_prk_step == tables; selection; count; rank_gather on one stream.
 * Key-column hand-over (optional, 4 consecutive 16-byte aligned key columns only): rank_gather can
 * also write y1[:, dims[0] .. dims[0]+3] packed into keycols_out [n_out, 4]; the NEXT step's
 * box_child_bounds / box_prk_count then take that array as `keycols` (it must describe their y0
 * exactly) and read 16 contiguous bytes per parent instead of one DRAM fill per 256-byte row. */
MCCB200_API size_t mccb200_box_prk_tables_bytes(int k, int n_dims, int bits);
MCCB200_API int mccb200_box_prk_tables(int k, const int32_t* dims_host, int n_dims, int bits, void* tables,
                                       size_t tables_bytes, void* stream);
MCCB200_API size_t mccb200_box_prk_selection_bytes(uint64_t in_per_part, uint64_t out_per_part);
MCCB200_API int mccb200_box_prk_selection(const uint32_t* idx_local, uint64_t out_per_part, uint64_t in_per_part,
                                          void* sel, size_t sel_bytes, void* stream);
MCCB200_API int mccb200_box_prk_count(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                      const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                                      const float* bounds, uint64_t in_per_part, const void* tables,
                                      const float* keycols /* optional */, void* workspace, size_t workspace_bytes,
                                      void* stream);
MCCB200_API int mccb200_box_prk_rank_gather(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                            const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                                            uint64_t out_per_part, uint64_t in_per_part, const void* tables,
                                            const void* sel, float* y1, float* keycols_out /* optional */,
                                            void* workspace, size_t workspace_bytes, void* stream);
/* Pass B alone, for particles sharded over several devices with ONE global partition -- what the reference's
 * PartitioningRecombinationKernel computes on the concatenated state (mccube/_kernels/base.py:152-177).
 * `key_start[key]` is the GLOBAL sorted position of this shard's first child with that key, so ranks, partitions
 * and output slots are global; the LOCAL id (parent * k + point) of every selected child of this shard is
 * written to idx_slots[global slot] (pre-filled with 0xFFFFFFFF by the caller, one entry per global output row).
 * `mccb200_box_prk_key_start_offset` locates the LOCAL key starts in the workspace mccb200_box_prk_count filled
 * ((size_t)-1: unsupported shape). */
MCCB200_API int mccb200_box_prk_rank_global(const float* y0, uint64_t n, int d, const mccb200_drift* drift, float dt,
                                            const mccb200_cubature* cub, const int32_t* dims_host, int n_dims, int bits,
                                            uint64_t out_per_part, uint64_t in_per_part, const void* tables,
                                            const void* sel, const uint32_t* key_start, uint32_t* idx_slots,
                                            void* workspace, size_t workspace_bytes, void* stream);
MCCB200_API size_t mccb200_box_prk_key_start_offset(uint64_t n, int k, int n_dims, int bits, uint64_t in_per_part,
                                                    uint64_t out_per_part, uint32_t* n_keys);

/* ------------------------------------------------------------------ A11: moments
 * One pass each: sums[d] (+ weight sum) and centred second moments [d, d] in fp64:
 *     mean = sum_r w_r p_r / sum_r w_r          (center_of_mass, mccube/_metrics.py:96)
 *     m2   = sum_r w_r (p_r - c)(p_r - c)^T     (jnp.cov, README.md:86), c = centre[d]
 * w == NULL -> unit weights.  Outputs are accumulated INTO the fp64 buffers (zero them
 * first; cross-GPU reduction = one all-reduce of d + d*d + 1 doubles). */
MCCB200_API int mccb200_moments_sum(const float* p, uint64_t n, int d, int ld, const float* w, int w_stride,
                        double* sum_out /*[d+1]: sums then weight total*/, void* stream);
MCCB200_API int mccb200_moments_m2(const float* p, uint64_t n, int d, int ld, const float* w, int w_stride,
                       const float* centre, double* m2_out /*[d*d]*/, void* stream);

/* Full-covariance Gaussian drift f = -(y - mean) @ prec (config 4) written to f[n, d];
 * 3xTF32-free fp32 SIMT tiles in this round. */
MCCB200_API int mccb200_gaussian_full_drift(const float* y, uint64_t n, int d, int ld, const float* mean,
                                const float* prec /*[d, d] row-major*/, float* f, void* stream);

/* The same drift on the tensor cores (tcgen05.mma kind::tf32, fp32 accumulation in TMEM) with
 * the 3xTF32 operand split, |error| ~ 1e-6 relative (csrc/gemm_drift.cu).  prepare splits prec
 * once into prec_split [2][d][d] (mccb200_gaussian_full_drift_tc_prepare_bytes).  Shapes:
 * d % 32 == 0, 32 <= d <= 1024, ld % 4 == 0, 16-byte aligned pointers; f is [n, d] contiguous. */
MCCB200_API int mccb200_gaussian_full_drift_tc_supported(uint64_t n, int d, int ld);
MCCB200_API size_t mccb200_gaussian_full_drift_tc_prepare_bytes(int d);
MCCB200_API int mccb200_gaussian_full_drift_tc_prepare(const float* prec, int d, float* prec_split, void* stream);
MCCB200_API int mccb200_gaussian_full_drift_tc(const float* y, uint64_t n, int d, int ld, const float* mean,
                                   const float* prec_split, float* f, void* stream);

/* ------------------------------------------------------------------ sharded global resample
 * Multi-GPU form of MonteCarloKernel (the reference is single-device).  Every rank
 * holds n_loc consecutive parents and the SAME global index vector idx[n_tot] (a pure function
 * of (key, t)).  owner(s) = parent(idx[s]) / n_loc can materialise slot s locally;
 * dest(s) = s / n_loc stores it.
 *   shard_owner_keys: keys[s] = owner(s); counts[owner * n_ranks + dest] = #slots (device, uint32)
 *   (sort_pairs(keys, iota, log2 n_ranks) then groups the slots by owner, ascending slot inside)
 *   shard_send_ids  : out[q] = idx[order[start + q]] - child_offset  (local child ids to expand)
 *   scatter_rows    : out[slots[r] - slot_offset, :] = in[r, :]      (placement after the exchange) */
MCCB200_API int mccb200_shard_owner_keys(const uint32_t* idx, uint64_t n_tot, uint64_t n_loc, int k, int n_ranks,
                                         uint32_t* keys, uint32_t* counts, void* stream);
MCCB200_API int mccb200_shard_send_ids(const uint32_t* idx, const uint32_t* order, uint64_t start, uint64_t count,
                                       uint64_t child_offset, uint32_t* out, void* stream);
MCCB200_API int mccb200_scatter_rows(const float* in, int ld, const uint32_t* slots, uint64_t slot_offset,
                                     uint64_t n_rows, float* out, void* stream);

/* Fused expand + exchange for the sharded resample: every selected child of this rank's parents is
 * written straight into the output shard of the rank that stores its slot,
 *     peer_out[slot / rows_per_rank][(slot % rows_per_rank), :] = child(idx[slot]),
 * for the slots order[start .. start + count) of this rank's group (start / count are read on the device
 * from the replicated counts matrix, so the host never synchronises).  peer_out is a DEVICE array of
 * n_ranks pointers into peer memory (CUDA IPC mappings over NVLink; the own shard is an ordinary
 * pointer).  The caller separates steps with a cross-rank barrier.  Hadamard cubature only. */
/* Peer shards for mccb200_expand_scatter_peers: cudaMalloc buffers exchanged between the ranks of a node as
 * 64-byte CUDA IPC handles; peer_open (importer's device current) maps the exporter's buffer for kernels of
 * the importing device. */
MCCB200_API int mccb200_peer_alloc(size_t bytes, void** ptr);
MCCB200_API int mccb200_peer_free(void* ptr);
MCCB200_API int mccb200_peer_export(const void* ptr, unsigned char* handle64);
MCCB200_API int mccb200_peer_open(const unsigned char* handle64, void** ptr);
MCCB200_API int mccb200_peer_close(void* ptr);
/* Pull form of the same exchange for with_replacement=True: this rank draws only the indices of the output
 * slots it stores (mccb200_mc_indices_slice: the draws are counter based) and the gather kernel reads every
 * selected parent from the shard of the rank that owns it,
 *     y1[q, :] = child(idx[q]) with parent row peer_in[p / rows_per_rank][p % rows_per_rank, :],  p = idx[q] / k
 * (NVLink peer loads).  No index replication, no grouping pass, per-rank work O(n_loc).  Hadamard, drift kinds
 * 0 / 2.  peer_in: DEVICE array of n_ranks pointers to the INPUT shards (all of rows_per_rank rows). */
MCCB200_API int mccb200_mc_indices_slice(const mccb200_rng* rng, uint64_t n_inputs, uint64_t count, uint64_t first,
                                         uint64_t slice, uint32_t* idx_out, void* workspace, size_t workspace_bytes,
                                         void* stream);
MCCB200_API int mccb200_expand_gather_peers(float* const* peer_in, int n_ranks, uint64_t rows_per_rank, int d,
                                            const mccb200_drift* drift, float dt, const mccb200_cubature* cub,
                                            const uint32_t* idx /* or NULL */, const uint64_t* idx64 /* or NULL */,
                                            uint64_t n_out, float* y1, void* stream);
/* Index slice with jax_enable_x64 semantics (int64 randint: 64-bit draws, see oracle/jax_random.py::randint64;
 * unpinned: no JAX in the image).  n_inputs = n_tot * k may exceed 2^32: BASELINE config 5 (n = 2^27, k = 128). */
MCCB200_API int mccb200_mc_indices_slice64(const mccb200_rng* rng, uint64_t n_inputs, uint64_t count, uint64_t first,
                                           uint64_t slice, uint64_t* idx_out, void* workspace, size_t workspace_bytes,
                                           void* stream);
/* Lets kernels of the current device dereference memory of `peer_device` (needed once per peer before
 * mccb200_expand_scatter_peers; IPC mappings opened by another library do not imply it). */
MCCB200_API int mccb200_enable_peer_access(int peer_device);
MCCB200_API int mccb200_expand_scatter_peers(const float* y0, uint64_t n_loc, int d, const mccb200_drift* drift, float dt,
                                             const mccb200_cubature* cub, const uint32_t* idx, const uint32_t* order,
                                             const uint32_t* counts, int rank, int n_ranks, uint64_t rows_per_rank,
                                             float* const* peer_out, void* stream);
/* The same kernel driven by a list of global output slots (global Box partition over shards): idx_slots[slot] is
 * the LOCAL child id this rank contributes to `slot`, slot_list the slots it contributes to (any order),
 * counts[rank * n_ranks + 0..] their number (other rows zero).  mccb200_compact_slots builds slot_list / the count
 * from idx_slots on the device (entries != 0xFFFFFFFF; *cursor is zeroed by the caller; a full list traps). */
MCCB200_API int mccb200_expand_scatter_slots(const float* y0, uint64_t n_loc, int d, const mccb200_drift* drift, float dt,
                                             const mccb200_cubature* cub, const uint32_t* idx_slots,
                                             const uint32_t* slot_list, const uint32_t* counts, int rank, int n_ranks,
                                             uint64_t rows_per_rank, float* const* peer_out, void* stream);
/* Id exchange of the sharded global resample with replacement (SURVEY.md 8e "owner-computes"; the semantics are
 * jr.choice(..., replace=True) of /root/reference/mccube/_kernels/random.py:49-68 on the concatenated state): `idx` is
 * THIS rank's slice of the global index vector (child ids, uint32, or uint64 when idx_is_64), entry q belonging to
 * output slot first_slot + q.  slots_out[n] / ids_out[n] = (slot, child id relative to the owner's first parent),
 * grouped by owner rank = child / (n_loc * k) and ascending in slot inside a group -- the send buffers of the
 * exchange; counts_out[n_ranks] = group sizes.  n_loc * k <= 2^32 and slots < 2^32.  ids_out == NULL packs the two
 * into slots_out[n][2] = (slot, id) instead (one all-to-all; mccb200_unzip_pairs splits them on the receiving side). */
MCCB200_API size_t mccb200_shard_bucket_ids_workspace_bytes(int n_ranks);
MCCB200_API int mccb200_shard_bucket_ids(const void* idx, int idx_is_64, uint64_t n, uint64_t first_slot, uint64_t n_loc,
                                         int k, int n_ranks, uint32_t* slots_out, uint32_t* ids_out, uint32_t* counts_out,
                                         void* workspace, size_t workspace_bytes, void* stream);
/* The same with every owner's group in its own fixed-capacity region (group g starts at entry g * region_cap; 0 = back to
 * back): regions of a known size can be pushed to their owners with plain device-to-device copies before any count
 * is known on the host.  The caller checks counts_out[g] <= region_cap afterwards. */
MCCB200_API int mccb200_shard_bucket_ids_regions(const void* idx, int idx_is_64, uint64_t n, uint64_t first_slot,
                                                 uint64_t n_loc, int k, int n_ranks, uint64_t region_cap,
                                                 uint32_t* slots_out, uint32_t* ids_out, uint32_t* counts_out,
                                                 void* workspace, size_t workspace_bytes, void* stream);
/* cudaMemcpyAsync(dst, src, bytes, DeviceToDevice) on `stream`; dst may be a peer mapping (mccb200_peer_open). */
MCCB200_API int mccb200_peer_copy(void* dst, const void* src, size_t bytes, void* stream);
MCCB200_API int mccb200_unzip_pairs(const uint32_t* pairs /*[n][2]*/, uint64_t n, uint32_t* slots, uint32_t* ids, void* stream);
MCCB200_API int mccb200_compact_slots(const uint32_t* idx_slots, uint64_t n_slots, uint32_t* list, uint64_t capacity,
                                      uint32_t* cursor, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCCUBE_B200_H_ */
