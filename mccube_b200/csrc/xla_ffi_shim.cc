// XLA FFI custom-call shim: one handler per fused op of include/mccube_b200.h.
//
// This is the layer the brief asks for between the reference's Python (`MCCSolver.step`,
// /root/reference/mccube/_solvers.py:115-148, running inside diffrax's jitted while_loop) and the C ABI:
// `jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(lib.<name>), platform="CUDA")` +
// `jax.ffi.ffi_call(name, out_types)(*arrays, **attrs)` (integration/mccube/_b200.py shows every call).
//
// STATUS: UNVERIFIED UNDER REAL JAX.  jax / jaxlib and the XLA FFI headers are not in this image, so the file is
// not part of libmccube_b200.so.  What IS checked here (tests/test_ffi_shim.py, CPU): every handler body compiles
// against include/mccube_b200.h with a small stand-in for "xla/ffi/api/ffi.h" (tests/ffi_stub), so each C-ABI
// call is type-checked argument by argument.  Build where JAX exists:
//   nvcc -shared -Xcompiler -fPIC -I$(python -c 'import jax; print(jax.ffi.include_dir())') -Iinclude
//        mccube_b200/csrc/xla_ffi_shim.cc -Lmccube_b200 -lmccube_b200 -o libmccube_b200_ffi.so      (one command)
//
// Conventions: XLA owns every buffer; inputs are read-only, outputs and scratch are `Ret` buffers
// sized in Python from the `*_bytes()` queries; every handler only enqueues on XLA's stream and returns, so all
// of them are marked kCmdBufferCompatible (capturable in the while_loop body's command buffer).
// Drift encoding: attribute drift_kind (0 none, 1 array f[n,d], 2 diagonal Gaussian) and three buffers f, mean,
// inv_var; the ones a kind does not use are 1-element dummies.
#include <cuda_runtime.h>

#include <cstdint>

#include "xla/ffi/api/ffi.h"

#include "mccube_b200.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Fail(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInvalidArgument, mccb200_last_error());
}

using F32 = ffi::Buffer<ffi::F32>;
using F64 = ffi::Buffer<ffi::F64>;
using U32 = ffi::Buffer<ffi::U32>;
using U8 = ffi::Buffer<ffi::U8>;
template <typename B>
using Out = ffi::Result<B>;
using Dims = ffi::Span<const int32_t>;

mccb200_drift MakeDrift(int64_t kind, const F32& f, const F32& mean, const F32& inv_var) {
  mccb200_drift d{};
  d.kind = (int32_t)kind;
  d.f = kind == 1 ? f.typed_data() : nullptr;
  d.mean = kind == 2 ? mean.typed_data() : nullptr;
  d.inv_var = kind == 2 ? inv_var.typed_data() : nullptr;
  return d;
}

mccb200_cubature Hadamard(int64_t k, float scale) {
  mccb200_cubature c{};
  c.k = (int32_t)k;
  c.points = nullptr;
  c.scale = scale;
  c.lam = nullptr;
  return c;
}

mccb200_rng MakeRng(const U32& key, const F32& t, int64_t leaf, int64_t n_leaves, int64_t partitionable) {
  mccb200_rng r{};
  r.key_dev = key.typed_data();   // jr.key_data(self.key): 2 words on the device
  r.t_dev = t.typed_data();       // the traced float32 step time
  r.leaf = (uint32_t)leaf;
  r.n_leaves = (uint32_t)n_leaves;
  r.partitionable = (int32_t)partitionable;
  r.fold_time = 1;
  return r;
}

// ---- A8: idx[count] = jr.choice(key_at(t), P, (count,), replace) as indices   (mccube/_kernels/random.py:56-65)
ffi::Error McIndicesImpl(cudaStream_t s, U32 key, F32 t, int64_t n_inputs, int64_t replace, int64_t leaf,
                         int64_t n_leaves, int64_t partitionable, Out<U32> idx, Out<U8> workspace) {
  const mccb200_rng rng = MakeRng(key, t, leaf, n_leaves, partitionable);
  return Fail(mccb200_mc_indices(&rng, (uint64_t)n_inputs, (uint64_t)idx->element_count(), (int)replace,
                                 idx->typed_data(), workspace->typed_data(), workspace->size_bytes(), s));
}

// ---- A1-A5: y1[n_out, ld] = fused expand + select   (mccube/_solvers.py:125-144, _term.py:70-77)
ffi::Error ExpandSelectImpl(cudaStream_t s, F32 y0, F32 f, F32 mean, F32 inv_var, U32 idx, int64_t drift_kind,
                            int64_t weighted, int64_t k, float dt, float scale, Out<F32> y1) {
  const auto dims = y0.dimensions();
  const int ld = (int)dims[1], d = ld - (weighted ? 1 : 0);
  const mccb200_drift drift = MakeDrift(drift_kind, f, mean, inv_var);
  const mccb200_cubature cub = Hadamard(k, scale);
  return Fail(mccb200_expand_select(y0.typed_data(), (uint64_t)dims[0], d, (int)weighted, &drift, dt, &cub,
                                    idx.typed_data(), nullptr, 1, 1, (uint64_t)idx.element_count(), y1->typed_data(), s));
}

// ---- A1-A5 + A8 in one kernel: with_replacement=True child ids drawn inside the gather   (random.py:45-66)
ffi::Error ExpandSelectDrawImpl(cudaStream_t s, F32 y0, F32 f, F32 mean, F32 inv_var, U32 key, F32 t, int64_t drift_kind,
                                int64_t k, float dt, float scale, int64_t partitionable, Out<F32> y1, Out<U8> workspace) {
  const auto dims = y0.dimensions();
  const mccb200_drift drift = MakeDrift(drift_kind, f, mean, inv_var);
  const mccb200_cubature cub = Hadamard(k, scale);
  const mccb200_rng rng = MakeRng(key, t, 0, 1, partitionable);
  return Fail(mccb200_expand_select_draw(y0.typed_data(), (uint64_t)dims[0], (int)dims[1], &drift, dt, &cub, &rng,
                                         (uint64_t)y1->dimensions()[0], y1->typed_data(), nullptr,
                                         workspace->typed_data(), workspace->size_bytes(), s));
}

// ---- BinaryTreePartitioningKernel without the host callback   (mccube/_kernels/tree.py:44-58)
ffi::Error KdPartitionImpl(cudaStream_t s, F32 y, int64_t d, int64_t levels, Out<U32> idx, Out<U8> workspace) {
  const auto dims = y.dimensions();
  return Fail(mccb200_kd_partition(y.typed_data(), (uint64_t)dims[0], (int)d, (int)dims[1], (int)levels,
                                   idx->typed_data(), workspace->typed_data(), workspace->size_bytes(), s));
}

// ---- generic path: the whole [n*k, ld] expansion (arbitrary kernels, weights)   (_solvers.py:132-141)
ffi::Error ExpandAllImpl(cudaStream_t s, F32 y0, F32 f, F32 mean, F32 inv_var, int64_t drift_kind, int64_t weighted,
                         int64_t k, float dt, float scale, Out<F32> y1) {
  const auto dims = y0.dimensions();
  const int ld = (int)dims[1], d = ld - (weighted ? 1 : 0);
  const mccb200_drift drift = MakeDrift(drift_kind, f, mean, inv_var);
  const mccb200_cubature cub = Hadamard(k, scale);
  return Fail(mccb200_expand_all(y0.typed_data(), (uint64_t)dims[0], d, (int)weighted, &drift, dt, &cub,
                                 y1->typed_data(), s));
}

// ---- Box partition: bounding box of the implicit children -> [lo..., inv_h...]
ffi::Error BoxChildBoundsImpl(cudaStream_t s, F32 y0, F32 f, F32 mean, F32 inv_var, F32 keycols, int64_t drift_kind,
                              int64_t k, float dt, float scale, int64_t bits, int64_t has_keycols, Dims dims,
                              Out<F32> bounds, Out<U8> workspace) {
  const auto sh = y0.dimensions();
  const mccb200_drift drift = MakeDrift(drift_kind, f, mean, inv_var);
  const mccb200_cubature cub = Hadamard(k, scale);
  return Fail(mccb200_box_child_bounds(y0.typed_data(), (uint64_t)sh[0], (int)sh[1], (int)sh[1], &drift, dt, &cub,
                                       dims.begin(), (int)dims.size(), (int)bits,
                                       has_keycols ? keycols.typed_data() : nullptr, bounds->typed_data(),
                                       workspace->typed_data(), workspace->size_bytes(), s));
}

// ---- Box + PRK, part 1 (once per (k, dims, bits)): class tables
ffi::Error BoxPrkTablesImpl(cudaStream_t s, int64_t k, int64_t bits, Dims dims, Out<U8> tables) {
  return Fail(mccb200_box_prk_tables((int)k, dims.begin(), (int)dims.size(), (int)bits, tables->typed_data(),
                                     tables->size_bytes(), s));
}

// ---- part 2 (per step, may run on a second stream): selection tables of the shared local index vector
ffi::Error BoxPrkSelectionImpl(cudaStream_t s, U32 idx_local, int64_t in_per_part, Out<U8> sel) {
  return Fail(mccb200_box_prk_selection(idx_local.typed_data(), (uint64_t)idx_local.element_count(),
                                        (uint64_t)in_per_part, sel->typed_data(), sel->size_bytes(), s));
}

// ---- part 3: per-segment key histogram + scan (integer key-sort without the sort)
ffi::Error BoxPrkCountImpl(cudaStream_t s, F32 y0, F32 f, F32 mean, F32 inv_var, F32 bounds, U8 tables, F32 keycols,
                           int64_t drift_kind, int64_t k, float dt, float scale, int64_t bits, int64_t in_per_part,
                           int64_t has_keycols, Dims dims, Out<U8> workspace) {
  const auto sh = y0.dimensions();
  const mccb200_drift drift = MakeDrift(drift_kind, f, mean, inv_var);
  const mccb200_cubature cub = Hadamard(k, scale);
  return Fail(mccb200_box_prk_count(y0.typed_data(), (uint64_t)sh[0], (int)sh[1], &drift, dt, &cub, dims.begin(),
                                    (int)dims.size(), (int)bits, bounds.typed_data(), (uint64_t)in_per_part,
                                    tables.typed_data(), has_keycols ? keycols.typed_data() : nullptr,
                                    workspace->typed_data(), workspace->size_bytes(), s));
}

// ---- part 4: stable ranks + selection + gather (the fused ranking / gather kernel); `workspace` is the buffer
//      part 3 filled (input_output_aliases={workspace: workspace} on the Python side)
ffi::Error BoxPrkRankGatherImpl(cudaStream_t s, F32 y0, F32 f, F32 mean, F32 inv_var, U8 tables, U8 sel,
                                U8 workspace_in, int64_t drift_kind, int64_t k, float dt, float scale, int64_t bits,
                                int64_t out_per_part, int64_t in_per_part, int64_t want_keycols, Dims dims,
                                Out<F32> y1, Out<F32> keycols_out, Out<U8> workspace) {
  (void)workspace_in;   // aliased with `workspace`
  const auto sh = y0.dimensions();
  const mccb200_drift drift = MakeDrift(drift_kind, f, mean, inv_var);
  const mccb200_cubature cub = Hadamard(k, scale);
  return Fail(mccb200_box_prk_rank_gather(y0.typed_data(), (uint64_t)sh[0], (int)sh[1], &drift, dt, &cub, dims.begin(),
                                          (int)dims.size(), (int)bits, (uint64_t)out_per_part, (uint64_t)in_per_part,
                                          tables.typed_data(), sel.typed_data(), y1->typed_data(),
                                          want_keycols ? keycols_out->typed_data() : nullptr, workspace->typed_data(),
                                          workspace->size_bytes(), s));
}

// ---- parts 1-4 on one stream   (mccube/_kernels/base.py:152-177 with BoxPartitionKernel + MonteCarloKernel)
ffi::Error BoxPrkStepImpl(cudaStream_t s, F32 y0, F32 f, F32 mean, F32 inv_var, F32 bounds, U32 idx_local,
                          int64_t drift_kind, int64_t k, float dt, float scale, int64_t bits, int64_t in_per_part,
                          Dims dims, Out<F32> y1, Out<U8> workspace) {
  const auto sh = y0.dimensions();
  const mccb200_drift drift = MakeDrift(drift_kind, f, mean, inv_var);
  const mccb200_cubature cub = Hadamard(k, scale);
  return Fail(mccb200_box_prk_step(y0.typed_data(), (uint64_t)sh[0], (int)sh[1], &drift, dt, &cub, dims.begin(),
                                   (int)dims.size(), (int)bits, bounds.typed_data(), idx_local.typed_data(),
                                   (uint64_t)idx_local.element_count(), (uint64_t)in_per_part, y1->typed_data(),
                                   workspace->typed_data(), workspace->size_bytes(), s));
}

// ---- A6, config 4: full-covariance drift on the tensor cores (tcgen05 kind::tf32, 3xTF32)
ffi::Error FullDriftPrepareImpl(cudaStream_t s, F32 prec, Out<F32> prec_split) {
  return Fail(mccb200_gaussian_full_drift_tc_prepare(prec.typed_data(), (int)prec.dimensions()[0],
                                                     prec_split->typed_data(), s));
}
ffi::Error FullDriftImpl(cudaStream_t s, F32 y, F32 mean, F32 prec_split, Out<F32> f) {
  const auto sh = y.dimensions();
  return Fail(mccb200_gaussian_full_drift_tc(y.typed_data(), (uint64_t)sh[0], (int)sh[1], (int)sh[1], mean.typed_data(),
                                             prec_split.typed_data(), f->typed_data(), s));
}
// any shape / alignment (fp32 SIMT tiles)
ffi::Error FullDriftSimtImpl(cudaStream_t s, F32 y, F32 mean, F32 prec, Out<F32> f) {
  const auto sh = y.dimensions();
  return Fail(mccb200_gaussian_full_drift(y.typed_data(), (uint64_t)sh[0], (int)sh[1], (int)sh[1], mean.typed_data(),
                                          prec.typed_data(), f->typed_data(), s));
}

// ---- A11: weighted column sums / centred second moment in float64   (mccube/_metrics.py:79-96, README.md:85-86)
ffi::Error MomentsSumImpl(cudaStream_t s, F32 p, F32 w, int64_t d, int64_t weighted, Out<F64> sums) {
  const auto sh = p.dimensions();
  cudaMemsetAsync(sums->typed_data(), 0, sums->size_bytes(), s);   // the kernel accumulates
  return Fail(mccb200_moments_sum(p.typed_data(), (uint64_t)sh[0], (int)d, (int)sh[1], weighted ? w.typed_data() : nullptr,
                                  1, sums->typed_data(), s));
}
ffi::Error MomentsM2Impl(cudaStream_t s, F32 p, F32 w, F32 centre, int64_t d, int64_t weighted, Out<F64> m2) {
  const auto sh = p.dimensions();
  cudaMemsetAsync(m2->typed_data(), 0, m2->size_bytes(), s);
  return Fail(mccb200_moments_m2(p.typed_data(), (uint64_t)sh[0], (int)d, (int)sh[1], weighted ? w.typed_data() : nullptr,
                                 1, centre.typed_data(), m2->typed_data(), s));
}

// ---- A10: StratifiedPartitioningKernel keys   (mccube/_kernels/stratified.py:36-45)
ffi::Error StratifiedKeysImpl(cudaStream_t s, F32 p, F32 com, int64_t d, Out<U32> keys) {
  const auto sh = p.dimensions();
  return Fail(mccb200_stratified_keys(p.typed_data(), (uint64_t)sh[0], (int)d, (int)sh[1], com.typed_data(),
                                      keys->typed_data(), s));
}

// ---- lax.sort_key_val(is_stable=True) / jnp.argsort on bits [0, key_bits)
ffi::Error SortPairsImpl(cudaStream_t s, U32 keys, U32 values, int64_t key_bits, int64_t has_values, Out<U32> keys_out,
                         Out<U32> values_out, Out<U8> workspace) {
  return Fail(mccb200_sort_pairs(keys.typed_data(), has_values ? values.typed_data() : nullptr,
                                 (uint64_t)keys.element_count(), (int)key_bits, keys_out->typed_data(),
                                 values_out->typed_data(), workspace->typed_data(), workspace->size_bytes(), s));
}

// ---- p[idx]   (mccube/_kernels/random.py:65, base.py:170-177)
ffi::Error GatherRowsImpl(cudaStream_t s, F32 p, U32 idx, Out<F32> out) {
  return Fail(mccb200_gather_rows(p.typed_data(), (int)p.dimensions()[1], idx.typed_data(), (uint64_t)idx.element_count(),
                                  out->typed_data(), s));
}

// ---- weight renormalisation after recombination   (mccube/_solvers.py:146, _utils.py:47-51)
ffi::Error NormaliseWeightsImpl(cudaStream_t s, F32 y_in, Out<F32> y, Out<F64> scratch) {
  (void)y_in;   // aliased with y (input_output_aliases={0: 0})
  const auto sh = y->dimensions();
  return Fail(mccb200_normalise_weights(y->typed_data(), (uint64_t)sh[0], (int)sh[1], scratch->typed_data(), s));
}

}  // namespace

#define MCCB_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
#define MCCB_DRIFT_ARGS .Arg<F32>().Arg<F32>().Arg<F32>() /* f, mean, inv_var */

XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_mc_indices_ffi, McIndicesImpl,
    MCCB_STREAM.Arg<U32>().Arg<F32>().Attr<int64_t>("n_inputs").Attr<int64_t>("replace").Attr<int64_t>("leaf")
        .Attr<int64_t>("n_leaves").Attr<int64_t>("partitionable").Ret<U32>().Ret<U8>(),
    {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_expand_select_ffi, ExpandSelectImpl,
    MCCB_STREAM.Arg<F32>() MCCB_DRIFT_ARGS.Arg<U32>().Attr<int64_t>("drift_kind").Attr<int64_t>("weighted")
        .Attr<int64_t>("k").Attr<float>("dt").Attr<float>("scale").Ret<F32>(),
    {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_expand_select_draw_ffi, ExpandSelectDrawImpl,
    MCCB_STREAM.Arg<F32>() MCCB_DRIFT_ARGS.Arg<U32>().Arg<F32>().Attr<int64_t>("drift_kind").Attr<int64_t>("k")
        .Attr<float>("dt").Attr<float>("scale").Attr<int64_t>("partitionable").Ret<F32>().Ret<U8>(),
    {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_kd_partition_ffi, KdPartitionImpl,
    MCCB_STREAM.Arg<F32>().Attr<int64_t>("d").Attr<int64_t>("levels").Ret<U32>().Ret<U8>(),
    {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_expand_all_ffi, ExpandAllImpl,
    MCCB_STREAM.Arg<F32>() MCCB_DRIFT_ARGS.Attr<int64_t>("drift_kind").Attr<int64_t>("weighted").Attr<int64_t>("k")
        .Attr<float>("dt").Attr<float>("scale").Ret<F32>(),
    {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_box_child_bounds_ffi, BoxChildBoundsImpl,
    MCCB_STREAM.Arg<F32>() MCCB_DRIFT_ARGS.Arg<F32>().Attr<int64_t>("drift_kind").Attr<int64_t>("k").Attr<float>("dt")
        .Attr<float>("scale").Attr<int64_t>("bits").Attr<int64_t>("has_keycols").Attr<Dims>("dims").Ret<F32>().Ret<U8>(),
    {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_box_prk_tables_ffi, BoxPrkTablesImpl,
    MCCB_STREAM.Attr<int64_t>("k").Attr<int64_t>("bits").Attr<Dims>("dims").Ret<U8>(),
    {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_box_prk_selection_ffi, BoxPrkSelectionImpl,
    MCCB_STREAM.Arg<U32>().Attr<int64_t>("in_per_part").Ret<U8>(), {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_box_prk_count_ffi, BoxPrkCountImpl,
    MCCB_STREAM.Arg<F32>() MCCB_DRIFT_ARGS.Arg<F32>().Arg<U8>().Arg<F32>().Attr<int64_t>("drift_kind").Attr<int64_t>("k")
        .Attr<float>("dt").Attr<float>("scale").Attr<int64_t>("bits").Attr<int64_t>("in_per_part")
        .Attr<int64_t>("has_keycols").Attr<Dims>("dims").Ret<U8>(),
    {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_box_prk_rank_gather_ffi, BoxPrkRankGatherImpl,
    MCCB_STREAM.Arg<F32>() MCCB_DRIFT_ARGS.Arg<U8>().Arg<U8>().Arg<U8>().Attr<int64_t>("drift_kind").Attr<int64_t>("k")
        .Attr<float>("dt").Attr<float>("scale").Attr<int64_t>("bits").Attr<int64_t>("out_per_part")
        .Attr<int64_t>("in_per_part").Attr<int64_t>("want_keycols").Attr<Dims>("dims").Ret<F32>().Ret<F32>().Ret<U8>(),
    {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_box_prk_step_ffi, BoxPrkStepImpl,
    MCCB_STREAM.Arg<F32>() MCCB_DRIFT_ARGS.Arg<F32>().Arg<U32>().Attr<int64_t>("drift_kind").Attr<int64_t>("k")
        .Attr<float>("dt").Attr<float>("scale").Attr<int64_t>("bits").Attr<int64_t>("in_per_part").Attr<Dims>("dims")
        .Ret<F32>().Ret<U8>(),
    {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_gaussian_full_drift_tc_prepare_ffi, FullDriftPrepareImpl,
    MCCB_STREAM.Arg<F32>().Ret<F32>(), {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_gaussian_full_drift_tc_ffi, FullDriftImpl,
    MCCB_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>(), {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_gaussian_full_drift_ffi, FullDriftSimtImpl,
    MCCB_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Ret<F32>(), {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_moments_sum_ffi, MomentsSumImpl,
    MCCB_STREAM.Arg<F32>().Arg<F32>().Attr<int64_t>("d").Attr<int64_t>("weighted").Ret<F64>(),
    {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_moments_m2_ffi, MomentsM2Impl,
    MCCB_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Attr<int64_t>("d").Attr<int64_t>("weighted").Ret<F64>(),
    {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_stratified_keys_ffi, StratifiedKeysImpl,
    MCCB_STREAM.Arg<F32>().Arg<F32>().Attr<int64_t>("d").Ret<U32>(), {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_sort_pairs_ffi, SortPairsImpl,
    MCCB_STREAM.Arg<U32>().Arg<U32>().Attr<int64_t>("key_bits").Attr<int64_t>("has_values").Ret<U32>().Ret<U32>().Ret<U8>(),
    {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_gather_rows_ffi, GatherRowsImpl,
    MCCB_STREAM.Arg<F32>().Arg<U32>().Ret<F32>(), {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(mccb200_normalise_weights_ffi, NormaliseWeightsImpl,
    MCCB_STREAM.Arg<F32>().Ret<F32>().Ret<F64>(), {ffi::Traits::kCmdBufferCompatible});
