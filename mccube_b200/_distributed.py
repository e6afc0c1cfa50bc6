"""Multi-GPU host logic (one process per GPU, torch.distributed).

The reference has no multi-device code at all.  What shards naturally:
  * expansion + drift: embarrassingly parallel over parents -> contiguous row shards;
  * partitioned recombination: each rank partitions and recombines the children of its own
    shard ("device-local partitioning", a labelled variant of the reference's single global
    partition) -- no data-path collective; every rank draws the SAME local index vector, which
    is what quirk Q4 prescribes for every partition anyway;
  * moments / W2: per-rank fp64 partial sums + ONE all-reduce of d + 1 (+ d*d) doubles.
These helpers are device agnostic (they only touch torch.distributed and small tensors), so the
N > 1 logic is covered by world_size-2 gloo tests on CPU (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous row range [lo, hi) of `rank` when n rows are split over `world` ranks
    (first n % world ranks get one extra row)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def global_mean(local_sums: torch.Tensor, group=None):
    """local_sums = [sum_r w_r x_r (d entries), sum_r w_r] of this rank's shard (fp64) ->
    (global mean[d], global weight total) after one all-reduce."""
    tot = all_reduce_sum_(local_sums.clone(), group)
    return tot[:-1] / tot[-1], tot[-1]


def global_cov(local_m2: torch.Tensor, local_sums: torch.Tensor, centre: torch.Tensor, mean: torch.Tensor,
               wtot: torch.Tensor, ddof: int = 1, group=None, local_w2=None):
    """local_m2 = sum_r w_r (x_r - centre)(x_r - centre)^T over this rank's shard with a centre
    COMMON to all ranks (fp64).  One all-reduce, then the exact shift to the global mean:
    sum w (x-m)(x-m)^T = sum w (x-c)(x-c)^T - W (m-c)(m-c)^T.
    Normalisation as numpy / jnp.cov: `wtot - ddof` without weights, `wtot - ddof * sum(w^2) / wtot`
    with aweights (`local_w2` = this rank's sum of squared weights, reduced in the same all-reduce)."""
    del local_sums
    d = local_m2.shape[0]
    packed = torch.cat([local_m2.reshape(-1), (local_w2 if local_w2 is not None else local_m2.new_zeros(())).reshape(1)])
    packed = all_reduce_sum_(packed, group)
    m2 = packed[:-1].reshape(d, d)
    delta = mean - centre.to(mean.dtype)
    m2 = m2 - wtot * torch.outer(delta, delta)
    fact = wtot - ddof if local_w2 is None else wtot - ddof * packed[-1] / wtot
    return m2 / fact


def distributed_mean_and_cov(particles: torch.Tensor, weights=None, ddof: int = 1, group=None):
    """jnp.mean / jnp.cov(aweights=weights, ddof=ddof) of the GLOBAL particle set from per-rank shards: two GPU
    reductions (mccb200_moments_sum, mccb200_moments_m2) and two tiny all-reduces."""
    from . import _lib

    d = particles.shape[1]
    if weights is not None:
        weights = torch.as_tensor(weights, device=particles.device).to(torch.float32)   # the kernels read float32
    mean_local, wtot_local = _lib.column_means(particles.contiguous(), d, weights)
    sums = torch.cat([mean_local * wtot_local, wtot_local.reshape(1)])
    mean, wtot = global_mean(sums, group)
    centre = mean.to(torch.float32).contiguous()
    m2 = _lib.second_moment(particles.contiguous(), d, centre, weights)
    w2 = None if weights is None else (weights.double() * weights.double()).sum()
    cov = global_cov(m2, sums, centre, mean, wtot, ddof, group, local_w2=w2)
    return mean, cov


# ------------------------------------------------------------------ sharded global resample
def plan_blocks(counts, rank: int):
    """Pure host arithmetic on the replicated counts matrix counts[owner, dest] (number of output
    slots stored by `dest` whose selected parent lives on `owner`).  With the slots stably
    grouped by owner (ascending slot inside a group, hence ascending dest):
      send = (start, count) of the block this rank materialises, and its per-dest split sizes;
      recv = per-source split sizes and, for each source o, the (start, count) of the sub-block
             of o's group that is destined to this rank (arrival order == ascending slot)."""
    import numpy as np

    counts = np.asarray(counts, dtype=np.int64)
    per_owner = counts.sum(axis=1)
    owner_start = np.concatenate([[0], np.cumsum(per_owner)[:-1]])
    send = dict(start=int(owner_start[rank]), count=int(per_owner[rank]), splits=[int(c) for c in counts[rank]])
    recv_blocks = [(int(owner_start[o] + counts[o, :rank].sum()), int(counts[o, rank])) for o in range(counts.shape[0])]
    recv = dict(splits=[int(c) for c in counts[:, rank]], blocks=recv_blocks, count=int(counts[:, rank].sum()))
    return send, recv


def plan_rebalance(per_owner, n_loc: int):
    """Deterministic rebalancing of owner-computed rows (pure host arithmetic, identical on every
    rank).  per_owner[r] = rows rank r materialises from its own parents (sums to G * n_loc).
    Every rank keeps its first min(per_owner[r], n_loc) rows; ranks with a surplus give their
    LAST rows, in rank order, to the ranks with a deficit, in rank order.
    Returns transfers[src][dst] = row count, and for every src the offset (within its surplus
    rows) at which the rows for dst begin."""
    import numpy as np

    per_owner = np.asarray(per_owner, dtype=np.int64)
    G = per_owner.shape[0]
    surplus = np.maximum(per_owner - n_loc, 0)
    deficit = np.maximum(n_loc - per_owner, 0)
    assert surplus.sum() == deficit.sum()
    transfers = np.zeros((G, G), dtype=np.int64)
    src, dst = 0, 0
    s_left, d_left = surplus.copy(), deficit.copy()
    while True:
        while src < G and s_left[src] == 0:
            src += 1
        while dst < G and d_left[dst] == 0:
            dst += 1
        if src >= G or dst >= G:
            break
        c = min(s_left[src], d_left[dst])
        transfers[src, dst] += c
        s_left[src] -= c
        d_left[dst] -= c
    return transfers


def exchange_rows(send_buf: torch.Tensor, send_splits, recv_splits, group=None) -> torch.Tensor:
    """Variable-size all-to-all of rows (NCCL over NVLink on GPUs; gloo on CPU for the tests)."""
    recv = send_buf.new_empty((sum(recv_splits),) + tuple(send_buf.shape[1:]))
    dist.all_to_all_single(recv, send_buf.contiguous(), output_split_sizes=list(recv_splits),
                           input_split_sizes=list(send_splits), group=group)
    return recv


class PeerShards:
    """Two ping-pong output shards per rank, every rank's shards mapped into every other rank's address
    space (cudaMalloc buffers + CUDA IPC handles gathered over the process group; the importer opens
    them with cudaIpcMemLazyEnablePeerAccess, i.e. NVLink peer loads / stores from its kernels).
    `ptrs[b]` is an int64 device tensor with the base pointer of buffer b on every rank."""

    def __init__(self, n_loc: int, d: int, rank: int, world: int, device, group=None):
        from . import _lib

        self.rank, self.world, self.group = rank, world, group
        self.local = [_lib.peer_alloc((n_loc, d), device) for _ in range(2)]
        handles = [_lib.peer_export(t) for t in self.local]
        gathered = [None] * world
        dist.all_gather_object(gathered, handles, group=group)
        self.opened, self.ptrs = [], []
        for b in range(2):
            row = []
            for r in range(world):
                if r == rank:
                    row.append(self.local[b].data_ptr())
                else:
                    ptr = _lib.peer_open(gathered[r][b], device)
                    self.opened.append(ptr)
                    row.append(ptr)
            self.ptrs.append(torch.tensor(row, dtype=torch.int64, device=device))
        self.flag = torch.zeros(1, dtype=torch.float32, device=device)
        self.turn = 0

    def load(self, y_shard: torch.Tensor):
        """Copies this rank's state into the current buffer (start of a `step_pull` trajectory)."""
        self.local[self.turn].copy_(y_shard)
        self.barrier()
        return self.local[self.turn]

    def barrier(self):
        """Stream-ordered cross-rank barrier (a 4-byte all-reduce: synchronisation only, no data path)."""
        dist.all_reduce(self.flag, group=self.group)

    def close(self):
        """Unmaps the peers' buffers and frees the own ones (every rank must have stopped using them)."""
        from . import _lib

        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for ptr in self.opened:
            _lib.peer_close(ptr)
        self.opened = []
        dist.barrier(group=self.group)
        for t in self.local:
            _lib.peer_free(t)
        self.local = []


class PeerIdBuffers:
    """Receive side of the copy-engine id exchange: per rank, two (ping-pong) buffers of G fixed-capacity regions of
    (slot, id) pairs -- region `src` is written by rank `src` -- and a [G, G] counts matrix, all mapped into every
    other rank's address space (CUDA IPC, as `PeerShards`).  Plus a local staging buffer of the same shape."""

    def __init__(self, n_loc: int, rank: int, world: int, device, group=None):
        import math

        from . import _lib

        self.rank, self.world, self.group = rank, world, group
        mean = n_loc / world
        self.cap = int(mean + 8.0 * math.sqrt(mean) + 1024)           # 8 sigma of the binomial group size
        self.cap = (self.cap + 63) // 64 * 64
        words = 2 * world * self.cap                                   # int32 words of one pairs buffer
        self.stage = torch.empty((world * self.cap, 2), dtype=torch.int32, device=device)
        self.stage_counts = torch.empty(world, dtype=torch.int32, device=device)
        self.local = [_lib.peer_alloc((words + world * world,), device) for _ in range(2)]   # pairs | counts matrix
        handles = [_lib.peer_export(t) for t in self.local]
        gathered = [None] * world
        dist.all_gather_object(gathered, handles, group=group)
        self.opened, self.base = [], []
        for b in range(2):
            row = []
            for r in range(world):
                if r == rank:
                    row.append(self.local[b].data_ptr())
                else:
                    ptr = _lib.peer_open(gathered[r][b], device)
                    self.opened.append(ptr)
                    row.append(ptr)
            self.base.append(row)
        self.words = words
        self.flag = torch.zeros(1, dtype=torch.float32, device=device)
        self.host_counts = [torch.empty((world, world), dtype=torch.int32, pin_memory=True) for _ in range(2)]
        self.turn = 0

    def pairs(self, b: int) -> torch.Tensor:
        return self.local[b][:self.words].view(torch.int32).view(self.world * self.cap, 2)

    def counts(self, b: int) -> torch.Tensor:
        return self.local[b][self.words:].view(torch.int32).view(self.world, self.world)

    def close(self):
        from . import _lib

        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for ptr in self.opened:
            _lib.peer_close(ptr)
        self.opened = []
        dist.barrier(group=self.group)
        for t in self.local:
            _lib.peer_free(t)
        self.local = []


class ShardedMonteCarloStep:
    """One MCCSolver step with a GLOBAL MonteCarloKernel resample over row-sharded particles.

    Every rank draws the same global index vector (bit-exact with the single-GPU / reference
    one), materialises the selected children of ITS OWN parents with the fused expand+select
    kernel (no parent row ever crosses the link), and one variable-size all-to-all moves each
    child row to the rank that stores its output slot.  The concatenation of the per-rank results
    is bit-identical to the single-device step on the concatenated state.
    """

    def __init__(self, solver, terms, n_ranks: int):
        from ._kernels import MonteCarloKernel

        if type(solver.recombination_kernel) is not MonteCarloKernel or solver.weighted or solver.n_substeps != 1:
            raise NotImplementedError("sharded step: unweighted MonteCarloKernel with n_substeps=1 only")
        self.solver, self.terms, self.n_ranks = solver, terms, n_ranks

    def prepare(self, t0, t1, y_shard: torch.Tensor, rank: int):
        """Local work before the exchange: (send_buf, send_splits, recv_splits, placement)."""
        import math

        from . import _lib

        solver, kernel = self.solver, self.solver.recombination_kernel
        n_loc, d = y_shard.shape
        n_tot = n_loc * self.n_ranks
        if kernel.recombination_count != n_tot:
            raise ValueError(f"MonteCarloKernel.recombination_count must be the GLOBAL particle count {n_tot}")
        drift, cub, dt, k, _ = solver._describe(self.terms, y_shard, *solver.substep_times(t0, t1)[0], None)
        idx = kernel.indices(t0, n_tot * k, n_tot, y_shard.device)            # replicated, deterministic
        keys, counts_dev = _lib.shard_owner_keys(idx, n_loc, k, self.n_ranks)
        order = _lib.sort_pairs(keys, None, max(1, math.ceil(math.log2(self.n_ranks))))
        counts = counts_dev.cpu().numpy()                                     # host split sizes for the collective
        send, recv = plan_blocks(counts, rank)
        ids = _lib.shard_send_ids(idx, order, send["start"], send["count"], rank * n_loc * k)
        send_buf = _lib.expand_select(y_shard, d, False, drift, dt, cub, ids, n_out=send["count"])
        slots = torch.cat([order[s:s + c] for (s, c) in recv["blocks"]]) if recv["count"] else order[:0]
        return send_buf, send["splits"], recv["splits"], (slots.contiguous(), rank * n_loc, n_loc, d)

    @staticmethod
    def place(recv_buf: torch.Tensor, placement) -> torch.Tensor:
        from . import _lib

        slots, slot_offset, n_loc, d = placement
        y1 = torch.empty((n_loc, d), dtype=torch.float32, device=recv_buf.device)
        return _lib.scatter_rows_(y1, recv_buf.contiguous(), slots, slot_offset)

    def step(self, t0, t1, y_shard: torch.Tensor, rank: int, group=None) -> torch.Tensor:
        send_buf, send_splits, recv_splits, placement = self.prepare(t0, t1, y_shard.contiguous(), rank)
        if self.n_ranks == 1:
            return self.place(send_buf, placement)
        return self.place(exchange_rows(send_buf, send_splits, recv_splits, group), placement)

    # ---- fused expand + exchange over peer memory: no send buffer, no collective in the data path
    def step_p2p(self, t0, t1, y_shard: torch.Tensor, rank: int, shards: PeerShards) -> torch.Tensor:
        """Same result as `step` (shards bit-identical to the single-device step), but the gather kernel
        writes every selected child directly into the output shard of the rank that stores its slot
        (NVLink peer stores), and nothing is synchronised with the host: the block of slots this rank
        materialises is read on the device from the replicated counts matrix.  `y_shard` must not be
        the ping-pong buffer this call writes (it alternates; feed the returned tensor back in)."""
        import math

        from . import _lib

        solver, kernel = self.solver, self.solver.recombination_kernel
        n_loc, d = y_shard.shape
        n_tot = n_loc * self.n_ranks
        if kernel.recombination_count != n_tot:
            raise ValueError(f"MonteCarloKernel.recombination_count must be the GLOBAL particle count {n_tot}")
        drift, cub, dt, k, _ = solver._describe(self.terms, y_shard, *solver.substep_times(t0, t1)[0], None)
        idx = kernel.indices(t0, n_tot * k, n_tot, y_shard.device)            # replicated, deterministic
        keys, counts_dev = _lib.shard_owner_keys(idx, n_loc, k, self.n_ranks)
        order = _lib.sort_pairs(keys, None, max(1, math.ceil(math.log2(self.n_ranks))))
        b = shards.turn
        out = shards.local[b]
        if out.data_ptr() == y_shard.data_ptr():
            raise ValueError("step_p2p: the input is the buffer this step writes; feed the previous result back in")
        _lib.expand_scatter_peers(y_shard.contiguous(), d, drift, dt, cub, idx, order, counts_dev, rank, self.n_ranks,
                                  n_loc, shards.ptrs[b])
        shards.barrier()          # every rank's rows have landed (and nobody still reads the other buffer)
        shards.turn ^= 1
        return out

    def step_pull(self, t0, t1, rank: int, shards: PeerShards) -> torch.Tensor:
        """with_replacement=True only.  The state lives in `shards` (load it once with `shards.load`): this
        rank draws just the indices of the output slots it stores and the gather kernel reads every selected
        parent from the shard of the rank that owns it (NVLink peer loads), writing its own next shard.
        Per-rank work is O(n_loc): no replicated index vector, no grouping pass, no collective in the data
        path, no host synchronisation.  Shards are bit-identical to the single-device step."""
        from . import _lib

        solver, kernel = self.solver, self.solver.recombination_kernel
        cur, nxt = shards.turn, shards.turn ^ 1
        y_in, y_out = shards.local[cur], shards.local[nxt]
        n_loc, d = y_in.shape
        n_tot = n_loc * self.n_ranks
        if kernel.recombination_count != n_tot:
            raise ValueError(f"MonteCarloKernel.recombination_count must be the GLOBAL particle count {n_tot}")
        drift, cub, dt, k, _ = solver._describe(self.terms, y_in, *solver.substep_times(t0, t1)[0], None)
        idx = kernel.indices_slice(t0, n_tot * k, n_tot, rank * n_loc, n_loc, y_in.device)   # int64 when kernel.x64
        _lib.expand_gather_peers(shards.ptrs[cur], self.n_ranks, n_loc, d, drift, dt, cub, idx, y_out)
        shards.barrier()          # everybody has finished reading the `cur` buffers
        shards.turn = nxt
        return y_out

    # ---- owner-computes variant: no row leaves its rank except the binomial imbalance
    def prepare_owner(self, t0, t1, y_shard: torch.Tensor, rank: int):
        """Every rank materialises the selected children of its own parents in place (slot order)
        and only the surplus over n_loc rows is exchanged, so the collective moves O(sqrt(n_loc))
        rows instead of (G-1)/G of the state.  The global multiset of particles equals the
        reference's; the row ORDER differs by a known permutation, returned as `slot_map`
        (global output slot of every local row) so parity can be checked exactly.
        Returns (y1 with the kept rows filled, send_buf, send_splits, recv_splits, keep, slot_map_parts)."""
        import math

        from . import _lib

        solver, kernel = self.solver, self.solver.recombination_kernel
        n_loc, d = y_shard.shape
        G = self.n_ranks
        n_tot = n_loc * G
        drift, cub, dt, k, _ = solver._describe(self.terms, y_shard, *solver.substep_times(t0, t1)[0], None)
        idx = kernel.indices(t0, n_tot * k, n_tot, y_shard.device)
        keys, counts_dev = _lib.shard_owner_keys(idx, n_loc, k, G)
        order = _lib.sort_pairs(keys, None, max(1, math.ceil(math.log2(G))))
        counts = counts_dev.cpu().numpy()
        per_owner = counts.sum(axis=1)
        starts = [int(x) for x in ([0] + list(per_owner.cumsum()[:-1]))]
        transfers = plan_rebalance(per_owner, n_loc)
        keep = int(min(per_owner[rank], n_loc))
        y1 = torch.empty((n_loc, d), dtype=torch.float32, device=y_shard.device)
        base = rank * n_loc * k
        if keep:
            ids = _lib.shard_send_ids(idx, order, starts[rank], keep, base)
            _lib.expand_select(y_shard, d, False, drift, dt, cub, ids, n_out=keep, out=y1[:keep])
        n_give = int(transfers[rank].sum())
        ids = _lib.shard_send_ids(idx, order, starts[rank] + n_loc, n_give, base)
        send_buf = (_lib.expand_select(y_shard, d, False, drift, dt, cub, ids, n_out=n_give) if n_give
                    else y1.new_empty((0, d)))
        # slot map: kept rows, then received rows in giver order (the giver's surplus rows are
        # order[starts[g] + n_loc + offset ...], offset = rows it gave to lower-ranked takers)
        parts = [order[starts[rank]:starts[rank] + keep]]
        for g in range(G):
            c = int(transfers[g, rank])
            if c:
                off = int(transfers[g, :rank].sum())
                parts.append(order[starts[g] + n_loc + off: starts[g] + n_loc + off + c])
        return (y1, send_buf, [int(c) for c in transfers[rank]], [int(c) for c in transfers[:, rank]], keep,
                torch.cat(parts))

    def step_owner(self, t0, t1, y_shard: torch.Tensor, rank: int, group=None):
        """Returns (y1_shard, slot_map): y1_shard[q] is the reference's output row slot_map[q]."""
        y1, send_buf, send_splits, recv_splits, keep, slot_map = self.prepare_owner(t0, t1, y_shard.contiguous(), rank)
        if self.n_ranks > 1:
            recv = exchange_rows(send_buf, send_splits, recv_splits, group)
            if recv.shape[0]:
                y1[keep:keep + recv.shape[0]] = recv
        return y1, slot_map


    # ---- id exchange + owner-computes: O(n_loc) memory per rank, 64-bit child ids, no row leaves its owner
    def owner_ids_bucket(self, t0, n_loc: int, rank: int, device, packed: bool = False):
        """Phase 1 (local): this rank draws ONLY the index slice of the output slots [rank * n_loc, (rank+1) * n_loc)
        (`with_replacement=True`: the draws are counter based) and groups (slot, LOCAL child id) by the rank that owns
        the selected parent, stably (ascending slot inside a group).  Depends on (key, t0) only, not on the state.
        Returns (slots [n_loc] int32, ids [n_loc] int32, per-owner counts as a device tensor [G])."""
        from . import _lib

        kernel = self.solver.recombination_kernel
        if not kernel.with_replacement:
            raise ValueError("step_owner_ids needs with_replacement=True (the shuffle is not sliceable)")
        G = self.n_ranks
        n_tot = n_loc * G
        if kernel.recombination_count != n_tot:
            raise ValueError(f"MonteCarloKernel.recombination_count must be the GLOBAL particle count {n_tot}")
        k = self.terms.cde.control.gaussian_cubature.point_count
        if n_loc * k > 2 ** 32:
            raise ValueError("step_owner_ids: a shard's children need 32-bit local ids (n_loc * k <= 2^32)")
        idx = kernel.indices_slice(t0, n_tot * k, n_tot, rank * n_loc, n_loc, device)
        return _lib.shard_bucket_ids(idx, rank * n_loc, n_loc, k, G, packed=packed)

    def owner_ids_materialise(self, t0, t1, y_shard: torch.Tensor, rank: int, recv_pairs: torch.Tensor, counts):
        """Phase 2 (local, after the id exchange): `recv_pairs` [T, 2] int32 = (output slot, local child id) of every
        output row whose selected parent lives on this rank, in sender order; `counts[src][owner]` = the replicated
        matrix of pair counts.  Materialises the first n_loc of them into this rank's next shard and the surplus
        into a send buffer for the ranks with a deficit (`plan_rebalance`, O(sqrt(n_loc)) rows); the gather reads the
        id column of the pairs in place.
        Returns (y1 [n_loc, d] with rows [:keep] filled, slots [keep], send_rows, send_slots, send_splits,
        recv_splits, keep)."""
        import numpy as np

        from . import _lib

        solver = self.solver
        n_loc, d = y_shard.shape
        drift, cub, dt, k, _ = solver._describe(self.terms, y_shard, *solver.substep_times(t0, t1)[0], None)
        counts = np.asarray(counts, dtype=np.int64)
        per_owner = counts.sum(axis=0)                       # rows every rank materialises
        transfers = plan_rebalance(per_owner, n_loc)
        total = int(per_owner[rank])
        assert recv_pairs.shape[0] == total
        keep = min(total, n_loc)
        y1 = torch.empty((n_loc, d), dtype=torch.float32, device=y_shard.device)
        if keep:
            _lib.expand_select_pairs(y_shard, d, drift, dt, cub, recv_pairs, n_out=keep, out=y1[:keep])
        n_give = total - keep
        send_rows = (_lib.expand_select_pairs(y_shard, d, drift, dt, cub, recv_pairs, n_out=n_give, first=keep)
                     if n_give else y1.new_empty((0, d)))
        return (y1, recv_pairs[:keep, 0], send_rows, recv_pairs[keep:, 0], [int(c) for c in transfers[rank]],
                [int(c) for c in transfers[:, rank]], keep)

    def _ids_launch(self, t, n_loc: int, rank: int, device, group):
        """First half of the id exchange of the step whose outer time is `t` (current stream, nothing waits on the
        host): draw + bucket + all-gather of the per-owner counts."""
        G = self.n_ranks
        pairs, counts_dev = self.owner_ids_bucket(t, n_loc, rank, device, packed=True)
        if G == 1:
            return pairs, counts_dev, None
        allc = [torch.empty_like(counts_dev) for _ in range(G)]
        dist.all_gather(allc, counts_dev, group=group)
        return pairs, counts_dev, allc

    def _ids_finish(self, launched, rank: int, group):
        """Second half: the step's one host synchronisation (the replicated counts matrix sizes the variable
        all-to-all) and one packed all-to-all.  Returns (pairs received [T, 2], counts[src][owner])."""
        pairs, counts_dev, allc = launched
        if allc is None:
            return pairs, counts_dev.cpu().numpy().reshape(1, 1)
        counts = torch.stack(allc).cpu().numpy()
        return exchange_rows(pairs, [int(c) for c in counts[rank]], [int(c) for c in counts[:, rank]], group), counts

    # ---- the same exchange through the copy engines: fixed-capacity regions pushed into peer memory
    def _dma_launch(self, t, n_loc: int, rank: int, device, pib: PeerIdBuffers, b: int):
        """Draw + bucket into fixed-capacity regions + one device-to-device copy per peer (region `owner` of the
        staging buffer -> region `rank` of the owner's receive buffer `b`, and this rank's counts row into every
        peer's counts matrix).  Copy engines over NVLink: nothing here needs an SM once the two kernels are done,
        and nothing waits on the host."""
        from . import _lib

        kernel = self.solver.recombination_kernel
        G = self.n_ranks
        n_tot = n_loc * G
        k = self.terms.cde.control.gaussian_cubature.point_count
        import os

        with _lib.side_residency(int(os.environ.get("MCCB200_IDS_SIDE_CTAS", "0"))):   # (co-residency knobs: measured no gain, see profiles/README.md)
            idx = kernel.indices_slice(t, n_tot * k, n_tot, rank * n_loc, n_loc, device)
            _lib.shard_bucket_regions(idx, rank * n_loc, n_loc, k, G, pib.cap, pib.stage, pib.stage_counts)
        for o in range(G):
            _lib.peer_copy(pib.base[b][o] + 8 * pib.cap * rank, pib.stage[o * pib.cap:], 8 * pib.cap)
            _lib.peer_copy(pib.base[b][o] + 4 * pib.words + 4 * G * rank, pib.stage_counts, 4 * G)

    def _dma_seal(self, pib: PeerIdBuffers, b: int, after: "torch.cuda.Event"):
        """Cross-rank barrier (every rank's pushes into buffer `b` have landed; every rank is done reading the
        buffer the NEXT pushes will overwrite: `after` = this rank's gather) + the counts matrix to pinned host
        memory.  Returns the event the next step waits for."""
        side = torch.cuda.current_stream()
        side.wait_event(after)
        dist.all_reduce(pib.flag, group=self._ids_group)
        pib.host_counts[b].copy_(pib.counts(b), non_blocking=True)
        event = torch.cuda.Event()
        event.record(side)
        return event

    def _materialise_regions(self, t0, t1, y_shard, rank: int, pib: PeerIdBuffers, b: int, counts):
        """`owner_ids_materialise` reading the G receive regions of buffer `b` in place (sender order)."""
        import numpy as np

        from . import _lib

        solver = self.solver
        n_loc, d = y_shard.shape
        G = self.n_ranks
        drift, cub, dt, k, _ = solver._describe(self.terms, y_shard, *solver.substep_times(t0, t1)[0], None)
        counts = np.asarray(counts, dtype=np.int64)
        per_owner = counts.sum(axis=0)
        transfers = plan_rebalance(per_owner, n_loc)
        total = int(per_owner[rank])
        keep = min(total, n_loc)
        n_give = total - keep
        pairs = pib.pairs(b)
        y1 = torch.empty((n_loc, d), dtype=torch.float32, device=y_shard.device)
        slot_map = torch.empty(n_loc, dtype=torch.int32, device=y_shard.device)
        send_rows = y1.new_empty((n_give, d))
        send_slots = torch.empty(n_give, dtype=torch.int32, device=y_shard.device)
        pos = 0
        for src in range(G):
            c = int(counts[src, rank])
            first = src * pib.cap
            n1 = max(0, min(c, keep - pos))
            if n1:
                _lib.expand_select_pairs(y_shard, d, drift, dt, cub, pairs, n_out=n1, out=y1[pos:pos + n1], first=first)
                _lib.unzip_slots_(pairs, first, n1, slot_map[pos:pos + n1])
            if c - n1:
                sp = pos + n1 - keep
                _lib.expand_select_pairs(y_shard, d, drift, dt, cub, pairs, n_out=c - n1, out=send_rows[sp:sp + c - n1],
                                         first=first + n1)
                _lib.unzip_slots_(pairs, first + n1, c - n1, send_slots[sp:sp + c - n1])
            pos += c
        return (y1, slot_map, send_rows, send_slots, [int(c) for c in transfers[rank]],
                [int(c) for c in transfers[:, rank]], keep)

    def step_owner_ids(self, t0, t1, y_shard: torch.Tensor, rank: int, group=None, prefetch: bool = True, dma=None):
        """One step with a GLOBAL `MonteCarloKernel(with_replacement=True)` resample in which no parent row and
        (up to the binomial imbalance) no child row crosses the link: ranks exchange 8 bytes per output row --
        (slot, child id) -- and every rank materialises the selected children of its OWN parents with the local
        fused gather.  Works with 64-bit global child ids (`kernel.x64`, BASELINE config 5: 2^34 children) in
        O(n_loc) memory per rank.  The index vector depends on (key, t) only, so with `prefetch` the draw, the
        bucketing and the id exchange of the NEXT step (outer time t1) are issued on a high-priority side stream
        before this step's HBM-bound gather and run under it: with `dma` (default when prefetching) the pairs go
        into fixed-capacity regions that the copy engines push into the owners' memory (no collective, no host
        synchronisation until the next step starts) and the gather leaves a quarter of every SM to the draw and
        bucketing kernels; otherwise through an all-gather of counts and one variable NCCL all-to-all.  Returns
        (y1_shard, slot_map): y1_shard[q] is row slot_map[q] of the single-device step on the concatenated state
        (same multiset, known permutation)."""
        import os

        from . import _lib

        G = self.n_ranks
        y_shard = y_shard.contiguous()
        n_loc, d = y_shard.shape
        dev = y_shard.device
        main = torch.cuda.current_stream(dev)
        if dma is None:
            dma = os.environ.get("MCCB200_IDS_DMA", "1") != "0"
        use_dma = bool(prefetch and dma and G > 1)
        ready = getattr(self, "_ids_ready", None)
        self._ids_ready = None
        regions = None
        if ready is not None and ready[0] == (float(t0), n_loc, rank):
            if ready[1] == "dma":
                _, _, event, b = ready
                event.synchronize()                           # the step's one host wait: the counts matrix
                counts = self._pib.host_counts[b].numpy().copy()
                if int(counts.max()) <= self._pib.cap:
                    regions = b
                else:                                         # a group outgrew its region (8 sigma): redo through NCCL
                    recv_pairs, counts = self._ids_finish(self._ids_launch(t0, n_loc, rank, dev, group), rank, group)
            else:
                _, _, event, recv_pairs, counts = ready
                main.wait_event(event)
                recv_pairs.record_stream(main)
        else:
            recv_pairs, counts = self._ids_finish(self._ids_launch(t0, n_loc, rank, dev, group), rank, group)
        trace = getattr(self, "_trace", None)                 # experiments: [(label, event)] with timing events

        def mark(label, stream):
            if trace is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                trace.append((label, ev))

        mark("step_begin(main)", main)
        launched = side = None
        if prefetch and G > 1:
            if getattr(self, "_ids_group", None) is None:     # its own communicator: independent of the data path's
                self._ids_group = dist.new_group(list(range(G))) if group is None else group
                self._ids_stream = torch.cuda.Stream(device=dev, priority=-1)   # its CTAs go ahead of the gather's
            side = self._ids_stream
            if use_dma and (getattr(self, "_pib", None) is None or self._pib_n_loc != n_loc):
                if getattr(self, "_pib", None) is not None:
                    self._pib.close()
                self._pib, self._pib_n_loc = PeerIdBuffers(n_loc, rank, G, dev, self._ids_group), n_loc
            with torch.cuda.stream(side):
                mark("prefetch_begin(side)", side)
                if use_dma:
                    nb = self._pib.turn
                    self._pib.turn ^= 1
                    self._dma_launch(t1, n_loc, rank, dev, self._pib, nb)
                else:
                    launched = self._ids_launch(t1, n_loc, rank, dev, self._ids_group)
                mark("prefetch_launched(side)", side)
        if regions is not None:
            with _lib.expand_residency(int(os.environ.get("MCCB200_IDS_GATHER_CTAS", "0"))):   # (0 = default residency)
                y1, slot_map, send_rows, send_slots, send_splits, recv_splits, keep = self._materialise_regions(
                    t0, t1, y_shard, rank, self._pib, regions, counts)
        else:
            y1, slots, send_rows, send_slots, send_splits, recv_splits, keep = self.owner_ids_materialise(
                t0, t1, y_shard, rank, recv_pairs, counts)
            slot_map = torch.empty(n_loc, dtype=torch.int32, device=dev)
            slot_map[:keep] = slots
        gather_done = torch.cuda.Event()
        gather_done.record(main)
        mark("gather_done(main)", main)
        if G > 1 and (sum(send_splits) or sum(recv_splits)):
            # the surplus rows travel with their slot as one more column (its int32 bit pattern): one all-to-all
            packed = torch.empty((send_rows.shape[0], d + 1), dtype=torch.float32, device=dev)
            packed[:, :d] = send_rows
            packed[:, d] = send_slots.contiguous().view(torch.float32)
            got = exchange_rows(packed, send_splits, recv_splits, group)
            if got.shape[0]:
                y1[keep:keep + got.shape[0]] = got[:, :d]
                slot_map[keep:keep + got.shape[0]] = got[:, d].contiguous().view(torch.int32)
        if prefetch and G > 1:                                # the gather is enqueued: now the side stream may wait
            with torch.cuda.stream(side):
                if use_dma:
                    self._ids_ready = ((float(t1), n_loc, rank), "dma", self._dma_seal(self._pib, nb, gather_done), nb)
                else:
                    nxt = self._ids_finish(launched, rank, self._ids_group)
                    event = torch.cuda.Event()
                    event.record(side)
                    self._ids_ready = ((float(t1), n_loc, rank), "nccl", event) + nxt
                mark("prefetch_sealed(side)", side)
        mark("step_end(main)", main)
        return y1, slot_map

    def close(self):
        """Frees the peer-mapped id buffers (collective: every rank calls it)."""
        if getattr(self, "_pib", None) is not None:
            self._pib.close()
            self._pib = None
        self._ids_ready = None


class ShardedBoxStep:
    """One MCCSolver step with `PartitioningRecombinationKernel(BoxPartitionKernel, MonteCarloKernel)` over
    row-sharded particles and ONE GLOBAL partition: what the reference computes on the concatenated state
    (mccube/_kernels/base.py:152-177), shards bit-identical to the single-device step.

    The global stable sort by box key never happens.  Per rank: the min / max of the implicit children (one
    all-reduce of 2 * dims floats), the per-key child counts of its own parents (`box_prk_count`; one all-gather of
    n_keys integers tells every rank where its children of each key start in the global order), the ranking pass
    with those global key starts (`box_prk_rank_global`: the shard's hits land in a sparse [n_tot] slot array),
    an on-device compaction of the hit slots, and the fused expand + scatter kernel that writes every selected
    child straight into the shard of the rank storing its output row (NVLink peer stores).  No sort, no row
    buffer, no host synchronisation; the data path has no collective.  `BoxPartitionKernel.partition_count` and
    `MonteCarloKernel.recombination_count` are the GLOBAL partition count and the per-partition output count.

    The phases are separate methods so that tests can emulate the ranks on one device."""

    def __init__(self, solver, terms, n_ranks: int):
        from ._kernels import BoxPartitionKernel, MonteCarloKernel, PartitioningRecombinationKernel

        kernel = solver.recombination_kernel
        if (type(kernel) is not PartitioningRecombinationKernel or type(kernel.partitioning_kernel) is not BoxPartitionKernel
                or type(kernel.recombination_kernel) is not MonteCarloKernel or solver.weighted or solver.n_substeps != 1):
            raise NotImplementedError("sharded Box step: unweighted Box + PRK + MonteCarloKernel with n_substeps=1 only")
        self.solver, self.terms, self.n_ranks = solver, terms, n_ranks

    # ---- shapes shared by all phases
    def _plan(self, t0, t1, y_shard):
        from . import _lib

        from ._drift import GaussianDiagDrift

        # the phases of one step share one description; only when the drift is the analytic one (an array drift is
        # evaluated on the shard's current contents, which a raw-pointer kernel may have rewritten in place)
        if not isinstance(self.terms.ode.vector_field, GaussianDiagDrift):
            return self._plan_uncached(t0, t1, y_shard)
        tag = (float(t0), float(t1), y_shard.data_ptr(), tuple(y_shard.shape))
        cached = getattr(self, "_plan_cache", None)
        if cached is not None and cached[0] == tag:
            return cached[1]
        pl = self._plan_uncached(t0, t1, y_shard)
        self._plan_cache = (tag, pl)
        return pl

    def _plan_uncached(self, t0, t1, y_shard):
        from . import _lib

        solver = self.solver
        part = solver.recombination_kernel.partitioning_kernel
        mc = solver.recombination_kernel.recombination_kernel
        n_loc, d = y_shard.shape
        drift, cub, dt, k, _ = solver._describe(self.terms, y_shard, *solver.substep_times(t0, t1)[0], None)
        n_tot = n_loc * self.n_ranks
        m = part.partition_count
        if (n_tot * k) % m:
            raise ValueError(f"partition_count {m} does not divide the {n_tot * k} expanded particles")
        in_per_part, out_per_part = n_tot * k // m, mc.recombination_count
        if m * out_per_part != n_tot:
            raise ValueError(f"partition_count * recombination_count = {m * out_per_part} must equal the global particle count {n_tot}")
        if n_tot * k > 2 ** 32:
            raise ValueError("sharded Box step: more than 2^32 children in total (32-bit sorted positions)")
        dims = part.resolve_dims(d)
        if not _lib.box_prk_supported(n_loc, d, cub, dims, part.bits, in_per_part):
            raise NotImplementedError("sharded Box step: shape not supported by the fused Box kernel")
        return dict(part=part, mc=mc, drift=drift, cub=cub, dt=dt, k=k, d=d, n_loc=n_loc, n_tot=n_tot,
                    in_per_part=in_per_part, out_per_part=out_per_part, dims=dims, bits=part.bits)

    def local_minmax(self, t0, t1, y_shard):
        """Phase 1: [min..., max...] of the key dims over this shard's implicit children (None with fixed bounds)."""
        from . import _lib

        pl = self._plan(t0, t1, y_shard)
        if pl["part"].fixed_bounds(pl["dims"], y_shard.device) is not None:
            return None
        return _lib.box_child_minmax(y_shard, pl["d"], pl["drift"], pl["dt"], pl["cub"], pl["dims"], pl["bits"])

    def bounds_from(self, t0, t1, y_shard, minmax):
        """Bounds of the global partition from the reduced min / max: (lo, 2^bits / (hi - lo)) in float32, the
        arithmetic of the single-device bounds kernel."""
        pl = self._plan(t0, t1, y_shard)
        fixed = pl["part"].fixed_bounds(pl["dims"], y_shard.device)
        if fixed is not None:
            return fixed
        nd = len(pl["dims"])
        lo, hi = minmax[:nd], minmax[nd:]
        rng = hi - lo
        cells = torch.full_like(rng, float(1 << pl["bits"]))
        inv_h = torch.where(rng > 0, cells / rng, torch.zeros_like(rng))
        return torch.cat([lo, inv_h]).contiguous()

    def count(self, t0, t1, y_shard, bounds):
        """Phase 2: counting pass over this shard; returns (workspace for phase 3, per-key child counts [n_keys])."""
        from . import _lib

        pl = self._plan(t0, t1, y_shard)
        tables = _lib.box_prk_tables(pl["k"], pl["dims"], pl["bits"], y_shard.device)
        ws = _lib.box_prk_count(y_shard, pl["d"], pl["drift"], pl["dt"], pl["cub"], pl["dims"], pl["bits"], bounds,
                                pl["in_per_part"], tables)
        starts = _lib.box_prk_local_key_starts(ws, pl["n_loc"], pl["k"], len(pl["dims"]), pl["bits"], pl["in_per_part"],
                                               pl["out_per_part"]).to(torch.int64) & 0xFFFFFFFF
        ends = torch.cat([starts[1:], torch.full((1,), pl["n_loc"] * pl["k"], dtype=torch.int64, device=starts.device)])
        return ws, (ends - starts)

    # ---- the two collectives of a step (tiny: 2 * dims floats, n_keys integers); no data path collective
    @staticmethod
    def reduce_minmax(minmax: torch.Tensor, group=None) -> torch.Tensor:
        """Global [min..., max...] from every rank's local one with ONE all-reduce: min over (lo, -hi)."""
        nd = minmax.numel() // 2
        both = torch.cat([minmax[:nd], -minmax[nd:]])
        dist.all_reduce(both, op=dist.ReduceOp.MIN, group=group)
        return torch.cat([both[:nd], -both[nd:]])

    @staticmethod
    def gather_totals(totals: torch.Tensor, n_ranks: int, group=None) -> torch.Tensor:
        """[n_ranks, n_keys] matrix of every rank's per-key child counts."""
        gathered = [torch.empty_like(totals) for _ in range(n_ranks)]
        dist.all_gather(gathered, totals.contiguous(), group=group)
        return torch.stack(gathered)

    @staticmethod
    def key_starts_for(rank: int, totals_all: torch.Tensor) -> torch.Tensor:
        """Global sorted position of rank's first child of every key from the [n_ranks, n_keys] count matrix:
        children of smaller keys on all ranks + children of the same key on lower ranks (uint32 bit patterns)."""
        per_key = totals_all.sum(0)
        key_base = torch.cumsum(per_key, 0) - per_key
        below = totals_all[:rank].sum(0)
        v = (key_base + below) & 0xFFFFFFFF
        return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous()

    def rank_and_scatter(self, t0, t1, y_shard, rank: int, ws, key_start, peer_ptrs):
        """Phase 3: global ranks -> this shard's hits -> fused expand + scatter into the owners' shards."""
        from . import _lib

        pl = self._plan(t0, t1, y_shard)
        dev = y_shard.device
        tables = _lib.box_prk_tables(pl["k"], pl["dims"], pl["bits"], dev)
        sel = self._selection(t0, pl, dev)
        idx_slots = torch.full((pl["n_tot"],), -1, dtype=torch.int32, device=dev)
        _lib.box_prk_rank_global(y_shard, pl["d"], pl["drift"], pl["dt"], pl["cub"], pl["dims"], pl["bits"],
                                 pl["out_per_part"], pl["in_per_part"], tables, sel, ws, key_start, idx_slots)
        G, n_loc = self.n_ranks, pl["n_loc"]
        counts = torch.zeros(G * G, dtype=torch.int32, device=dev)
        # hard bound on this shard's hits (every output slot, or every local child): the compaction kernel traps on
        # overflow, and the per-rank hit counts are correlated (all partitions share one local index vector)
        slot_list = torch.empty(min(pl["n_tot"], n_loc * pl["k"]), dtype=torch.int32, device=dev)
        _lib.compact_slots(idx_slots, slot_list, counts[rank * G:rank * G + 1])
        _lib.expand_scatter_slots(y_shard, pl["d"], pl["drift"], pl["dt"], pl["cub"], idx_slots, slot_list, counts,
                                  rank, G, n_loc, peer_ptrs)

    def _selection(self, t0, pl, dev):
        """The shared local index vector and its selection tables (a function of (key, t0) only); `step` draws
        them on a side stream ahead of time."""
        from . import _lib

        ready = getattr(self, "_sel_ready", None)
        if ready is not None and ready[0] == (float(t0), pl["in_per_part"], pl["out_per_part"]):
            self._sel_ready = None
            torch.cuda.current_stream(dev).wait_stream(ready[3])
            ready[1].record_stream(torch.cuda.current_stream(dev))
            ready[2].record_stream(torch.cuda.current_stream(dev))
            return ready[2]
        idx_local = pl["mc"].indices(t0, pl["in_per_part"], pl["out_per_part"], dev)
        return _lib.box_prk_selection(idx_local, pl["in_per_part"])

    def step(self, t0, t1, rank: int, shards: PeerShards, group=None) -> torch.Tensor:
        """The state lives in `shards` (load it once with `shards.load`); returns this rank's next shard."""
        from . import _lib
        from ._solvers import _side_stream

        cur, nxt = shards.turn, shards.turn ^ 1
        y_in, y_out = shards.local[cur], shards.local[nxt]
        pl = self._plan(t0, t1, y_in)
        side = _side_stream(y_in.device)
        with torch.cuda.stream(side):                            # overlaps the bounds / counting passes
            idx_local = pl["mc"].indices(t0, pl["in_per_part"], pl["out_per_part"], y_in.device)
            sel = _lib.box_prk_selection(idx_local, pl["in_per_part"])
        self._sel_ready = ((float(t0), pl["in_per_part"], pl["out_per_part"]), idx_local, sel, side)
        mm = self.local_minmax(t0, t1, y_in)
        if mm is not None and self.n_ranks > 1:
            mm = self.reduce_minmax(mm, group)
        bounds = self.bounds_from(t0, t1, y_in, mm)
        ws, totals = self.count(t0, t1, y_in, bounds)
        totals_all = self.gather_totals(totals, self.n_ranks, group) if self.n_ranks > 1 else totals[None]
        key_start = self.key_starts_for(rank, totals_all)
        self.rank_and_scatter(t0, t1, y_in, rank, ws, key_start, shards.ptrs[nxt])
        shards.barrier()          # every rank's rows have landed (and nobody still reads the other buffer)
        shards.turn = nxt
        return y_out
