"""Identity-sharded galleries across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  Rank r owns gallery rows
[shard_bounds(N, R, r)); queries are replicated; every rank runs the local fused search with
idx_base = its first row, so candidates carry GLOBAL row ids.  The exchange is ONE kernel over NVLink peer memory
(frb_exchange_topk_merge: every CTA stores its queries' candidates into all ranks' buffers, flags them, waits for
the same CTA of every rank and merges); where CUDA IPC is not available — and under gloo in the CPU tests — it is
ONE all-gather of the packed [Q, k] (id, score) records followed by frb_topk_merge_strided.  Ties -> lowest
global id, so the answer is identical for any R and for either transport.  Nothing else crosses NVLink: shards are loaded once and never move.

`local_search` and `merge` are injectable so the host-side plumbing (bounds, id offsets, gather layout)
can be exercised with the gloo backend on a CPU-only box; the product wiring below binds them to the
CUDA kernels and there is no CPU implementation in this package.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) of rank `rank`: ceil(N/R) rows per rank, the last ranks may be short or empty."""
    per = (n_rows + world_size - 1) // world_size
    lo = min(rank * per, n_rows)
    return lo, min(lo + per, n_rows)


def balanced_bounds(n_rows: int, weights, rank: int, align: int = 256) -> Tuple[int, int]:
    """Rows [lo, hi) of `rank` when rank r's share of the gallery is proportional to weights[r] — for boxes whose GPUs
    do not run at the same speed (under the power cap the GPUs of one chassis differ by 10-30 %, and a sharded step
    waits for its slowest rank: profiles/r2_rank_skew.txt).  Inner boundaries are rounded to `align` rows (a gallery
    tile); every row belongs to exactly one rank; answers do not depend on the split (global row ids, ties -> lowest).
    Used by `bench.py --balance` (opt-in; exercised at 2 and 4 GPUs); the engine classes shard equally (shard_bounds)."""
    w = [max(float(x), 0.0) for x in weights]
    total = sum(w)
    if total <= 0.0:
        return shard_bounds(n_rows, len(w), rank)
    edges, cum = [0], 0.0
    for r in range(len(w) - 1):
        cum += w[r]
        e = int(round(n_rows * cum / total / align)) * align
        edges.append(min(max(e, edges[-1]), n_rows))
    edges.append(n_rows)
    return edges[rank], edges[rank + 1]


def measure_rank_weights(step: Callable[[], None], sync: Callable[[], None], seconds: float = 1.0,
                         group: Optional[dist.ProcessGroup] = None, clamp: float = 0.3):
    """Collective calibration for balanced_bounds: every rank runs `step` (the same probe workload on each GPU) back to
    back for `seconds` at the same time — so each GPU is measured in the state a balanced job keeps it in, busy all the
    time — and the ranks' steps-per-second are all-gathered.  Returns one weight per rank, mean 1, limited to
    1 +- clamp so that a noisy probe cannot starve a rank."""
    import time
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return [1.0]
    for _ in range(3):
        step()
    sync()
    dist.barrier(group=group)
    t0, n = time.perf_counter(), 0
    while True:
        for _ in range(8):
            step()
        n += 8
        sync()
        if time.perf_counter() - t0 >= seconds:
            break
    rate = n / (time.perf_counter() - t0)
    rates = [None] * world
    dist.all_gather_object(rates, float(rate), group=group)
    mean = sum(rates) / world
    return [min(max(r / mean, 1.0 - clamp), 1.0 + clamp) for r in rates]


def _record_layout(n_query: int, k: int) -> Tuple[int, int, int]:
    """Byte layout of one rank's candidate record: ids i64 [Q, k] first, then scores f32 [Q, k], padded to 16."""
    idx_bytes = n_query * k * 8
    rec = idx_bytes + n_query * k * 4
    return idx_bytes, rec, (rec + 15) // 16 * 16


class ShardedSearch:
    """local top-k -> ONE all_gather -> merge.  Works for cosine (largest=True) and chi-square (largest=False).

    Each rank packs its [Q, k] ids and scores into one byte record, so a single collective moves both; the merge
    reads the gathered [R, record] buffer in place (frb_topk_merge_strided).  `local_search` / `merge` are the
    injectable tensor-level forms used by the CPU (gloo) tests; `local_into` / `merge_packed` are the zero-copy
    product forms."""

    def __init__(self, local_search: Optional[Callable[[torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]]],
                 merge: Optional[Callable[[torch.Tensor, torch.Tensor, bool], Tuple[torch.Tensor, torch.Tensor]]],
                 largest: bool, group: Optional[dist.ProcessGroup] = None,
                 local_into: Optional[Callable[[torch.Tensor, int, torch.Tensor, torch.Tensor], None]] = None,
                 merge_packed: Optional[Callable[[torch.Tensor, int, int, bool], Tuple[torch.Tensor, torch.Tensor]]] = None,
                 peer_exchange: bool = False):
        self.local_search, self.merge, self.largest, self.group = local_search, merge, largest, group
        self.local_into, self.merge_packed = local_into, merge_packed
        # peer_exchange: use the fused NVLink exchange + merge kernel (frb_exchange_topk_merge) instead of the
        # all-gather; set up lazily, and only if every rank managed to map every peer's buffer (CUDA IPC)
        self.peer_exchange, self._exchange, self._exchange_failed = peer_exchange, None, False
        # product wiring may add: local search over queries that are already normalised bf16 (cosine), and the
        # normaliser that produces them (see cosine_sharded)
        self.local_search_bf16: Optional[Callable[[torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]]] = None
        self.normalize_bf16: Optional[Callable[[torch.Tensor, torch.Tensor], None]] = None
        self._graphs = {}      # (data_ptr, shape, dtype, k) -> (CUDAGraph, scores, idx): the replayed step

    def _get_exchange(self, n_query: int, k: int, device: torch.device):
        from . import _native as N
        from . import ops
        ex = self._exchange
        if ex is not None and n_query <= ex.max_query and k <= ex.max_k:
            return ex
        if self._exchange_failed:
            return None
        world, rank = self.world(), dist.get_rank(self.group)
        ok, new = 1, None
        try:
            new = ops.Exchange(world, rank, max(n_query, ex.max_query if ex else 0), max(k, ex.max_k if ex else 0), device)
            mine = torch.tensor(list(new.handle), dtype=torch.uint8, device=device)
        except N.FrbError:
            ok, mine = 0, torch.zeros(N.FRB_IPC_HANDLE_BYTES, dtype=torch.uint8, device=device)
        handles = torch.empty(world * N.FRB_IPC_HANDLE_BYTES, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(handles, mine, group=self.group)
        if ok:
            try:
                new.open(bytes(handles.cpu().tolist()))
            except N.FrbError:
                ok = 0
        agree = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)    # all ranks take the same path
        if int(agree.item()) == 0:
            self._exchange_failed = True
            if new is not None:
                new.close()
            if rank == 0:
                print("facerecognition_b200: peer-memory exchange unavailable (CUDA IPC), using the all-gather path")
            return None
        if ex is not None:
            torch.cuda.synchronize(device)
            dist.barrier(group=self.group)       # nobody is still reading the old buffers
            ex.close()
        self._exchange = new
        return new

    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _local(self, queries: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if queries.dtype == torch.bfloat16:
            if self.local_search_bf16 is None:
                raise ValueError("this sharded search has no bf16-query form (cosine over a bf16 shard only)")
            return self.local_search_bf16(queries, k)
        return self.local_search(queries, k)

    def search(self, queries: torch.Tensor, k: int, graph: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """Merged top-k of the replicated `queries` over all shards, on every rank.  bf16 queries (cosine) are taken as
        already normalised (see gather_normalized).  graph=True: the step (local search kernels + the exchange kernel)
        is captured into a CUDA graph on first use for this exact input buffer and replayed afterwards — `queries`
        must then be the same tensor every call (refill it in place) and the returned tensors are overwritten by the
        next call with the same buffer."""
        world = self.world()
        n_query = queries.shape[0]
        if graph and queries.is_cuda and n_query > 0:
            ex = self._get_exchange(n_query, k, queries.device) if (world > 1 and self.peer_exchange) else None
            if world == 1 or ex is not None:
                return self._replay(queries, k, ex)
        if world == 1:
            return self._local(queries, k)                     # [Q, k] with global ids
        if self.peer_exchange and queries.is_cuda:
            ex = self._get_exchange(n_query, k, queries.device)
            if ex is not None:
                scores, idx = self._local(queries, k)
                return ex.topk_merge(scores, idx, self.largest)   # ONE kernel: peer stores + flags + merge
        if queries.dtype == torch.bfloat16:
            scores, idx = self._local(queries, k)
            return self._gather_merge(scores, idx)
        return self._gather_merge(None, None, queries, k)

    def _replay(self, queries: torch.Tensor, k: int, ex) -> Tuple[torch.Tensor, torch.Tensor]:
        key = (queries.data_ptr(), tuple(queries.shape), queries.dtype, k)
        hit = self._graphs.get(key)
        if hit is None:
            def step():
                s, i = self._local(queries, k)
                return ex.topk_merge(s, i, self.largest) if ex is not None else (s, i)
            step()                                             # eager once: lazy per-device setup must not be captured
            torch.cuda.synchronize(queries.device)
            if ex is not None:
                dist.barrier(group=self.group)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = step()
            hit = (g, out[0], out[1])
            if len(self._graphs) >= 8:                         # a serving loop cycles through a few staging buffers
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = hit
        hit[0].replay()
        return hit[1], hit[2]

    def gather_normalized(self, my_rows: torch.Tensor, out: torch.Tensor, q0: int) -> torch.Tensor:
        """This rank's fp32 slice [q1 - q0, D] of the batch -> L2-normalised bf16 rows written into out[q0:q1], then ONE
        all-gather replicates every rank's slice into `out` (bf16 [Q, D]): half the NVLink bytes of gathering the fp32
        batch, and each query is normalised once in the job instead of once per rank.  Equal slices on every rank."""
        if self.normalize_bf16 is None:
            raise ValueError("this sharded search has no bf16-query form (cosine over a bf16 shard only)")
        mine = out[q0:q0 + my_rows.shape[0]]
        self.normalize_bf16(my_rows, mine)
        if self.world() > 1:
            dist.all_gather_into_tensor(out, mine, group=self.group)
        return out

    def resync(self) -> None:
        """Collective: after any rank skipped or failed a step, restart the exchange epochs (frb_exchange_reset between
        two barriers) and drop the captured graphs' outputs' validity (graphs themselves stay valid)."""
        if self._exchange is None:
            return
        dev = self._exchange.device
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)
        self._exchange.reset()
        dist.barrier(group=self.group)

    def _gather_merge(self, scores, idx, queries: Optional[torch.Tensor] = None, k: int = 0):
        world = self.world()
        if scores is not None:
            n_query, k = scores.shape
            device = scores.device
        else:
            n_query, device = queries.shape[0], queries.device
        idx_bytes, rec, rec_pad = _record_layout(n_query, k)
        rank = dist.get_rank(self.group)
        gathered = torch.empty((world, rec_pad), dtype=torch.uint8, device=device)
        mine = gathered[rank]                                   # in-place all-gather: my record sits in my slot
        my_idx = mine[:idx_bytes].view(torch.int64).view(n_query, k)
        my_scores = mine[idx_bytes:rec].view(torch.float32).view(n_query, k)
        if scores is not None:
            my_scores.copy_(scores)
            my_idx.copy_(idx)
        elif self.local_into is not None:
            self.local_into(queries, k, my_scores, my_idx)
        else:
            scores, idx = self.local_search(queries, k)
            my_scores.copy_(scores)
            my_idx.copy_(idx)
        dist.all_gather_into_tensor(gathered.view(-1), mine, group=self.group)   # the ONE collective of the path
        if self.merge_packed is not None:
            return self.merge_packed(gathered, n_query, k, self.largest)
        all_i = gathered[:, :idx_bytes].contiguous().view(torch.int64).view(world, n_query, k)
        all_s = gathered[:, idx_bytes:rec].contiguous().view(torch.float32).view(world, n_query, k)
        return self.merge(all_s, all_i, self.largest)           # [R, Q, k] -> [Q, k]


def cosine_sharded(gallery_shard: torch.Tensor, row_offset: int, *, qnorm_mode: int = 0,
                   group: Optional[dist.ProcessGroup] = None, peer_exchange: bool = True) -> ShardedSearch:
    """Product wiring for K1: `gallery_shard` = this rank's rows (fp32 or bf16, CUDA), `row_offset` = global id of row 0."""
    from . import _native as N
    from . import ops

    def local(q: torch.Tensor, k: int):
        return ops.cosine_topk(q, gallery_shard, k, score_mode=N.FRB_SCORE_IP, qnorm_mode=qnorm_mode, idx_base=row_offset)

    def local_into(q: torch.Tensor, k: int, scores: torch.Tensor, idx: torch.Tensor):
        ops.cosine_topk(q, gallery_shard, k, score_mode=N.FRB_SCORE_IP, qnorm_mode=qnorm_mode, idx_base=row_offset,
                        out=(scores, idx))

    ss = ShardedSearch(local, ops.topk_merge, True, group, local_into=local_into, merge_packed=ops.topk_merge_packed,
                       peer_exchange=peer_exchange)
    if gallery_shard.dtype == torch.bfloat16 and qnorm_mode != 0:
        ss.local_search_bf16 = lambda q16, k: ops.cosine_topk_bf16q(q16, gallery_shard, k, idx_base=row_offset)
        ss.normalize_bf16 = lambda rows, out: ops.normalize_rows(rows, qnorm_mode, out=out)
    return ss


def chisq_sharded(hist_shard: torch.Tensor, cell_px: int, row_offset: int, *, q_cell_px: Optional[int] = None,
                  group: Optional[dist.ProcessGroup] = None, peer_exchange: bool = True) -> ShardedSearch:
    """Product wiring for K3: histogram shard of u16 counts, or u8 (ops.compact_histograms) when cell_px <= 255;
    queries are u16 histograms [Q, L]."""
    from . import ops

    def local(q_hist: torch.Tensor, k: int):
        return ops.chisq_topk(q_hist, q_cell_px or cell_px, hist_shard, cell_px, k, idx_base=row_offset)

    return ShardedSearch(local, ops.topk_merge, False, group, merge_packed=ops.topk_merge_packed, peer_exchange=peer_exchange)


class HostBatchPipeline:
    """Double-buffered host -> GPU -> host ingestion around a (sharded) search, for throughput serving.

    A synchronous call pays, every batch, an H2D copy of the queries, the search, and a D2H copy of the answers one
    after the other.  Here the copies of batch i + 1 / i - 1 ride the two copy engines while batch i is searched:
    `submit()` enqueues H2D on a copy stream, the search (and, with several ranks, the all-gather that replicates
    this rank's slice of the batch) on the CALLER's stream, the D2H on a second copy stream, and returns a ticket;
    `result(ticket)` blocks until that batch's answers are in pinned host memory.  All searches stay on one stream
    in submission order, so the cross-rank exchange kernels never overlap one another.

    queries per batch: `n_query` rows of `dim` fp32 for the whole job; this rank uploads rows [q0, q1) (its slice of
    the batch; the whole batch on one GPU) and downloads the answers of the same rows.
    """

    def __init__(self, search, n_query: int, dim: int,
                 k: int, device: torch.device, *, rows: Optional[Tuple[int, int]] = None, depth: int = 2,
                 group: Optional[dist.ProcessGroup] = None, graph: bool = True):
        if device.type != "cuda":
            raise RuntimeError("HostBatchPipeline needs a CUDA device: there is no CPU implementation in this package")
        # `search`: a ShardedSearch (preferred: its step is graph-replayed, and with several ranks each rank normalises
        # its own slice and the all-gather moves bf16 rows) or any callable (stage fp32 [Q, D], k) -> (scores, idx)
        self.sharded = search if isinstance(search, ShardedSearch) else None
        self.search = search.search if self.sharded is not None else search
        self.k, self.device, self.depth, self.group = k, device, depth, group
        self.graph = graph and self.sharded is not None
        self.q0, self.q1 = rows if rows is not None else (0, n_query)
        self.whole = (self.q0, self.q1) == (0, n_query)
        n_mine = self.q1 - self.q0
        self.group = group
        self.copy_in, self.copy_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
        self.bf16_gather = (self.sharded is not None and self.sharded.normalize_bf16 is not None and not self.whole)
        if self.bf16_gather:
            self.stage = [torch.empty((n_mine, dim), dtype=torch.float32, device=device) for _ in range(depth)]
            self.stage16 = [torch.empty((n_query, dim), dtype=torch.bfloat16, device=device) for _ in range(depth)]
        else:
            self.stage = [torch.empty((n_query, dim), dtype=torch.float32, device=device) for _ in range(depth)]
        self.out_s = [torch.empty((n_mine, k), dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.out_i = [torch.empty((n_mine, k), dtype=torch.int64).pin_memory() for _ in range(depth)]
        self.h2d_done = [torch.cuda.Event() for _ in range(depth)]
        self.searched = [torch.cuda.Event() for _ in range(depth)]
        self.d2h_done = [torch.cuda.Event() for _ in range(depth)]
        self.busy = [False] * depth
        self.n = 0

    def submit(self, q_host: torch.Tensor) -> int:
        """q_host: this rank's rows of the batch, pinned fp32 [q1 - q0, dim].  Returns the ticket for result()."""
        slot = self.n % self.depth
        self.n += 1
        if self.busy[slot]:
            raise RuntimeError("HostBatchPipeline: result() of the batch submitted `depth` calls ago has not been taken")
        main = torch.cuda.current_stream(self.device)
        stage = self.stage[slot]
        with torch.cuda.stream(self.copy_in):
            self.copy_in.wait_event(self.searched[slot])        # the search that last read this staging buffer is done
            (stage if self.bf16_gather else stage[self.q0:self.q1]).copy_(q_host, non_blocking=True)
            if self.bf16_gather:
                # still on the ingest stream, so it overlaps the search of the previous batch: normalise my slice -> bf16,
                # ONE all-gather of bf16 rows (every rank issues its collectives in submission order on this stream)
                q16 = self.sharded.gather_normalized(stage, self.stage16[slot], self.q0)
            self.h2d_done[slot].record(self.copy_in)
        main.wait_event(self.h2d_done[slot])
        if self.bf16_gather:
            s, i = self.sharded.search(q16, self.k, graph=self.graph)   # tensor-core search on the gathered batch
        else:
            if not self.whole:
                dist.all_gather_into_tensor(stage, stage[self.q0:self.q1], group=self.group)
            s, i = self.sharded.search(stage, self.k, graph=self.graph) if self.sharded is not None else self.search(stage, self.k)
        self.searched[slot].record(main)
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(self.searched[slot])
            self.out_s[slot].copy_(s[self.q0:self.q1], non_blocking=True)
            self.out_i[slot].copy_(i[self.q0:self.q1], non_blocking=True)
            self.d2h_done[slot].record(self.copy_out)
        if not self.graph:                                       # graph outputs are static buffers owned by the graph
            s.record_stream(self.copy_out)
            i.record_stream(self.copy_out)
        self.busy[slot] = True
        return slot

    def result(self, ticket: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(scores fp32 [q1 - q0, k], ids int64 [q1 - q0, k]) in pinned host memory; valid until the slot is reused,
        i.e. until `depth` more submit() calls."""
        self.d2h_done[ticket].synchronize()
        self.busy[ticket] = False
        return self.out_s[ticket], self.out_i[ticket]
