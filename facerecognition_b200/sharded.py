"""Identity-sharded galleries across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  Rank r owns gallery rows
[shard_bounds(N, R, r)); queries are replicated; every rank runs the local fused search with
idx_base = its first row, so candidates carry GLOBAL row ids; ONE all-gather of the [Q, k] (score, id)
lists follows and every rank merges the R lists with frb_topk_merge (ties -> lowest global id, so the
answer is identical for any R).  Nothing else crosses NVLink: shards are loaded once and never move.

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


class ShardedSearch:
    """local top-k -> all_gather -> merge.  Works for cosine (largest=True) and chi-square (largest=False)."""

    def __init__(self, local_search: Callable[[torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]],
                 merge: Callable[[torch.Tensor, torch.Tensor, bool], Tuple[torch.Tensor, torch.Tensor]],
                 largest: bool, group: Optional[dist.ProcessGroup] = None):
        self.local_search, self.merge, self.largest, self.group = local_search, merge, largest, group

    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def search(self, queries: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        scores, idx = self.local_search(queries, k)          # [Q, k] with global ids
        world = self.world()
        if world == 1:
            return scores, idx
        all_s = torch.empty((world,) + tuple(scores.shape), dtype=scores.dtype, device=scores.device)
        all_i = torch.empty((world,) + tuple(idx.shape), dtype=idx.dtype, device=idx.device)
        # one all-gather per payload; the output slices alias all_s / all_i, so the [R, Q, k] layout
        # frb_topk_merge expects is produced in place (works for NCCL and for gloo in the CPU tests)
        dist.all_gather(list(all_s.unbind(0)), scores.contiguous(), group=self.group)
        dist.all_gather(list(all_i.unbind(0)), idx.contiguous(), group=self.group)
        return self.merge(all_s, all_i, self.largest)          # [R, Q, k] -> [Q, k]


def cosine_sharded(gallery_shard: torch.Tensor, row_offset: int, *, qnorm_mode: int = 0,
                   group: Optional[dist.ProcessGroup] = None) -> ShardedSearch:
    """Product wiring for K1: `gallery_shard` = this rank's rows (fp32 or bf16, CUDA), `row_offset` = global id of row 0."""
    from . import _native as N
    from . import ops

    def local(q: torch.Tensor, k: int):
        return ops.cosine_topk(q, gallery_shard, k, score_mode=N.FRB_SCORE_IP, qnorm_mode=qnorm_mode, idx_base=row_offset)

    return ShardedSearch(local, ops.topk_merge, True, group)


def chisq_sharded(hist_shard: torch.Tensor, cell_px: int, row_offset: int, *, q_cell_px: Optional[int] = None,
                  group: Optional[dist.ProcessGroup] = None) -> ShardedSearch:
    """Product wiring for K3: u16 histogram shard; queries are u16 histograms [Q, L]."""
    from . import ops

    def local(q_hist: torch.Tensor, k: int):
        return ops.chisq_topk(q_hist, q_cell_px or cell_px, hist_shard, cell_px, k, idx_base=row_offset)

    return ShardedSearch(local, ops.topk_merge, False, group)
