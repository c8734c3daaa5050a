"""Tensor-level wrappers over the libfrb200 C ABI.

PyTorch is used here for device memory, streams and nothing else: every function takes CUDA
tensors, passes raw device pointers + the current stream to the C entry point named in its
docstring, and returns CUDA tensors.  No arithmetic is done in torch; the one place torch kernels run at all is the
index plumbing (nonzero / gather / scatter of a few rows) when `cosine_topk_exact` re-runs queries whose candidate list
could not be proven complete.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _native as N

_I64 = ctypes.c_int64


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError("facerecognition_b200 kernels take CUDA tensors (there is no CPU path)")
        if not t.is_contiguous():
            raise ValueError("tensor must be contiguous")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def _stream(dev: torch.device):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


_FRB_DTYPE = {torch.float32: N.FRB_F32, torch.bfloat16: N.FRB_BF16, torch.float16: N.FRB_F16}


def row_norms(x: torch.Tensor) -> torch.Tensor:
    """frb_row_norms_f32: fp32 [R, D] -> fp32 [R] L2 norms."""
    dev = _require_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2
    out = torch.empty(x.shape[0], dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        N.call("frb_row_norms_f32", _p(x), _I64(x.shape[0]), x.shape[1], _p(out), _stream(dev))
    return out


def normalize_rows(x: torch.Tensor, mode: int = N.FRB_QNORM_EPS, out_dtype: torch.dtype = torch.float32,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """frb_normalize_rows: fp32 [R, D] -> fp32|bf16 [R, D], x/max(|x|,1e-12) (CLAMP) or x/(|x|+1e-8) (EPS).
    `out`: write into this contiguous [R, D] tensor (its dtype wins) instead of allocating."""
    dev = _require_cuda(x, out)
    assert x.dtype == torch.float32 and x.dim() == 2
    if out is None:
        assert out_dtype in _FRB_DTYPE
        out = torch.empty(x.shape, dtype=out_dtype, device=dev)
    else:
        assert out.shape == x.shape and out.dtype in _FRB_DTYPE
        out_dtype = out.dtype
    with torch.cuda.device(dev):
        N.call("frb_normalize_rows", _p(x), _I64(x.shape[0]), x.shape[1], mode, _p(out), _FRB_DTYPE[out_dtype], _stream(dev))
    return out


def cosine_topk(queries: torch.Tensor, gallery: torch.Tensor, k: int, *, score_mode: int = N.FRB_SCORE_IP,
                qnorm_mode: int = N.FRB_QNORM_NONE, q_norms: Optional[torch.Tensor] = None,
                g_norms: Optional[torch.Tensor] = None, idx_base: int = 0,
                out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """frb_cosine_topk: fp32 queries [Q, D] x fp32|bf16|fp16 gallery [N, D] -> (scores fp32 [Q, k], idx int64 [Q, k])."""
    dev = _require_cuda(queries, gallery, q_norms, g_norms)
    assert queries.dtype == torch.float32 and queries.dim() == 2 and gallery.dim() == 2
    assert gallery.dtype in _FRB_DTYPE
    assert gallery.shape[0] == 0 or gallery.shape[1] == queries.shape[1]
    q, d, n = queries.shape[0], queries.shape[1], gallery.shape[0]
    gdt = _FRB_DTYPE[gallery.dtype]
    if out is None:
        scores = torch.empty((q, k), dtype=torch.float32, device=dev)
        idx = torch.empty((q, k), dtype=torch.int64, device=dev)
    else:
        scores, idx = out
    with torch.cuda.device(dev):
        ws_bytes = N.lib.frb_cosine_topk_workspace_bytes(q, n, d, gdt, k)
        ws = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=dev)
        N.call("frb_cosine_topk", _p(queries), _I64(q), _p(gallery), gdt, _I64(n), d, _p(q_norms), _p(g_norms),
               score_mode, qnorm_mode, k, _I64(idx_base), _p(scores), _p(idx), _p(ws), ctypes.c_size_t(ws.numel()),
               _stream(dev))
    return scores, idx


def cosine_topk_bf16q(queries_bf16: torch.Tensor, gallery_bf16: torch.Tensor, k: int, *, idx_base: int = 0,
                      out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """frb_cosine_topk_bf16q: ALREADY normalised bf16 queries [Q, D] x bf16 gallery [N, D] -> (scores, idx); the
    tensor-core kernel reads the queries in place (no prologue launch).  The sharded search calls this after
    all-gathering the ranks' normalised query slices."""
    dev = _require_cuda(queries_bf16, gallery_bf16)
    assert queries_bf16.dtype == torch.bfloat16 and gallery_bf16.dtype == torch.bfloat16
    assert queries_bf16.dim() == 2 and gallery_bf16.dim() == 2
    assert gallery_bf16.shape[0] == 0 or gallery_bf16.shape[1] == queries_bf16.shape[1]
    q, d, n = queries_bf16.shape[0], queries_bf16.shape[1], gallery_bf16.shape[0]
    if out is None:
        scores = torch.empty((q, k), dtype=torch.float32, device=dev)
        idx = torch.empty((q, k), dtype=torch.int64, device=dev)
    else:
        scores, idx = out
    with torch.cuda.device(dev):
        ws_bytes = N.lib.frb_cosine_topk_workspace_bytes(q, n, d, N.FRB_BF16, k)
        ws = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=dev)
        N.call("frb_cosine_topk_bf16q", _p(queries_bf16), _I64(q), _p(gallery_bf16), _I64(n), d, k, _I64(idx_base), _p(scores),
               _p(idx), _p(ws), ctypes.c_size_t(ws.numel()), _stream(dev))
    return scores, idx


# Exact fp32 top-k through the tensor cores: a 16-bit first pass proposes kp candidates per query, frb_cosine_rescore_topk
# re-scores them in fp32 under the reference's rule and proves every list complete (see include/frb200.h).
#   eps_abs >= |true cosine - first-pass score| for ANY pair of unit vectors rounded to the first-pass format: fp16 keeps
#   11 significant bits, so 2 * 2^-11 (Cauchy-Schwarz) + subnormal tails (512 * 2 * 2^-25) + fp32 accumulation (6e-5) <
#   1.1e-3; bf16 (8 bits): 2 * 2^-8 + ... < 8.0e-3.
#   eps_rel >= |reference score - true cosine| / |cosine|: cosine_similarity()'s raw-dot branch returns cos * |q| |g| with
#   both norms within 1e-3 of 1 (inference/recognition_engine.py:57-58), i.e. at most 2.001e-3; 0 for inner products.
REFINE_EPS_F16 = 1.1e-3
REFINE_EPS_BF16 = 8.0e-3
REFINE_EPS_REL = 2.001e-3
# measured on B200 (profiles/r2_refine.txt): from 8 queries x 16k rows up the first pass + re-score beats the fp32 kernels
REFINE_MIN_QUERIES = 8
REFINE_MIN_ROWS = 16384
REFINE_MIN_ROWS_SINGLE = 750_000      # 1-3 queries: 1M rows 0.31 vs 0.36-0.38 ms, 4M 0.78 vs 1.25-1.31 ms, 400k 0.21 vs 0.19 ms


def refine_applicable(n_query: int, n_rows: int, dim: int, k: int) -> bool:
    """Shapes where ops.cosine_topk_exact beats the fp32 kernels (profiles/r2_refine.txt): any batch of >= 8 queries over
    >= 16k rows (2-40x), batches of >= 512 over galleries as small as 4k rows (4096 x 10k: 0.21 vs 1.64 ms), and the
    5-7-query batches (the fp32 FFMA-tiled kernel's smallest) from 64k rows up (100k rows: 0.15 vs 0.21 ms, 1M: 0.31 vs
    0.91 ms); 4 queries from 256k rows (1M: 0.31 vs 0.57 ms); 1-3 queries from REFINE_MIN_ROWS_SINGLE rows, where
    streaming the half-size fp16 copy outweighs the extra launches (the first pass stays on the tcgen05 kernel: the
    row-streaming kernel with a 16-slot list per warp measured 0.49 ms at 1M rows against 0.30)."""
    if not refine_list_length(k) or dim % 64 != 0 or dim > 512:
        return False
    if n_query >= REFINE_MIN_QUERIES:
        return n_rows >= REFINE_MIN_ROWS or (n_rows >= 4096 and n_query >= 512)
    if n_query >= 4:
        return (n_query >= 5 and n_rows >= 65536) or n_rows >= 262144
    return n_rows >= REFINE_MIN_ROWS_SINGLE


def refine_list_length(k: int) -> int:
    """Candidates fetched by the first pass for a final top-k (0: k too large for a provable margin).  With fp16's
    1.1e-3 bound a 16-slot list proves the reference's top-5 for all but a few in a thousand queries (those are re-run
    exactly); longer results get the longest list the kernels keep."""
    return 16 if k <= 5 else (N.FRB_MAX_K if k <= 16 else 0)


def cosine_topk_refined(queries: torch.Tensor, gallery: torch.Tensor, gallery_unit16: torch.Tensor, k: int, *,
                        score_mode: int = N.FRB_SCORE_REF_COSINE, q_norms: Optional[torch.Tensor] = None,
                        g_norms: Optional[torch.Tensor] = None, idx_base: int = 0
                        ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """First pass (frb_cosine_topk on `gallery_unit16`, the unit-norm fp16 or bf16 copy of `gallery`) ->
    frb_cosine_rescore_topk.  For the reference's cosine rule (FRB_SCORE_REF_COSINE), or inner products of unit-norm
    queries against unit-norm rows (which rank like the cosine).  Returns (scores, idx, fail_count int32 [1],
    fail_flags int32 [Q]): flagged queries could not be proven complete and must be answered by `cosine_topk` on the
    fp32 gallery (`cosine_topk_exact` does that)."""
    dev = _require_cuda(queries, gallery, gallery_unit16, q_norms, g_norms)
    kp = refine_list_length(k)
    assert kp > 0 and gallery.dtype == torch.float32 and gallery_unit16.dtype in (torch.bfloat16, torch.float16)
    assert gallery.shape == gallery_unit16.shape
    q, d, n = queries.shape[0], queries.shape[1], gallery.shape[0]
    approx, cand = cosine_topk(queries, gallery_unit16, kp, qnorm_mode=N.FRB_QNORM_CLAMP)
    scores = torch.empty((q, k), dtype=torch.float32, device=dev)
    idx = torch.empty((q, k), dtype=torch.int64, device=dev)
    fail = torch.zeros(1, dtype=torch.int32, device=dev)
    flags = torch.empty(q, dtype=torch.int32, device=dev)
    eps = REFINE_EPS_F16 if gallery_unit16.dtype == torch.float16 else REFINE_EPS_BF16
    eps_rel = REFINE_EPS_REL if score_mode == N.FRB_SCORE_REF_COSINE else 0.0
    with torch.cuda.device(dev):
        N.call("frb_cosine_rescore_topk", _p(queries), _I64(q), _p(gallery), _I64(n), d, _p(q_norms), _p(g_norms), score_mode,
               _p(cand), _p(approx), kp, k, ctypes.c_float(eps), ctypes.c_float(eps_rel), _I64(idx_base), _p(scores), _p(idx),
               _p(fail), _p(flags), _stream(dev))
    return scores, idx, fail, flags


def cosine_topk_exact(queries: torch.Tensor, gallery: torch.Tensor, gallery_unit16: torch.Tensor, k: int, *,
                      q_norms: torch.Tensor, g_norms: torch.Tensor, idx_base: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's exact fp32 top-k (cosine_similarity() rule, ties -> lowest row) at tensor-core speed:
    cosine_topk_refined, then the fp32 kernels for exactly the queries whose candidate list could not be proven
    complete (one host read of the failure count; the index plumbing of the re-run is torch gather / scatter)."""
    s, i, fail, flags = cosine_topk_refined(queries, gallery, gallery_unit16, k, score_mode=N.FRB_SCORE_REF_COSINE,
                                            q_norms=q_norms, g_norms=g_norms, idx_base=idx_base)
    if int(fail.item()):
        rows = torch.nonzero(flags).flatten()
        s2, i2 = cosine_topk(queries.index_select(0, rows).contiguous(), gallery, k, score_mode=N.FRB_SCORE_REF_COSINE,
                             q_norms=q_norms.index_select(0, rows).contiguous(), g_norms=g_norms, idx_base=idx_base)
        s.index_copy_(0, rows, s2)
        i.index_copy_(0, rows, i2)
    return s, i


def topk_merge(cand_scores: torch.Tensor, cand_idx: torch.Tensor, largest: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """frb_topk_merge: [R, Q, k] candidate lists -> best [Q, k] (ties -> lowest idx; idx < 0 is padding)."""
    dev = _require_cuda(cand_scores, cand_idx)
    assert cand_scores.dtype == torch.float32 and cand_idx.dtype == torch.int64
    assert cand_scores.dim() == 3 and cand_scores.shape == cand_idx.shape
    r, q, k = cand_scores.shape
    scores = torch.empty((q, k), dtype=torch.float32, device=dev)
    idx = torch.empty((q, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        N.call("frb_topk_merge", _p(cand_scores), _p(cand_idx), r, _I64(q), k, 1 if largest else 0, _p(scores), _p(idx),
               _stream(dev))
    return scores, idx


def group_mean_renorm(emb: torch.Tensor, order: torch.Tensor, offsets: torch.Tensor, want_bf16: bool = False
                      ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """frb_group_mean_renorm: f32 [M, D] rows grouped by `order`/`offsets` -> (f32 [G, D], bf16 [G, D] or None)."""
    dev = _require_cuda(emb, order, offsets)
    assert emb.dtype == torch.float32 and emb.dim() == 2 and order.dtype == torch.int64 and offsets.dtype == torch.int64
    g, d = offsets.shape[0] - 1, emb.shape[1]
    out = torch.empty((g, d), dtype=torch.float32, device=dev)
    out16 = torch.empty((g, d), dtype=torch.bfloat16, device=dev) if want_bf16 else None
    with torch.cuda.device(dev):
        N.call("frb_group_mean_renorm", _p(emb), _p(order), _p(offsets), _I64(g), d, _p(out), _p(out16), _stream(dev))
    return out, out16


def packed_candidates(n_query: int, k: int, device: torch.device, n_lists: int = 1
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """One byte buffer [n_lists, record] whose record is {ids i64 [Q, k] | scores f32 [Q, k]} (ids first, so both
    views are naturally aligned), plus the (scores, idx) views of list 0 when n_lists == 1.  A single all-gather of
    such records feeds `topk_merge_packed`."""
    rec = n_query * k * 12
    rec_pad = (rec + 15) // 16 * 16
    buf = torch.empty((n_lists, rec_pad), dtype=torch.uint8, device=device)
    idx = buf[0, :n_query * k * 8].view(torch.int64).view(n_query, k)
    scores = buf[0, n_query * k * 8:rec].view(torch.float32).view(n_query, k)
    return buf, scores, idx


def topk_merge_packed(buf: torch.Tensor, n_query: int, k: int, largest: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """frb_topk_merge_strided over the [R, record] buffer produced by `packed_candidates` + one all-gather."""
    dev = _require_cuda(buf)
    assert buf.dtype == torch.uint8 and buf.dim() == 2 and buf.shape[1] % 16 == 0 and buf.shape[1] >= n_query * k * 12
    r, rec = buf.shape
    scores = torch.empty((n_query, k), dtype=torch.float32, device=dev)
    idx = torch.empty((n_query, k), dtype=torch.int64, device=dev)
    base = buf.data_ptr()
    with torch.cuda.device(dev):
        N.call("frb_topk_merge_strided", ctypes.c_void_p(base + n_query * k * 8), ctypes.c_void_p(base), _I64(rec // 4),
               _I64(rec // 8), r, _I64(n_query), k, 1 if largest else 0, _p(scores), _p(idx), _stream(dev))
    return scores, idx


def bgr_to_gray(frames: torch.Tensor) -> torch.Tensor:
    """frb_bgr2gray_u8: u8 [..., 3] interleaved BGR -> u8 [...] gray, bit-exact with cv2.cvtColor(COLOR_BGR2GRAY)."""
    dev = _require_cuda(frames)
    assert frames.dtype == torch.uint8 and frames.dim() >= 2 and frames.shape[-1] == 3
    out = torch.empty(frames.shape[:-1], dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.call("frb_bgr2gray_u8", _p(frames), _I64(out.numel()), _p(out), _stream(dev))
    return out


def resize_linear(frames: torch.Tensor, dsize: Tuple[int, int], to_gray: bool = False) -> torch.Tensor:
    """frb_resize_linear_u8: u8 [B, H, W] or [B, H, W, 3] -> u8 [B, rows, cols(, 3)], bit-exact with
    cv2.resize(img, dsize) (default INTER_LINEAR); dsize = (cols, rows) as in cv2.  to_gray=True (BGR input) also
    applies cv2.cvtColor(COLOR_BGR2GRAY) to the resized pixels and returns u8 [B, rows, cols]."""
    dev = _require_cuda(frames)
    assert frames.dtype == torch.uint8 and frames.dim() in (3, 4) and frames.is_contiguous()
    ch = 1 if frames.dim() == 3 else int(frames.shape[3])
    cols, rows = int(dsize[0]), int(dsize[1])
    shape = (frames.shape[0], rows, cols) if (ch == 1 or to_gray) else (frames.shape[0], rows, cols, ch)
    out = torch.empty(shape, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.call("frb_resize_linear_u8", _p(frames), _I64(frames.shape[0]), int(frames.shape[1]), int(frames.shape[2]), ch,
               _p(out), rows, cols, int(bool(to_gray)), _stream(dev))
    return out


class Exchange:
    """frb_exchange_*: this rank's peer-memory exchange buffer (see include/frb200.h).  `handle` is the 64-byte CUDA
    IPC handle to all-gather; `open(handles)` maps the peers; `topk_merge` is the fused exchange + merge kernel."""

    def __init__(self, world: int, rank: int, max_query: int, max_k: int, device: torch.device):
        self.world, self.rank, self.max_query, self.max_k, self.device = world, rank, max_query, max_k, device
        self._ctx = ctypes.c_void_p(0)
        buf = (ctypes.c_ubyte * N.FRB_IPC_HANDLE_BYTES)()
        with torch.cuda.device(device):
            N.call("frb_exchange_create", world, rank, _I64(max_query), max_k, ctypes.byref(self._ctx), buf)
        self.handle = bytes(buf)

    def open(self, handles: bytes) -> None:
        assert len(handles) == self.world * N.FRB_IPC_HANDLE_BYTES
        with torch.cuda.device(self.device):
            N.call("frb_exchange_open", self._ctx, ctypes.c_char_p(handles))

    def topk_merge(self, scores: torch.Tensor, idx: torch.Tensor, largest: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        dev = _require_cuda(scores, idx)
        assert scores.dtype == torch.float32 and idx.dtype == torch.int64 and scores.shape == idx.shape and scores.dim() == 2
        q, k = scores.shape
        out_s = torch.empty_like(scores)
        out_i = torch.empty_like(idx)
        with torch.cuda.device(dev):
            N.call("frb_exchange_topk_merge", self._ctx, _p(scores), _p(idx), _I64(q), k, 1 if largest else 0, _p(out_s), _p(out_i),
                   _stream(dev))
        return out_s, out_i

    def status(self) -> Tuple[int, int]:
        """(kernels that gave up waiting for a peer, steps finished); synchronises with the device."""
        t, e = ctypes.c_int(0), ctypes.c_uint(0)
        with torch.cuda.device(self.device):
            N.call("frb_exchange_status", self._ctx, ctypes.byref(t), ctypes.byref(e))
        return t.value, e.value

    def reset(self) -> None:
        """frb_exchange_reset: clear flags / epoch / timeouts.  Every rank calls it between two host barriers."""
        with torch.cuda.device(self.device):
            N.call("frb_exchange_reset", self._ctx)

    def close(self) -> None:
        if self._ctx:
            N.lib.frb_exchange_destroy(self._ctx)
            self._ctx = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def exchange_emulate(ranks: Sequence["Exchange"], scores: torch.Tensor, idx: torch.Tensor, largest: bool
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """frb_exchange_emulate: [R, Q, k] local lists -> [R, Q, k] merged lists, all ranks in one launch on one GPU."""
    dev = _require_cuda(scores, idx)
    r, q, k = scores.shape
    assert len(ranks) == r and idx.shape == scores.shape
    arr = (ctypes.c_void_p * r)(*[x._ctx for x in ranks])
    out_s, out_i = torch.empty_like(scores), torch.empty_like(idx)
    with torch.cuda.device(dev):
        N.call("frb_exchange_emulate", arr, r, _p(scores), _p(idx), _I64(q), k, 1 if largest else 0, _p(out_s), _p(out_i), _stream(dev))
    return out_s, out_i


def lbp_codes(images: torch.Tensor, radius: int = 1, neighbors: int = 8) -> torch.Tensor:
    """frb_lbp_codes_u8: u8 [B, H, W] -> u8 [B, H-2r, W-2r] LBP codes (OpenCV elbp_ semantics; radius 1 / 8 neighbours
    take the tuned kernel, other settings with up to 8 neighbours the general one)."""
    dev = _require_cuda(images)
    assert images.dtype == torch.uint8 and images.dim() == 3
    b, h, w = images.shape
    out = torch.empty((b, max(h - 2 * radius, 0), max(w - 2 * radius, 0)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.call("frb_lbp_codes_u8", _p(images), _I64(b), h, w, radius, neighbors, _p(out), _stream(dev))
    return out


def lbp_cell_px(rows: int, cols: int, grid_x: int = 8, grid_y: int = 8, radius: int = 1) -> int:
    """Pixels per LBP grid cell (OpenCV spatial_histogram: floor((cols-2r)/grid_x) x floor((rows-2r)/grid_y))."""
    return max((cols - 2 * radius) // grid_x, 0) * max((rows - 2 * radius) // grid_y, 0)


def lbp_hist(images: torch.Tensor, radius: int = 1, neighbors: int = 8, grid_x: int = 8, grid_y: int = 8,
             counts8: bool = False) -> Tuple[torch.Tensor, int]:
    """frb_lbp_hist_u8: u8 [B, H, W] -> (u16 [B, grid_x*grid_y*256] cell histograms, pixels per cell).
    counts8=True: frb_lbp_hist_u8_counts8 — the same histograms written as u8 counts (the gallery form) when a cell
    has <= 255 pixels; larger cells still come back as u16."""
    dev = _require_cuda(images)
    assert images.dtype == torch.uint8 and images.dim() == 3
    b, h, w = images.shape
    as8 = counts8 and lbp_cell_px(h, w, grid_x, grid_y, radius) <= 255
    out = torch.empty((b, grid_x * grid_y * (1 << neighbors)), dtype=torch.uint8 if as8 else torch.uint16, device=dev)
    cell_px = ctypes.c_int(0)
    with torch.cuda.device(dev):
        N.call("frb_lbp_hist_u8_counts8" if as8 else "frb_lbp_hist_u8", _p(images), _I64(b), h, w, radius, neighbors, grid_x,
               grid_y, _p(out), ctypes.byref(cell_px), _stream(dev))
    return out, cell_px.value


def compact_histograms(hist_u16: torch.Tensor, cell_px: int) -> torch.Tensor:
    """frb_counts_u16_to_u8: u16 cell histograms -> the u8 gallery form the chi-square kernels also read (exact: every
    count <= cell_px <= 255); returns the input unchanged when the counts do not fit a byte or the row length is not a
    multiple of 16.  (LBPH train()/update() do not need it: lbp_hist(counts8=True) writes u8 directly.)"""
    if cell_px > 255 or hist_u16.shape[1] % 16 != 0 or hist_u16.dtype == torch.uint8:
        return hist_u16
    dev = _require_cuda(hist_u16)
    out = torch.empty(hist_u16.shape, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.call("frb_counts_u16_to_u8", _p(hist_u16), _I64(hist_u16.numel()), _p(out), _stream(dev))
    return out


def index_remap(idx: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """frb_index_remap: idx int64 [...] in place, idx >= 0 -> table[idx] (negative entries are padding and stay)."""
    dev = _require_cuda(idx, table)
    assert idx.dtype == torch.int64 and table.dtype == torch.int64
    with torch.cuda.device(dev):
        N.call("frb_index_remap", _p(idx), _I64(idx.numel()), _p(table), _I64(table.numel()), _stream(dev))
    return idx


# Batches at least this large against galleries at least this long go through the tensor-core candidate filter
# (frb_chisq_top1_filtered_g8): same answers as the exact scan.  Measured (profiles/r2_chisq_filter.txt): 1024 x 125k
# 20.7 vs 482 ms, 64 x 8192 1.4 vs 2.5 ms, 16 x 100k 3.8 vs 6.2 ms, 16 x 30k 1.4 vs 2.0 ms; at 8 queries the exact scan
# still wins (a filter unit multiplies a full 128-query tile whatever the batch holds).
FILTER_MIN_QUERIES = 16
FILTER_MIN_ROWS = 8192
FILTER_ENABLED = True


def chisq_filter_eligible(n_query: int, n_gallery: int, hist_len: int, q_cell_px: int, g_cell_px: int, k: int,
                          gallery_dtype: torch.dtype) -> bool:
    return (FILTER_ENABLED and k == 1 and gallery_dtype == torch.uint8 and q_cell_px == g_cell_px and g_cell_px <= 255
            and hist_len % 16 == 0 and hist_len <= 16384 and n_query >= FILTER_MIN_QUERIES and n_gallery >= FILTER_MIN_ROWS)


def chisq_top1_filtered(q_hist: torch.Tensor, gallery_u8: torch.Tensor, cell_px: int, idx_base: int = 0,
                        stats: Optional[torch.Tensor] = None, want_scores: bool = False):
    """frb_chisq_top1_filtered_g8: u16 queries [Q, L] vs u8 gallery [N, L] of the same cell size -> (dist fp32 [Q, 1],
    idx int64 [Q, 1]) identical to chisq_topk(k=1), through the fp16 tcgen05 candidate filter + exact re-score.
    stats: int32 [4] CUDA tensor that the call increments (fallback queries, survivors, raw candidates, -).
    want_scores: also return the approximate sum_j f of every pair, fp32 [Q, N] (tests / calibration)."""
    dev = _require_cuda(q_hist, gallery_u8, stats)
    assert q_hist.dtype == torch.uint16 and gallery_u8.dtype == torch.uint8 and q_hist.dim() == 2 and gallery_u8.dim() == 2
    q, hist_len, n = q_hist.shape[0], q_hist.shape[1], gallery_u8.shape[0]
    assert n == 0 or gallery_u8.shape[1] == hist_len
    assert stats is None or (stats.dtype == torch.int32 and stats.numel() >= 4)
    dist = torch.empty((q, 1), dtype=torch.float32, device=dev)
    idx = torch.empty((q, 1), dtype=torch.int64, device=dev)
    scores = torch.empty((q, n), dtype=torch.float32, device=dev) if want_scores else None
    with torch.cuda.device(dev):
        ws_bytes = N.lib.frb_chisq_filter_workspace_bytes(q, n, hist_len)
        ws = torch.empty(max(int(ws_bytes), 1024), dtype=torch.uint8, device=dev)
        N.call("frb_chisq_top1_filtered_g8", _p(q_hist), _I64(q), _p(gallery_u8), _I64(n), hist_len, cell_px, _I64(idx_base),
               _p(dist), _p(idx), _p(stats), _p(scores), _p(ws), ctypes.c_size_t(ws.numel()), _stream(dev))
    return (dist, idx, scores) if want_scores else (dist, idx)


def chisq_topk(q_hist: torch.Tensor, q_cell_px: int, gallery: torch.Tensor, g_cell_px: int, k: int = 1,
               idx_base: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """frb_chisq_topk / frb_chisq_topk_g8: u16 [Q, L] vs u16 or u8 [N, L] -> (dist fp32 [Q, k] ascending, idx int64 [Q, k]).
    Large k = 1 batches over a u8 gallery take frb_chisq_top1_filtered_g8 (same result, see chisq_filter_eligible)."""
    dev = _require_cuda(q_hist, gallery)
    assert q_hist.dtype == torch.uint16 and gallery.dtype in (torch.uint16, torch.uint8) and q_hist.dim() == 2 and gallery.dim() == 2
    q, hist_len, n = q_hist.shape[0], q_hist.shape[1], gallery.shape[0]
    assert n == 0 or gallery.shape[1] == hist_len
    if chisq_filter_eligible(q, n, hist_len, q_cell_px, g_cell_px, k, gallery.dtype):
        return chisq_top1_filtered(q_hist, gallery, g_cell_px, idx_base)
    dist = torch.empty((q, k), dtype=torch.float32, device=dev)
    idx = torch.empty((q, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        ws_bytes = N.lib.frb_chisq_topk_workspace_bytes(q, n, hist_len, k)
        ws = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=dev)
        N.call("frb_chisq_topk" if gallery.dtype == torch.uint16 else "frb_chisq_topk_g8", _p(q_hist), _I64(q), q_cell_px,
               _p(gallery), _I64(n), hist_len, g_cell_px, k,
               _I64(idx_base), _p(dist), _p(idx), _p(ws), ctypes.c_size_t(ws.numel()), _stream(dev))
    return dist, idx


def chisq_dist(q_hist: torch.Tensor, q_cell_px: int, gallery: torch.Tensor, g_cell_px: int) -> torch.Tensor:
    """frb_chisq_dist / frb_chisq_dist_g8: all distances, fp32 [Q, N]; the gallery may hold u16 or u8 counts."""
    dev = _require_cuda(q_hist, gallery)
    assert q_hist.dtype == torch.uint16 and gallery.dtype in (torch.uint16, torch.uint8)
    q, hist_len, n = q_hist.shape[0], q_hist.shape[1], gallery.shape[0]
    out = torch.empty((q, n), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        N.call("frb_chisq_dist" if gallery.dtype == torch.uint16 else "frb_chisq_dist_g8", _p(q_hist), _I64(q), q_cell_px,
               _p(gallery), _I64(n), hist_len, g_cell_px, _p(out),
               _stream(dev))
    return out
