"""Out-of-bounds WRITE check without compute-sanitizer (closed on this pool): every output and workspace buffer that
facerecognition_b200.ops allocates during a call is carved out of a larger allocation with a 4 KiB canary band on each
side; after the call every band must still hold its pattern.  Ragged shapes on purpose (tile remainders, k at the
list sizes, rows not a multiple of anything), through every kernel family: tcgen05 cosine (single-CTA and CTA-pair,
short and long lists), row-streaming, FFMA-tiled, exact re-score, LBP + histograms (u16 / u8, aligned / unaligned, other
radius), chi-square scans (u16 / u8, k = 1 / 5, mixed cell sizes), the tensor-core chi-square filter (incl. the forced
exact fallback), merges, gallery builders and the resize / gray front end."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 4096
PATTERN = 0xA5


class _GuardedTorch:
    """Stands in for the `torch` module inside ops.py: empty / zeros / empty_like hand out guarded CUDA buffers."""

    def __init__(self):
        self.raw = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def _alloc(self, shape, dtype, device, zero):
        if device is None or torch.device(device).type != "cuda":
            return (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
        n = int(np.prod(shape)) if len(shape) else 1
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        body = (nbytes + 255) // 256 * 256
        raw = torch.full((body + 2 * GUARD,), PATTERN, dtype=torch.uint8, device=device)
        self.raw.append((raw, nbytes))
        view = raw[GUARD:GUARD + nbytes].view(dtype).view(shape)
        if zero:
            view.zero_()
        return view

    @staticmethod
    def _shape(size):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            return tuple(int(x) for x in size[0])
        return tuple(int(x) for x in size)

    def empty(self, *size, dtype=torch.float32, device=None, **kw):
        return self._alloc(self._shape(size), dtype, device, False)

    def zeros(self, *size, dtype=torch.float32, device=None, **kw):
        return self._alloc(self._shape(size), dtype, device, True)

    def empty_like(self, t, **kw):
        return self._alloc(tuple(t.shape), t.dtype, t.device, False)

    def check(self, what):
        torch.cuda.synchronize()
        for raw, nbytes in self.raw:
            body = (nbytes + 255) // 256 * 256
            lo, hi = raw[:GUARD], raw[GUARD + nbytes:GUARD + body + GUARD]
            assert bool((lo == PATTERN).all()), f"{what}: write BELOW a {nbytes}-byte buffer"
            assert bool((hi == PATTERN).all()), f"{what}: write PAST a {nbytes}-byte buffer"
        n = len(self.raw)
        self.raw.clear()
        return n


@pytest.fixture
def guarded(monkeypatch):
    from facerecognition_b200 import ops
    g = _GuardedTorch()
    monkeypatch.setattr(ops, "torch", g)
    return g


def _unit(x):
    return x / x.norm(dim=1, keepdim=True)


def test_cosine_kernels_write_inside_their_buffers(guarded):
    from facerecognition_b200 import ops, _native as NV
    gen = torch.Generator(device="cuda").manual_seed(3)
    total = 0
    for n_rows, nq, k, dim in [(70_001, 4097, 5, 512), (33_333, 300, 16, 512), (20_011, 515, 64, 512), (5_003, 129, 33, 256),
                               (1_000, 77, 8, 512), (257, 3, 7, 512), (50_000, 1, 5, 512), (999, 128, 64, 64)]:
        gal = _unit(torch.randn((n_rows, dim), generator=gen, device="cuda"))
        q = torch.randn((nq, dim), generator=gen, device="cuda")
        for dt in (torch.bfloat16, torch.float16):
            g16 = ops.normalize_rows(gal, NV.FRB_QNORM_CLAMP, dt)
            s, i = ops.cosine_topk(q, g16, k, qnorm_mode=NV.FRB_QNORM_CLAMP, idx_base=11)
            total += guarded.check(f"cosine_topk {dt} {nq}x{n_rows} k={k} d={dim}")
            assert bool(((i >= 11) & (i < 11 + n_rows)).all()) or n_rows < k
        if dim == 512:
            q16 = ops.normalize_rows(q, NV.FRB_QNORM_CLAMP, torch.bfloat16)
            ops.cosine_topk_bf16q(q16, ops.normalize_rows(gal, NV.FRB_QNORM_CLAMP, torch.bfloat16), k)
            total += guarded.check(f"cosine_topk_bf16q {nq}x{n_rows} k={k}")
        if nq <= 300:                                            # fp32 kernels: FFMA-tiled / row-streaming, reference rule
            qn, gn = ops.row_norms(q), ops.row_norms(gal)
            ops.cosine_topk(q, gal, k, score_mode=NV.FRB_SCORE_REF_COSINE, q_norms=qn, g_norms=gn)
            total += guarded.check(f"fp32 cosine_topk {nq}x{n_rows} k={k} d={dim}")
            if k <= 16 and dim == 512:
                ops.cosine_topk_exact(q, gal, ops.normalize_rows(gal, NV.FRB_QNORM_CLAMP, torch.float16), k, q_norms=qn, g_norms=gn)
                total += guarded.check(f"cosine_topk_exact {nq}x{n_rows} k={k}")
    # merges of ragged candidate lists
    cs = torch.randn((3, 77, 5), generator=gen, device="cuda")
    ci = torch.randint(0, 1000, (3, 77, 5), generator=gen, device="cuda")
    ops.topk_merge(cs, ci, True)
    total += guarded.check("topk_merge")
    assert total > 100


def test_lbph_kernels_write_inside_their_buffers(guarded):
    from facerecognition_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(4)
    total = 0
    hist = {}
    for n, h, w in [(301, 112, 112), (77, 100, 100), (13, 57, 83), (5, 131, 129), (2, 19, 21)]:
        imgs = torch.randint(0, 256, (n, h, w), generator=gen, device="cuda", dtype=torch.uint8)
        ops.lbp_codes(imgs)
        total += guarded.check(f"lbp_codes {n}x{h}x{w}")
        for c8 in (False, True):
            hh, px = ops.lbp_hist(imgs, counts8=c8)
            total += guarded.check(f"lbp_hist {n}x{h}x{w} counts8={c8}")
            hist[(h, w, c8)] = (hh, px)
        ops.lbp_hist(imgs[:, 1:, 1:].contiguous(), counts8=False)                # odd size: images not 16-byte multiples (no TMA staging)
        total += guarded.check("lbp_hist on a cropped batch")
    imgs = torch.randint(0, 256, (9, 64, 64), generator=gen, device="cuda", dtype=torch.uint8)
    ops.lbp_hist(imgs, radius=2, neighbors=6, grid_x=5, grid_y=3)
    total += guarded.check("lbp_hist radius 2 / 6 neighbours / 5x3 grid")

    g16, px = hist[(112, 112, False)]
    g8, _ = hist[(112, 112, True)]
    q16 = g16[:37].contiguous()
    for k in (1, 5):
        ops.chisq_topk(q16, px, g16, px, k)
        total += guarded.check(f"chisq_topk u16 k={k}")
        ops.chisq_topk(q16, px, g8, px, k)
        total += guarded.check(f"chisq_topk u8 k={k}")
    small16, spx = hist[(100, 100, False)]
    ops.chisq_topk(small16[:9].contiguous(), spx, g16, px, 1)                    # mixed cell sizes (144 vs 169)
    total += guarded.check("chisq_topk mixed cell sizes")
    ops.chisq_dist(q16[:3].contiguous(), px, g16, px)
    total += guarded.check("chisq_dist")
    stats = torch.zeros(4, dtype=torch.int32, device="cuda")
    ops.chisq_top1_filtered(q16, g8, px, stats=stats)
    total += guarded.check("chisq_top1_filtered 37 x 301")
    big8 = g8.repeat(30, 1)[:9001].contiguous()                                 # three ragged gallery tiles per CTA and duplicates
    ops.chisq_top1_filtered(g16[:130].contiguous(), big8, px, stats=stats, want_scores=True)
    total += guarded.check("chisq_top1_filtered 130 x 9001 + scores")
    assert total > 40


def test_filter_fallback_and_front_end_write_inside_their_buffers(guarded, monkeypatch):
    from facerecognition_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(5)
    imgs = torch.randint(0, 256, (4200, 100, 100), generator=gen, device="cuda", dtype=torch.uint8)
    g8, px = ops.lbp_hist(imgs, counts8=True)
    q16, _ = ops.lbp_hist(imgs[:70].contiguous())
    guarded.check("setup")
    monkeypatch.setenv("FRB_CHISQ_FILTER_ACC_REL", "10")                        # window covers every row -> lists overflow -> exact fallback
    stats = torch.zeros(4, dtype=torch.int32, device="cuda")
    d, i = ops.chisq_top1_filtered(q16, g8, px, stats=stats)
    n = guarded.check("chisq_top1_filtered with every query in the exact fallback")
    monkeypatch.delenv("FRB_CHISQ_FILTER_ACC_REL")
    assert int(stats[0]) == 70 and bool((i[:, 0] == torch.arange(70, device="cuda")).all())
    frames = torch.randint(0, 256, (33, 181, 243, 3), generator=gen, device="cuda", dtype=torch.uint8)
    ops.bgr_to_gray(frames)
    n += guarded.check("bgr_to_gray")
    ops.resize_linear(frames, (112, 112), to_gray=True)
    n += guarded.check("resize_linear 181x243x3 -> 112x112 gray")
    ops.resize_linear(frames[:, :50, :60].contiguous(), (100, 100))
    n += guarded.check("resize_linear up-scaling, 3 channels")
    ops.resize_linear(frames[..., 0].contiguous(), (57, 83))
    n += guarded.check("resize_linear one channel")
    emb = torch.randn((1001, 512), generator=gen, device="cuda")
    order = torch.randperm(1001, generator=gen, device="cuda").to(torch.int64)
    offsets = torch.tensor([0, 1, 500, 500, 1001], dtype=torch.int64, device="cuda")
    ops.group_mean_renorm(emb, order, offsets, want_bf16=True)
    n += guarded.check("group_mean_renorm")
    assert n >= 8
