"""Every C-ABI entry point only enqueues work on the caller's stream (no allocation, no synchronisation, no host
read-back), so a caller can record a whole identification step into a CUDA graph and replay it on fresh inputs.
These tests capture each path once and check that the replay on NEW buffer contents equals a plain call."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _capture(fn):
    """Run fn once eagerly on a side stream (kernel attributes, lazy module load), then record it."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    return g, out


@pytest.mark.parametrize("n_query,rows,dtype,k", [
    (384, 70_000, torch.bfloat16, 5),      # tensor-core path: normalise -> warm-up pass -> main pass -> merge
    (1, 70_000, torch.bfloat16, 5),        # row-streaming path
    (48, 9_000, torch.float32, 3),         # tiled fp32 path
    (3, 9_000, torch.float32, 1),          # row-streaming fp32, one pass of 4
])
def test_cosine_step_replays_in_a_graph(n_query, rows, dtype, k):
    from facerecognition_b200 import ops, _native as NV
    g = torch.Generator(device="cuda").manual_seed(5)
    gal = ops.normalize_rows(torch.randn((rows, 512), generator=g, device="cuda"), NV.FRB_QNORM_CLAMP, dtype)
    q_static = torch.empty((n_query, 512), device="cuda")
    q_static.copy_(torch.randn((n_query, 512), generator=g, device="cuda"))
    graph, (s_out, i_out) = _capture(lambda: ops.cosine_topk(q_static, gal, k, qnorm_mode=NV.FRB_QNORM_CLAMP))
    for trial in range(3):
        fresh = gal[torch.randint(0, rows, (n_query,), generator=g, device="cuda")].float() \
            + 0.02 * torch.randn((n_query, 512), generator=g, device="cuda")
        q_static.copy_(fresh)
        graph.replay()
        torch.cuda.synchronize()
        s_ref, i_ref = ops.cosine_topk(fresh, gal, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
        assert torch.equal(i_out, i_ref), f"trial {trial}"
        assert torch.equal(s_out, s_ref), f"trial {trial}"


def test_lbph_step_replays_in_a_graph(oracle_lbph):
    """BGR frames -> gray -> LBP histograms -> chi-square nearest neighbour, recorded once, replayed on new frames."""
    from facerecognition_b200 import ops
    rng = np.random.default_rng(11)
    gal_faces = rng.integers(0, 256, (300, 112, 112), dtype=np.uint8)
    gal, px = ops.lbp_hist(torch.from_numpy(gal_faces).cuda())
    frames = torch.empty((16, 112, 112, 3), dtype=torch.uint8, device="cuda")
    frames.copy_(torch.from_numpy(rng.integers(0, 256, (16, 112, 112, 3), dtype=np.uint8)).cuda())

    def step():
        gray = ops.bgr_to_gray(frames)
        h, qpx = ops.lbp_hist(gray)
        return ops.chisq_topk(h, qpx, gal, px, k=2)

    graph, (d_out, i_out) = _capture(step)
    for trial in range(2):
        new = np.repeat(gal_faces[rng.integers(0, 300, 16)][..., None], 3, axis=3)   # gray frames: gallery faces themselves
        new[8:] = rng.integers(0, 256, (8, 112, 112, 3), dtype=np.uint8)
        frames.copy_(torch.from_numpy(np.ascontiguousarray(new)).cuda())
        graph.replay()
        torch.cuda.synchronize()
        d_ref, i_ref = step()
        assert torch.equal(i_out, i_ref) and torch.equal(d_out, d_ref), f"trial {trial}"
        assert (d_out[:8, 0] == 0).all()          # a gallery face matches itself at distance 0
        want = [oracle_lbph.c_chisq_scan_u16(gal.cpu().numpy(), px, hq, px).argmin() for hq in
                ops.lbp_hist(ops.bgr_to_gray(frames))[0].cpu().numpy()[8:]]
        assert i_out[8:, 0].cpu().tolist() == [int(w) for w in want]


def test_sharded_search_graph_replay_and_host_pipeline_on_one_gpu():
    """ShardedSearch.search(graph=True) (the replayed step bench.py and HostBatchPipeline use) on a single rank: replays on
    refilled input buffers equal plain calls, for fp32 queries and for pre-normalised bf16 queries; the host pipeline returns
    every batch's own answers with two batches in flight."""
    from facerecognition_b200 import ops, _native as NV
    from facerecognition_b200.sharded import HostBatchPipeline, cosine_sharded
    g = torch.Generator(device="cuda").manual_seed(11)
    gal = ops.normalize_rows(torch.randn((50_000, 512), generator=g, device="cuda"), NV.FRB_QNORM_CLAMP, torch.bfloat16)
    search = cosine_sharded(gal, 1000, qnorm_mode=NV.FRB_QNORM_CLAMP)
    q = torch.empty((300, 512), device="cuda")
    q16 = torch.empty((300, 512), dtype=torch.bfloat16, device="cuda")
    for trial in range(3):
        q.copy_(gal[torch.randint(0, 50_000, (300,), generator=g, device="cuda")].float() + 0.05 * torch.randn((300, 512), generator=g, device="cuda"))
        s, i = search.search(q, 5, graph=True)
        ws, wi = ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP, idx_base=1000)
        assert torch.equal(i, wi) and torch.equal(s, ws), trial
        search.gather_normalized(q, q16, 0)                       # world 1: normalise in place of the gather
        s2, i2 = search.search(q16, 5, graph=True)
        assert torch.equal(i2, wi) and torch.equal(s2, ws), trial
    pipe = HostBatchPipeline(search, 300, 512, 5, torch.device("cuda"))
    batches = [(gal[torch.randint(0, 50_000, (300,), generator=g, device="cuda")].float()).cpu().pin_memory() for _ in range(5)]
    tickets, got = [], []
    for b in batches:
        tickets.append(pipe.submit(b))
        if len(tickets) == pipe.depth:
            s, i = pipe.result(tickets.pop(0))
            got.append((s.clone(), i.clone()))
    for t in tickets:
        s, i = pipe.result(t)
        got.append((s.clone(), i.clone()))
    for b, (s, i) in zip(batches, got):
        ws, wi = ops.cosine_topk(b.cuda(), gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP, idx_base=1000)
        assert torch.equal(i, wi.cpu()) and torch.equal(s, ws.cpu())
