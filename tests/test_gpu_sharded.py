"""Sharding by identity: R local searches with global id offsets + frb_topk_merge give exactly the
unsharded answer for any R (single process, shards emulated on one GPU; the multi-process
all-gather plumbing is covered by tests/test_host_logic.py under gloo)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R", [2, 3, 8])
def test_cosine_shards_merge_to_unsharded_result(R):
    from facerecognition_b200 import ops, _native as NV
    from facerecognition_b200.sharded import shard_bounds
    gen = torch.Generator(device="cuda").manual_seed(R)
    N, Q, k = 50_001, 200, 5
    gal = ops.normalize_rows(torch.randn((N, 512), generator=gen, device="cuda"), NV.FRB_QNORM_CLAMP, torch.bfloat16)
    gal[N - 1] = gal[5]                                                    # tie across the first and last shard
    q = torch.randn((Q, 512), generator=gen, device="cuda")
    q[0] = gal[5].float()
    full_s, full_i = ops.cosine_topk(q, gal, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
    cs, ci = [], []
    for r in range(R):
        lo, hi = shard_bounds(N, R, r)
        s, i = ops.cosine_topk(q, gal[lo:hi].contiguous(), k, qnorm_mode=NV.FRB_QNORM_CLAMP, idx_base=lo)
        cs.append(s); ci.append(i)
    ms, mi = ops.topk_merge(torch.stack(cs).contiguous(), torch.stack(ci).contiguous(), largest=True)
    assert torch.equal(mi, full_i) and torch.equal(ms, full_s)
    assert mi[0, 0].item() == 5 and mi[0, 1].item() == N - 1
    # the packed-record layout one all-gather leaves behind (sharded.py) merges in place to the same answer
    buf, _, _ = ops.packed_candidates(Q, k, q.device, n_lists=R)
    for r in range(R):
        buf[r, :Q * k * 8].view(torch.int64).copy_(ci[r].reshape(-1))
        buf[r, Q * k * 8:Q * k * 12].view(torch.float32).copy_(cs[r].reshape(-1))
    ps, pi = ops.topk_merge_packed(buf, Q, k, largest=True)
    assert torch.equal(pi, full_i) and torch.equal(ps, full_s)


@pytest.mark.parametrize("R", [2, 8])
def test_chisq_shards_merge_to_unsharded_result(R):
    from facerecognition_b200 import ops
    from facerecognition_b200.sharded import shard_bounds
    rng = np.random.default_rng(R)
    N, Q, px = 3001, 9, 144
    gal = rng.multinomial(px, np.ones(256) / 256, size=(N, 64)).astype(np.uint16).reshape(N, -1)
    q = rng.multinomial(px, np.ones(256) / 256, size=(Q, 64)).astype(np.uint16).reshape(Q, -1)
    gal[N - 2] = gal[3]
    q[0] = gal[3]
    g, qd = torch.from_numpy(gal).cuda(), torch.from_numpy(q).cuda()
    full_d, full_i = ops.chisq_topk(qd, px, g, px, k=3)
    cd, ci = [], []
    for r in range(R):
        lo, hi = shard_bounds(N, R, r)
        d, i = ops.chisq_topk(qd, px, g[lo:hi].contiguous(), px, k=3, idx_base=lo)
        cd.append(d); ci.append(i)
    md, mi = ops.topk_merge(torch.stack(cd).contiguous(), torch.stack(ci).contiguous(), largest=False)
    assert torch.equal(mi, full_i) and torch.equal(md, full_d)
    assert [int(x) for x in mi[0, :2]] == [3, N - 2] and md[0, 0].item() == 0.0


@pytest.mark.parametrize("R,largest", [(2, True), (8, True), (3, False)])
def test_peer_exchange_protocol_emulated_on_one_gpu(R, largest):
    """frb_exchange_emulate: R rank contexts in one process, ONE launch with blockIdx.y = rank, runs the real
    store / release-flag / acquire-wait / merge kernel; every emulated rank must end with the unsharded merge.
    Two steps in a row exercise the epoch / parity double buffering, and a smaller second batch the re-layout."""
    from facerecognition_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(R)
    ranks = [ops.Exchange(R, r, 700, 5, torch.device("cuda")) for r in range(R)]
    try:
        for Q, k in [(700, 5), (333, 3), (700, 5)]:
            s = torch.randn((R, Q, k), generator=gen, device="cuda")
            s = torch.sort(s, dim=2, descending=largest).values
            i = torch.randint(0, 10_000, (R, Q, k), generator=gen, device="cuda")
            s[1, 0, 0] = s[0, 0, 0]                       # a tie across ranks: the lower id must win
            i[R - 1, 5, k - 1] = -1                        # padding entry
            want_s, want_i = ops.topk_merge(s.contiguous(), i.contiguous(), largest)
            got_s, got_i = ops.exchange_emulate(ranks, s.contiguous(), i.contiguous(), largest)
            for r in range(R):
                assert torch.equal(got_i[r], want_i) and torch.equal(got_s[r], want_s)
    finally:
        for x in ranks:
            x.close()


def test_host_batch_pipeline_matches_blocking_calls():
    """sharded.HostBatchPipeline (two batches in flight, copies on side streams) returns, batch by batch, exactly
    what a blocking search of the same host batch returns — including when the staging slots are reused."""
    from facerecognition_b200 import ops, _native as NV
    from facerecognition_b200.sharded import HostBatchPipeline, cosine_sharded
    g = torch.Generator(device="cuda").manual_seed(3)
    gal = ops.normalize_rows(torch.randn((50_000, 512), generator=g, device="cuda"), NV.FRB_QNORM_CLAMP, torch.bfloat16)
    search = cosine_sharded(gal, 0, qnorm_mode=NV.FRB_QNORM_CLAMP)
    batches = [(gal[torch.randint(0, 50_000, (300,), generator=g, device="cuda")].float()
                + 0.02 * torch.randn((300, 512), generator=g, device="cuda")).cpu().pin_memory() for _ in range(7)]
    pipe = HostBatchPipeline(search.search, 300, 512, 5, torch.device("cuda", 0))
    got, tickets = [], []
    for b in batches:
        tickets.append(pipe.submit(b))
        if len(tickets) == pipe.depth:
            s, i = pipe.result(tickets.pop(0))
            got.append((s.clone(), i.clone()))
    for t in tickets:
        s, i = pipe.result(t)
        got.append((s.clone(), i.clone()))
    assert len(got) == len(batches)
    for b, (s, i) in zip(batches, got):
        rs, ri = search.search(b.cuda(), 5)
        assert torch.equal(i, ri.cpu()) and torch.equal(s, rs.cpu())
    with pytest.raises(RuntimeError):
        pipe.submit(batches[0]); pipe.submit(batches[1]); pipe.submit(batches[2])   # third without taking a result
