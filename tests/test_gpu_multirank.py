"""Two NCCL ranks on two GPUs of one box (skipped with fewer): the fused NVLink exchange + merge kernel against the
unsharded answer, its CUDA-graph replay (device-side epochs), the bf16 query-slice gather, the resync after a skipped
step, and the sharded RecognitionEngine / LBPHFaceRecognizer classes.  The single-GPU emulation of the exchange
protocol is in test_gpu_sharded.py, the host plumbing under gloo in test_host_logic.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q_out):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        import facerecognition_b200 as F
        from facerecognition_b200 import ops, _native as NV
        from facerecognition_b200.sharded import cosine_sharded, shard_bounds

        # ---- cosine: sharded == unsharded, eager / graph replay / bf16 slice gather ----
        gen = torch.Generator(device=dev).manual_seed(1)
        N, Q, k = 40_003, 512, 5
        gal = ops.normalize_rows(torch.randn((N, 512), generator=gen, device=dev), NV.FRB_QNORM_CLAMP, torch.bfloat16)
        gal[N - 1] = gal[5]                                            # a tie across the two shards
        q = torch.randn((Q, 512), generator=gen, device=dev)
        q[0] = gal[5].float()
        want_s, want_i = ops.cosine_topk(q, gal, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
        lo, hi = shard_bounds(N, world, rank)
        search = cosine_sharded(gal[lo:hi].contiguous(), lo, qnorm_mode=NV.FRB_QNORM_CLAMP)
        s, i = search.search(q, k)
        assert search._exchange is not None, "peer-memory exchange (CUDA IPC) should be available on one box"
        assert torch.equal(i, want_i) and torch.equal(s, want_s)
        assert [int(x) for x in i[0, :2]] == [5, N - 1]
        for rep in range(4):                                           # graph replay: epochs advance on the device
            q2 = q.roll(rep, 0).contiguous()
            q.copy_(q2)
            s, i = search.search(q, k, graph=True)
            w_s, w_i = ops.cosine_topk(q, gal, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
            assert torch.equal(i, w_i) and torch.equal(s, w_s), f"graph replay {rep}"
        per = Q // world
        stage16 = torch.empty((Q, 512), dtype=torch.bfloat16, device=dev)
        q16 = search.gather_normalized(q[rank * per:(rank + 1) * per].contiguous(), stage16, rank * per)
        s, i = search.search(q16, k, graph=True)
        w_s, w_i = ops.cosine_topk(q, gal, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
        assert torch.equal(i, w_i) and torch.equal(s, w_s), "bf16 slice gather"
        timeouts, epoch = search._exchange.status()
        assert timeouts == 0 and epoch >= 6
        search.resync()                                                # collective reset: epochs restart, answers unchanged
        assert search._exchange.status() == (0, 0)
        s, i = search.search(q, k, graph=True)
        assert torch.equal(i, w_i) and torch.equal(s, w_s), "after resync"

        # ---- RecognitionEngine(group=True) == single-GPU engine ----
        rng = np.random.default_rng(3)
        g32 = rng.standard_normal((3001, 512)).astype(np.float32)
        g32 /= np.linalg.norm(g32, axis=1, keepdims=True)
        g32[2900] = g32[17]
        db = {f"id_{j:05d}": v for j, v in enumerate(g32)}
        qs = (g32[rng.integers(0, 3001, 40)] + 0.03 * rng.standard_normal((40, 512))).astype(np.float32)
        qs[0] = g32[17]
        single = F.RecognitionEngine(model_path=None, threshold=0.4, use_face_detection=False, device=f"cuda:{rank}")
        single.db = db
        sharded = F.RecognitionEngine(model_path=None, threshold=0.4, use_face_detection=False, device=f"cuda:{rank}", group=True)
        sharded.db = db
        a, b = single.recognize_embeddings(qs), sharded.recognize_embeddings(torch.from_numpy(qs).to(dev))
        assert sharded.gallery().rows.shape[0] in (1500, 1501)
        for x, y in zip(a, b):
            assert x[0] == y[0] and abs(x[1] - y[1]) <= 2e-6 and [t[0] for t in x[2]] == [t[0] for t in y[2]]
        assert [t[0] for t in b[0][2][:2]] == ["id_00017", "id_02900"]

        # ---- LBPHFaceRecognizer(group=True) == single-GPU model ----
        faces = rng.integers(0, 256, (41, 100, 100), dtype=np.uint8)
        faces[33] = faces[6]
        labels = np.arange(41, dtype=np.int32) + 500
        m1 = F.LBPHFaceRecognizer_create(device=f"cuda:{rank}")
        m1.train(list(faces), labels)
        m2 = F.LBPHFaceRecognizer_create(device=f"cuda:{rank}", group=True)
        m2.train(list(faces[:30]), labels[:30])
        m2.update(list(faces[30:]), labels[30:])
        probe = [faces[6], faces[33], faces[40], rng.integers(0, 256, (100, 100), dtype=np.uint8)]
        l1, d1 = m1.predict_batch(probe)
        l2, d2 = m2.predict_batch(probe)
        assert list(l1) == list(l2) and np.array_equal(d1, d2) and list(l2[:3]) == [506, 506, 540]
        assert m2.predict(faces[12]) == m1.predict(faces[12])
        dist.barrier()
        torch.cuda.synchronize()
        q_out.put((rank, True))
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        import traceback
        q_out.put((rank, traceback.format_exc()))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_two_rank_nccl_exchange_graph_gather_resync_and_sharded_classes():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)], results
