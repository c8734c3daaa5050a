"""LBPH parity on the GPU, through the C ABI (facerecognition_b200.ops / .lbph -> libfrb200.so):
LBP codes and u16 histograms bit-exact against the oracle; chi-square within 1e-5 relative of the
oracle (itself pinned on cv2.compareHist) and of cv2.compareHist directly; predict() semantics."""
import cv2
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-5          # north_star: chi-square within 1e-5 relative of OpenCV's compareHist
DBL_MAX = np.finfo(np.float64).max


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def synth_faces(seed, n, h, w):
    rng = np.random.default_rng(seed)
    out = np.zeros((n, h, w), np.uint8)
    for i in range(n):
        kind = i % 4
        if kind == 0:
            img = rng.integers(0, 255, (h, w), dtype=np.uint8)
            c = (i // 4) % max(1, h // 10)
            img[c * 10:(c + 1) * 10, :] = 255                       # test_lbph_logic.py:26-28 stripe
        elif kind == 1:
            img = cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (0, 0), 1.0 + (i % 5))
        elif kind == 2:
            img = np.zeros((h, w), np.uint8)
            for _ in range(20):
                y0, x0 = rng.integers(0, h), rng.integers(0, w)
                img[y0:y0 + rng.integers(3, 40), x0:x0 + rng.integers(3, 40)] = rng.choice([0, 1, 2, 3, 7, 127, 128, 254, 255])
        else:
            img = rng.integers(0, 4, (h, w), dtype=np.uint8)         # values 0..3: exercises the centre==1 threshold
        out[i] = img
    return out


@pytest.mark.parametrize("tag", ["s100", "s112", "s57x83"])
def test_codes_and_hist_bit_exact_on_golden(lbph_golden, oracle_lbph, tag):
    from facerecognition_b200 import ops
    faces = lbph_golden[f"{tag}_faces"]
    codes = ops.lbp_codes(dev(faces)).cpu().numpy()
    np.testing.assert_array_equal(codes[0].astype(np.int32), lbph_golden[f"{tag}_codes0"])
    for i, f in enumerate(faces):
        np.testing.assert_array_equal(codes[i].astype(np.int32), oracle_lbph.c_elbp(f))
    hist, px = ops.lbp_hist(dev(faces))
    assert px == int(lbph_golden[f"{tag}_cell_px"])
    np.testing.assert_array_equal(hist.cpu().numpy(), lbph_golden[f"{tag}_hist"])


@pytest.mark.parametrize("shape,n", [((100, 100), 257), ((112, 112), 130), ((3, 3), 5), ((10, 10), 9), ((64, 200), 33),
                                     ((131, 97), 17), ((18, 18), 300)])
def test_codes_and_hist_bit_exact_seeded(oracle_lbph, shape, n):
    from facerecognition_b200 import ops
    faces = synth_faces(11 + shape[0], n, *shape)
    want_hist, want_px = oracle_lbph.c_lbp_hist(faces)
    hist, px = ops.lbp_hist(dev(faces))
    assert px == want_px
    np.testing.assert_array_equal(hist.cpu().numpy(), want_hist)
    codes = ops.lbp_codes(dev(faces[:8])).cpu().numpy()
    for i in range(min(8, n)):
        np.testing.assert_array_equal(codes[i].astype(np.int32), oracle_lbph.c_elbp(faces[i]))


def test_flat_images_and_other_grids(oracle_lbph, lbph_golden):
    from facerecognition_b200 import ops
    flat = np.stack([np.full((12, 12), v, np.uint8) for v in range(256)])
    codes = ops.lbp_codes(dev(flat)).cpu().numpy()
    np.testing.assert_array_equal(codes[:, 3, 3].astype(np.int32), lbph_golden["flat_codes"])
    faces = synth_faces(3, 12, 100, 100)
    for gx, gy in [(4, 4), (8, 4), (1, 1), (7, 5)]:
        want, wpx = oracle_lbph.c_lbp_hist(faces, 1, 8, gx, gy)
        got, px = ops.lbp_hist(dev(faces), 1, 8, gx, gy)
        assert px == wpx
        np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_unaligned_image_batches(oracle_lbph):
    """rows*cols not a multiple of 16 -> the byte-wise staging path."""
    from facerecognition_b200 import ops
    faces = synth_faces(5, 7, 33, 35)
    want, _ = oracle_lbph.c_lbp_hist(faces)
    np.testing.assert_array_equal(ops.lbp_hist(dev(faces))[0].cpu().numpy(), want)


def test_chisq_distances_vs_oracle_and_cv2(oracle_lbph, lbph_golden):
    from facerecognition_b200 import ops
    for tag in ["s100", "s112", "s57x83"]:
        hist, px = lbph_golden[f"{tag}_hist"], int(lbph_golden[f"{tag}_cell_px"])
        d = ops.chisq_dist(dev(hist), px, dev(hist), px).cpu().numpy().astype(np.float64)
        ref = lbph_golden[f"{tag}_cv2_chisq"]                      # the real cv2.compareHist
        np.testing.assert_allclose(d, ref, rtol=REL, atol=0)
        assert np.all(np.diag(d) == 0.0)                           # identical histograms -> exactly 0, like OpenCV


def test_chisq_topk_matches_scan_and_first_wins_ties(oracle_lbph):
    from facerecognition_b200 import ops
    faces = synth_faces(21, 700, 100, 100)
    hist, px = oracle_lbph.c_lbp_hist(faces)
    gal = hist.copy()
    gal[650] = gal[40]                                             # duplicate rows far apart (different CTA chunks)
    gal[41] = gal[40]
    q = np.concatenate([hist[[40, 5, 699]], oracle_lbph.c_lbp_hist(synth_faces(99, 13, 100, 100))[0]])
    dist, idx = ops.chisq_topk(dev(q), px, dev(gal), px, k=5)
    dist, idx = dist.cpu().numpy().astype(np.float64), idx.cpu().numpy()
    for r in range(q.shape[0]):
        ref = oracle_lbph.c_chisq_scan_u16(gal, px, q[r], px)
        order = np.lexsort((np.arange(len(ref)), ref))[:5]
        np.testing.assert_allclose(dist[r], ref[order], rtol=REL)
        # identical except stated near-ties: positions whose reference gap is below the tolerance
        for j in range(5):
            if idx[r, j] != order[j]:
                assert abs(ref[idx[r, j]] - ref[order[j]]) <= REL * ref[order[j]]
    assert list(idx[0, :3]) == [40, 41, 650] and np.all(dist[0, :3] == 0.0)
    full = ops.chisq_dist(dev(q), px, dev(gal), px).cpu().numpy()
    np.testing.assert_array_equal(full.min(1), dist[:, 0].astype(np.float32))   # top-1 == min of the full scan


def test_chisq_mixed_cell_sizes_and_edge_cases(oracle_lbph):
    from facerecognition_b200 import ops
    g_hist, g_px = oracle_lbph.c_lbp_hist(synth_faces(1, 40, 112, 112))
    q_hist, q_px = oracle_lbph.c_lbp_hist(synth_faces(2, 6, 100, 100))
    assert g_px != q_px
    d = ops.chisq_dist(dev(q_hist), q_px, dev(g_hist), g_px).cpu().numpy().astype(np.float64)
    ref = np.stack([oracle_lbph.c_chisq_scan_u16(g_hist, g_px, q, q_px) for q in q_hist])
    np.testing.assert_allclose(d, ref, rtol=REL)
    # k larger than the gallery -> (+inf, -1) padding; empty gallery -> all padding
    dist, idx = ops.chisq_topk(dev(q_hist), q_px, dev(g_hist[:3]), g_px, k=5)
    assert np.all(idx.cpu().numpy()[:, 3:] == -1) and np.isinf(dist.cpu().numpy()[:, 3:]).all()
    dist, idx = ops.chisq_topk(dev(q_hist), q_px, torch.zeros((0, 16384), dtype=torch.uint16, device="cuda"), g_px, k=2)
    assert np.all(idx.cpu().numpy() == -1)
    # idx_base shifts ids (what a shard reports)
    _, i0 = ops.chisq_topk(dev(q_hist), q_px, dev(g_hist), g_px, k=3)
    _, i1 = ops.chisq_topk(dev(q_hist), q_px, dev(g_hist), g_px, k=3, idx_base=1000)
    assert torch.equal(i0 + 1000, i1)


def test_recognizer_protocol_matches_oracle(oracle_lbph, lbph_golden, tmp_path):
    import facerecognition_b200 as F
    faces = list(lbph_golden["s100_faces"])
    labels = np.arange(len(faces), dtype=np.int32) + 100
    model = F.train_lbph_model(faces, labels, 1, 8, 8, 8)           # models/lbphmodel/train_lbph.py signature
    ref = oracle_lbph.OracleLBPH()
    ref.train(faces, labels)
    for i, f in enumerate(faces):
        assert model.predict(f) == (100 + i, 0.0)
    fresh = synth_faces(77, 24, 100, 100)
    for f in fresh:
        lab, conf = model.predict(f)
        rlab, rconf = ref.predict(f)
        assert lab == rlab and abs(conf - rconf) <= REL * rconf
    labs, confs = model.predict_batch(list(fresh))
    assert [int(x) for x in labs] == [ref.predict(f)[0] for f in fresh]
    # recognize_face wrapper (inference_lbph.py:4-18): strict '<'
    lab, conf = model.predict(fresh[0])
    assert F.recognize_face(model, fresh[0], conf + 1)["status"] == "known"
    assert F.recognize_face(model, fresh[0], conf) == {"label": None, "confidence": conf, "status": "unknown"}
    # model threshold -> (-1, DBL_MAX)
    model.setThreshold(0.0)
    assert model.predict(faces[0]) == (-1, DBL_MAX)
    model.setThreshold(DBL_MAX)
    # update() appends; duplicates keep the first label
    model.update([faces[2]], np.array([999], np.int32))
    assert model.predict(faces[2]) == (102, 0.0)
    # getHistograms is OpenCV's float view
    np.testing.assert_array_equal(model.getHistograms()[1].ravel(), ref.hists[1])
    # evaluate_lbph / find_optimal_threshold mirrors
    acc, cov, used, confs = F.evaluate_lbph(model, faces, labels, 1.0)
    assert (acc, cov, used) == (1.0, 1.0, len(faces)) and np.all(confs == 0.0)
    thr, score, rows = F.find_optimal_threshold(model, faces, labels, min_coverage=0.3)
    assert thr == 40 and score == 1.0 and len(rows) == 17
    # save / read round trip through the OpenCV FileStorage layout
    p = str(tmp_path / "lbph_model.xml")
    model.save(p)
    assert "<opencv_lbphfaces>" in open(p).read(4096)
    m2 = F.LBPHFaceRecognizer_create()
    m2.read(p)
    assert m2.size == model.size and m2.predict(fresh[3]) == model.predict(fresh[3])
    # a model that was read back keeps the compact (u8) gallery form and can still be updated
    assert all(g.hist.dtype == torch.uint8 for g in m2._groups) and all(g.hist.dtype == torch.uint8 for g in model._groups)
    m2.update([fresh[5]], np.array([4242], np.int32))
    assert m2.predict(fresh[5]) == (4242, 0.0) and m2.size == model.size + 1
    np.testing.assert_array_equal(m2.get_histograms_u16()[0][:model.size], model.get_histograms_u16()[0])
    with pytest.raises(F.lbph.LBPHError):
        F.LBPHFaceRecognizer_create().predict(faces[0])
    with pytest.raises(F.lbph.LBPHError):
        model.train(faces, labels[:-1])


def test_mixed_size_gallery(oracle_lbph):
    import facerecognition_b200 as F
    a, b = synth_faces(31, 9, 100, 100), synth_faces(32, 8, 112, 112)
    faces = [a[0], b[0], a[1], b[1], a[2]] + list(a[3:]) + list(b[2:])
    labels = np.arange(len(faces), dtype=np.int32)
    model = F.train_lbph_model(faces, labels)
    ref = oracle_lbph.OracleLBPH()
    ref.train(faces, labels)
    for f in [a[1], b[1], synth_faces(33, 1, 100, 100)[0], synth_faces(34, 1, 112, 112)[0]]:
        lab, conf = model.predict(f)
        rlab, rconf = ref.predict(f)
        assert lab == rlab and abs(conf - rconf) <= REL * max(rconf, 1e-30)


def test_baseline_config1_shape_train_predict_1k(oracle_lbph):
    """BASELINE configs[0]: LBPH (1, 8, 8x8) train + predict on 1k synthetic 100x100 faces."""
    import facerecognition_b200 as F
    faces = synth_faces(2024, 1000, 100, 100)
    labels = (np.arange(1000) // 10).astype(np.int32)
    model = F.train_lbph_model(list(faces), labels)
    labs, confs = model.predict_batch(torch.from_numpy(faces).cuda())
    first = {}
    hist, _ = oracle_lbph.c_lbp_hist(faces)
    for i, h in enumerate(hist):
        first.setdefault(h.tobytes(), i)
    assert np.all(confs == 0.0)
    assert [int(x) for x in labs] == [int(labels[first[h.tobytes()]]) for h in hist]
    probe = synth_faces(4048, 64, 100, 100)
    labs, confs = model.predict_batch(list(probe))
    ph, px = oracle_lbph.c_lbp_hist(probe)
    for r in range(64):
        ref = oracle_lbph.c_chisq_scan_u16(hist, px, ph[r], px)
        j = int(np.argmin(ref))
        assert abs(confs[r] - ref[j]) <= REL * ref[j]
        assert labs[r] == labels[j] or abs(ref[j] - np.partition(ref, 1)[1]) <= REL * ref[j]


def test_bgr2gray_is_bit_exact_with_cv2():
    """frb_bgr2gray_u8 vs the REAL cv2.cvtColor(COLOR_BGR2GRAY) of the installed OpenCV core: a colour cube slab
    (every B, G with 16 R values), random frames of ragged sizes, and the batched video front end."""
    import cv2
    from facerecognition_b200 import ops
    b, g, r = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8),
                          np.arange(0, 256, 16, dtype=np.uint8) + 7, indexing="ij")
    cube = np.stack([b.ravel(), g.ravel(), r.ravel()], 1).reshape(1024, 1024, 3)
    got = ops.bgr_to_gray(torch.from_numpy(cube).cuda()).cpu().numpy()
    np.testing.assert_array_equal(got, cv2.cvtColor(cube, cv2.COLOR_BGR2GRAY))
    rng = np.random.default_rng(4)
    for shape in [(1, 1, 3), (3, 5, 3), (7, 112, 112, 3), (2, 101, 99, 3)]:
        x = rng.integers(0, 256, shape, dtype=np.uint8)
        if len(shape) == 4:
            ref = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in x])
        else:
            ref = cv2.cvtColor(x, cv2.COLOR_BGR2GRAY)
        np.testing.assert_array_equal(ops.bgr_to_gray(torch.from_numpy(x).cuda()).cpu().numpy(), ref)
    # unaligned base pointer -> scalar path
    x = rng.integers(0, 256, (64 * 3 + 1,), dtype=np.uint8)
    t = torch.from_numpy(x).cuda()[1:].view(64, 3)
    np.testing.assert_array_equal(ops.bgr_to_gray(t.contiguous()).cpu().numpy(),
                                  cv2.cvtColor(x[1:].reshape(1, 64, 3), cv2.COLOR_BGR2GRAY)[0])


@pytest.mark.parametrize("L", [256, 4096, 8960, 16384])
def test_chisq_other_histogram_lengths_and_large_counts(oracle_lbph, L):
    """K3 with other grids (hist_len 256 = 1x1, 4096 = 4x4, 8960 = 7x5, 16384 = 8x8): partial and full per-thread
    coverage, same and different cell sizes, counts up to the u16 limit, tail rows (gallery not a multiple of 12)."""
    from facerecognition_b200 import ops
    rng = np.random.default_rng(L)
    N, Q = 157, 5
    gal = rng.integers(0, 60_000, (N, L)).astype(np.uint16)
    gal[rng.random((N, L)) < 0.4] = 0                                   # empty bins on both sides
    q = gal[rng.integers(0, N, Q)].copy()
    q[1:] = np.where(rng.random((Q - 1, L)) < 0.1, rng.integers(0, 60_000, (Q - 1, L)), q[1:]).astype(np.uint16)
    for q_px, g_px in [(30_000, 30_000), (144, 169)]:
        ref = np.stack([oracle_lbph.c_chisq_scan_u16(gal, g_px, qq, q_px) for qq in q])
        d = ops.chisq_dist(dev(q), q_px, dev(gal), g_px).cpu().numpy().astype(np.float64)
        np.testing.assert_allclose(d, ref, rtol=REL, atol=0)
        dist, idx = ops.chisq_topk(dev(q), q_px, dev(gal), g_px, k=3)
        order = np.argsort(ref, axis=1, kind="stable")[:, :3]
        np.testing.assert_allclose(dist.cpu().numpy(), np.take_along_axis(ref, order, 1), rtol=REL)
        if q_px == g_px:
            assert float(dist[0, 0]) == 0.0                             # q[0] is a gallery row: exactly zero


def test_resize_front_end_is_bit_exact_with_cv2():
    """frb_resize_linear_u8 vs the REAL cv2.resize (default INTER_LINEAR) / cv2.cvtColor of the installed OpenCV core
    and vs oracle/resize.py: batches, 1 and 3 channels, up- and down-scaling, the exact-halving (INTER_AREA) case,
    1-pixel sides, and the fused resize + BGR2GRAY of _preprocess_image_for_lbph (train_lbph_script.py:67-72)."""
    import cv2
    from facerecognition_b200 import ops
    from oracle import resize as OR
    rng = np.random.default_rng(314)
    cases = [(480, 640, 100, 100), (250, 250, 112, 112), (50, 40, 100, 100), (200, 200, 100, 100), (224, 224, 112, 112),
             (100, 100, 100, 100), (37, 53, 112, 112), (720, 1280, 112, 112), (1, 1, 7, 5), (9, 1, 4, 6), (1, 13, 3, 40),
             (2, 2, 1, 1), (64, 48, 32, 24), (31, 29, 300, 517)]
    cases += [tuple(int(v) for v in rng.integers(1, 260, 4)) for _ in range(25)]
    for sh, sw, dh, dw in cases:
        n = 3
        for ch in (1, 3):
            imgs = rng.integers(0, 256, (n, sh, sw, ch) if ch == 3 else (n, sh, sw), dtype=np.uint8)
            if (sh + sw) % 3 == 0:
                imgs[1] = (imgs[1] // 64) * 85          # flat patches
            got = ops.resize_linear(torch.from_numpy(imgs).cuda(), (dw, dh)).cpu().numpy()
            for b in range(n):
                ref = cv2.resize(imgs[b], (dw, dh)).reshape(got[b].shape)
                np.testing.assert_array_equal(got[b], ref, err_msg=f"{sh}x{sw} -> {dh}x{dw}, {ch} ch")
                np.testing.assert_array_equal(got[b], OR.resize_linear_u8(imgs[b], dw, dh).reshape(got[b].shape))
            if ch == 3:
                gray = ops.resize_linear(torch.from_numpy(imgs).cuda(), (dw, dh), to_gray=True).cpu().numpy()
                for b in range(n):
                    want = cv2.cvtColor(cv2.resize(imgs[b], (dw, dh)).reshape(dh, dw, 3), cv2.COLOR_BGR2GRAY)
                    np.testing.assert_array_equal(gray[b], want, err_msg=f"fused gray {sh}x{sw} -> {dh}x{dw}")


def test_resize_front_end_unaligned_batch_and_last_pixels():
    """A batch that starts at an odd address (byte-load path) and the last pixels of the last image (the word-load
    path must not read past the buffer) give the same bytes as cv2."""
    import cv2
    from facerecognition_b200 import ops
    rng = np.random.default_rng(5)
    for off in (0, 1, 2, 3):
        n, sh, sw = 2, 33, 47
        flat = torch.from_numpy(rng.integers(0, 256, off + n * sh * sw * 3, dtype=np.uint8)).cuda()
        frames = flat[off:].view(n, sh, sw, 3)
        host = frames.cpu().numpy()
        for dsize in [(20, 15), (47, 33), (60, 70), (46, 32)]:
            got = ops.resize_linear(frames, dsize).cpu().numpy()
            gray = ops.resize_linear(frames, dsize, to_gray=True).cpu().numpy()
            for b in range(n):
                ref = cv2.resize(host[b], dsize)
                np.testing.assert_array_equal(got[b], ref, err_msg=f"offset {off} dsize {dsize}")
                np.testing.assert_array_equal(gray[b], cv2.cvtColor(ref, cv2.COLOR_BGR2GRAY))


def test_resize_front_end_large_destination_and_limits():
    """An up-scale whose tap tables need more than 48 KB of shared memory (opt-in path), and the documented limit."""
    import cv2
    from facerecognition_b200 import ops, _native as NV
    rng = np.random.default_rng(6)
    img = rng.integers(0, 256, (1, 40, 30, 3), dtype=np.uint8)
    got = ops.resize_linear(torch.from_numpy(img).cuda(), (3000, 1000)).cpu().numpy()
    np.testing.assert_array_equal(got[0], cv2.resize(img[0], (3000, 1000)))
    with pytest.raises(NV.FrbError):
        ops.resize_linear(torch.from_numpy(img).cuda(), (4097, 10))
    with pytest.raises(NV.FrbError):
        ops.resize_linear(torch.from_numpy(img[..., :2].copy()).cuda(), (10, 10))     # 2 channels


def test_predict_device_frames_equals_the_host_preprocessing(oracle_lbph):
    """Frames of another size -> device resize + gray -> LBPH predict == cv2.resize + cv2.cvtColor on the host, then
    predict (the reference's no-detector path, web_app.py:484-486 + :587)."""
    import cv2
    from facerecognition_b200.lbph import LBPHFaceRecognizer_create, preprocess_frames_device
    rng = np.random.default_rng(8)
    frames = rng.integers(0, 256, (12, 180, 240, 3), dtype=np.uint8)
    host = [cv2.cvtColor(cv2.resize(f, (100, 100)), cv2.COLOR_BGR2GRAY) for f in frames]
    dev_frames = torch.from_numpy(frames).cuda()
    np.testing.assert_array_equal(preprocess_frames_device(dev_frames, (100, 100)).cpu().numpy(), np.stack(host))
    model = LBPHFaceRecognizer_create()
    model.train(host[:8], np.arange(8, dtype=np.int32))
    dist, idx = model.predict_device_frames(dev_frames, (100, 100))
    for j in range(12):
        lab, conf = model.predict(host[j])
        assert int(model.getLabels()[int(idx[j, 0]), 0]) == lab and float(dist[j, 0]) == pytest.approx(conf, rel=1e-6, abs=1e-12)
    assert (dist[:8, 0] == 0).all()


@pytest.mark.parametrize("L,N,Q,px,q_px", [(16384, 700, 5, 169, 169), (16384, 300, 3, 144, 169), (4096, 257, 4, 255, 255),
                                           (2064, 90, 2, 36, 36), (16, 40, 3, 9, 9), (12288, 130, 2, 100, 121)])
def test_chisq_u8_gallery_matches_oracle_and_u16_path(oracle_lbph, L, N, Q, px, q_px):
    """frb_chisq_*_g8: the gallery stored as u8 counts (cell_px <= 255) gives the oracle's distances (<= 1e-5), the
    same nearest rows as the u16 gallery, exactly 0 for a self match, first row on ties; full and partial register
    coverage (L = 16384 / others), equal and different cell sizes."""
    from facerecognition_b200 import ops
    rng = np.random.default_rng(L + N)
    gal = rng.multinomial(px, np.ones(256) / 256, size=(N, (L + 255) // 256)).reshape(N, -1)[:, :L].astype(np.uint16) \
        if L >= 256 else rng.integers(0, px + 1, (N, L)).astype(np.uint16)
    gal[N // 2] = gal[3]                                    # duplicate row: the first one must win
    q = gal[rng.integers(0, N, Q)].copy()
    q[0] = gal[3]
    if q_px != px:
        q = np.minimum(q.astype(np.int64) * q_px // px, q_px).astype(np.uint16)
    g16, g8 = dev(gal), ops.compact_histograms(dev(gal), px)
    assert g8.dtype == torch.uint8
    ref = np.stack([oracle_lbph.c_chisq_scan_u16(gal, px, qq, q_px) for qq in q])
    d8 = ops.chisq_dist(dev(q), q_px, g8, px).cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(d8, ref, rtol=1e-5, atol=0)
    for k in (1, 4):
        dist8, idx8 = ops.chisq_topk(dev(q), q_px, g8, px, k=k, idx_base=1000)
        dist16, idx16 = ops.chisq_topk(dev(q), q_px, g16, px, k=k, idx_base=1000)
        order = np.argsort(ref, axis=1, kind="stable")[:, :k]
        gaps_ok = np.take_along_axis(ref, order, 1)
        np.testing.assert_allclose(dist8.cpu().numpy(), gaps_ok, rtol=1e-5, atol=0)
        assert torch.equal(idx8[:, 0], idx16[:, 0])
    if q_px == px:
        d, i = ops.chisq_topk(dev(q[:1]), q_px, g8, px, k=1)
        assert float(d[0, 0]) == 0.0 and int(i[0, 0]) == 3


def test_chisq_u8_gallery_limits():
    from facerecognition_b200 import ops, _native as NV
    q = torch.zeros((1, 4096), dtype=torch.int16, device="cuda").view(torch.uint16)
    g = torch.zeros((4, 4096), dtype=torch.uint8, device="cuda")
    with pytest.raises(NV.FrbError):
        ops.chisq_topk(q, 300, g, 300, k=1)                 # counts up to 300 do not fit a byte
    q24 = torch.zeros((1, 24), dtype=torch.int16, device="cuda").view(torch.uint16)
    with pytest.raises(NV.FrbError):
        ops.chisq_topk(q24, 9, torch.zeros((4, 24), dtype=torch.uint8, device="cuda"), 9, k=1)   # 24 % 16 != 0
    h16 = torch.zeros((2, 24), dtype=torch.int16, device="cuda").view(torch.uint16)
    assert ops.compact_histograms(h16, 9).dtype == torch.uint16 and ops.compact_histograms(h16, 300).dtype == torch.uint16


@pytest.mark.parametrize("shape,grid", [((100, 100), 8), ((112, 112), 8), ((61, 75), 4), ((130, 98), 8)])
def test_k2_u8_count_output_equals_the_u16_output(shape, grid):
    """frb_lbp_hist_u8_counts8 (the gallery form written straight by K2) == frb_lbp_hist_u8 narrowed; also the
    u16 -> u8 narrowing kernel and the row-id remap kernel that replaced the eager torch ops on the product path."""
    from facerecognition_b200 import ops
    rng = np.random.default_rng(shape[0] * 7 + grid)
    imgs = torch.from_numpy(rng.integers(0, 256, (37,) + shape, dtype=np.uint8)).cuda()
    imgs[3] = 200                                                          # flat image: one bin holds the whole cell
    h16, px = ops.lbp_hist(imgs, grid_x=grid, grid_y=grid)
    h8, px8 = ops.lbp_hist(imgs, grid_x=grid, grid_y=grid, counts8=True)
    assert px == px8
    if px <= 255:
        assert h8.dtype == torch.uint8 and np.array_equal(h8.cpu().numpy(), h16.cpu().numpy().astype(np.uint8))
        assert int(h8.cpu().numpy().astype(np.int64).sum()) == 37 * grid * grid * px
        assert torch.equal(ops.compact_histograms(h16, px), h8)
    else:
        assert h8.dtype == torch.uint16 and torch.equal(h8.view(torch.int16), h16.view(torch.int16))
    idx = torch.tensor([[0, 3, -1], [2, -1, 1]], dtype=torch.int64, device="cuda")
    table = torch.tensor([100, 200, 300, 400], dtype=torch.int64, device="cuda")
    assert ops.index_remap(idx, table).tolist() == [[100, 400, -1], [300, -1, 200]]


def test_mixed_size_gallery_and_web_ui_mapping(oracle_lbph):
    """Two image sizes in one model (two histogram groups, row ids remapped on the device) against the oracle, plus the
    web UI's confidence / Unknown mapping (web_app.py:597,605)."""
    import facerecognition_b200 as F
    rng = np.random.default_rng(91)
    a = [rng.integers(0, 256, (100, 100), dtype=np.uint8) for _ in range(9)]
    b = [rng.integers(0, 256, (112, 112), dtype=np.uint8) for _ in range(7)]
    faces = [a[0], b[0], a[1], b[1]] + a[2:] + b[2:]
    labels = np.arange(len(faces), dtype=np.int32) + 10
    model = F.train_lbph_model(faces, labels)
    ref = oracle_lbph.OracleLBPH()
    ref.train(faces, labels)
    for f in [faces[1], faces[2], faces[-1], rng.integers(0, 256, (112, 112), dtype=np.uint8)]:
        lab, dist = model.predict(f)
        rlab, rdist = ref.predict(f)
        assert lab == rlab and abs(dist - rdist) <= 1e-5 * max(rdist, 1e-30)
    assert F.web_confidence(0.0) == 1.0 and F.web_confidence(50.0) == 0.75 and F.web_confidence(200.0) == 0.0 and F.web_confidence(1e9) == 0.0
    r = F.recognize_face_web(model, faces[2], threshold=80.0, label_map={12: "carol"})
    assert r == {"identity": "carol", "confidence": 1.0, "distance": 0.0, "label": 12}
    far = rng.integers(0, 256, (100, 100), dtype=np.uint8)
    r = F.recognize_face_web(model, far, threshold=1.0)
    assert r["identity"] == "Unknown" and r["confidence"] == F.web_confidence(r["distance"])


@pytest.mark.parametrize("radius,neighbors,shape,grid", [(2, 8, (100, 100), 8), (1, 4, (64, 80), 4), (3, 6, (90, 70), 5),
                                                          (2, 3, (50, 50), 8), (1, 8, (100, 100), 8)])
def test_other_radius_and_neighbor_settings(oracle_lbph, radius, neighbors, shape, grid):
    """The reference exposes radius / neighbors as options (models/lbphmodel/train_lbph_script.py:353-363); settings other
    than its defaults (1, 8) take the general kernels: codes and histograms bit-exact against the oracle's elbp_,
    predict() == the oracle's (label, distance)."""
    import facerecognition_b200 as F
    from facerecognition_b200 import ops
    rng = np.random.default_rng(radius * 100 + neighbors)
    faces = rng.integers(0, 256, (14,) + shape, dtype=np.uint8)
    faces[2] = 77                                                          # flat image
    faces[3, :, ::2] = 255
    want, wpx = oracle_lbph.c_lbp_hist(faces, radius, neighbors, grid, grid)
    dev_faces = torch.from_numpy(faces).cuda()
    got, px = ops.lbp_hist(dev_faces, radius, neighbors, grid, grid)
    assert px == wpx and got.shape == want.shape and np.array_equal(got.cpu().numpy(), want)
    got8, _ = ops.lbp_hist(dev_faces, radius, neighbors, grid, grid, counts8=True)
    assert np.array_equal(got8.cpu().numpy().astype(np.uint16), want)
    codes = ops.lbp_codes(dev_faces[:2].contiguous(), radius, neighbors).cpu().numpy()
    for j in range(2):
        ref = np.zeros((shape[0] - 2 * radius, shape[1] - 2 * radius), np.int32)
        oracle_lbph.lib().frb_oracle_elbp(oracle_lbph._p(np.ascontiguousarray(faces[j])), shape[0], shape[1], radius, neighbors, oracle_lbph._p(ref))
        assert np.array_equal(codes[j], ref)
    if (grid * grid * (1 << neighbors)) % 16 == 0:
        labels = np.arange(14, dtype=np.int32) * 3
        model = F.train_lbph_model(list(faces[:10]), labels[:10], radius=radius, neighbors=neighbors, grid_x=grid, grid_y=grid)
        ref_model = oracle_lbph.OracleLBPH(radius, neighbors, grid, grid)
        ref_model.train(list(faces[:10]), labels[:10])
        for f in faces[8:14]:
            lab, dist = model.predict(f)
            rlab, rdist = ref_model.predict(f)
            assert lab == rlab and abs(dist - rdist) <= 1e-5 * max(rdist, 1e-30), (lab, dist, rlab, rdist)
