"""Shared checks for PsiInit / weight masks / Mul iteration vs the oracle; run against the host emulator (CPU) and the CUDA library (GPU)."""
import numpy as np
import pytest


def _raw_blend(oracle, ds, v):
    mn, mx = ds.boxes[v]
    return oracle.blending_weight(ds.dims_zyx, mn, mx, (0.0,) * 3, (12.0,) * 3)


def check_weights_on_device_match_oracle(lib, oracle, small_dataset):
    import mvrecon_b200 as m
    ds = small_dataset
    dv = m.DeconViews([m.DeconView(ds.images[v], None, ds.psfs[v], m.PSFTYPE.INDEPENDENT) for v in range(3)], library=lib)
    try:
        for v in range(3):
            mn, mx = ds.boxes[v]
            dv.makeBlendingWeights(v, mn, mx, (0.0,) * 3, (12.0,) * 3)
            assert np.array_equal(dv.getWeight(v), _raw_blend(oracle, ds, v))            # bit exact: same float/double promotion order
        dv.normalizeWeights(1.0, False)
        for v in range(3):
            assert np.array_equal(dv.getWeight(v), ds.weights[v])
    finally:
        dv.close()


def check_weight_normalisation_variants(lib, oracle, small_dataset, smooth, osem):
    import mvrecon_b200 as m
    ds = small_dataset
    raw = [_raw_blend(oracle, ds, v) for v in range(3)]
    ref = oracle.normalize_weights(raw, osem, smooth)
    dv = m.DeconViews([m.DeconView(ds.images[v], raw[v], ds.psfs[v]) for v in range(3)], library=lib)
    try:
        dv.normalizeWeights(osem, smooth)
        for v in range(3):
            assert np.array_equal(dv.getWeight(v), ref[v])
    finally:
        dv.close()


def check_psi_init_variants(lib, oracle, small_dataset):
    import mvrecon_b200 as m
    ds = small_dataset
    mk = lambda: m.DeconViews([m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v]) for v in range(3)], library=lib)
    # FUSED_BLURRED
    ref_psi, ref_max, ref_avg = oracle.psi_init_blurred_fused(ds.images, ds.weights, 5.0)
    dv = mk()
    try:
        init = m.PsiInitBlurredFused(5.0)
        dec = m.MultiViewDeconvolutionSeq(dv, 0, init)
        assert np.array_equal(init.getMax(), ref_max)
        assert abs(init.getAvg() - ref_avg) < 1e-9 * ref_avg
        assert oracle.rel_l2(dec.getPSI(), ref_psi) < 1e-6          # FFT Gaussian vs separable direct sum
    finally:
        dv.close()
    # AVG
    _, ref_max, ref_avg = oracle.psi_init_avg_precise(ds.images)
    dv = mk()
    try:
        init = m.PsiInitAvgPrecise()
        dec = m.MultiViewDeconvolutionSeq(dv, 0, init)
        assert np.array_equal(init.getMax(), ref_max) and abs(init.getAvg() - ref_avg) < 1e-9 * ref_avg
        assert np.all(dec.getPSI() == np.float32(ref_avg))
    finally:
        dv.close()
    # APPROX_AVG (getAvg() == -1 like the reference)
    ref_psi, ref_max, ref_avg = oracle.psi_init_avg_approx(ds.images)
    dv = mk()
    try:
        init = m.PsiInitAvgApprox()
        dec = m.MultiViewDeconvolutionSeq(dv, 0, init)
        assert init.getAvg() == -1.0 and np.array_equal(init.getMax(), ref_max)
        assert np.allclose(dec.getPSI(), ref_psi, rtol=1e-6)
    finally:
        dv.close()
    # no view covers the volume -> error like the reference
    z = np.zeros_like(ds.images[0])
    dv = m.DeconViews([m.DeconView(z, ds.weights[0], ds.psfs[0])], library=lib)
    try:
        with pytest.raises(m.MvdError):
            m.MultiViewDeconvolutionSeq(dv, 0, m.PsiInitBlurredFused())
    finally:
        dv.close()


def check_mul_iteration_matches_oracle(lib, oracle, small_dataset, lam):
    import mvrecon_b200 as m
    ds = small_dataset
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.INDEPENDENT)
    dv = m.DeconViews([m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v], m.PSFTYPE.INDEPENDENT) for v in range(3)], lambda_=lam, library=lib)
    try:
        dec = m.MultiViewDeconvolutionMul(dv, 2, m.PsiInitFromRAI(psi0, [v.max_intensity for v in views]))
        psi = psi0
        for it in range(2):
            st = dec.runNextIteration()[0]
            psi, s, mx = oracle.iteration_mul_whole(psi, views, lam, dtype=np.float64)
            assert oracle.rel_l2(dec.getPSI(), psi) < 4e-6
            assert abs(st.sumChange - s) <= 2e-4 * abs(s) + 0.5 and abs(st.maxChange - mx) <= 2e-3 * abs(mx) + 1e-3
    finally:
        dv.close()


def check_affine_blending_weights(lib, oracle):
    """rotated + scaled view box: device weights == oracle restatement of TransformWeight.transformBlending, bit for bit"""
    import mvrecon_b200 as m
    dims = (30, 36, 44)
    th = np.deg2rad(27.0)
    fwd = np.array([[np.cos(th), 0.0, np.sin(th) * 1.7, 9.3], [0.0, 1.0, 0.0, -2.25], [-np.sin(th), 0.0, np.cos(th) * 1.7, 14.1], [0, 0, 0, 1.0]])
    inv = np.linalg.inv(fwd)[:3].ravel()
    img_min, img_max, off = (0, 0, 0), (39, 35, 19), (-3, 1, 2)
    ref = oracle.blending_weight_affine(dims, off, inv, img_min, img_max, (2.0, 1.0, 0.5), (12.0, 10.0, 6.0))
    assert 0 < np.count_nonzero(ref) < ref.size and ref.max() == 1.0
    img = np.ones(dims, np.float32)
    dv = m.DeconViews([m.DeconView(img, None, np.ones((3, 3, 3), np.float32))], library=lib)
    try:
        dv.makeBlendingWeightsAffine(0, img_min, img_max, inv, off, (2.0, 1.0, 0.5), (12.0, 10.0, 6.0))
        got = dv.getWeight(0)
    finally:
        dv.close()
    assert np.array_equal(got, ref)
    # identity transform reduces to the axis-aligned case
    ident = np.eye(4)[:3].ravel()
    a = oracle.blending_weight_affine(dims, (0, 0, 0), ident, (4, 3, 2), (40, 30, 25), (0, 0, 0), (12, 12, 12))
    b = oracle.blending_weight(dims, (4, 3, 2), (40, 30, 25), (0, 0, 0), (12, 12, 12))
    assert np.array_equal(a, b)


# ---- the step before the loop: view materialisation and PSF preparation -----------------------------------------
def _models():
    """two rotated / anisotropically scaled views and their inverses (row-packed 3x4)"""
    out = []
    for th_deg, sz, tr in ((27.0, 1.7, (9.3, -2.25, 14.1)), (-14.0, 2.1, (4.2, 1.5, -3.75))):
        th = np.deg2rad(th_deg)
        fwd = np.array([[np.cos(th), 0.0, np.sin(th) * sz, tr[0]], [0.0, 1.0, 0.0, tr[1]], [-np.sin(th), 0.0, np.cos(th) * sz, tr[2]], [0, 0, 0, 1.0]])
        out.append((fwd[:3].ravel(), np.linalg.inv(fwd)[:3].ravel()))
    return out


def check_fuse_group(lib, oracle, interpolation, with_blending):
    """ProcessInputImages.fuseGroups on the device == oracle restatement, bit for bit: two rotated raw views fused into one virtual view
    (image and summed weight), then used as the input of a view update"""
    import mvrecon_b200 as m
    rng = np.random.default_rng(11)
    dims, bbox_min = (30, 36, 44), (-3, 1, 2)
    raws = [(rng.random(s) * 300).astype(np.float32) for s in ((20, 36, 40), (18, 30, 38))]
    raws[0][5:9, 10:20, 10:30] = 0.25                                   # below minValueImg: clamped to 1 inside the image
    models = _models()
    fb = [((2.0, 1.0, 0.5), (12.0, 10.0, 6.0)), ((1.0, 1.0, 1.0), (8.0, 8.0, 4.0))] if with_blending else None
    db = [((-3.0, -3.0, -1.0), (12.0, 10.0, 6.0)), ((0.0, 0.0, 0.0), (6.0, 6.0, 3.0))] if with_blending else None
    ref_img, ref_w = oracle.fuse_group(raws, [mi for _, mi in models], bbox_min, dims, interpolation, fb, db)
    assert 0 < np.count_nonzero(ref_img) < ref_img.size
    rv = [m.RawView(raws[j], models[j][1], interpolation, None if fb is None else fb[j], None if db is None else db[j]) for j in range(2)]
    psf = oracle.synth_psf(0, 1, (5, 5, 5), (1.0, 1.0, 1.2))
    dv = m.DeconViews([m.DeconView(m.FusedGroup(rv, bbox_min, dims), None, psf)], library=lib)
    try:
        got_img, got_w = dv.getImage(0), dv.getWeight(0)
        assert np.array_equal(got_w, ref_w)
        assert np.array_equal(got_img, ref_img)
        # the materialised view drives the loop like an uploaded one
        dv.normalizeWeights(1.0, False)
        w = oracle.normalize_weights([ref_w])[0]
        psi0 = np.full(dims, 50.0, np.float32)
        dec = m.MultiViewDeconvolutionSeq(dv, 1, m.PsiInitFromRAI(psi0, [float(ref_img.max())]))
        dec.runIterations()
        k1, k2 = oracle.derive_kernels([psf], oracle.INDEPENDENT)
        view = oracle.OracleView(ref_img, w, k1[0], k2[0], float(ref_img.max()))
        want, _, _ = oracle.view_update_whole(psi0, view, 0.0, dtype=np.float64)
        assert oracle.rel_l2(dec.getPSI(), want) < 2e-6
    finally:
        dv.close()


def check_psf_preparation(lib, oracle):
    """PSFPreparation.loadGroupTransformPSFs pieces vs the oracle restatement (host code: bit exact)"""
    rng = np.random.default_rng(12)
    psfs = [rng.random(s).astype(np.float32) for s in ((9, 11, 13), (11, 11, 11))]
    models = _models()
    t = [lib.psf_transform(p, a, ia) for p, (a, ia) in zip(psfs, models)]
    for got, p, (a, ia) in zip(t, psfs, models):
        want = oracle.psf_transform(p, a, ia)
        assert got.shape == want.shape and all(s % 2 == 1 for s in got.shape)
        assert np.array_equal(got, want)
    for use_max in (False, True):
        assert np.array_equal(lib.psf_average(t, use_max), oracle.psf_average(t, use_max))
    avg = lib.psf_average(t)
    assert np.array_equal(lib.psf_make_same_size(avg, (41, 15, 33)), oracle.psf_make_same_size(avg, (41, 15, 33)))
    assert np.array_equal(lib.psf_make_same_size(avg, (5, 7, 3)), oracle.psf_make_same_size(avg, (5, 7, 3)))
    import mvrecon_b200 as m
    groups = [[(psfs[0], *models[0]), (psfs[1], *models[1])], [(psfs[1], *models[0])]]
    got = m.loadGroupTransformPSFs(groups, True, lib)
    a = oracle.psf_average([oracle.psf_transform(p, f, i) for p, f, i in groups[0]])
    b = oracle.psf_average([oracle.psf_transform(p, f, i) for p, f, i in groups[1]])
    size = tuple(max(a.shape[d], b.shape[d]) for d in range(3))
    assert np.array_equal(got[0], oracle.psf_make_same_size(a, size)) and np.array_equal(got[1], oracle.psf_make_same_size(b, size))


# ---- TIFF stacks at the boundary: PsiInitFromFile / result export ----------------------------------------------------------------------
def check_tiff_io_and_psi_init_from_file(lib, oracle, small_dataset, tmp_path):
    import mvrecon_b200 as m
    from PIL import Image
    ds = small_dataset
    rng = np.random.default_rng(21)
    vol = (rng.random(ds.dims_zyx) * 500).astype(np.float32)
    # writer -> independent reader (PIL)
    p1 = str(tmp_path / "ours.tif")
    lib.tiff_write(p1, vol)
    with Image.open(p1) as im:
        assert im.n_frames == vol.shape[0] and im.mode == "F"
        for z in (0, vol.shape[0] // 2, vol.shape[0] - 1):
            im.seek(z)
            assert np.array_equal(np.asarray(im, dtype=np.float32), vol[z])
    assert np.array_equal(lib.tiff_read(p1), vol)
    # independent writer (PIL) -> reader: float, 16-bit and 8-bit stacks are opened "as 32 bit"
    for dtype, mode in ((np.float32, "F"), (np.uint16, "I;16"), (np.uint8, "L")):
        src = vol.astype(dtype)
        frames = [Image.fromarray(src[z]) for z in range(src.shape[0])]
        assert frames[0].mode == mode
        p2 = str(tmp_path / f"pil_{mode.replace(';', '')}.tif")
        frames[0].save(p2, save_all=True, append_images=frames[1:], compression=None)
        got = lib.tiff_read(p2)
        assert got.dtype == np.float32 and np.array_equal(got, src.astype(np.float32))
    # big-endian, hand-built single-slice file
    be = bytearray(b"MM\x00\x2a\x00\x00\x00\x08")
    w, h = 3, 2
    entries = [(256, 3, 1, w << 16), (257, 3, 1, h << 16), (258, 3, 1, 16 << 16), (259, 3, 1, 1 << 16), (273, 4, 1, 8 + 2 + 12 * 7 + 4),
               (277, 3, 1, 1 << 16), (279, 4, 1, w * h * 2)]
    be += len(entries).to_bytes(2, "big")
    for tag, typ, cnt, val in entries:
        be += tag.to_bytes(2, "big") + typ.to_bytes(2, "big") + cnt.to_bytes(4, "big") + val.to_bytes(4, "big")
    be += (0).to_bytes(4, "big") + b"".join(int(v).to_bytes(2, "big") for v in (1, 2, 3, 40000, 5, 6))
    p3 = str(tmp_path / "be.tif")
    open(p3, "wb").write(bytes(be))
    assert np.array_equal(lib.tiff_read(p3), np.array([[[1, 2, 3], [40000, 5, 6]]], np.float32))
    # errors are messages, not crashes
    open(tmp_path / "junk.tif", "wb").write(b"not a tiff at all")
    with pytest.raises(m.MvdError, match="TIFF"):
        lib.tiff_read(str(tmp_path / "junk.tif"))
    # PsiInitFromFile: psi = file, statistics from the views without touching psi
    dv = m.DeconViews([m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v]) for v in range(3)], library=lib)
    try:
        for precise in (True, False):
            init = m.PsiInitFromFile(p1, precise)
            dec = m.MultiViewDeconvolutionSeq(dv, 0, init)
            assert dec.initWasSuccessful()
            assert np.array_equal(dec.getPSI(), vol)
            if precise:
                _, ref_max, ref_avg = oracle.psi_init_avg_precise(ds.images, set_img_to_avg=False, psi=vol.copy())
                assert abs(init.getAvg() - ref_avg) < 1e-9 * ref_avg
            else:
                _, ref_max, _ = oracle.psi_init_avg_approx(ds.images, set_img_to_avg=False, psi=vol.copy())
                assert init.getAvg() == -1
            assert np.array_equal(init.getMax(), ref_max)
        bad = m.PsiInitFromFile(p3, True)                       # wrong dimensions: the reference returns false
        assert not m.MultiViewDeconvolutionSeq(dv, 0, bad).initWasSuccessful() and "dimensions" in bad.error
        assert not m.MultiViewDeconvolutionSeq(dv, 0, m.PsiInitFromFile(str(tmp_path / "missing.tif"), True)).initWasSuccessful()
    finally:
        dv.close()


def check_skip_empty_tiles(lib, oracle):
    """DeconView.filterBlocksForContent on the resident path: tiles in which a view has no weight are not computed; the result and the
    statistics are those of the unfiltered run (a zero weight leaves psi untouched, DeconvolutionMethods.java:356)."""
    import mvrecon_b200 as m
    ds = oracle.make_synthetic((64, 36, 40), 2, seed=5, psf_size_xyz=(5, 5, 7), psf_sigma_xyz=(1.0, 1.0, 1.6), bead_density=1024)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    mx = [v.max_intensity for v in views]
    weights = [w.copy() for w in ds.weights]
    weights[0][20:] = 0.0                               # view 0 contributes nothing to the upper part of the volume
    results = []
    for on in (False, True):
        dv = m.DeconViews([m.DeconView(ds.images[v], weights[v], ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(2)],
                          max_fft_len=48, library=lib)
        try:
            assert dv.tile_info()["num_tiles"] >= 2
            dec = m.MultiViewDeconvolutionSeq(dv, 2, m.PsiInitFromRAI(psi0, mx))
            skipped = dv.filterBlocksForContent(on)
            assert (skipped >= 1) if on else (skipped == 0)
            dec.runIterations()
            results.append((dec.getPSI(), [(s.sumChange, s.maxChange) for it in dec.stats for s in it], skipped))
            if on:                                       # a change of the weights re-evaluates the filter
                dv.lib.check(dv.lib.dll.mvd_set_view(dv._ctx, 0, m._fp(ds.images[0]), m._fp(ds.weights[0])))
                assert dv.filterBlocksForContent(True) == 0
        finally:
            dv.close()
    assert np.array_equal(results[0][0], results[1][0])
    for a, b in zip(results[0][1], results[1][1]):
        assert abs(a[0] - b[0]) <= 1e-9 * max(1.0, abs(a[0])) and a[1] == b[1]
    ov = [oracle.OracleView(ds.images[v], weights[v], views[v].kernel1, views[v].kernel2, mx[v]) for v in range(2)]
    ref, _ = oracle.run_iterations_seq(psi0, ov, 2, 0.0, dtype=np.float64)
    assert oracle.rel_l2(results[1][0], ref) < 4e-6


def check_filter_blocks_mirror():
    """Python mirror of DeconView.filterBlocksForContent / blockContainsContent (DeconView.java:204-274) used by the L2 block driver."""
    import mvrecon_b200 as m
    w = np.zeros((40, 36, 32), dtype=np.float32)
    w[:10, :, :] = 1.0
    blocks = m.divideIntoBlocks((32, 36, 40), (32, 32, 32), (9, 9, 9))
    batches = m.sortBlocksBySmallestFootprint(blocks, (32, 36, 40))
    n_before = sum(len(b) for b in batches)
    expect_keep = sum(1 for b in blocks if b.offset[2] < 10 and b.offset[2] + b.blockSize[2] > 0)
    removed, removed_batches = m.filterBlocksForContent(batches, w)
    assert removed == n_before - expect_keep and removed > 0
    assert sum(len(b) for b in batches) == expect_keep and all(len(b) > 0 for b in batches)
    assert all(m.blockContainsContent(b, w) for batch in batches for b in batch)


# ---- N5 datasets (PointSpreadFunction.load / save, N5 export) -----------------------------------------------------------------------
def _n5_write_reference(path, vol, block_xyz, dtype=">f4", compression="gzip"):
    """An independent writer of the published N5 file-system format (numpy + gzip + struct), used to check the library's reader."""
    import gzip, json, os, struct
    nz, ny, nx = vol.shape
    os.makedirs(path, exist_ok=True)
    name = {">f4": "float32", ">u2": "uint16", ">u1": "uint8", ">f8": "float64", ">i2": "int16"}[dtype]
    comp = {"type": "gzip", "useZlib": False, "level": 1} if compression == "gzip" else {"type": "raw"}
    with open(os.path.join(path, "attributes.json"), "w") as f:
        json.dump({"dimensions": [nx, ny, nz], "blockSize": list(block_xyz), "dataType": name, "compression": comp}, f)
    bx, by, bz = block_xyz
    for gz in range(-(-nz // bz)):
        for gy in range(-(-ny // by)):
            for gx in range(-(-nx // bx)):
                blk = vol[gz * bz:(gz + 1) * bz, gy * by:(gy + 1) * by, gx * bx:(gx + 1) * bx]
                if gz == 1 and gy == 0 and gx == 0 and not blk.any():
                    continue                                        # a missing block reads as zeros
                payload = np.ascontiguousarray(blk).astype(dtype).tobytes()        # x fastest, big endian
                if compression == "gzip":
                    payload = gzip.compress(payload, 1)
                hdr = struct.pack(">HHIII", 0, 3, blk.shape[2], blk.shape[1], blk.shape[0])
                d = os.path.join(path, str(gx), str(gy))
                os.makedirs(d, exist_ok=True)
                with open(os.path.join(d, str(gz)), "wb") as f:
                    f.write(hdr + payload)


def _n5_read_reference(path):
    import gzip, json, os, struct
    a = json.load(open(os.path.join(path, "attributes.json")))
    nx, ny, nz = a["dimensions"]
    bx, by, bz = a["blockSize"]
    assert a["dataType"] == "float32"
    out = np.zeros((nz, ny, nx), dtype=np.float32)
    for gz in range(-(-nz // bz)):
        for gy in range(-(-ny // by)):
            for gx in range(-(-nx // bx)):
                fn = os.path.join(path, str(gx), str(gy), str(gz))
                if not os.path.exists(fn):
                    continue
                raw = open(fn, "rb").read()
                mode, nd, sx, sy, sz = struct.unpack(">HHIII", raw[:16])
                assert mode == 0 and nd == 3
                payload = raw[16:]
                if a["compression"]["type"] == "gzip":
                    assert payload[:2] == b"\x1f\x8b"               # a gzip member, as N5's GzipCompression writes
                    payload = gzip.decompress(payload)
                blk = np.frombuffer(payload, dtype=">f4").reshape(sz, sy, sx)
                out[gz * bz:gz * bz + sz, gy * by:gy * by + sy, gx * bx:gx * bx + sx] = blk
    return out, a


def check_n5_io(lib, tmp_path):
    """mvd_n5_read / mvd_n5_write against an independent implementation of the N5 file-system format: several blocks with truncated edge
    blocks, raw and gzip, integer and floating element types, a missing block; and the PSF round trip of PointSpreadFunction.save / load."""
    import os
    rng = np.random.default_rng(11)
    vol = rng.random((21, 19, 37), dtype=np.float32) * 1000 - 200
    vol[16:, :8, :16] = 0.0                                          # block (0, 0, 1) of the 16 x 8 x 16 grid is empty -> not written
    # reader: files written by the independent writer
    for k, (dtype, comp) in enumerate([(">f4", "gzip"), (">f4", "raw"), (">u2", "gzip"), (">f8", "raw"), (">i2", "gzip"), (">u1", "raw")]):
        d = os.path.join(str(tmp_path), f"ds{k}")
        src = vol if dtype in (">f4", ">f8") else np.clip(np.round(vol), 0 if dtype != ">i2" else -200, 255 if dtype == ">u1" else 800)
        _n5_write_reference(d, src, (16, 8, 16), dtype, comp)
        got = lib.n5_read(d)
        assert got.shape == vol.shape
        assert np.array_equal(got, src.astype(dtype).astype(np.float32))
    # writer: read back by the independent reader and by the library
    for level in (1, -1, 6):
        d = os.path.join(str(tmp_path), f"out{level}")
        lib.n5_write(d, vol, (16, 8, 16), level)
        back, attrs = _n5_read_reference(d)
        assert np.array_equal(back, vol) and attrs["dimensions"] == [37, 19, 21] and attrs["blockSize"] == [16, 8, 16]
        assert attrs["compression"]["type"] == ("raw" if level < 0 else "gzip")
        assert np.array_equal(lib.n5_read(d), vol)
    # PointSpreadFunction.save -> load: one 128^3 gzip-1 block for a PSF
    psf = rng.random((25, 19, 25), dtype=np.float32)
    d = os.path.join(str(tmp_path), "psf.n5", "psf_t0_v3")
    lib.n5_write(d, psf)
    assert np.array_equal(lib.n5_read(d), psf) and os.path.exists(os.path.join(d, "0", "0", "0"))
    with pytest.raises(Exception):
        lib.n5_read(os.path.join(str(tmp_path), "nothing_here"))


def check_debug_interval(lib, oracle, small_dataset):
    """MultiViewDeconvolution.setDebug / setDebugInterval (MultiViewDeconvolution.java:119-122,153-191): psi is copied out before the
    iterations it with (it - 1) % debugInterval == 0 (Java remainder), and the run itself is unchanged."""
    import mvrecon_b200 as m
    ds = small_dataset
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    mx = [v.max_intensity for v in views]
    res = []
    for interval in (None, 1, 2):
        dv = m.DeconViews([m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(3)], library=lib)
        try:
            dec = m.MultiViewDeconvolutionSeq(dv, 4, m.PsiInitFromRAI(psi0, mx))
            if interval is not None:
                dec.setDebug(True)
                dec.setDebugInterval(interval)
            dec.runIterations()
            res.append((dec.getPSI(), [it for it, _ in dec.getDebugImage()], [p for _, p in dec.getDebugImage()]))
        finally:
            dv.close()
    assert res[0][1] == [] and res[1][1] == [0, 1, 2, 3] and res[2][1] == [1, 3]
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][0], res[2][0])
    assert np.array_equal(res[1][2][0], psi0) and np.array_equal(res[1][2][1], res[2][2][0])


def check_zarr_export(lib, tmp_path):
    """mvd_zarr_write: an OME-Zarr 0.4 group read back with nothing but json / gzip / numpy (Zarr v2 layout: full-size C-order
    little-endian chunks, edge chunks padded with the fill value, "/" as dimension separator)."""
    import gzip, json, os
    rng = np.random.default_rng(13)
    vol = rng.random((21, 19, 37), dtype=np.float32) * 50
    for level in (1, -1):
        root = os.path.join(str(tmp_path), f"out{level}.zarr")
        lib.zarr_write(root, vol, (16, 8, 16), level, voxel_size_xyz=(0.5, 0.5, 2.0))
        assert json.load(open(os.path.join(root, ".zgroup"))) == {"zarr_format": 2}
        ms = json.load(open(os.path.join(root, ".zattrs")))["multiscales"][0]
        assert ms["version"] == "0.4" and [a["name"] for a in ms["axes"]] == ["z", "y", "x"] and ms["datasets"][0]["path"] == "0"
        assert ms["datasets"][0]["coordinateTransformations"][0] == {"type": "scale", "scale": [2.0, 0.5, 0.5]}
        za = json.load(open(os.path.join(root, "0", ".zarray")))
        assert za["shape"] == [21, 19, 37] and za["chunks"] == [16, 8, 16] and za["dtype"] == "<f4" and za["order"] == "C"
        assert za["dimension_separator"] == "/" and za["fill_value"] == 0 and za["zarr_format"] == 2
        assert za["compressor"] == (None if level < 0 else {"id": "gzip", "level": level})
        back = np.zeros((32, 24, 48), dtype=np.float32)
        for iz in range(2):
            for iy in range(3):
                for ix in range(3):
                    raw = open(os.path.join(root, "0", str(iz), str(iy), str(ix)), "rb").read()
                    if level >= 0:
                        raw = gzip.decompress(raw)
                    back[iz * 16:(iz + 1) * 16, iy * 8:(iy + 1) * 8, ix * 16:(ix + 1) * 16] = np.frombuffer(raw, dtype="<f4").reshape(16, 8, 16)
        assert np.array_equal(back[:21, :19, :37], vol) and not back[21:].any() and not back[:, 19:].any() and not back[:, :, 37:].any()
