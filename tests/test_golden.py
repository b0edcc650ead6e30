"""Golden vectors (tests/golden/small_case.npz, produced by the float64 oracle -- see make_golden.py): the float32 oracle and the
kernel bodies (host emulation here, CUDA in test_gpu_parity.py) must reproduce them within the stated tolerance."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_case.npz")
# tolerance of BASELINE.md section 6: relL2 <= max(4*eps_it, 2e-6*it), max-abs <= 1e-3 * max(psi)
REL_TOL = lambda it: max(4 * 1.5e-7, 2e-6 * (it + 1))


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _views(o, g, ptype):
    k = [(g[f"k1_t{ptype}_v{v}"], g[f"k2_t{ptype}_v{v}"]) for v in range(3)]
    return [o.OracleView(g[f"img{v}"], g[f"weight{v}"], k[v][0], k[v][1], float(g["max"][v])) for v in range(3)]


def test_oracle_float32_reproduces_golden(oracle, gold):
    views = _views(oracle, gold, 2)
    psi = gold["psi0"]
    lam = float(gold["lambda"])
    k = 0
    for it in range(5):
        for v in range(3):
            psi, s, m = oracle.view_update_whole(psi, views[v], lam, dtype=np.float32)
            if it < 2:
                ref = gold[f"psi_it{it}_v{v}"]
                assert oracle.rel_l2(psi, ref) <= REL_TOL(it)
                assert np.abs(psi - ref).max() <= 1e-3 * ref.max()
            it_, v_, s_ref, m_ref = gold["stats"][k]
            assert abs(s - s_ref) <= 1e-4 * abs(s_ref) + 0.5 and abs(m - m_ref) <= 1e-3 * abs(m_ref) + 1e-3
            k += 1
    assert oracle.rel_l2(psi, gold["psi_it4"]) <= REL_TOL(4)


def test_kernel_derivation_reproduces_golden(oracle, gold):
    psfs = [gold[f"psf{v}"] for v in range(3)]
    for ptype in range(4):
        k1, k2 = oracle.derive_kernels(psfs, ptype, quirk_threads=int(gold["quirk_threads"]))
        for v in range(3):
            assert oracle.rel_l2(k1[v], gold[f"k1_t{ptype}_v{v}"]) < 1e-6
            assert oracle.rel_l2(k2[v], gold[f"k2_t{ptype}_v{v}"]) < 2e-6


def test_hostemu_reproduces_golden(hostemu_lib, oracle, gold):
    import mvrecon_b200 as m
    views = [m.DeconView(gold[f"img{v}"], gold[f"weight{v}"], gold[f"psf{v}"], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(3)]
    dv = m.DeconViews(views, lambda_=float(gold["lambda"]), norm_quirk_threads=int(gold["quirk_threads"]), library=hostemu_lib)
    try:
        dec = m.MultiViewDeconvolutionSeq(dv, 5, m.PsiInitFromRAI(gold["psi0"], gold["max"]))
        k = 0
        for it in range(5):
            for v in range(3):
                st = (m.C.c_double * 2)()
                assert dv.lib.dll.mvd_run_view_update(dv._ctx, v, st) == 0
                if it < 2:
                    assert oracle.rel_l2(dec.getPSI(), gold[f"psi_it{it}_v{v}"]) <= REL_TOL(it)
                _, _, s_ref, m_ref = gold["stats"][k]
                assert abs(st[0] - s_ref) <= 1e-4 * abs(s_ref) + 0.5 and abs(st[1] - m_ref) <= 1e-3 * abs(m_ref) + 1e-3
                k += 1
        assert oracle.rel_l2(dec.getPSI(), gold["psi_it4"]) <= REL_TOL(4)
    finally:
        dv.close()


# ---- the step before the loop (tests/golden/prep_case.npz, made by tests/golden/make_golden_prep.py) --------------------------------
@pytest.fixture(scope="module")
def prep():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prep_case.npz"))


def _prep_inputs(g):
    dims, bb = tuple(int(x) for x in g["dims_zyx"]), tuple(int(x) for x in g["bbox_min_xyz"])
    fb = [(tuple(g["fusion_blending"][j][0]), tuple(g["fusion_blending"][j][1])) for j in range(2)]
    db = [(tuple(g["decon_blending"][j][0]), tuple(g["decon_blending"][j][1])) for j in range(2)]
    return dims, bb, fb, db


def test_oracle_reproduces_prep_golden(oracle, prep):
    g = prep
    dims, bb, fb, db = _prep_inputs(g)
    for j in range(2):
        assert np.array_equal(oracle.transform_view(g[f"raw{j}"], g[f"inv_affine{j}"], bb, dims, 1), g[f"view_linear{j}"])
        assert np.array_equal(oracle.transform_view(g[f"raw{j}"], g[f"inv_affine{j}"], bb, dims, 0), g[f"view_nearest{j}"])
        assert np.array_equal(oracle.psf_transform(g[f"psf{j}"], g[f"affine{j}"], g[f"inv_affine{j}"]), g[f"psf_t{j}"])
    img, w = oracle.fuse_group([g["raw0"], g["raw1"]], [g["inv_affine0"], g["inv_affine1"]], bb, dims, 1, fb, db)
    assert np.array_equal(img, g["fused_img"]) and np.array_equal(w, g["fused_weight"])
    assert np.array_equal(oracle.psf_average([g["psf_t0"], g["psf_t1"]]), g["psf_avg"])
    assert np.array_equal(oracle.psf_make_same_size(g["psf_avg"], (25, 13, 17)), g["psf_same"])
    raw_w = [g["blend0"], g["blend1"]]
    for tag, osem, smooth in (("hard", 1.0, False), ("smooth", 1.0, True), ("osem", 2.0, False)):
        for j, nw in enumerate(oracle.normalize_weights(raw_w, osem, smooth)):
            assert np.array_equal(nw, g[f"norm_{tag}{j}"])


def test_hostemu_reproduces_prep_golden(hostemu_lib, oracle, prep):
    import mvrecon_b200 as m
    g = prep
    dims, bb, fb, db = _prep_inputs(g)
    rv = [m.RawView(g[f"raw{j}"], g[f"inv_affine{j}"], 1, fb[j], db[j]) for j in range(2)]
    dv = m.DeconViews([m.DeconView(m.FusedGroup(rv, bb, dims), None, np.ones((3, 3, 3), np.float32))], library=hostemu_lib)
    try:
        assert np.array_equal(dv.getImage(0), g["fused_img"]) and np.array_equal(dv.getWeight(0), g["fused_weight"])
    finally:
        dv.close()
    t = [hostemu_lib.psf_transform(g[f"psf{j}"], g[f"affine{j}"], g[f"inv_affine{j}"]) for j in range(2)]
    assert np.array_equal(t[0], g["psf_t0"]) and np.array_equal(t[1], g["psf_t1"])
    assert np.array_equal(hostemu_lib.psf_average(t), g["psf_avg"])
    assert np.array_equal(hostemu_lib.psf_make_same_size(g["psf_avg"], (25, 13, 17)), g["psf_same"])
    dv = m.DeconViews([m.DeconView(np.ones(dims, np.float32), g[f"blend{j}"], np.ones((3, 3, 3), np.float32)) for j in range(2)], library=hostemu_lib)
    try:
        dv.normalizeWeights(1.0, True)
        for j in range(2):
            assert np.array_equal(dv.getWeight(j), g[f"norm_smooth{j}"])
    finally:
        dv.close()


# ---- BASELINE config c1 in full (tests/golden/c1_case.npz, made by tests/golden/make_golden_c1.py) ---------------------------------
def test_oracle_float32_reproduces_c1_golden_first_iteration(oracle):
    """the float32 oracle (the CPU baseline bench.py times) against the float64 golden of config c1: iteration 1 on the full
    256 x 256 x 128 volume with the 25 x 19 x 25 PSFs; pins the seeded inputs, the kernels' (quirky) sums and the statistics"""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_c1", os.path.join(here, "golden", "make_golden_c1.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = np.load(os.path.join(here, "golden", "c1_case.npz"))
    ds, views, psi0, avg = gen.inputs()
    assert np.allclose([float(im.sum(dtype=np.float64)) for im in ds.images], g["img_sums"], rtol=1e-12)
    assert np.allclose([float(w.sum(dtype=np.float64)) for w in ds.weights], g["weight_sums"], rtol=1e-12)
    assert np.allclose([float(v.kernel1.sum(dtype=np.float64)) for v in views], g["k1_sums"], atol=1e-7)
    assert np.allclose([float(v.kernel2.sum(dtype=np.float64)) for v in views], g["k2_sums"], atol=1e-6)
    assert abs(g["k1_sums"][0] - 1.0) > 5e-3                        # the reference's kernels do not sum to 1 (AdjustInput.java:115-119): view 0 is off by 0.77 %
    assert abs(avg - float(g["avg"])) <= 1e-9 * abs(avg) and np.array_equal(np.array([v.max_intensity for v in views], np.float32), g["max"])
    psi = psi0
    for v in range(4):
        psi, s, m = oracle.view_update_whole(psi, views[v], 0.0, dtype=np.float32)
        _, _, s_ref, m_ref = g["stats"][v]
        assert abs(s - s_ref) <= 1e-4 * abs(s_ref) + 0.5 and abs(m - m_ref) <= 1e-3 * abs(m_ref) + 1e-3
    step = int(g["lattice_step"])
    assert oracle.rel_l2(psi[::step, ::step, ::step], g["psi_it1"]) <= REL_TOL(0)
