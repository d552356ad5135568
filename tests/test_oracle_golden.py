"""Pins the CPU oracle against everything the reference's own tests hold for this path:
GGX positivity proptests + the saved regression case + the fixed direction pairs, the Fresnel inputs, the
2-D importance-sampling integral, plus structural properties of the restated math."""
import ctypes as ct
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F = ct.c_float


def _v(a):
    return (F * 3)(*[float(x) for x in a])


@pytest.fixture(scope="module")
def lib(oracle):
    oracle.rpto_ggx_bsdf.argtypes = [F, F, F, F, ct.c_int, ct.c_void_p, ct.c_void_p, ct.POINTER(F), ct.POINTER(F)]
    oracle.rpto_ggx_generate_and_evaluate.argtypes = [F, F, F, F, ct.c_int, F, F, ct.c_void_p, ct.c_void_p, ct.POINTER(F), ct.POINTER(F)]
    oracle.rpto_fresnel_dielectric.argtypes = [F, F, F]
    oracle.rpto_fresnel_dielectric.restype = F
    oracle.rpto_fresnel_conductor.argtypes = [F, F, F, F]
    oracle.rpto_fresnel_conductor.restype = F
    oracle.rpto_cdf_sample.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_uint32, F, F, ct.c_int, F, F, ct.POINTER(F), ct.POINTER(F)]
    oracle.rpto_linear_curve_eval.argtypes = [ct.c_void_p, ct.c_uint32, F, F, ct.c_int, F]
    oracle.rpto_linear_curve_eval.restype = F
    oracle.rpto_frame_roundtrip.argtypes = [ct.c_void_p] * 4
    oracle.rpto_philox.argtypes = [ct.c_uint64, ct.c_uint32, ct.c_uint32, ct.c_uint32, ct.c_void_p]
    return oracle


def glass_eta(lam):
    """ggx_glass of the reference's tests: eta = cauchy(1.5, 10000), eta_o = cie_e(1), kappa = void (ggx.rs:630-635)."""
    return np.float32(1.5) + np.float32(10000.0) / (np.float32(lam) * np.float32(lam))


def bsdf(lib, alpha, lam, wi, wo):
    f, p = F(), F()
    lib.rpto_ggx_bsdf(alpha, float(glass_eta(lam)), 1.0, 0.0, 0, _v(wi), _v(wo), ct.byref(f), ct.byref(p))
    return f.value, p.value


def generate(lib, alpha, lam, s, wi):
    wo = (F * 3)()
    f, p = F(), F()
    lib.rpto_ggx_generate_and_evaluate(alpha, float(glass_eta(lam)), 1.0, 0.0, 0, s[0], s[1], _v(wi), wo, ct.byref(f), ct.byref(p))
    return np.array(list(wo), dtype=np.float32), f.value, p.value


def test_ggx_saved_proptest_regression(lib):
    """proptest-regressions/materials/ggx.txt:7 — the case proptest saved because it FAILED test_ggx
    (ggx.rs:637-685; the file's neighbours are marked "TODO: debug this failure case"). With roughness 8.7 the
    sampled microfacet normal is almost horizontal, the "reflection" lands in the opposite hemisphere, and the
    swapped evaluation hits total internal reflection: F = 1 so the transmitted f is exactly 0. The oracle must
    reproduce that arithmetic: everything finite, the sampled order positive, the swapped order exactly 0."""
    wi = [0.54826164, 0.0, -0.83630687]
    wo, fg, pg = generate(lib, 8.736748, 400.0, (0.0, 0.0), wi)
    assert np.isfinite(wo).all() and fg > 0 and pg > 0
    assert wo[2] > 0 > wi[2]
    f1, p1 = bsdf(lib, 8.736748, 400.0, wi, wo)
    f2, p2 = bsdf(lib, 8.736748, 400.0, wo, wi)
    assert f1 > 0 and p1 >= 0 and f2 == 0.0 and p2 >= 0


def test_ggx_positivity_proptest(lib):
    """test_ggx (ggx.rs:637-685), 1000 seeded cases: roughness in (0, inf) like props.rs:10-14, unit wi,
    lambda in [400, 800): generate returns a direction; f > 0 and pdf >= 0 in both argument orders."""
    rng = np.random.default_rng(1234)
    positive = 0
    for _ in range(1000):
        alpha = float(np.exp(rng.uniform(np.log(1e-3), np.log(50.0))))
        wi = rng.normal(size=3)
        wi /= np.linalg.norm(wi)
        lam = float(rng.uniform(400, 800))
        s = (float(rng.uniform(0, 1)), float(rng.uniform(0, 1)))
        wo, _, _ = generate(lib, alpha, lam, s, wi)
        assert np.isfinite(wo).all()
        f1, p1 = bsdf(lib, alpha, lam, wi, wo)
        f2, p2 = bsdf(lib, alpha, lam, wo, wi)
        # the swapped order may be exactly 0 under total internal reflection (see the saved regression above)
        assert f1 >= 0 and p1 >= 0 and f2 >= 0 and p2 >= 0, (alpha, wi, lam, s, wo, f1, p1, f2, p2)
        assert np.isfinite([f1, p1, f2, p2]).all()
        positive += f1 > 0 and f2 > 0
    assert positive >= 900, positive


def test_ggx2_arbitrary_pairs_nonnegative(lib):
    """test_ggx2 (ggx.rs:687-755): arbitrary wi, wo: f >= 0, pdf >= 0 both orders."""
    rng = np.random.default_rng(99)
    for _ in range(1000):
        alpha = float(np.exp(rng.uniform(np.log(1e-3), np.log(50.0))))
        wi, wo = rng.normal(size=3), rng.normal(size=3)
        wi /= np.linalg.norm(wi)
        wo /= np.linalg.norm(wo)
        lam = float(rng.uniform(400, 800))
        for a, b in ((wi, wo), (wo, wi)):
            f, p = bsdf(lib, alpha, lam, a, b)
            assert f >= 0 and p >= 0


def test_ggx_fixed_vectors_match_committed_golden(lib):
    """The reference's fixed direction pairs (ggx.rs:825-826, 888-890, 905-906, 923-924, 933-935) print but do
    not assert; their oracle values are committed (tests/golden/ggx_fixed.json, made by tools/make_golden.py)
    so a change to the restatement is caught. Sign properties are asserted as in test_ggx_functions."""
    cases = json.load(open(os.path.join(GOLDEN, "ggx_fixed.json")))
    for c in cases:
        f, p = bsdf(lib, c["alpha"], c["lambda"], c["wi"], c["wo"])
        assert f >= 0 and p >= 0
        assert np.isclose(f, c["f"], rtol=1e-5, atol=1e-30) and np.isclose(p, c["pdf"], rtol=1e-5, atol=1e-30), c


def test_fresnel_inputs(lib):
    """test_fresnel (ggx.rs:613-628): eta_o 1.004 / eta 1.45 at cos = +-0.76048267 and +-0.00871551."""
    for c in (0.76048267, 0.00871551):
        a = lib.rpto_fresnel_dielectric(1.004, 1.45, -c)
        b = lib.rpto_fresnel_dielectric(1.004, 1.45, c)
        assert 0.0 <= a <= 1.0 and 0.0 <= b <= 1.0
    # normal incidence closed form ((n1-n2)/(n1+n2))^2 and grazing -> 1
    assert np.isclose(lib.rpto_fresnel_dielectric(1.0, 1.5, 1.0), 0.04, rtol=1e-5)
    assert lib.rpto_fresnel_dielectric(1.0, 1.5, 1e-6) > 0.999
    # conductor with k = 0 reduces to the dielectric formula
    for c in (0.2, 0.7, 1.0):
        assert np.isclose(lib.rpto_fresnel_conductor(1.0, 1.5, 0.0, c), lib.rpto_fresnel_dielectric(1.0, 1.5, c), rtol=2e-4)


def test_2d_importance_sampling_integral(lib):
    """test_2d_importance_sampling (world/importance_map.rs:798-942): MC estimate of the integral of
    exp(-x^2-y^2) over [-2,2]^2 through CurveWithCDF sampling = 3.11227031972 within 1e-3 (100x100 map,
    65 536 samples). This is the one numeric pin the reference holds on Curve::to_cdf / sample_power_and_pdf."""
    res = 100
    xs = (np.arange(res, dtype=np.float32) / np.float32(res))
    tr = lambda x: np.float32(4.0) * (x - np.float32(0.5))
    rows_pdf, rows_cdf, integrals = [], [], []
    for yi in range(res):
        sig = np.exp(-(tr(xs) ** 2 + tr(xs[yi]) ** 2)).astype(np.float32)  # Curve::from_function, Linear mode
        cdf = np.cumsum(sig.astype(np.float64) * (1.0 / res))
        integrals.append(cdf[-1])
        rows_pdf.append(sig)
        rows_cdf.append((cdf / cdf[-1]).astype(np.float32))
    marg_pdf = np.asarray(integrals, dtype=np.float32)
    mc = np.cumsum(marg_pdf.astype(np.float64) * (1.0 / res))
    marg_int = float(mc[-1])
    marg_cdf = (mc / mc[-1]).astype(np.float32)
    rng = np.random.default_rng(7)
    n = 256 * 256
    est = 0.0
    u, pu, v, pv = F(), F(), F(), F()
    vp = lambda a: a.ctypes.data_as(ct.c_void_p)
    for s in rng.uniform(size=(n, 2)).astype(np.float32):
        lib.rpto_cdf_sample(vp(marg_pdf), vp(marg_cdf), res, 0.0, 1.0, 0, marg_int, float(s[1]), ct.byref(u), ct.byref(pu))
        row = min(int(u.value * res), res - 1)
        lib.rpto_cdf_sample(vp(rows_pdf[row]), vp(rows_cdf[row]), res, 0.0, 1.0, 0, float(integrals[row]), float(s[0]), ct.byref(v), ct.byref(pv))
        assert 0.0 <= u.value < 1.0 and 0.0 <= v.value < 1.0
        pdf = pu.value * pv.value / 16.0
        est += float(np.exp(-(tr(np.float32(u.value)) ** 2 + tr(np.float32(v.value)) ** 2))) / pdf / n
    assert abs(est - 3.11227031972) / 3.11227031972 < 1e-3, est


def test_tangent_frame_is_orthonormal_and_invertible(lib):
    rng = np.random.default_rng(3)
    for _ in range(200):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        v = rng.normal(size=3)
        loc, back = (F * 3)(), (F * 3)()
        lib.rpto_frame_roundtrip(_v(n), _v(v), loc, back)
        assert np.allclose(list(back), v, atol=1e-5)
        assert np.isclose(loc[2], float(np.dot(n.astype(np.float32), v.astype(np.float32))), atol=1e-5)
        assert np.isclose(np.linalg.norm(list(loc)), np.linalg.norm(v), rtol=1e-5)


def test_philox_known_answers_and_uniformity(lib):
    out = (F * 4)()
    lib.rpto_philox(0, 0, 0, 0, out)
    golden = json.load(open(os.path.join(GOLDEN, "philox.json")))
    for g in golden:
        lib.rpto_philox(g["seed"], g["pixel"], g["sample"], g["block"], out)
        assert list(out) == g["out"]
    vals = []
    for i in range(4096):
        lib.rpto_philox(42, i, 3, 1, out)
        vals.extend(list(out))
    vals = np.asarray(vals)
    assert vals.min() >= 0.0 and vals.max() < 1.0
    assert abs(vals.mean() - 0.5) < 0.01 and abs(vals.var() - 1 / 12) < 0.005


def test_oracle_cornell_golden_film(pkg):
    """The oracle's own output on a tiny Cornell render is pinned by a committed fixture, so the checker cannot
    drift silently (tests/golden/cornell_oracle_16x9.npy, made by tools/make_golden.py)."""
    import parity

    world, st, flat = parity.load_scene("cornell", 16, 9, 4)
    sc = parity.oracle_scene(flat)
    film, cnt = sc.render_pt(st.params(seed=11))
    gold = np.load(os.path.join(GOLDEN, "cornell_oracle_16x9.npy"))
    assert np.allclose(film, gold, rtol=1e-4, atol=1e-7)
    sc.close()


def test_ggx_sampling_weights_conserve_energy(lib):
    """The reference's test_integral (ggx.rs) only prints; the property it eyeballs is asserted here: with VNDF sampling the
    estimator weight f |cos| / pdf of a non-absorbing dielectric never exceeds 1 and averages to 1 minus the single-scatter
    loss (which grows with roughness and towards grazing incidence); a conductor stays below its Fresnel reflectance."""
    lib.rpto_ggx_generate_and_evaluate.argtypes = [F, F, F, F, ct.c_int, F, F, ct.c_void_p, ct.c_void_p, ct.POINTER(F), ct.POINTER(F)]
    rng = np.random.default_rng(1)
    for metallic, eta, kappa in ((0, 1.5, 0.0), (1, 0.2, 3.5)):
        means = {}
        for alpha in (0.05, 0.2, 0.5):
            for cz in (0.95, 0.6, 0.25):
                wi, wo, f, p = (F * 3)(float(np.sqrt(1 - cz * cz)), 0.0, cz), (F * 3)(), F(), F()
                w = []
                for sx, sy in rng.random((1500, 2)):
                    lib.rpto_ggx_generate_and_evaluate(alpha, eta, 1.0, kappa, metallic, float(sx), float(sy), wi, wo, ct.byref(f), ct.byref(p))
                    w.append(f.value * abs(wo[2]) / p.value if p.value > 0 else 0.0)
                assert max(w) <= 1.0 + 1e-3, (metallic, alpha, cz, max(w))
                means[(alpha, cz)] = float(np.mean(w))
        if metallic:
            assert all(0.6 < m < 0.97 for m in means.values()), means
        else:
            assert all(0.85 < m <= 1.0 + 1e-3 for m in means.values()), means
            assert means[(0.05, 0.95)] > 0.999 and means[(0.5, 0.25)] < means[(0.05, 0.25)]


def test_estimator_does_not_depend_on_the_number_of_light_samples(pkg):
    """pt.rs:590-599 divides each light sample's contribution by light_samples: the converged mean must be the same for
    1, 2 and 4 samples per vertex (3 % at this sample count; the seeds are fixed). It is NOT the mean of BSDF-only sampling (light_samples = 0):
    the reference mixes the balance heuristic for NEE with the power heuristic at light hits, so its MIS
    weights do not sum to one; the restatement reproduces that bias rather than correcting it, and this test records it."""
    import parity

    world, st0, flat = parity.load_scene("cornell", 32, 18, 2048)
    sc = parity.oracle_scene(flat)
    mean = {}
    for L in (0, 1, 2, 4):
        st = pkg.renderer.PTSettings.from_dict(st0.to_dict())
        st.light_samples = L
        film, _ = sc.render_pt(st.params(seed=100 + L))
        mean[L] = float(film[..., 1].mean())
    sc.close()
    assert abs(mean[1] - mean[2]) / mean[2] < 0.03 and abs(mean[4] - mean[2]) / mean[2] < 0.03, mean
    assert mean[0] > 1.1 * mean[2], mean  # F10: documented, deliberate


def test_oracle_golden_films_all_scenes(pkg):
    """One tiny oracle film per scene blob plus a small importance-map bake, pinned by tests/golden/oracle_films_16x12.npz
    (tools/make_golden.py): any change to the checker's arithmetic on any material / light / environment / camera path shows
    up here, on the CPU, before it can silently move the target the GPU is compared with."""
    import parity

    gold = np.load(os.path.join(GOLDEN, "oracle_films_16x12.npz"))
    for name in gold.files:
        if name.startswith("hdri_imap"):
            continue
        world, st, flat = parity.load_scene(name, 16, 12, 4)
        sc = parity.oracle_scene(flat)
        film, _ = sc.render_pt(st.params(seed=11))
        sc.close()
        assert np.allclose(film, gold[name], rtol=1e-4, atol=1e-7), name
    world, st, flat = parity.load_scene("hdri", 16, 12, 1)
    sc = parity.oracle_scene(flat)
    lum, basis = pkg.importance_map.bake_curve_tables(world, pkg.curves.y_bar_curve(), st.wavelength_bounds)
    bk = sc.bake_importance_map(12, 20, lum, basis, st.wavelength_bounds)
    sc.close()
    assert np.allclose(bk["row_cdf"], gold["hdri_imap_row_cdf_12x20"], rtol=1e-6, atol=0)
    assert np.allclose(bk["marginal_cdf"], gold["hdri_imap_marginal_cdf_12"], rtol=1e-6, atol=0)


def furnace_expectation(n):
    """What the REFERENCE's estimator returns for the exact furnace (Constant environment of radiance 1, unit sphere of
    albedo 1, p_env = 1) at a first vertex with unit normal n, by quadrature. It is not 1 (BASELINE.md "Furnace"):
      * the walk's own escape to the environment is weighted power_heuristic(bsdf_psa, nee_psa) with BOTH pdfs divided by the
        cosine once more (Q9): ((1/pi) / c)^2 / (((1/pi) / c)^2 + ((1/4pi) / c)^2) = 16/17, whatever the direction;
      * NEE draws uv uniformly in [0,1]^2 but claims pdf 1/4pi (Q16, environment.rs:308-310), weights with the balance
        heuristic (F10): contribution (c/pi) * [(1/4pi) / (1/4pi + c/pi)] / (1/4pi) = 4c / (1 + 4c), averaged over uv;
      * its shadow ray starts at p + n * 0.001 * sign(dir.z) - WORLD z (Q12, pt.rs:256): a sample with dir.z < 0 starts
        inside the sphere and is occluded by it."""
    N = 1500
    u = (np.arange(N) + 0.5) / N
    U, V = np.meshgrid(u, u, indexing="ij")
    th, ph = (U - 0.5) * 2 * np.pi, np.pi * V
    w = np.stack([np.sin(ph) * np.cos(th), np.sin(ph) * np.sin(th), np.cos(ph)], -1)
    c = np.maximum(w @ np.asarray(n, dtype=np.float64), 0.0)
    nee = np.where((c > 0) & (w[..., 2] > 0), 4 * c / (1 + 4 * c), 0.0).mean()
    return 16.0 / 17.0 + nee


def test_oracle_furnace_energy(pkg):
    """Exact furnace (SURVEY A9 ii), BASELINE.json check (b) "returns 1.0 within 1e-3": background pixels see E = 1 directly,
    so their Y is the mean of y_bar over the wavelength range (asserted to 2 %); sphere pixels return what the reference's
    estimator returns, which is NOT 1: 1.072 at the centre (normal -x), 1.27 towards the +z pole, 0.94 towards the -z pole
    (furnace_expectation above: quirks Q9 / Q12 / Q16 / F10). Asserted against that quadrature within 3 sigma of the
    Monte-Carlo noise (single-wavelength sampling: relative std 1.2 / sqrt(samples))."""
    import parity

    assert abs(furnace_expectation([-1, 0, 0]) - 1.0723) < 2e-3 and abs(furnace_expectation([0, 0, -1]) - 16 / 17) < 1e-9
    world, st, flat = parity.load_scene("furnace_exact", 64, 64, 512)
    sc = parity.oracle_scene(flat)
    film, _ = sc.render_pt(st.params(seed=1))
    sc.close()
    ybar_mean = float(np.mean(flat.cie_lut[1]))
    corner = float(np.concatenate([film[:6, :6, 1].ravel(), film[:6, -6:, 1].ravel(), film[-6:, :6, 1].ravel(), film[-6:, -6:, 1].ravel()]).mean())
    assert abs(corner - ybar_mean) / ybar_mean < 0.02
    centre = float(film[28:36, 28:36, 1].mean())
    sigma = 1.2 / np.sqrt(64 * 512) + 1.2 / np.sqrt(144 * 512)
    assert abs(centre / corner - furnace_expectation([-1, 0, 0])) < 3 * sigma + 0.01, centre / corner
    # camera at -x looking +x with z up: image rows run from +z (top) to -z (bottom)
    rows = film[:, 28:36, 1].mean(axis=1) / corner
    covered = np.flatnonzero(np.abs(rows - 1.0) > 0.03)
    top, bottom = rows[covered[0] + 2 : covered[0] + 8].mean(), rows[covered[-1] - 7 : covered[-1] - 1].mean()
    assert top > 1.15 and bottom < 1.0, (top, bottom)  # NEE towards +z survives (1.2-1.27), towards -z is self-occluded (0.94-0.98)


def test_oracle_output_film_known_answers(pkg):
    """output_film restatement vs closed forms: Clamp per-channel = ceil(255 * OETF(M * clip(XYZ * 2^e))) and the Reinhard
    log-average l_w = exp(mean ln(0.001 + Y)) / factor (tonemap/clamp.rs:76-101, reinhard0.rs:48-73, mod.rs:147-205,314-331)."""
    import parity

    world, st, flat = parity.load_scene("cornell", 8, 8, 1)
    sc = parity.oracle_scene(flat)
    rng = np.random.default_rng(0)
    film = np.zeros((8, 8, 4), dtype=np.float32)
    film[..., :3] = rng.uniform(0.0, 1.2, size=(8, 8, 3)).astype(np.float32)
    o = pkg.ffi.RptOutputSettings()
    o.tonemapper, o.luminance_only, o.exposure, o.colorspace, o.factor = 0, 0, 0.0, 0, 1.0
    rgb, rgba, _ = sc.output_film(o, film, 8, 8)
    M = np.array([[3.24096994, -1.53738318, -0.49861076], [-0.96924364, 1.8759675, 0.04155506], [0.05563008, -0.20397696, 1.05697151]], dtype=np.float32)
    lin = np.clip(film[..., :3], 0, 1) @ M.T
    enc = np.where(lin < 0.0031308, 323.0 / 25.0 * lin, 211.0 / 200.0 * np.power(np.maximum(lin, 0), 5.0 / 12.0) - 11.0 / 200.0)
    want = np.clip(np.ceil(enc * 255.0), 0, 255).astype(np.uint8)
    assert np.abs(rgba[..., :3].astype(int) - want.astype(int)).max() <= 1
    assert np.allclose(rgb, film[..., :3] @ M.T, rtol=1e-5, atol=1e-6)
    o.tonemapper, o.luminance_only, o.key_value, o.factor = 1, 1, 0.18, 2.0
    _, _, lw = sc.output_film(o, film, 8, 8)
    assert np.isclose(lw[1], np.exp(np.mean(np.log(0.001 + film[..., 1].astype(np.float64)))) / 2.0, rtol=1e-5)
    sc.close()
