"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs. BASELINE.json's three checks: (a) primary-hit ids >= 99.99 %, (b) furnace within 1e-3,
(c) converged images within a stated relMSE tolerance in linear XYZ."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu

ALL_SCENES = ["cornell", "furnace", "furnace_exact", "gem", "hdri", "test_nee_sphere", "orb_caustic", "sun_test",
              "parallel_prism", "lighting_north", "rtiow2", "kitchen_sink"]


@pytest.fixture(scope="module")
def scenes():
    cache = {}

    def get(name, w, h, spp):
        key = (name, w, h, spp)
        if key not in cache:
            world, st, flat = parity.load_scene(name, w, h, spp)
            cache[key] = (st, parity.cuda_scene(flat), parity.oracle_scene(flat))
        return cache[key]

    yield get
    for st, cs, os_ in cache.values():
        cs.close()
        os_.close()


@pytest.mark.parametrize("name", ALL_SCENES + ["instanced_monkeys"])
def test_primary_hit_ids(scenes, name):
    """(a) primary-ray hit (instance, primitive) ids match on >= 99.99 % of pixels."""
    w, h = (320, 180) if name != "instanced_monkeys" else (256, 144)
    st, cs, os_ = scenes(name, w, h, 1)
    p = st.params(seed=3)
    gi, gp, gt = cs.trace_primary(p)
    oi, op, ot = os_.trace_primary(p)
    same = (gi == oi) & (gp == op)
    frac = float(np.mean(same))
    assert frac >= 0.9999, f"{name}: only {frac:.6f} of primary hits match"
    both = same & (gi != 0xFFFFFFFF)
    if both.any():
        assert np.allclose(gt[both], ot[both], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["cornell", "gem", "test_nee_sphere", "instanced_monkeys", "kitchen_sink"])
def test_random_rays(scenes, name):
    """Closest hits of random rays from inside the scene bounds match the oracle (ids bit-exact)."""
    st, cs, os_ = scenes(name, 64, 64, 1)
    rng = np.random.default_rng(11)
    n = 20000
    scale = {"instanced_monkeys": 40.0, "kitchen_sink": 3.0}.get(name, 1.0)
    o = (rng.uniform(-0.9, 0.9, size=(n, 3)) * scale).astype(np.float32)
    if name == "cornell":
        o = (rng.uniform(0.01, 0.54, size=(n, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    tmax = np.full(n, np.inf, dtype=np.float32)
    gi, gp, gt = cs.trace_rays(o, d, tmax)
    oi, op, ot = os_.trace_rays(o, d, tmax)
    frac = float(np.mean((gi == oi) & (gp == op)))
    assert frac >= 0.9999, f"{name}: {frac:.6f}"


def axis_aligned_rays(rng, n, lo, hi):
    """Rays with one or two exactly-zero direction components (+0 and -0), origins off the planes x/y/z = 0
    (b * inf - o * inf is NaN or +-inf depending on signs; the slab test must leave such an axis unconstrained
    like aabb.rs:41-45 does)."""
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    kind = rng.integers(0, 3, size=n)          # 0: one zero component, 1: two zero components, 2: all nonzero (control)
    ax = rng.integers(0, 3, size=n)
    neg = rng.integers(0, 2, size=n).astype(bool)
    zero = np.where(neg, np.float32(-0.0), np.float32(0.0))
    for i in range(n):
        if kind[i] == 0:
            d[i, ax[i]] = zero[i]
        elif kind[i] == 1:
            d[i, (ax[i] + 1) % 3] = zero[i]
            d[i, (ax[i] + 2) % 3] = -zero[i]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


@pytest.mark.parametrize("name", ["cornell", "gem", "test_nee_sphere", "instanced_monkeys", "kitchen_sink", "rtiow2"])
def test_axis_aligned_rays(scenes, name):
    """Rays with exactly-zero direction components hit what the oracle hits (bit-exact ids); a third of the batch is
    a control group of general rays through the same code path."""
    st, cs, os_ = scenes(name, 64, 64, 1)
    rng = np.random.default_rng(17)
    n = 12000
    scale = {"instanced_monkeys": 40.0, "kitchen_sink": 3.0}.get(name, 1.0)
    lo, hi = (-0.9 * scale, 0.9 * scale) if name != "cornell" else (0.01, 0.54)
    o, d = axis_aligned_rays(rng, n, lo, hi)
    tmax = np.full(n, np.inf, dtype=np.float32)
    gi, gp, gt = cs.trace_rays(o, d, tmax)
    oi, op, ot = os_.trace_rays(o, d, tmax)
    same = (gi == oi) & (gp == op)
    assert (oi != 0xFFFFFFFF).mean() > 0.2, "the batch should hit something"
    assert same.mean() >= 0.9999, f"{name}: {same.mean():.6f}; first mismatches {np.flatnonzero(~same)[:5]}"


@pytest.mark.parametrize("name", ALL_SCENES + ["instanced_monkeys"])
def test_same_stream_images(scenes, name):
    """Same Philox streams on both sides: the low-spp images agree far below the noise floor.
    Tolerance: relMSE(XYZ) <= 2e-3 and mean-Y within 2e-3 (divergence only where fp32 rounding flips a
    branch: a russian-roulette decision, a grazing hit)."""
    st, cs, os_ = scenes(name, 96, 54, 8)
    p = st.params(seed=5)
    fg, cg = cs.render_pt(p)
    fo, co = os_.render_pt(p)
    assert np.isfinite(fg).all()
    assert parity.mean_rel_diff(fg, fo) < 2e-3, (name, fg[..., 1].mean(), fo[..., 1].mean())
    assert parity.rel_mse(fg, fo) < 2e-3, (name, parity.rel_mse(fg, fo))
    # ... and the pixels that differ at all are a handful (a systematic difference - a wrong weight, a visibility rule -
    # touches percent-level fractions of the image while staying under the relMSE bound above)
    yg, yo = fg[..., 1], fo[..., 1]
    differing = float(np.mean(np.abs(yg - yo) > 1e-4 * np.maximum(np.abs(yo), 1e-6)))
    assert differing <= 5e-3, (name, differing)
    # Profile counters (profile.rs): identical up to the rare divergent paths
    for k in ("camera_rays",):
        assert getattr(cg, k) == getattr(co, k)
    for k in ("bounce_rays", "shadow_rays", "env_hits", "segments"):
        a, b = getattr(cg, k), getattr(co, k)
        assert abs(a - b) <= max(4, 2e-3 * b), (name, k, a, b)


def test_furnace_exact(scenes):
    """(b) white furnace: Constant env 1.0, unit Lambertian sphere with albedo clamped to 1, p_env = 1.
    Pixels covered by the sphere and background pixels both return 1.0 within 1e-3 ... of the oracle; the
    absolute value is reported next to 1.0 (the reference's uniform-uv env sampling with a
    1/4pi pdf biases NEE, MIS only partly hides it)."""
    st, cs, os_ = scenes("furnace_exact", 128, 128, 256)
    p = st.params(seed=9)
    fg, _ = cs.render_pt(p)
    fo, _ = os_.render_pt(p)
    yg, yo = float(fg[..., 1].mean()), float(fo[..., 1].mean())
    assert abs(yg - yo) / yo < 1e-3, (yg, yo)
    # background pixels see the env directly: exactly the CIE-weighted mean of a unit spectrum
    corner_g, corner_o = float(fg[:8, :8, 1].mean()), float(fo[:8, :8, 1].mean())
    assert abs(corner_g - corner_o) / corner_o < 1e-3
    centre_g = float(fg[56:72, 56:72, 1].mean())
    print(f"furnace_exact: mean Y gpu {yg:.6f} oracle {yo:.6f}; centre/corner = {centre_g / corner_g:.4f} (1.0 = energy conserving)")
    # BASELINE.json check (b) says "returns 1.0 within 1e-3"; the REFERENCE's estimator returns 1.0723 at the centre of this
    # sphere (quadrature + derivation: tests/test_oracle_golden.py::furnace_expectation, BASELINE.md "Furnace"), and so must we
    from test_oracle_golden import furnace_expectation

    sigma = 1.2 / np.sqrt(256 * 256) + 1.2 / np.sqrt(64 * 256)
    assert abs(centre_g / corner_g - furnace_expectation([-1, 0, 0])) < 3 * sigma + 0.005, centre_g / corner_g


def test_furnace_shipped(scenes):
    """(b) the shipped white_furnace scene (rough glass sphere, camera inside): GPU == oracle within 1e-3 on mean Y."""
    st, cs, os_ = scenes("furnace", 96, 96, 128)
    p = st.params(seed=2)
    fg, _ = cs.render_pt(p)
    fo, _ = os_.render_pt(p)
    assert parity.mean_rel_diff(fg, fo) < 1e-3


@pytest.mark.parametrize("name", ["cornell", "test_nee_sphere", "orb_caustic"])
def test_converged_images_independent_seeds(scenes, name):
    """(c) converged images from INDEPENDENT seeds: relMSE(GPU seed A, oracle seed B) must sit at the
    oracle's own two-seed noise floor relMSE(oracle seed B, oracle seed C) (within 1.5x)."""
    st, cs, os_ = scenes(name, 64, 36, 512)
    fg, _ = cs.render_pt(st.params(seed=101))
    fo1, _ = os_.render_pt(st.params(seed=202))
    fo2, _ = os_.render_pt(st.params(seed=303))
    floor = parity.rel_mse(fo2, fo1)
    got = parity.rel_mse(fg, fo1)
    assert got < 1.5 * floor + 1e-5, (name, got, floor)
    assert parity.mean_rel_diff(fg, fo1) < 3 * parity.mean_rel_diff(fo2, fo1) + 5e-3


def test_spp_split_equals_whole(scenes):
    """Multi-GPU partition property: rendering spp as two offset halves and summing equals the whole."""
    st, cs, _ = scenes("cornell", 96, 54, 8)
    whole, _ = cs.render_pt(st.params(seed=4, spp=8, spp_offset=0, spp_total=8))
    a, _ = cs.render_pt(st.params(seed=4, spp=4, spp_offset=0, spp_total=8))
    b, _ = cs.render_pt(st.params(seed=4, spp=4, spp_offset=4, spp_total=8))
    assert np.allclose(a + b, whole, rtol=1e-4, atol=1e-7)


VARIANTS = [
    dict(only_direct=True), dict(light_samples=0), dict(light_samples=5), dict(min_bounces=0), dict(min_bounces=6),
    dict(max_bounces=1), dict(max_bounces=3), dict(wavelength_bounds=(450.0, 650.0)), dict(light_samples=0, only_direct=True),
]


@pytest.mark.parametrize("name", ["cornell", "kitchen_sink"])
def test_render_parameter_variants(name):
    """Every field of RptRenderParams that changes the integrator's control flow (integrator/mod.rs:59-105, pt.rs:425-428,
    519-523,583): only_direct, light_samples (0 = BSDF sampling only), the russian-roulette start index min_bounces,
    max_bounces, the wavelength bounds. Same-stream film and counters against the oracle for each."""
    world, st0, flat = parity.load_scene(name, 64, 36, 8)
    cs, os_ = parity.cuda_scene(flat), parity.oracle_scene(flat)
    seen = []
    for v in VARIANTS:
        st = parity.pkg().renderer.PTSettings.from_dict(st0.to_dict())
        for k, val in v.items():
            setattr(st, k, val)
        p = st.params(seed=31)
        fg, cg = cs.render_pt(p)
        fo, co = os_.render_pt(p)
        assert np.isfinite(fg).all(), v
        assert parity.rel_mse(fg, fo) < 2e-3 and parity.mean_rel_diff(fg, fo) < 2e-3, (name, v, parity.rel_mse(fg, fo))
        for k in ("camera_rays", "bounce_rays", "shadow_rays", "env_hits", "segments"):
            a, b = getattr(cg, k), getattr(co, k)
            assert abs(a - b) <= max(4, 2e-3 * b), (name, v, k, a, b)
        seen.append((co.segments, co.shadow_rays, round(float(fo[..., 1].mean()), 6)))
    assert len(set(seen)) >= 7  # the variants really take different paths
    cs.close()
    os_.close()


def test_against_committed_golden_films(pkg):
    """The CUDA path against the COMMITTED oracle fixtures (tests/golden/oracle_films_16x12.npz, tools/make_golden.py): the
    same comparison as test_same_stream_images, but with a target that cannot move with the oracle library of the day."""
    import os

    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_films_16x12.npz"))
    for name in gold.files:
        if name.startswith("hdri_imap"):
            continue
        world, st, flat = parity.load_scene(name, 16, 12, 4)
        cs = parity.cuda_scene(flat)
        film, _ = cs.render_pt(st.params(seed=11))
        cs.close()
        assert parity.rel_mse(film, gold[name]) < 2e-3 and parity.mean_rel_diff(film, gold[name]) < 2e-3, (name, parity.rel_mse(film, gold[name]))
    world, st, flat = parity.load_scene("hdri", 16, 12, 1)
    cs = parity.cuda_scene(flat)
    lum, basis = pkg.importance_map.bake_curve_tables(world, pkg.curves.y_bar_curve(), st.wavelength_bounds)
    bk = cs.bake_importance_map(12, 20, lum, basis, st.wavelength_bounds)
    cs.close()
    assert np.array_equal(bk["row_cdf"], gold["hdri_imap_row_cdf_12x20"]) and np.array_equal(bk["marginal_cdf"], gold["hdri_imap_marginal_cdf_12"])


def test_full_size_headline_config(pkg):
    """BASELINE configs[0] at its full size (Cornell 1920x1080 @ 16 spp, the bench workload: one wave of 33.2 M paths, queues
    of tens of millions of entries, every warp claiming tiles dynamically): primary hit ids over all 2 M pixels, the
    same-stream film and the Profile counters against the oracle (~15 s of CPU), and the spp-split linearity property."""
    world, st, flat = parity.load_scene("cornell")
    assert (st.width, st.height, st.min_samples) == (1920, 1080, 16)
    cs, os_ = parity.cuda_scene(flat), parity.oracle_scene(flat)
    p = st.params(seed=77)
    gi, gp, gt = cs.trace_primary(p)
    oi, op, ot = os_.trace_primary(p)
    assert float(np.mean((gi == oi) & (gp == op))) >= 0.9999
    fg, cg = cs.render_pt(p)
    fo, co = os_.render_pt(p)
    assert np.isfinite(fg).all()
    assert parity.rel_mse(fg, fo) < 1e-6, parity.rel_mse(fg, fo)
    assert abs(float(fg[..., 1].mean()) - float(fo[..., 1].mean())) / float(fo[..., 1].mean()) < 1e-5
    assert cg.camera_rays == co.camera_rays == 1920 * 1080 * 16
    for k in ("bounce_rays", "shadow_rays", "env_hits", "segments"):
        a, b = getattr(cg, k), getattr(co, k)
        assert abs(a - b) <= 1e-5 * b, (k, a, b)
    # linearity: 16 spp == 6 spp + 10 spp with continued sample indices (what the multi-GPU split relies on)
    a, _ = cs.render_pt(st.params(seed=77, spp=6, spp_offset=0, spp_total=16))
    b, _ = cs.render_pt(st.params(seed=77, spp=10, spp_offset=6, spp_total=16))
    assert np.allclose(a + b, fg, rtol=1e-4, atol=1e-7)
    cs.close()
    os_.close()


def test_empty_and_edge_cases(scenes, pkg):
    st, cs, _ = scenes("cornell", 96, 54, 8)
    film, cnt = cs.render_pt(st.params(seed=1, spp=0))
    assert not film.any() and cnt.camera_rays == 0
    with pytest.raises(pkg.ffi.RptError):
        p = st.params(seed=1)
        p.camera = 7
        cs.render_pt(p)
    # ragged: 1x1 film, 1 spp; odd sizes that do not fill a warp
    for (w, h) in ((1, 1), (33, 7)):
        world, st2, flat = parity.load_scene("cornell", w, h, 3)
        c2, o2 = parity.cuda_scene(flat), parity.oracle_scene(flat)
        fg, _ = c2.render_pt(st2.params(seed=8))
        fo, _ = o2.render_pt(st2.params(seed=8))
        assert fg.shape == (h, w, 4) and np.isfinite(fg).all()
        assert np.allclose(fg, fo, rtol=2e-2, atol=1e-4) or parity.rel_mse(fg, fo) < 5e-2
        c2.close()
        o2.close()


def test_public_api_render_writes_exr_and_png(pkg, tmp_path):
    """Renderer::render end to end: film -> output_film on the device -> `<filename>.exr` (linear RGB) and `<filename>.png`
    on the host, as renderer/mod.rs:24-80 does. The EXR holds exactly rpt_output_film's linear payload."""
    import os

    world, st, flat = parity.load_scene("cornell", 48, 27, 4)
    rs = pkg.loader.RenderSettings(filename="beauty", width=48, height=27, integrator_type="PT", light_samples=st.light_samples,
                                   medium_aware=False, min_bounces=st.min_bounces, max_bounces=st.max_bounces, hwss=False, threads=1,
                                   min_samples=4, camera_id="main", russian_roulette=True, only_direct=False, wavelength_bounds=None,
                                   premultiply=None)
    rs.raw = {"tonemap_settings": {"type": "Reinhard1", "luminance_only": False, "key_value": 0.18, "white_point": 1.0},
              "colorspace_settings": {"type": "sRGB"}}
    cfg = pkg.loader.Config(scene_file="", renderer={"type": "Cuda"}, render_settings=[rs], camera_names_to_index={"main": 0})
    r = pkg.CudaRenderer(device=0, seed=2)
    films = r.render(world, cfg, output_dir=str(tmp_path))
    exr_path, png_path = os.path.join(tmp_path, "beauty.exr"), os.path.join(tmp_path, "beauty.png")
    assert os.path.exists(exr_path)
    rgb = pkg.exr.read_exr_rgb(exr_path)
    sc = r.make_scene(world, st.wavelength_bounds)
    want, rgba, _ = sc.output_film(pkg.renderer.output_settings(rs), films["beauty"], 48, 27)
    sc.close()
    assert np.array_equal(rgb, want)
    if os.path.exists(png_path):
        from PIL import Image

        assert np.array_equal(np.asarray(Image.open(png_path)), rgba)


def test_public_api_render(pkg):
    """The reference-facing call: CudaRenderer.render(world, config) -> one mean-XYZ film per render setting
    (mirror of `trait Renderer::render`, src/renderer/mod.rs:107-112), checked against the oracle."""
    import parity

    world, st, flat = parity.load_scene("cornell", 64, 36, 4)
    rs = pkg.loader.RenderSettings(filename="beauty", width=64, height=36, integrator_type="PT", light_samples=st.light_samples,
                                   medium_aware=False, min_bounces=st.min_bounces, max_bounces=st.max_bounces, hwss=False, threads=1,
                                   min_samples=4, camera_id="main", russian_roulette=True, only_direct=False, wavelength_bounds=None,
                                   premultiply=None)
    cfg = pkg.loader.Config(scene_file="", renderer={"type": "Cuda"}, render_settings=[rs], camera_names_to_index={"main": 0})
    films = pkg.CudaRenderer(device=0, seed=13).render(world, cfg)
    assert set(films) == {"beauty"} and films["beauty"].shape == (36, 64, 4)
    os_ = parity.oracle_scene(flat)
    fo, _ = os_.render_pt(st.params(seed=13))
    assert parity.rel_mse(films["beauty"], fo) < 1e-6
    with pytest.raises(ValueError):
        rs.integrator_type = "LT"
        pkg.CudaRenderer(device=0).render(world, cfg)
    os_.close()


@pytest.mark.parametrize("name", ["kitchen_sink", "cornell"])
def test_panorama_camera(pkg, name):
    """PanoramaCamera (N4; panorama_camera.rs:68-91) in k_raygen: primary hit ids and the same-stream image against the oracle."""
    import parity

    world, st, flat = parity.load_scene(name, 128, 64, 8)
    look_from, look_at = ((0.0, -0.5, 1.0), (1.0, 0.2, 0.2)) if name == "kitchen_sink" else ((0.278, 0.273, 0.3), (0.278, 0.273, 0.0))
    world.cameras = [pkg.world.Camera.new_panorama("pano", look_from, look_at, (0.0, 0.0, 1.0) if name == "kitchen_sink" else (0.0, 1.0, 0.0), 360.0, 170.0)]
    flat = pkg.ffi.FlatScene(world, st.wavelength_bounds[0], st.wavelength_bounds[1], 1024)
    cs, os_ = parity.cuda_scene(flat), parity.oracle_scene(flat)
    p = st.params(seed=8)
    gi, gp, gt = cs.trace_primary(p)
    oi, op, ot = os_.trace_primary(p)
    assert np.mean((gi == oi) & (gp == op)) >= 0.9999
    assert len(np.unique(oi)) >= 3
    fg, cg = cs.render_pt(p)
    fo, co = os_.render_pt(p)
    assert np.isfinite(fg).all()
    assert parity.rel_mse(fg, fo) < 2e-3 and parity.mean_rel_diff(fg, fo) < 2e-3
    assert cg.camera_rays == co.camera_rays and abs(cg.segments - co.segments) <= max(4, 2e-3 * co.segments)
    cs.close()
    os_.close()


def test_importance_map_bake_on_device(pkg):
    """N3: rpt_scene_bake_importance_map from the scene's resident environment texels == the oracle's scalar restatement,
    bit for bit (same f32 operation order, no FMA contraction), for the map size of hdri_test.toml and an odd one; the
    baked tables are installed: renders that follow sample the environment through them on both sides."""
    import parity

    world, st, flat = parity.load_scene("hdri", 96, 54, 8)
    e = world.environment
    host = {"row_pdf": e.imap_row_pdf, "row_cdf": e.imap_row_cdf, "marginal_pdf": e.imap_marginal_pdf, "marginal_cdf": e.imap_marginal_cdf}
    cs, os_ = parity.cuda_scene(flat), parity.oracle_scene(flat)
    lum, basis = pkg.importance_map.bake_curve_tables(world, pkg.curves.y_bar_curve(), st.wavelength_bounds)
    for rows, cols in ((e.imap_row_pdf.shape), (77, 130)):
        g = cs.bake_importance_map(rows, cols, lum, basis, st.wavelength_bounds)
        o = os_.bake_importance_map(rows, cols, lum, basis, st.wavelength_bounds)
        for k in ("row_pdf", "row_cdf", "marginal_pdf", "marginal_cdf"):
            assert np.array_equal(g[k], o[k]), (rows, cols, k, np.max(np.abs(g[k] - o[k])))
        assert g["marginal_integral"] == o["marginal_integral"]
        if (rows, cols) == e.imap_row_pdf.shape:
            for k, r in host.items():  # and the vectorised f64 host bake agrees to f32 accumulation error
                assert np.allclose(g[k], r, rtol=2e-5, atol=1e-9), k
        p = st.params(seed=21)
        fg, cg = cs.render_pt(p)
        fo, co = os_.render_pt(p)
        assert parity.rel_mse(fg, fo) < 1e-6, (rows, cols, parity.rel_mse(fg, fo))
        assert cg.shadow_rays == co.shadow_rays
    # device-only bake (nothing downloaded) leaves the same tables installed
    f_before, _ = cs.render_pt(st.params(seed=22))
    cs.bake_importance_map(77, 130, lum, basis, st.wavelength_bounds, download=False)
    f_after, _ = cs.render_pt(st.params(seed=22))
    assert parity.rel_mse(f_after, f_before) < 1e-10  # (atomic accumulation order: not bit-identical run to run)
    cs.close()
    os_.close()
    # a scene without an HDR environment refuses
    world2, st2, flat2 = parity.load_scene("cornell", 32, 18, 1)
    c2 = parity.cuda_scene(flat2)
    with pytest.raises(pkg.ffi.RptError):
        c2.bake_importance_map(8, 8, lum, basis[:400], st2.wavelength_bounds)
    c2.close()


def test_public_api_bakes_unbaked_importance_map_on_device(pkg):
    """CudaRenderer.make_scene repeats phase 2 of NaiveRenderer::render (naive.rs:469-487): an HDR environment whose
    importance map is still Unbaked is baked before rendering - on the device. Same film as with the host-baked tables
    up to the f32-vs-f64 bake difference (a handful of CDF bins move by an ulp)."""
    import parity

    world, st, flat = parity.load_scene("hdri", 64, 36, 8)
    e = world.environment
    r = pkg.CudaRenderer(device=0, seed=4)
    s_host = r.make_scene(world, st.wavelength_bounds)
    f_host, _ = r.render_sampled(s_host, st)
    s_host.close()
    rows, cols = e.imap_row_pdf.shape
    e.imap_row_pdf = e.imap_row_cdf = e.imap_marginal_pdf = e.imap_marginal_cdf = None
    e.imap_request = (rows, cols, pkg.curves.y_bar_curve())
    s_dev = r.make_scene(world, st.wavelength_bounds)
    f_dev, _ = r.render_sampled(s_dev, st)
    s_dev.close()
    assert abs(f_dev[..., 1].mean() - f_host[..., 1].mean()) / f_host[..., 1].mean() < 2e-2
    # without a map the environment would be sampled uniformly: the films differ visibly from the importance-sampled one
    e.imap_request = None
    s_uni = r.make_scene(world, st.wavelength_bounds)
    f_uni, _ = r.render_sampled(s_uni, st)
    s_uni.close()
    assert parity.rel_mse(f_dev, f_host) < parity.rel_mse(f_uni, f_host)


@pytest.mark.parametrize("tm", ["Clamp", "Reinhard0", "Reinhard1"])
@pytest.mark.parametrize("lum_only", [True, False])
@pytest.mark.parametrize("cs", ["sRGB", "Rec2020"])
def test_output_film(scenes, pkg, tm, lum_only, cs):
    """N2: device output_film (tonemapper initialize + map, XYZ->RGB, OETF, bytes) vs the oracle's sequential restatement.
    Tolerance: linear RGB rtol 1e-5; log-average rel 1e-4 (parallel double sum vs sequential f32/f64); PNG bytes equal on
    >= 99.5 % of channels and never more than 1 apart (powf / logf ulps on a ceil() boundary)."""
    st, cs_gpu, cs_or = scenes("cornell", 160, 90, 16)
    film, _ = cs_or.render_pt(st.params(seed=31))
    film[3, 5, :3] = np.nan  # a poisoned pixel exercises the MAUVE substitution
    o = pkg.ffi.RptOutputSettings()
    o.tonemapper, o.luminance_only = pkg.renderer.TONEMAPPERS[tm], int(lum_only)
    o.exposure, o.key_value, o.white_point = 4.0, 0.18, 2.0
    o.colorspace, o.factor = pkg.renderer.COLORSPACES[cs], 1.5
    rg, bg, lg = cs_gpu.output_film(o, film, st.width, st.height)
    ro, bo, lo = cs_or.output_film(o, film, st.width, st.height)
    ok = np.isfinite(ro)
    assert np.allclose(rg[ok], ro[ok], rtol=1e-5, atol=1e-7)
    assert np.allclose(lg, lo, rtol=1e-4)
    d = np.abs(bg.astype(np.int32) - bo.astype(np.int32))
    assert d.max() <= 1 and float(np.mean(d == 0)) >= 0.995, (d.max(), float(np.mean(d == 0)))
    assert (bg[..., 3] == 255).all() and bg[..., :3].max() > 0


def test_output_film_from_device_resident_film(scenes, pkg):
    """film=None tonemaps the film the last render left on the device: same bytes as passing the downloaded film."""
    st, cs_gpu, _ = scenes("cornell", 160, 90, 16)
    film, _ = cs_gpu.render_pt(st.params(seed=5))
    o = pkg.ffi.RptOutputSettings()
    o.tonemapper, o.luminance_only, o.key_value, o.white_point, o.colorspace, o.factor = 2, 0, 0.18, 1.0, 2, 1.0
    _, b_host, _ = cs_gpu.output_film(o, film, st.width, st.height)
    cs_gpu.render_pt(st.params(seed=5))
    _, b_dev, _ = cs_gpu.output_film(o, None, st.width, st.height)
    assert float(np.mean(b_host == b_dev)) > 0.999  # energy atomics reorder between the two renders


def test_tma_tile_variant(monkeypatch):
    """k_trace with TMA-staged queue tiles (cp.async.bulk + mbarrier, RPT_TMA_TILES=1) is bit-identical to the default."""
    world, st, flat = parity.load_scene("cornell", 160, 90, 4)
    base = parity.cuda_scene(flat)
    f0, c0 = base.render_pt(st.params(seed=21))
    monkeypatch.setenv("RPT_TMA_TILES", "1")
    tma = parity.cuda_scene(flat)
    f1, c1 = tma.render_pt(st.params(seed=21))
    assert c0.segments == c1.segments and c0.shadow_rays == c1.shadow_rays
    assert np.allclose(f0, f1, rtol=1e-6, atol=1e-9)  # only the order of the energy atomics may differ
    base.close()
    tma.close()


@pytest.mark.parametrize("name", ["cornell", "test_nee_sphere", "rtiow2", "orb_caustic", "sun_test", "furnace", "hdri"])
def test_small_scene_mode_equals_bvh(monkeypatch, name):
    """RPT_SMALL=1: scenes of <= 64 leaves skip the BVH (warp-uniform walk of the leaf list out of shared memory, SmallTrav).
    Both modes answer the same query, so hit ids, counters and films are identical (energy atomics reorder). The mode is
    opt-in: it runs 32 of 32 lanes but more instructions per ray than the BVH walk (profiles/r02_small_vs_bvh.md)."""
    world, st, flat = parity.load_scene(name, 160, 90, 4)
    bvh = parity.cuda_scene(flat)
    monkeypatch.setenv("RPT_SMALL", "1")
    small = parity.cuda_scene(flat)
    monkeypatch.delenv("RPT_SMALL")
    p = st.params(seed=9, flags=2)
    fs, cs_ = small.render_pt(p)
    fb, cb = bvh.render_pt(p)
    assert cs_.walk_nodes == 0 and (cb.walk_nodes > 0 or len(world.instances) == 1), "the two scenes must really run the two modes"
    for k in ("segments", "bounce_rays", "shadow_rays", "shadow_rays_traced", "env_hits"):
        assert getattr(cs_, k) == getattr(cb, k), (name, k)
    assert np.allclose(fs, fb, rtol=1e-5, atol=1e-9), (name, float(np.abs(fs - fb).max()))
    gi, gp, gt = small.trace_primary(p)
    bi, bp, bt = bvh.trace_primary(p)
    assert np.array_equal(gi, bi) and np.array_equal(gp, bp) and np.array_equal(gt, bt)
    small.close()
    bvh.close()


@pytest.mark.parametrize("name", ["cornell", "gem", "instanced_monkeys", "kitchen_sink", "hdri2", "sun_test"])
def test_lane_refill_mode_equals_tile_mode(monkeypatch, name):
    """TRAV_BVH_REFILL (RPT_REFILL=1: finished lanes are re-armed with the next ray of the queue while the rest of the warp
    keeps walking; opt-in after measurement, profiles/r02_refill_vs_tile.md) answers the same queries as the tile-at-a-time walk:
    same hit ids, same counters, same film up to the order of the energy atomics."""
    world, st, flat = parity.load_scene(name, 192, 108, 4)
    monkeypatch.setenv("RPT_REFILL", "0")
    tile = parity.cuda_scene(flat)
    monkeypatch.setenv("RPT_REFILL", "1")
    refill = parity.cuda_scene(flat)
    monkeypatch.delenv("RPT_REFILL")
    p = st.params(seed=19, flags=2)
    ft, ct = tile.render_pt(p)
    fr, cr = refill.render_pt(p)
    for k in ("segments", "bounce_rays", "shadow_rays", "shadow_rays_traced", "env_hits", "walk_tris", "walk_insts"):
        assert getattr(ct, k) == getattr(cr, k), (name, k, getattr(ct, k), getattr(cr, k))
    ok = np.isfinite(ft)
    assert np.array_equal(ok, np.isfinite(fr))
    assert np.allclose(ft[ok], fr[ok], rtol=1e-5, atol=1e-9), (name, float(np.abs(ft[ok] - fr[ok]).max()))
    ti, tp, tt = tile.trace_primary(p)
    ri, rp, rt = refill.trace_primary(p)
    assert np.array_equal(ti, ri) and np.array_equal(tp, rp) and np.array_equal(tt, rt)
    tile.close()
    refill.close()


@pytest.mark.parametrize("name", ["cornell", "gem", "instanced_monkeys", "kitchen_sink", "hdri2", "sun_test", "furnace", "test_nee_sphere"])
def test_bvh4_mode_equals_bvh2(monkeypatch, name):
    """TRAV_BVH4 (RPT_BVH4=1: the same trees collapsed to four children per node, nearest-first walk, TravT<true>) answers the
    same queries as the two-wide walk: same hit ids (rendered primaries, random and axis-aligned rays through rpt_trace_rays),
    same counters, same film up to the order of the energy atomics; it visits fewer nodes."""
    world, st, flat = parity.load_scene(name, 192, 108, 4)
    monkeypatch.setenv("RPT_BVH4", "0")
    two = parity.cuda_scene(flat)
    monkeypatch.setenv("RPT_BVH4", "1")
    four = parity.cuda_scene(flat)
    monkeypatch.delenv("RPT_BVH4")
    p = st.params(seed=29, flags=2)
    f2, c2 = two.render_pt(p)
    f4, c4 = four.render_pt(p)
    for k in ("segments", "bounce_rays", "shadow_rays", "shadow_rays_traced", "env_hits", "nee_vertices"):
        assert getattr(c2, k) == getattr(c4, k), (name, k, getattr(c2, k), getattr(c4, k))
    assert c4.walk_nodes < c2.walk_nodes or c2.walk_nodes == 0, (name, c2.walk_nodes, c4.walk_nodes)
    ok = np.isfinite(f2)
    assert np.array_equal(ok, np.isfinite(f4))
    assert np.allclose(f2[ok], f4[ok], rtol=1e-5, atol=1e-9), (name, float(np.abs(f2[ok] - f4[ok]).max()))
    a = two.trace_primary(p)
    b = four.trace_primary(p)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    rng = np.random.default_rng(5)
    scale = {"instanced_monkeys": 40.0, "kitchen_sink": 3.0}.get(name, 1.0)
    lo, hi = (-0.9 * scale, 0.9 * scale) if name != "cornell" else (0.01, 0.54)
    o1 = rng.uniform(lo, hi, size=(12000, 3)).astype(np.float32)
    d1 = rng.normal(size=(12000, 3)).astype(np.float32)
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    o2, d2 = axis_aligned_rays(rng, 6000, lo, hi)
    o, d = np.concatenate([o1, o2]), np.concatenate([d1, d2])
    tmax = np.full(len(o), np.inf, np.float32)
    a = two.trace_rays(o, d, tmax)
    b = four.trace_rays(o, d, tmax)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    two.close()
    four.close()


@pytest.mark.parametrize("name", ["hdri2", "kitchen_sink", "instanced_monkeys", "cornell", "sun_test", "hdri"])
def test_nee_samples_sorted_by_kind_equal_unsorted(monkeypatch, name):
    """k_nee<., NEE_LIGHT> + k_nee<., NEE_ENV>: in scenes that sample both the environment and lights (0 < p_env < 1) every (vertex, sample) pair is
    classified first and the pairs are drawn 32 of a kind at a time (default there; RPT_NEE_SORT=0 / 1 forces either form).
    The samples are the same samples: identical counters, film equal up to the order of the energy atomics."""
    world, st, flat = parity.load_scene(name, 192, 108, 4)
    monkeypatch.setenv("RPT_NEE_SORT", "0")
    plain = parity.cuda_scene(flat)
    monkeypatch.setenv("RPT_NEE_SORT", "1")
    srt = parity.cuda_scene(flat)
    monkeypatch.delenv("RPT_NEE_SORT")
    p = st.params(seed=31, flags=2)
    f0, c0 = plain.render_pt(p)
    f1, c1 = srt.render_pt(p)
    for k in ("segments", "bounce_rays", "shadow_rays", "shadow_rays_traced", "env_hits", "nee_vertices", "walk_nodes", "walk_tris"):
        assert getattr(c0, k) == getattr(c1, k), (name, k, getattr(c0, k), getattr(c1, k))
    assert c0.shadow_rays > 0
    ok = np.isfinite(f0)
    assert np.array_equal(ok, np.isfinite(f1))
    assert np.allclose(f0[ok], f1[ok], rtol=1e-5, atol=1e-9), (name, float(np.abs(f0[ok] - f1[ok]).max()))
    plain.close()
    srt.close()


def env_roundtrip_inputs(rng, n):
    """(u, v) pairs for the uv -> direction -> uv round trip: uniform, importance-map grid points (1000 and 4096 x 2048), near
    and at both poles, both sides of the azimuth seam. 9 regimes of n pairs."""
    f32 = np.float32
    parts = [
        (rng.random(n, dtype=f32), rng.random(n, dtype=f32)),
        ((rng.integers(0, 1001, n) / 1000).astype(f32), (rng.integers(0, 1001, n) / 1000).astype(f32)),
        ((rng.integers(0, 4097, n) / 4096).astype(f32), (rng.integers(0, 2049, n) / 2048).astype(f32)),
        (rng.random(n, dtype=f32), (rng.random(n, dtype=f32) * f32(2e-3)).astype(f32)),
        (rng.random(n, dtype=f32), (rng.random(n, dtype=f32) * f32(2e-5)).astype(f32)),
        (rng.random(n, dtype=f32), (f32(1) - rng.random(n, dtype=f32) * f32(1e-3)).astype(f32)),
        ((rng.random(n, dtype=f32) * f32(1e-4)).astype(f32), rng.random(n, dtype=f32)),
        ((f32(1) - rng.random(n, dtype=f32) * f32(1e-4)).astype(f32), rng.random(n, dtype=f32)),
        (rng.random(n, dtype=f32), rng.integers(0, 2, n).astype(f32)),  # the poles themselves (x = y = +-0: atan2 goes by the signs)
    ]
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])


def test_env_roundtrip_fast_path():
    """uv_roundtrip_unrotated_cr (the HDR environment's uv -> direction -> uv round trip without f64 atan2 / acos, used when the
    environment is unrotated) returns the f32 values of the libm path it replaces: 18 M inputs over every regime; the two device
    paths may differ only where the two f64 evaluations straddle an f32 rounding boundary (both are ~1 ulp of f64 from the exact
    value), which is allowed for at most 1 input in a million and by at most one f32 ulp."""
    p = parity.pkg()
    lib = p.ffi.load_library()
    u, v = env_roundtrip_inputs(np.random.default_rng(3), 2_000_000)
    u0, v0 = p.ffi.debug_env_roundtrip(lib, 0, u, v, fast=False)
    u1, v1 = p.ffi.debug_env_roundtrip(lib, 0, u, v, fast=True)
    assert np.isfinite(u1).all() and np.isfinite(v1).all()
    for a, b, what in ((u0, u1, "u"), (v0, v1, "v")):
        diff = a != b
        assert diff.mean() <= 1e-6, (what, int(diff.sum()))
        if diff.any():
            ulp = np.abs(a[diff].view(np.int32).astype(np.int64) - b[diff].view(np.int32).astype(np.int64))
            assert ulp.max() <= 1, (what, int(ulp.max()))


@pytest.mark.parametrize("name", ["hdri2", "hdri"])
def test_env_fast_roundtrip_renders_the_same_film(monkeypatch, name):
    """RPT_ENV_FAST=0 keeps the libm round trip: same counters, same film (identical texels are read)."""
    world, st, flat = parity.load_scene(name, 192, 108, 8)
    monkeypatch.setenv("RPT_ENV_FAST", "0")
    slow = parity.cuda_scene(flat)
    monkeypatch.delenv("RPT_ENV_FAST")
    fast = parity.cuda_scene(flat)
    p = st.params(seed=37)
    f0, c0 = slow.render_pt(p)
    f1, c1 = fast.render_pt(p)
    for k in ("segments", "bounce_rays", "shadow_rays", "shadow_rays_traced", "env_hits", "nee_vertices"):
        assert getattr(c0, k) == getattr(c1, k), (name, k)
    ok = np.isfinite(f0)
    assert np.array_equal(ok, np.isfinite(f1)) and np.allclose(f0[ok], f1[ok], rtol=1e-5, atol=1e-9)
    slow.close()
    fast.close()


@pytest.mark.parametrize("name", ["cornell", "furnace", "hdri2", "test_nee_sphere", "rtiow2", "sun_test", "orb_caustic"])
def test_one_level_walk_equals_two_level(monkeypatch, name):
    """TRAV_BVH_FLAT (scenes without a transformed mesh instance get a walk with the instance-local ray and the BLAS bookkeeping
    compiled out; RPT_FLAT=0 keeps the two-level walk) visits the same nodes in the same order: identical hit ids, BVH work
    counters and path counters, film equal up to the order of the energy atomics."""
    world, st, flat = parity.load_scene(name, 192, 108, 4)
    monkeypatch.setenv("RPT_FLAT", "0")
    two = parity.cuda_scene(flat)
    monkeypatch.delenv("RPT_FLAT")
    one = parity.cuda_scene(flat)
    p = st.params(seed=41, flags=2)
    f2, c2 = two.render_pt(p)
    f1, c1 = one.render_pt(p)
    for k in ("segments", "bounce_rays", "shadow_rays", "shadow_rays_traced", "env_hits", "nee_vertices", "walk_nodes", "walk_tris", "walk_insts",
              "shadow_nodes", "shadow_tris", "shadow_insts"):
        assert getattr(c2, k) == getattr(c1, k), (name, k, getattr(c2, k), getattr(c1, k))
    ok = np.isfinite(f2)
    assert np.array_equal(ok, np.isfinite(f1)) and np.allclose(f2[ok], f1[ok], rtol=1e-5, atol=1e-9)
    a, b = two.trace_primary(p), one.trace_primary(p)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    rng = np.random.default_rng(7)
    lo, hi = (-0.9, 0.9) if name != "cornell" else (0.01, 0.54)
    o, d = axis_aligned_rays(rng, 9000, lo, hi)
    tmax = np.full(len(o), np.inf, np.float32)
    a, b = two.trace_rays(o, d, tmax), one.trace_rays(o, d, tmax)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    two.close()
    one.close()


@pytest.mark.parametrize("name", ["hdri2", "hdri"])
def test_imap_guide_tables_do_not_change_the_samples(monkeypatch, name):
    """RPT_IMAP_GUIDES=0 inverts the importance map's CDFs with the plain binary search over the whole row; the guide tables
    (default) bracket the same search: identical counters, film equal up to the order of the energy atomics."""
    world, st, flat = parity.load_scene(name, 192, 108, 8)
    monkeypatch.setenv("RPT_IMAP_GUIDES", "0")
    plain = parity.cuda_scene(flat)
    monkeypatch.delenv("RPT_IMAP_GUIDES")
    guided = parity.cuda_scene(flat)
    p = st.params(seed=47)
    f0, c0 = plain.render_pt(p)
    f1, c1 = guided.render_pt(p)
    for k in ("segments", "bounce_rays", "shadow_rays", "shadow_rays_traced", "env_hits", "nee_vertices"):
        assert getattr(c0, k) == getattr(c1, k), (name, k)
    ok = np.isfinite(f0)
    assert np.array_equal(ok, np.isfinite(f1)) and np.allclose(f0[ok], f1[ok], rtol=1e-5, atol=1e-9)
    plain.close()
    guided.close()


@pytest.mark.parametrize("name,w,h", [("cornell", 192, 108), ("instanced_monkeys", 192, 108), ("hdri2", 160, 88), ("kitchen_sink", 64, 4), ("gem", 8, 64)])
def test_tiled_slots_render_the_same_film(monkeypatch, name, w, h):
    """Slots of a frame run over 8 x 4 pixel tiles when the film's width is a multiple of 8 and its height of 4 (RPT_TILED=0:
    row-major). Samples are keyed by (pixel, sample) and accumulated per pixel: identical counters, film equal up to the order
    of the energy atomics, for one wave and for several."""
    world, st, flat = parity.load_scene(name, w, h, 6)
    monkeypatch.setenv("RPT_TILED", "0")
    rows = parity.cuda_scene(flat)
    f0, c0 = rows.render_pt(st.params(seed=53))
    monkeypatch.delenv("RPT_TILED")
    f1, c1 = rows.render_pt(st.params(seed=53))
    monkeypatch.setenv("RPT_WAVE_SLOTS_MAX", str(2 * w * h))  # three waves of 2 spp
    f2, c2 = rows.render_pt(st.params(seed=53))
    for c in (c1, c2):
        for k in ("camera_rays", "segments", "bounce_rays", "shadow_rays", "shadow_rays_traced", "env_hits", "nee_vertices"):
            assert getattr(c0, k) == getattr(c, k), (name, k)
    ok = np.isfinite(f0)
    for f in (f1, f2):
        assert np.array_equal(ok, np.isfinite(f)) and np.allclose(f0[ok], f[ok], rtol=1e-5, atol=1e-9), name
    rows.close()


@pytest.mark.parametrize("name", ["cornell", "kitchen_sink", "hdri2"])
def test_two_stream_half_waves_equal_single_stream(monkeypatch, name):
    """RPT_OVERLAP=1 (a wave cut into two half-waves on two streams; opt-in after measurement, profiles/r02_overlap.md) renders
    the same samples: identical counters, film equal up to the order of the f32 sums."""
    world, st, flat = parity.load_scene(name, 160, 90, 5)
    single = parity.cuda_scene(flat)
    f0, c0 = single.render_pt(st.params(seed=23))
    single.close()
    monkeypatch.setenv("RPT_OVERLAP", "1")
    two = parity.cuda_scene(flat)
    f1, c1 = two.render_pt(st.params(seed=23))
    two.close()
    for k in ("camera_rays", "segments", "bounce_rays", "shadow_rays", "shadow_rays_traced", "env_hits", "nee_vertices"):
        assert getattr(c0, k) == getattr(c1, k), (name, k)
    assert c1.kernel_launches > c0.kernel_launches  # (two launches per kernel and bounce)
    ok = np.isfinite(f0)
    assert np.array_equal(ok, np.isfinite(f1)) and np.allclose(f0[ok], f1[ok], rtol=1e-5, atol=1e-9)


def test_reference_parameter_ranges(scenes):
    """The reference takes any u16 for light_samples and max_bounces (parsing/config.rs:22-23); so does the library
    (round 1 rejected light_samples > 8 and max_bounces > 64)."""
    st, cs, os_ = scenes("cornell", 48, 27, 2)
    p = st.params(seed=4)
    p.light_samples, p.max_bounces = 11, 100
    fg, cg = cs.render_pt(p)
    fo, co = os_.render_pt(p)
    assert cg.camera_rays == co.camera_rays and abs(cg.segments - co.segments) <= 4
    assert parity.rel_mse(fg, fo) < 2e-3 and parity.mean_rel_diff(fg, fo) < 2e-3


def test_kernel_times_and_stats(scenes):
    """Instrumentation is a run-time opt-in (RptRenderParams.flags): off -> no per-kernel times, no BVH work counters;
    on -> both, and the film is the same."""
    st, cs, _ = scenes("cornell", 96, 54, 8)
    f0, c0 = cs.render_pt(st.params(seed=1))
    assert cs.kernel_times() == [] or all(k["ms"] == 0.0 for k in cs.kernel_times())
    assert c0.walk_nodes == 0 and c0.walk_tris == 0 and c0.kernel_launches > 0 and c0.device_ms > 0
    f1, c1 = cs.render_pt(st.params(seed=1, flags=3))
    names = [k["name"] for k in cs.kernel_times()]
    assert "k_trace" in names and "k_shadow" in names and "k_film" in names
    assert c1.walk_tris + c1.walk_nodes > 0 and c1.segments == c0.segments
    assert np.allclose(f0, f1, rtol=1e-6, atol=1e-9)
    s = cs.stats()
    assert s["triangles"] == 30 and s["instances"] == 4


def _device_count(pkg):
    import ctypes as ct

    n = ct.c_int()
    assert pkg.ffi.load_library().rpt_device_count(ct.byref(n)) == 0
    return n.value


@pytest.mark.parametrize("method", ["peer", "nccl"])
def test_multi_device_render_behind_the_c_abi(pkg, monkeypatch, method):
    """rpt_multi_*: the spp split, one host thread per device and the film exchange live inside the library (the reference is a
    single process: src/bin/main.rs:59-68,170). N devices x k spp reproduce the samples of one device x N k spp, so the
    multi-device film equals the single-device film up to the order of the f32 sums (test_spp_split_equals_whole through
    the C ABI). Runs on every device count available: with one GPU it covers the n = 1 path, with >= 2 both exchanges
    (fused NVLink peer kernel; NCCL reduce)."""
    ndev = _device_count(pkg)
    devices = list(range(min(ndev, 8)))
    if method == "nccl":
        if ndev < 2:
            pytest.skip("the NCCL exchange needs >= 2 devices")
        monkeypatch.setenv("RPT_MULTI_REDUCE", "nccl")
    world, st, flat = parity.load_scene("cornell", 160, 90, 12)
    single = parity.cuda_scene(flat, 0)
    f1, c1 = single.render_pt(st.params(seed=77))
    single.close()
    ms = pkg.ffi.MultiScene(pkg.ffi.load_library(), flat, devices)
    fm, cm = ms.render_pt(st.params(seed=77))
    t = ms.times.as_dict()
    assert t["devices"] == len(devices) and (len(devices) == 1 or t["method"] == method)
    assert cm.camera_rays == c1.camera_rays and cm.segments == c1.segments and cm.shadow_rays == c1.shadow_rays
    assert np.allclose(fm, f1, rtol=2e-5, atol=1e-8), float(np.abs(fm - f1).max())
    assert (fm[..., 3] == 0).all()
    # the one-shot form and an uneven split (12 spp over the devices with a different total divisor)
    p = st.params(seed=78, spp=7, spp_total=7)
    f7, _ = ms.render_pt(p)
    import ctypes as ct

    out = np.zeros_like(f7)
    arr = (ct.c_int * len(devices))(*devices)
    cnt = pkg.ffi.RptCounters()
    rc = ms.lib.rpt_render_pt_multi(ct.byref(flat.desc), arr, len(devices), ct.byref(p), out.ctypes.data_as(ct.c_void_p), ct.byref(cnt))
    assert rc == 0, ms.lib.rpt_last_error()
    assert np.allclose(out, f7, rtol=2e-5, atol=1e-8) and cnt.camera_rays == 160 * 90 * 7
    ms.close()


def test_multi_device_importance_map_and_errors(pkg):
    """An Unbaked HDR importance map is baked on every replica (CudaRenderer.make_multi_scene); bad device lists fail loudly."""
    ndev = _device_count(pkg)
    devices = list(range(min(ndev, 4)))
    world, st, flat = parity.load_scene("hdri2", 96, 54, 8)
    r = pkg.CudaRenderer(device=0, seed=3)
    ms = r.make_multi_scene(world, st.wavelength_bounds, devices)
    fm, cm = ms.render_pt(st.params(seed=3))
    ms.close()
    sc = parity.cuda_scene(flat, 0)
    f1, c1 = sc.render_pt(st.params(seed=3))
    sc.close()
    assert cm.segments == c1.segments and parity.rel_mse(fm, f1) < 1e-8
    with pytest.raises(pkg.ffi.RptError):
        pkg.ffi.MultiScene(pkg.ffi.load_library(), flat, [0, 0])
    with pytest.raises(pkg.ffi.RptError):
        pkg.ffi.MultiScene(pkg.ffi.load_library(), flat, [ndev + 3])


@pytest.mark.parametrize("name", ["gem", "hdri2", "instanced_monkeys"])
def test_same_stream_at_size_binned_queues_and_two_waves(monkeypatch, name):
    """BASELINE configs #3-#5 at a size that exercises what the 96x54 parity renders never reach: class queues of
    more than BIN_MIN_ITEMS = 4 M entries (the origin-cell-binned shadow-queue appends of k_shade_surface) and a render split
    into two waves (spp_chunk < spp; RPT_WAVE_SLOTS_MAX caps a wave at 8 spp of this 1280x720 film = 7.4 M paths, the job is
    16 spp). Same Philox streams on both sides; counters equal up to the handful of paths where an fp32 rounding flips a
    branch; at most 0.5 % of the pixels differ at all and the mean agrees to 2e-3, as in test_same_stream_images. relMSE
    <= 2e-3 as there, except for the gem: its GGX has alpha = 0.0004 (data/lib_materials.toml:78; ggx_d ~ 1 / alpha^2,
    SURVEY §7 hard part v), so the few samples whose path flips on a last-ulp difference of sincosf / powf land on caustic
    fireflies thousands of times the pixel mean, and relMSE - a sum of squares - is set by a handful of them: bound 5e-2."""
    w, h, spp = 1280, 720, 16
    world, st, flat = parity.load_scene(name, w, h, spp)
    monkeypatch.setenv("RPT_WAVE_SLOTS_MAX", str(w * h * 8))
    cs, os_ = parity.cuda_scene(flat), parity.oracle_scene(flat)
    p = st.params(seed=13)
    fg, cg = cs.render_pt(p)
    fo, co = os_.render_pt(p)
    cs.close()
    os_.close()
    assert cg.camera_rays == co.camera_rays == w * h * spp
    ok = np.isfinite(fo).all(axis=2)
    assert np.isfinite(fg[ok]).all() and ok.mean() > 0.9999
    yg, yo = fg[ok][..., 1], fo[ok][..., 1]
    differing = float(np.mean(np.abs(yg - yo) > 1e-4 * np.maximum(np.abs(yo), 1e-6)))
    r = parity.rel_mse(fg[ok], fo[ok])
    print(f"{name}: relMSE {r:.3e}, pixels differing {differing:.2e}, mean-Y rel diff {parity.mean_rel_diff(fg[ok], fo[ok]):.2e}")
    assert differing <= 5e-3, (name, differing)
    assert parity.mean_rel_diff(fg[ok], fo[ok]) < 2e-3, (name, yg.mean(), yo.mean())
    assert r < (5e-2 if name == "gem" else 2e-3), (name, r)
    for k in ("bounce_rays", "shadow_rays", "env_hits", "segments"):
        a, b = getattr(cg, k), getattr(co, k)
        assert abs(a - b) <= max(4, 2e-4 * b), (name, k, a, b)


CONVERGED = ["cornell", "furnace", "gem", "hdri2", "instanced_monkeys"]


@pytest.mark.parametrize("name", CONVERGED)
def test_converged_against_committed_oracle(name):
    """Parity check (c) of BASELINE.json at the planned scale: a GPU render of the full view at 256 x 256 with
    4096 spp and its OWN seed against the committed 4096-spp oracle film of the same view (tests/golden/converged_<scene>.npz,
    tools/make_converged.py: two independent 2048-spp oracle halves; `floor` = relMSE between the halves).
    Independent estimates of the same image: relMSE(GPU, oracle) is expected at floor / 2 (each half carries twice the
    variance of a 4096-spp estimate). Stated tolerance: relMSE <= 0.75 x floor (1.5x the expectation), and the mean of
    every XYZ channel within 1 % (hdri2, whose noise floor is firefly-dominated: 3 %)."""
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"converged_{name}.npz")
    if not os.path.exists(path):
        pytest.skip("converged fixture not generated (tools/make_converged.py)")
    gold = np.load(path)
    ref, floor = gold["film"], float(gold["floor"])
    world, st, flat = parity.load_scene(name, 256, 256, 4096)
    cs = parity.cuda_scene(flat)
    fg, cg = cs.render_pt(st.params(seed=303))
    cs.close()
    ok = np.isfinite(ref).all(axis=2) & np.isfinite(fg[..., :3]).all(axis=2)
    assert ok.mean() > 0.9995  # (a NaN sample poisons its pixel on either side: the reference paints those MAUVE)
    r = parity.rel_mse(fg[ok][..., :3], ref[ok])
    mean_g, mean_o = fg[ok][..., :3].mean(axis=0), ref[ok].mean(axis=0)
    print(f"{name}: relMSE(GPU 4096 spp, oracle 4096 spp) = {r:.3e}, noise floor (oracle half vs half) = {floor:.3e}, ratio {r / floor:.3f}; "
          f"mean XYZ gpu {mean_g} oracle {mean_o}")
    assert r <= 0.75 * floor, (name, r, floor)
    tol = 0.03 if name == "hdri2" else 0.01
    assert np.all(np.abs(mean_g - mean_o) <= tol * np.abs(mean_o)), (name, mean_g, mean_o)
