#!/usr/bin/env python3
"""Bakes portable scene blobs (scenes/*.npz) from the reference's own TOML / CSV / OBJ files plus the
synthesised fixtures, so that the GPU box — which has no /root/reference — renders the same scenes.

Run here (the container with /root/reference): `python tools/bake_scenes.py`.
Each blob = flattened World + the PT settings of its config (rust-pathtracer_b200/blob.py).

BASELINE.json configs:
  C1 cornell        data/config_test_cornell_box.toml (camera_id fixed to "main")
  C2 furnace        data/config_test_whitefurnace.toml; furnace_exact = the exact-1.0 variant (SURVEY A9 ii)
  C3 gem            data/scenes/cornell_box_diamond_gem.toml (synthetic low-res HDR, p_env = 0)
  C4 hdri           data/scenes/hdri_test.toml with a synthetic HDR + baked importance map
  C5 instanced      2388 instances of data/meshes/monkey.obj (10.0 M triangles), generated scene
plus in-tree scenes that cover every material / light / environment kind.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
from rust_pathtracer_b200 import blob, loader  # noqa: E402
from rust_pathtracer_b200.renderer import PTSettings  # noqa: E402

OUT = os.path.join(ROOT, "scenes")
GEN = os.path.join(ROOT, "fixtures", "data", "scenes")


def make_config(scene_file, width, height, spp, min_bounces, max_bounces, light_samples, only_direct=False, wavelength_bounds=None):
    rs = loader.RenderSettings(
        filename="beauty", width=width, height=height, integrator_type="PT", light_samples=light_samples, medium_aware=False,
        min_bounces=min_bounces, max_bounces=max_bounces, hwss=False, threads=None, min_samples=spp, camera_id="main",
        russian_roulette=True, only_direct=only_direct, wavelength_bounds=wavelength_bounds, premultiply=None)
    return loader.Config(scene_file=scene_file, renderer={"type": "Naive"}, render_settings=[rs])


def from_reference_config(path, **overrides):
    cfg = loader.get_config(path)
    for rs in cfg.render_settings:
        rs.camera_id = "main"  # the shipped configs name cameras no scene defines
        for k, v in overrides.items():
            setattr(rs, k, v)
    return cfg


def write_generated_scenes():
    os.makedirs(GEN, exist_ok=True)
    libs = 'curves = "data/lib_curves.toml"\ntextures = "data/lib_textures.toml"\nmaterials = "data/lib_materials.toml"\nmeshes = "data/lib_meshes.toml"\n'
    # exact furnace (SURVEY A9 ii): Constant flat_one env, unit Lambertian sphere with albedo clamped to 1
    with open(os.path.join(GEN, "furnace_exact.toml"), "w") as f:
        f.write('curves = "data/lib_curves.toml"\nmeshes = "data/lib_meshes.toml"\n' + """
env_sampling_probability = 1.0
[environment]
type = "Constant"
strength = 1.0
color = "flat_one"

[materials.lambertian_unit]
type = "Lambertian"
texture_id = "unit_albedo"

[[textures.unit_albedo]]
type = "Texture1"
filename = "data/textures/single_pixel.png"
curve = { type = "Flat", strength = 2.0 }

[[instances]]
material_name = "lambertian_unit"
[instances.aggregate]
type = "Sphere"
radius = 1.0
origin = [0.0, 0.0, 0.0]

[[cameras]]
type = "SimpleCamera"
name = "main"
look_from = [-5.0, 0.0, 0.0]
look_at = [0.0, 0.0, 0.0]
aperture_diameter = 0.001
aperture = { type = "Circular" }
focal_distance = 5.0
vfov = 30.0
""")
    # kitchen sink: every primitive / light / material kind the GPU path supports, with transforms, in one small scene
    with open(os.path.join(GEN, "kitchen_sink.toml"), "w") as f:
        f.write('curves = "data/lib_curves.toml"\nmeshes = "data/lib_meshes.toml"\n' + """
env_sampling_probability = 0.3
[environment]
type = "Constant"
strength = 0.004
color = "D65"

[[textures.checker]]
type = "Texture4"
filename = "data/textures/simple.png"
curves = ["srgb_r", "srgb_g", "srgb_b", "flat_zero"]
[[textures.lambertian_white]]
type = "Texture1"
filename = "data/textures/single_pixel.png"
curve = "cornell_white"
[[textures.lambertian_red]]
type = "Texture1"
filename = "data/textures/single_pixel.png"
curve = "cornell_red"

# floor: textured two-sided rect (uv lookup through Texture4)
[[instances]]
material_name = "lambertian_textured_simple"
[instances.aggregate]
type = "Rect"
size = [8.0, 8.0]
origin = [0.0, 0.0, -1.0]
normal = "Z"
two_sided = true

# a rotated, non-uniformly scaled rect light (transform on an analytic light: instance.rs:134-170)
[[instances]]
material_name = "diffuse_light_warm"
[instances.aggregate]
type = "Rect"
size = [1.0, 1.0]
origin = [0.0, 0.0, 0.0]
normal = "Y"
two_sided = true
[instances.transform]
scale = [1.5, 1.0, 0.75]
translate = [0.0, 3.0, 1.0]
[[instances.transform.rotate]]
axis = [0.0, 0.0, 1.0]
angle = 25

# a sphere light and a one-sided disk light with a sharp lobe
[[instances]]
material_name = "diffuse_light_xenon"
[instances.aggregate]
type = "Sphere"
radius = 0.3
origin = [-2.0, -1.0, 1.5]

[[instances]]
material_name = "sharp_light_warm"
[instances.aggregate]
type = "Disk"
radius = 0.6
origin = [0.8, -1.0, 1.6]
two_sided = false

# non-light disks: one-sided and two-sided, one of them tilted
[[instances]]
material_name = "lambertian_red"
[instances.aggregate]
type = "Disk"
radius = 0.8
origin = [2.0, 1.0, -0.2]
two_sided = true
[[instances]]
material_name = "ggx_copper"
[instances.aggregate]
type = "Disk"
radius = 0.7
origin = [0.0, 0.0, 0.0]
two_sided = false
[instances.transform]
translate = [-2.2, 1.6, 0.3]
[[instances.transform.rotate]]
axis = [1.0, 0.0, 0.0]
angle = 40

# spheres: metal, rough glass, dispersive glass, an ellipsoid through non-uniform scale
[[instances]]
material_name = "ggx_gold"
[instances.aggregate]
type = "Sphere"
radius = 0.6
origin = [0.3, 1.3, -0.4]
[[instances]]
material_name = "ggx_glass_rough"
[instances.aggregate]
type = "Sphere"
radius = 0.5
origin = [1.4, 0.2, -0.5]
[[instances]]
material_name = "ggx_glass_dispersive"
[instances.aggregate]
type = "Sphere"
radius = 0.5
origin = [0.0, 0.0, 0.0]
[instances.transform]
scale = [1.0, 0.6, 1.4]
translate = [-1.2, -0.4, -0.3]
[[instances.transform.rotate]]
axis = [0.0, 1.0, 0.0]
angle = 30

# meshes with shading normals: transformed (two-level traversal) and untransformed (flattened into the TLAS)
[[instances]]
material_name = "ggx_moissanite"
[instances.aggregate]
type = "Mesh"
name = "brilliant_diamond"
[instances.transform]
scale = [0.5, 0.5, 0.5]
translate = [0.6, -1.2, -0.6]
[[instances.transform.rotate]]
axis = [1.0, 1.0, 0.0]
angle = 35
[[instances]]
material_name = "lambertian_white"
[instances.aggregate]
type = "Mesh"
name = "prism"

[materials.lambertian_textured_simple]
type = "Lambertian"
texture_id = "checker"
[materials.lambertian_white]
type = "Lambertian"
texture_id = "lambertian_white"
[materials.lambertian_red]
type = "Lambertian"
texture_id = "lambertian_red"
[materials.diffuse_light_warm]
type = "DiffuseLight"
sidedness = "Dual"
emit_color = "blackbody_3000k_x5"
bounce_color = "flat_78"
[materials.diffuse_light_xenon]
type = "DiffuseLight"
sidedness = "Forward"
emit_color = "xenon_x5"
bounce_color = "flat_78"
[materials.sharp_light_warm]
type = "SharpLight"
sidedness = "Reverse"
sharpness = 12.0
emit_color = "blackbody_3000k_x5"
bounce_color = "flat_78"
[materials.ggx_gold]
type = "GGX"
permeability = 0.0
alpha = 0.05
eta_o = "air_ior"
eta = { type = "TabulatedCSV", filename = "data/curves/csv/gold.csv", column = 1, domain_mapping = { x_scale = 1000.0 }, interpolation_mode = "Cubic" }
kappa = { type = "TabulatedCSV", filename = "data/curves/csv/gold.csv", column = 2, domain_mapping = { x_scale = 1000.0 }, interpolation_mode = "Cubic" }
[materials.ggx_copper]
type = "GGX"
permeability = 0.0
alpha = 0.1
eta_o = "air_ior"
eta = { type = "TabulatedCSV", filename = "data/curves/csv/copper-mcpeak.csv", column = 1, domain_mapping = { x_scale = 1000.0 }, interpolation_mode = "Cubic" }
kappa = { type = "TabulatedCSV", filename = "data/curves/csv/copper-mcpeak.csv", column = 2, domain_mapping = { x_scale = 1000.0 }, interpolation_mode = "Cubic" }
[materials.ggx_glass_rough]
type = "GGX"
permeability = 1.0
alpha = 0.2
kappa = "flat_zero"
eta_o = "air_ior"
eta = { type = "Cauchy", a = 1.4, b = 10000.0 }
[materials.ggx_glass_dispersive]
type = "GGX"
permeability = 1.0
alpha = 0.01
kappa = "flat_zero"
eta_o = "air_ior"
eta = { type = "Cauchy", a = 1.4, b = 30000.0 }
[materials.ggx_moissanite]
type = "GGX"
permeability = 1.0
alpha = 0.01
kappa = "flat_zero"
eta_o = "air_ior"
eta = { type = "Cauchy", a = 2.4, b = 34000.0 }

[[cameras]]
type = "SimpleCamera"
name = "main"
look_from = [-6.0, -3.0, 2.5]
look_at = [0.0, 0.0, -0.2]
aperture_diameter = 0.05
aperture = { type = "Circular" }
focal_distance = 6.5
vfov = 50.0

[[cameras]]
type = "PanoramaCamera"
name = "pano"
look_from = [0.0, -0.5, 1.0]
look_at = [1.0, 0.2, 0.2]
fov = [360.0, 160.0]
""")
    # C3 as BASELINE.json states it ("dispersive Cauchy dielectric and spectral metals"): the shipped gem scene
    # (data/scenes/cornell_box_diamond_gem.toml, verbatim) plus the five tabulated-eta/kappa metal spheres of
    # data/scenes/cornell_box_metals_and_dielectrics.toml:89-119 (gold, iron, copper, platinum, lead), the quad at that scene's
    # own positions and the gold sphere moved from the centre (where the gem sits) to behind it.
    ref_gem = open("/root/reference/data/scenes/cornell_box_diamond_gem.toml").read()
    head, cams = ref_gem.split("[[cameras]]", 1)
    spheres = "".join(f"""
[[instances]]
material_name = "{m}"
[instances.aggregate]
type = "Sphere"
radius = 0.2
origin = [{x}, {y}, -0.79]
""" for m, x, y in (("ggx_gold", 0.55, 0.0), ("ggx_iron", 0.51, 0.51), ("ggx_copper", 0.51, -0.51), ("ggx_platinum", -0.51, 0.71), ("ggx_lead", -0.51, -0.71)))
    with open(os.path.join(GEN, "gem_metals.toml"), "w") as f:
        f.write(head + spheres + "\n[[cameras]]" + cams)

    # instanced monkeys (C5): 48 x 50 grid (minus 12) = 2388 instances, Philox-free numpy RNG seed 5
    rng = np.random.default_rng(5)
    mats = ["lambertian_white", "ggx_gold", "ggx_copper", "ggx_glass"]
    lines = [libs.replace('meshes = "data/lib_meshes.toml"\n', ""), """
env_sampling_probability = 0.5
[environment]
type = "Sun"
strength = 30.0
angular_diameter = 0.05
sun_direction = [0.3, 0.2, 1.0]
color = "blackbody_5000k"

[meshes.monkey]
filename = "data/meshes/monkey.obj"

[[instances]]
material_name = "lambertian_white"
[instances.aggregate]
type = "Rect"
size = [140.0, 140.0]
origin = [0.0, 0.0, -1.2]
normal = "Z"
two_sided = true

[[instances]]
material_name = "diffuse_light"
[instances.aggregate]
type = "Rect"
size = [40.0, 40.0]
origin = [0.0, 0.0, 40.0]
normal = "Z"
two_sided = true
"""]
    count = 0
    for gy in range(50):
        for gx in range(48):
            if count >= 2388:
                break
            axis = rng.normal(size=3)
            axis /= np.linalg.norm(axis)
            ang = float(rng.uniform(0, 360))
            sc = float(rng.uniform(0.8, 1.2))
            lines.append(f"""
[[instances]]
material_name = "{mats[count % 4]}"
[instances.transform]
scale = [{sc:.5f}, {sc:.5f}, {sc:.5f}]
rotate = [{{ axis = [{axis[0]:.5f}, {axis[1]:.5f}, {axis[2]:.5f}], angle = {ang:.3f} }}]
translate = [{(gx - 23.5) * 2.8:.4f}, {(gy - 24.5) * 2.8:.4f}, 0.0]
[instances.aggregate]
type = "Mesh"
name = "monkey"
""")
            count += 1
    lines.append("""
[[cameras]]
type = "SimpleCamera"
name = "main"
look_from = [-95.0, -60.0, 45.0]
look_at = [0.0, 0.0, 0.0]
aperture_diameter = 0.001
aperture = { type = "Circular" }
focal_distance = 100.0
vfov = 50.0
""")
    with open(os.path.join(GEN, "instanced_monkeys.toml"), "w") as f:
        f.write("".join(lines))


def bake(name, cfg, scene_file=None, num_lambda=1024, bake_importance_map=True):
    world = loader.construct_world(cfg, scene_file, bake_importance_map=bake_importance_map)
    rs = cfg.render_settings[0]
    st = PTSettings.from_render_settings(rs, cfg.camera_names_to_index[rs.camera_id])
    path = os.path.join(OUT, name + ".npz")
    blob.save_world(path, world, st.to_dict(), st.wavelength_bounds[0], st.wavelength_bounds[1], num_lambda)
    tris = sum(len(world.meshes[i.mesh].indices) for i in world.instances if i.mesh >= 0)
    print(f"{name:28s} {os.path.getsize(path) / 1024:9.1f} KiB  instances={len(world.instances)} tris(instanced)={tris} lights={len(world.lights)}")


def bake_hdri_small():
    import make_fixtures

    path = os.path.join(ROOT, "fixtures", "data", "hdri", "machine_shop_03_4k.hdr")
    make_fixtures.write_synthetic_map(path, 1024, 41)
    os.remove(path + ".recipe.json")  # texels stored in the blob
    try:
        bake("hdri", from_reference_config("data/config_test_lighting_hdri.toml", width=3840, height=2160))
    finally:
        make_fixtures.write_synthetic_map(path, 4096, 41)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    write_generated_scenes()
    # the synthetic HDRs must exist before the HDRI scenes are parsed (small ones for the blobs)
    import subprocess
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_fixtures.py"), "--hdri"])

    jobs = {
        # BASELINE configs[0]: Cornell, PT, 1080p @ 16 spp (the config file says 1080x1080 @ 128)
        "cornell": lambda: bake("cornell", from_reference_config("data/config_test_cornell_box.toml", width=1920, height=1080, min_samples=16)),
        "furnace": lambda: bake("furnace", from_reference_config("data/config_test_whitefurnace.toml")),
        "furnace_exact": lambda: bake("furnace_exact", make_config("data/scenes/furnace_exact.toml", 256, 256, 64, 1, 8, 4)),
        "gem": lambda: bake("gem", make_config("data/scenes/gem_metals.toml", 1920, 1080, 1024, 4, 16, 2)),
        # BASELINE configs[3]: data/config_test_lighting_hdri.toml at 4K. Its own scene file (hdri_test.toml: one Lambertian
        # sphere) is "hdri"; "hdri2" is data/scenes/hdri_test_2.toml (GGX gold / copper / dispersive glass spheres, p_env 0.5,
        # 1000 x 1000 importance map) under the same render settings - the scene the config's description names. Both sample a
        # synthetic 4096 x 2048 map (134 MB of RGBA f32 texels, stored in the blob as a recipe) and leave the importance map
        # Unbaked: it is baked on the device when the scene is created (naive.rs:469-487).
        # ("hdri" is the round-1 blob: the same scene file over a 1024 x 512 map with the importance map baked on the host
        # and stored; the parity / golden tests of the importance-map code keep using it)
        "hdri": bake_hdri_small,
        "hdri_4k": lambda: bake("hdri_4k", from_reference_config("data/config_test_lighting_hdri.toml", width=3840, height=2160), bake_importance_map=False),
        "hdri2": lambda: bake("hdri2", from_reference_config("data/config_test_lighting_hdri.toml", width=3840, height=2160),
                              scene_file="data/scenes/hdri_test_2.toml", bake_importance_map=False),
        "instanced_monkeys": lambda: bake("instanced_monkeys", make_config("data/scenes/instanced_monkeys.toml", 3840, 2160, 16, 2, 6, 2)),
        # in-tree scenes covering the remaining materials / lights / environments
        "test_nee_sphere": lambda: bake("test_nee_sphere", make_config("data/scenes/test_nee_sphere.toml", 512, 512, 32, 2, 8, 2)),
        "orb_caustic": lambda: bake("orb_caustic", make_config("data/scenes/cornell_box_single_orb_caustic.toml", 512, 512, 32, 2, 10, 2)),
        "sun_test": lambda: bake("sun_test", make_config("data/scenes/sun_test.toml", 512, 512, 32, 2, 8, 2)),
        "parallel_prism": lambda: bake("parallel_prism", make_config("data/scenes/cornell_box_parallel_prism.toml", 512, 512, 32, 2, 10, 2)),
        "lighting_north": lambda: bake("lighting_north", from_reference_config("data/config_test_lighting_north.toml")),
        "kitchen_sink": lambda: bake("kitchen_sink", make_config("data/scenes/kitchen_sink.toml", 512, 384, 32, 2, 10, 3)),
        "rtiow2": lambda: bake("rtiow2", make_config("data/scenes/test_rtiow_scene_2.toml", 512, 512, 32, 2, 8, 2)),
    }
    for name, job in jobs.items():
        if args.only and name not in args.only:
            continue
        try:
            job()
        except Exception as e:  # keep going: some in-tree scenes reference assets that are not shipped
            print(f"{name:28s} SKIPPED: {type(e).__name__}: {e}")


if __name__ == "__main__":
    main()
