"""Small helpers shared by the host-logic tests."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def world_from_scene(pkg, scene_file, width=64, height=64, spp=4):
    import bake_scenes

    cfg = bake_scenes.make_config(scene_file, width, height, spp, 2, 8, 2)
    world = pkg.loader.construct_world(cfg)
    st = pkg.PTSettings.from_render_settings(cfg.render_settings[0], 0)
    return world, st
