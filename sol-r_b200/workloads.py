"""The configurations BASELINE.json names, as bench.py and the parity tests run them (synthetic scenes of the named sizes:
sol-r_b200/scenes.py; the reference's media files do not travel)."""
from . import scenes, wire

WORKLOADS = {
    # configs[1]: the configuration the headline metric is quoted on
    "config2": dict(name="config2_molecule_216k_primitives_1920x1080_glFull_3_bounces", size=(1920, 1080), nit=3,
                    scene=lambda: scenes.config2(), iterations=[0], limits=None, capacity=None, camera=wire.CT_PERSPECTIVE),
    # configs[2]: ~1 M triangles, reflections / refractions depth 5 (capacity lifted so that every triangle is kept; the
    # reference's 2.5 M-box array holds 741 k of them)
    "config3": dict(name="config3_mesh_1M_triangles_1920x1080_glFull_5_bounces", size=(1920, 1080), nit=5,
                    scene=lambda: scenes.triangle_mesh(1_000_000), iterations=[0], limits=None,
                    capacity=(16_000_000, 4_000_000), camera=wire.CT_PERSPECTIVE),
    # configs[3]: 1 M spheres at 4K, 4 accumulated samples = iterations 10..13 (10 starts the accumulation, 11-13 add a jittered sample each)
    "config4": dict(name="config4_1M_spheres_3840x2160_glFull_3_bounces_accumulated_samples_iterations_10_to_13", size=(3840, 2160), nit=3,
                    scene=lambda: scenes.config4(), iterations=[10, 11, 12, 13], limits=(3840, 2160),
                    capacity=(16_000_000, 4_000_000), camera=wire.CT_PERSPECTIVE),
    # configs[4]: anaglyph 4K render of the molecule scene, 16-frame progressive refinement (iterations 0..15)
    "config5": dict(name="config5_anaglyph_molecule_3840x2160_16_frame_progressive_iterations_0_to_15", size=(3840, 2160), nit=3,
                    scene=lambda: scenes.config2(), iterations=list(range(16)), limits=(3840, 2160), capacity=None,
                    camera=wire.CT_ANAGLYPH),
}


def scene_info(key, width=None, height=None):
    wl = WORKLOADS[key]
    W, H = (width, height) if width else wl["size"]
    si = wire.default_scene_info(W, H, graphics_level=wire.GL_FULL, nb_ray_iterations=wl["nit"])
    si.cameraType = wl["camera"]
    if wl["camera"] == wire.CT_ANAGLYPH:
        si.eyeSeparation = 380.0
    si.maxPathTracingIterations = max(wl["iterations"]) + 1
    return si
