"""Small deterministic cases shared by the golden generator (tests/golden/make_golden.py, which runs the
REFERENCE) and the tests that check the oracle / the CUDA engine against those goldens.

Each case: a scene (setter calls), a SceneInfo, a camera, a random table spec and a list of frames
(pathTracingIteration values rendered in sequence with the per-pixel state carried over)."""
import numpy as np

from _solr_b200_import import solr_b200  # noqa: F401
from solr_b200 import scenes, wire

W, H = 96, 72
RANDOM_TABLE = wire.REF_MAX_BITMAP_SIZE


def randoms(seed):
    """Same distribution as GPUKernel::render_begin's table (GPUKernel.cpp:2724-2726), fixed seed."""
    if seed is None:
        return np.zeros(RANDOM_TABLE, np.float32)
    rng = np.random.Generator(np.random.PCG64(seed))
    return (0.000005 * (rng.integers(0, 2000, size=RANDOM_TABLE) - 1000)).astype(np.float32)


def _pack(name, mats, prims, **kw):
    return scenes._pack(name, mats, prims, **kw)


def _light(mats):
    mats.append(scenes._light_material())
    return (wire.PT_SPHERE, (-5000.0, 5000.0, -15000.0) + (0,) * 6 + (1.0, 0, 0), len(mats) - 1)


def mixed_scene():
    """Every primitive type of the path: sphere, ellipsoid, cylinder, cone, triangle, planes, checkerboard."""
    M = scenes.material
    mats = [M(0.9, 0.2, 0.2, reflection=0.4, propagation=50000.0), M(0.2, 0.8, 0.3, propagation=50000.0),
            M(0.3, 0.4, 0.9, transparency=0.6, refraction=1.33, opacity=0.2, reflection=0.3, propagation=50000.0),
            M(0.8, 0.8, 0.2, propagation=50000.0, procedural=1), M(0.5, 0.5, 0.5, propagation=50000.0),
            M(0.7, 0.3, 0.7, transparency=0.5, refraction=1.1, fast_transparency=1, propagation=50000.0)]
    prims = [
        (wire.PT_SPHERE, (-2500, 800, 0) + (0,) * 6 + (900, 0, 0), 0),
        (wire.PT_SPHERE, (0, 600, -1500) + (0,) * 6 + (800, 0, 0), 2),
        (wire.PT_SPHERE, (2400, 900, 300) + (0,) * 6 + (700, 0, 0), 3),
        (wire.PT_ELLIPSOID, (0, 2600, 500) + (0,) * 6 + (1400, 500, 700), 1),
        (wire.PT_CYLINDER, (-3500, -1500, -500, -1500, 2500, 500) + (0,) * 3 + (300, 0, 0), 1),
        (wire.PT_CONE, (3500, -1500, -500, 1800, 2400, 600) + (0,) * 3 + (250, 0, 0), 0),
        (wire.PT_TRIANGLE, (-4000, -2000, 2500, 4000, -2000, 2500, 0, 3500, 3000, 0, 0, 0), 4),
        (wire.PT_TRIANGLE, (-1000, -1800, -2500, 1000, -1800, -2500, 0, -200, -2300, 0, 0, 0), 2),
        (wire.PT_SPHERE, (1200, -800, -2600) + (0,) * 6 + (450, 0, 0), 5),
        (wire.PT_SPHERE, (1500, -700, -1800) + (0,) * 6 + (400, 0, 0), 5),
        (wire.PT_XZPLANE, (0, -2200, 0) + (0,) * 6 + (6000, 0, 6000), 4),
        (wire.PT_YZPLANE, (-5000, 0, 0) + (0,) * 6 + (0, 4000, 4000), 1),
        (wire.PT_XYPLANE, (0, 0, 4500) + (0,) * 6 + (6000, 4000, 0), 0),
        (wire.PT_CHECKBOARD, (0, -2600, 0) + (0,) * 6 + (5000, 0, 5000), 4),
    ]
    prims.append(_light(mats))
    return _pack("mixed", mats, prims)


def textured_scene():
    rng = np.random.Generator(np.random.PCG64(77))
    tex0 = rng.integers(0, 256, size=(16, 32, 3), dtype=np.uint8)   # diffuse
    tex1 = rng.integers(0, 256, size=(16, 32, 3), dtype=np.uint8)   # normal / specular / ao ...
    tex2 = rng.integers(0, 256, size=(8, 8, 3), dtype=np.uint8)     # skybox
    M = scenes.material
    N = wire.TEXTURE_NONE
    mats = [M(1, 1, 1, propagation=50000.0, textures=[0, 1, N, 1, N, N, N]),
            M(1, 1, 1, reflection=0.5, transparency=0.3, refraction=1.2, propagation=50000.0, textures=[0, N, 1, N, 1, 1, 1]),
            M(0.4, 0.6, 0.8, propagation=50000.0, textures=[wire.TEXTURE_NONE - 1, N, N, N, N, N, N]),  # mandelbrot
            M(0.8, 0.6, 0.4, propagation=50000.0, textures=[wire.TEXTURE_NONE - 2, N, N, N, N, N, N]),  # julia
            M(0.2, 0.3, 0.5, propagation=50000.0, textures=[2, N, N, N, N, N, N])]                      # skybox
    prims = [
        (wire.PT_SPHERE, (-2000, 500, 0) + (0,) * 6 + (1200, 0, 0), 0),
        (wire.PT_SPHERE, (2000, 500, 0) + (0,) * 6 + (1200, 0, 0), 1),
        (wire.PT_TRIANGLE, (-4000, -2500, 1500, 4000, -2500, 1500, 0, 3000, 2500, 0, 0, 0), 2),
        (wire.PT_TRIANGLE, (-3000, -2600, -1500, 3000, -2600, -1500, 0, -2600, 2500, 0, 0, 0), 0),
        (wire.PT_XZPLANE, (0, -2900, 0) + (0,) * 6 + (5000, 0, 5000), 3),
    ]
    prims.append(_light(mats))
    sc = _pack("textured", mats, prims, textures=[(0, tex0), (1, tex1), (2, tex2)])
    return sc


def spheres_scene():
    return scenes.random_spheres(60, 4000.0, 300.0, 700.0, scenes.SEED + 11, "spheres60", ground_y=-5000.0)


def molecule_scene():
    return scenes.molecule(cells=1, atoms_per_cell=60, seed=scenes.SEED + 12, name="molecule1")


def mesh_scene():
    return scenes.triangle_mesh(n_target=400, seed=scenes.SEED + 13, name="mesh400")


def _si(**kw):
    si = wire.default_scene_info(W, H, graphics_level=kw.pop("gl", wire.GL_FULL), nb_ray_iterations=kw.pop("nit", 3))
    for k, v in kw.items():
        if k == "backgroundColor":
            si.backgroundColor = wire.Float4(*v)
        else:
            setattr(si, k, v)
    return si


# name -> dict(scene=fn, si=SceneInfo kwargs, frames=[iterations], randoms seed, camera overrides)
CASES = {
    "spheres_full": dict(scene=spheres_scene, si=dict(nit=3), frames=[0]),
    "spheres_noshading": dict(scene=spheres_scene, si=dict(gl=wire.GL_NO_SHADING, nit=1), frames=[0]),
    "spheres_phong": dict(scene=spheres_scene, si=dict(gl=wire.GL_PHONG_BLINN, nit=2), frames=[0]),
    "spheres_progressive": dict(scene=spheres_scene, si=dict(nit=2, maxPathTracingIterations=14), frames=list(range(0, 14)), randoms=5),
    "spheres_rotated": dict(scene=spheres_scene, si=dict(nit=3), frames=[0], angles=(0.3, -0.4, 0.2, 6400.0)),
    "spheres_ortho": dict(scene=spheres_scene, si=dict(nit=2, cameraType=wire.CT_ORTHOGRAPHIC), frames=[0]),
    "spheres_aa": dict(scene=spheres_scene, si=dict(nit=2, cameraType=wire.CT_ANTIALIASED), frames=[0]),
    "spheres_anaglyph": dict(scene=spheres_scene, si=dict(nit=2, cameraType=wire.CT_ANAGLYPH, maxPathTracingIterations=12), frames=[0, 10, 11], randoms=6),
    "spheres_anaglyph_rotated": dict(scene=spheres_scene, si=dict(nit=2, cameraType=wire.CT_ANAGLYPH), frames=[0], angles=(0.2, -0.3, 0.1, 6400.0)),
    "spheres_aa_rotated": dict(scene=spheres_scene, si=dict(nit=2, cameraType=wire.CT_ANTIALIASED), frames=[0], angles=(-0.15, 0.25, -0.1, 6400.0)),
    "spheres_fog_gradient": dict(scene=spheres_scene, si=dict(nit=2, atmosphericEffect=1, gradientBackground=1, viewDistance=21000.0,
                                                              backgroundColor=(0.3, 0.5, 0.9, 0.4)), frames=[0]),
    "spheres_boxes": dict(scene=spheres_scene, si=dict(nit=2, renderBoxes=1), frames=[0]),
    "spheres_gi": dict(scene=spheres_scene, si=dict(nit=2, advancedIllumination=wire.AI_FULL, maxPathTracingIterations=12), frames=[0, 10, 11], randoms=7),
    "spheres_random_illum": dict(scene=spheres_scene, si=dict(nit=2, advancedIllumination=wire.AI_RANDOM), frames=[0], randoms=8),
    "molecule_full": dict(scene=molecule_scene, si=dict(nit=3), frames=[0], eye=(0.0, 0.0, -9000.0)),
    "mesh_full": dict(scene=mesh_scene, si=dict(nit=5), frames=[0]),
    "mesh_doublesided": dict(scene=mesh_scene, si=dict(nit=3, doubleSidedTriangles=1), frames=[0]),
    "mesh_noextended": dict(scene=mesh_scene, si=dict(nit=3, extendedGeometry=0), frames=[0]),
    "mixed_full": dict(scene=mixed_scene, si=dict(nit=4, transparentColor=2.0), frames=[0]),
    "mixed_default_transparent": dict(scene=mixed_scene, si=dict(nit=3), frames=[0]),
    "mixed_timestamp": dict(scene=mixed_scene, si=dict(nit=3, transparentColor=2.0, timestamp=1234), frames=[0, 1]),
    "textured_skybox": dict(scene=textured_scene, si=dict(nit=3, transparentColor=2.0, skyboxMaterialId=4, skyboxRadius=45000), frames=[0]),
    # remaining cameras: side-by-side stereo (k_3DVisionRenderer) and 360-degree panorama (k_fishEyeRenderer)
    "spheres_vr": dict(scene=spheres_scene, si=dict(nit=2, cameraType=wire.CT_VR, eyeSeparation=380.0), frames=[0, 1],
                       angles=(0.1, -0.2, 0.0, 6400.0)),
    "spheres_panoramic": dict(scene=spheres_scene, si=dict(nit=2, cameraType=wire.CT_PANORAMIC, maxPathTracingIterations=12),
                              frames=[0, 10, 11], randoms=31, eye=(0.0, 0.0, -3000.0), target=(0.0, 0.0, 0.0), post=(wire.PPE_NONE, 0.0, 0.0002, 0)),
    # post-processing effects of cudaRender's second pass (type, param1 = focus depth, param2 = strength, param3 = samples / filter)
    "spheres_pp_dof": dict(scene=spheres_scene, si=dict(nit=2), frames=[0], randoms=21, post=(wire.PPE_DEPTH_OF_FIELD, 14000.0, 4000.0, 24)),
    "spheres_pp_ao": dict(scene=spheres_scene, si=dict(nit=2), frames=[0], randoms=22, post=(wire.PPE_AMBIENT_OCCLUSION, 0.0, 9000.0, 0)),
    "spheres_pp_radiosity": dict(scene=spheres_scene, si=dict(nit=2, maxPathTracingIterations=13), frames=[0, 10, 11, 12], randoms=23,
                                 post=(wire.PPE_RADIOSITY, 0.0, 6000.0, 16)),
    "spheres_pp_emboss": dict(scene=spheres_scene, si=dict(nit=2), frames=[0], post=(wire.PPE_FILTER, 0.0, 0.0, 0)),
    "spheres_pp_subtle_sharpen": dict(scene=spheres_scene, si=dict(nit=2, maxPathTracingIterations=12), frames=[0, 10, 11], randoms=24,
                                      post=(wire.PPE_FILTER, 0.0, 0.0, 5)),
    "spheres_pp_cartoon": dict(scene=spheres_scene, si=dict(nit=2), frames=[0], post=(wire.PPE_CARTOON, -60000.0, 0.0, 0)),
}


def case_post(name):
    """PostProcessingInfo of a case (ppe_none when the case has no post-processing effect)."""
    pp = wire.PostProcessingInfo()
    p = CASES[name].get("post")
    if p is not None:
        pp.type, pp.param1, pp.param2, pp.param3 = int(p[0]), float(p[1]), float(p[2]), int(p[3])
    return pp


def case_setup(name):
    c = CASES[name]
    sc = c["scene"]()
    si = _si(**dict(c["si"]))
    eye = c.get("eye", sc.eye)
    target = c.get("target", sc.target)
    angles = c.get("angles", sc.angles)
    return sc, si, eye, target, angles, randoms(c.get("randoms")), c["frames"]
