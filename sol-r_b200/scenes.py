"""Synthetic scenes of the sizes BASELINE.json names, expressed as setter calls.

A scene is plain data (numpy arrays) that is replayed through any object exposing the reference's
setter interface (GPUKernel::addPrimitive/setPrimitive/setMaterial, GPUKernel.cpp:495-684,1780-1909):
the product's SceneHost, or the reference itself through oracle/ref_build/ref_harness.cpp in tests.
Recipes and constants follow SURVEY.md §8(d); every RNG is numpy's PCG64 seeded 20261017 + config#.
"""
from dataclasses import dataclass, field

import numpy as np

from . import wire

SEED = 20261017

# material float columns: r,g,b,noise,reflection,refraction,transparency,opacity,specValue,specPower,specCoef,
#                         innerIllumination,illuminationDiffusion,illuminationPropagation
# material int columns:   procedural,wireframe,wireframeWidth,diffuse,normal,bump,specular,reflection,
#                         transparent,ambientOcclusion,fastTransparency
MAT_F = 14
MAT_I = 11


def material(r, g, b, reflection=0.0, refraction=0.0, transparency=0.0, opacity=0.0, spec_value=1.0,
             spec_power=100.0, spec_coef=0.0, inner=0.0, diffusion=0.0, propagation=0.0, noise=0.0,
             procedural=0, wireframe=0, wireframe_width=0, fast_transparency=0, textures=None):
    """textures: optional 7 ids (diffuse, normal, bump, specular, reflection, transparent, ambient occlusion)."""
    f = np.array([r, g, b, noise, reflection, refraction, transparency, opacity, spec_value, spec_power,
                  spec_coef, inner, diffusion, propagation], dtype=np.float32)
    tex = list(textures) if textures is not None else [wire.TEXTURE_NONE] * 7
    i = np.array([procedural, wireframe, wireframe_width] + tex + [fast_transparency], dtype=np.int32)
    return f, i


@dataclass
class Scene:
    name: str
    mat_f: np.ndarray            # [M, 14] float32
    mat_i: np.ndarray            # [M, 11] int32
    prim_type: np.ndarray        # [N] int32
    prim_v: np.ndarray           # [N, 12] float32: p0, p1, p2, size
    prim_mat: np.ndarray         # [N] int32
    normals: dict = field(default_factory=dict)   # index -> 9 floats
    textures: list = field(default_factory=list)  # [(index, uint8[h, w, depth])], set before the materials
    bulk_normals: object = None                   # float32[n, 9] for primitives 0..n-1
    eye: tuple = (0.0, 0.0, -15000.0)
    target: tuple = (0.0, 0.0, 0.0)
    angles: tuple = (0.0, 0.0, 0.0, 6400.0)

    @property
    def nb_primitives(self):
        return int(self.prim_type.shape[0])

    def replay(self, builder):
        """Replay through an object with add_materials/add_primitives/set_normals/compact_boxes."""
        for idx, texels in self.textures:
            builder.set_texture(idx, texels)
        builder.add_materials(self.mat_f, self.mat_i)
        builder.add_primitives(self.prim_type, self.prim_v, self.prim_mat)
        if self.bulk_normals is not None:
            builder.set_normals_bulk(0, self.bulk_normals)
        for idx, n in self.normals.items():
            builder.set_normals(idx, np.asarray(n, dtype=np.float32))
        return builder.compact_boxes()

    def texture_atlas(self):
        """The flat texel array as GPUKernel::processTextureOffsets lays it out (GPUKernel.cpp:2691-2705)."""
        if not self.textures:
            return None
        parts = [np.ascontiguousarray(t, np.uint8).reshape(-1) for _, t in sorted(self.textures, key=lambda x: x[0])]
        return np.concatenate(parts + [np.zeros(16, np.uint8)])


def _pack(name, mats, prims, **kw):
    mat_f = np.stack([m[0] for m in mats]).astype(np.float32)
    mat_i = np.stack([m[1] for m in mats]).astype(np.int32)
    t = np.array([p[0] for p in prims], dtype=np.int32)
    v = np.array([p[1] for p in prims], dtype=np.float32).reshape(-1, 12)
    m = np.array([p[2] for p in prims], dtype=np.int32)
    return Scene(name, mat_f, mat_i, t, v, m, **kw)


def _light_material(view_distance=50000.0):
    # DEFAULT_LIGHT_MATERIAL (apps/scenes/Scene.cpp:628-633): white, innerIllumination.x = 2; lamp range
    # (innerIllumination.z) >= viewDistance so no pixel falls outside it.
    return material(1.0, 1.0, 1.0, inner=2.0, diffusion=0.0, propagation=2.0 * view_distance, spec_value=0.0,
                    spec_power=0.0)


def _palette(n, rng, reflective_half=True, view_distance=50000.0):
    mats = []
    for k in range(n):
        r, g, b = rng.uniform(0.2, 1.0, size=3)
        refl = 0.3 if (reflective_half and k % 2 == 1) else 0.0
        mats.append(material(r, g, b, reflection=refl, spec_value=1.0, spec_power=100.0,
                             propagation=view_distance))
    return mats


def _ground(y, half, mat):
    # two triangles at height y (configs 1,2,4: "ground = 2 triangles")
    a = (-half, y, -half); b = (half, y, -half); c = (half, y, half); d = (-half, y, half)
    return [(wire.PT_TRIANGLE, a + c + b + (0, 0, 0), mat), (wire.PT_TRIANGLE, a + d + c + (0, 0, 0), mat)]


def random_spheres(n, extent, rmin, rmax, seed, name, ground_y=None, n_materials=8):
    """Configs 1 and 4: n spheres, centres uniform in [-extent, extent]^3, radii uniform [rmin, rmax]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    mats = _palette(n_materials, rng)
    light_id = len(mats)
    mats.append(_light_material())
    c = rng.uniform(-extent, extent, size=(n, 3)).astype(np.float32)
    r = rng.uniform(rmin, rmax, size=n).astype(np.float32)
    m = rng.integers(0, n_materials, size=n)
    t = np.full(n, wire.PT_SPHERE, dtype=np.int32)
    v = np.zeros((n, 12), dtype=np.float32)
    v[:, 0:3] = c
    v[:, 9] = r
    extra = [(wire.PT_SPHERE, (-5000.0, 5000.0, -15000.0) + (0,) * 6 + (1.0, 0, 0), light_id)]
    if ground_y is not None:
        extra += _ground(ground_y, 4.0 * extent, 0)
    et = np.array([e[0] for e in extra], dtype=np.int32)
    ev = np.array([e[1] for e in extra], dtype=np.float32).reshape(-1, 12)
    em = np.array([e[2] for e in extra], dtype=np.int32)
    mat_f = np.stack([x[0] for x in mats]).astype(np.float32)
    mat_i = np.stack([x[1] for x in mats]).astype(np.int32)
    return Scene(name, mat_f, mat_i, np.concatenate([t, et]), np.concatenate([v, ev]),
                 np.concatenate([m.astype(np.int32), em]))


def config1(n=1000):
    """1 000 random spheres + 2 ground triangles + 1 light, for 1024x768 (BASELINE.json configs[0])."""
    return random_spheres(n, 4000.0, 100.0, 300.0, SEED + 1, "config1_spheres", ground_y=-5000.0)


def molecule(cells=6, atoms_per_cell=486, seed=SEED + 2, name="config2_molecule"):
    """Config 2 stand-in: a synthetic B-DNA-like double helix (spheres = atoms, cylinders = bonds) per
    cell, replicated on a cells^3 lattice: 6^3 x (486 atoms + bonds) ~ 100k+ primitives.  The reference's
    medias/pdb/1BNA.pdb has 486 ATOM records (SURVEY finding 5) and does not travel to the GPU box, so the
    geometry is generated; only setters are used, as the PDB loader does (PDBReader.cpp:386-707)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    mats = []
    cpk = [(0.8, 0.8, 0.8), (0.56, 0.56, 0.56), (0.19, 0.31, 0.97), (1.0, 0.05, 0.05), (1.0, 0.5, 0.0),
           (1.0, 1.0, 0.19)]
    for k, (r, g, b) in enumerate(cpk):
        mats.append(material(r, g, b, reflection=0.2 if k in (2, 3) else 0.0, spec_value=1.0, spec_power=100.0,
                             propagation=50000.0))
    stick = len(mats)
    mats.append(material(0.6, 0.6, 0.6, spec_value=1.0, spec_power=100.0, propagation=50000.0))
    light_id = len(mats)
    mats.append(_light_material())

    # one cell: two strands, atoms_per_cell/2 atoms each
    per = atoms_per_cell // 2
    k = np.arange(per, dtype=np.float64)
    rise, twist, radius = 34.0 / per * 12.0, 2.0 * np.pi / (per / 12.0), 100.0
    pos = []
    for phase in (0.0, 2.2):
        ang = k * twist + phase
        p = np.stack([radius * np.cos(ang), k * rise - per * rise / 2.0, radius * np.sin(ang)], axis=1)
        p += rng.normal(0.0, 6.0, size=p.shape)
        pos.append(p)
    cell = np.concatenate(pos).astype(np.float32)
    kinds = rng.integers(0, len(cpk), size=cell.shape[0])
    radii = (14.0 + 6.0 * (kinds % 3)).astype(np.float32)
    bonds = [(i, i + 1) for s in (0, per) for i in range(s, s + per - 1)]
    bonds += [(i, per + i) for i in range(0, per, 8)]

    extent = float(np.abs(cell).max()) * 2.0 * 1.2
    offs = (np.arange(cells) - (cells - 1) / 2.0) * extent
    prim_t, prim_v, prim_m = [], [], []
    for ox in offs:
        for oy in offs:
            for oz in offs:
                o = np.array([ox, oy, oz], dtype=np.float32)
                c = cell + o
                v = np.zeros((c.shape[0], 12), dtype=np.float32)
                v[:, 0:3] = c
                v[:, 9] = radii
                prim_t.append(np.full(c.shape[0], wire.PT_SPHERE, dtype=np.int32))
                prim_v.append(v)
                prim_m.append(kinds.astype(np.int32))
                b = np.zeros((len(bonds), 12), dtype=np.float32)
                ia = np.array([x[0] for x in bonds]); ib = np.array([x[1] for x in bonds])
                b[:, 0:3] = c[ia]
                b[:, 3:6] = c[ib]
                b[:, 9] = 5.0
                prim_t.append(np.full(len(bonds), wire.PT_CYLINDER, dtype=np.int32))
                prim_v.append(b)
                prim_m.append(np.full(len(bonds), stick, dtype=np.int32))
    scale = 9000.0 / (extent * cells)  # fit the lattice in +-4500 like the viewer's molecule scale
    t = np.concatenate(prim_t); v = np.concatenate(prim_v) * np.float32(scale); m = np.concatenate(prim_m)
    extra = [(wire.PT_SPHERE, (-5000.0, 5000.0, -15000.0) + (0,) * 6 + (1.0, 0, 0), light_id)]
    extra += _ground(-5000.0, 20000.0, stick)
    et = np.array([e[0] for e in extra], dtype=np.int32)
    ev = np.array([e[1] for e in extra], dtype=np.float32).reshape(-1, 12)
    em = np.array([e[2] for e in extra], dtype=np.int32)
    mat_f = np.stack([x[0] for x in mats]).astype(np.float32)
    mat_i = np.stack([x[1] for x in mats]).astype(np.int32)
    return Scene(name, mat_f, mat_i, np.concatenate([t, et]), np.concatenate([v, ev]), np.concatenate([m, em]))


def config2():
    return molecule(6)


def config4(n=1_000_000):
    """1 M random spheres in [-20000, 20000]^3, radii [20, 60], for 3840x2160 (BASELINE.json configs[3])."""
    return random_spheres(n, 20000.0, 20.0, 60.0, SEED + 4, "config4_spheres", ground_y=None)


def triangle_mesh(n_target=1_000_000, seed=SEED + 3, name="config3_mesh"):
    """Config 3 stand-in: a torus tessellated to ~n_target triangles with per-vertex normals, half
    reflective (0.5), half refractive (transparency 0.7, refraction 1.33); direct ptTriangle setters."""
    mats = [material(0.9, 0.4, 0.2, reflection=0.5, propagation=50000.0),
            material(0.6, 0.8, 1.0, transparency=0.7, refraction=1.33, opacity=0.0, propagation=50000.0),
            material(0.5, 0.5, 0.5, propagation=50000.0)]
    light_id = len(mats)
    mats.append(_light_material())
    nu = int(np.sqrt(n_target / 2.0 * 2.0)); nv = max(3, n_target // (2 * nu))
    R, r = 3000.0, 1200.0
    u = np.linspace(0, 2 * np.pi, nu, endpoint=False); v = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    P = np.stack([(R + r * np.cos(vv)) * np.cos(uu), r * np.sin(vv), (R + r * np.cos(vv)) * np.sin(uu)], -1)
    N = np.stack([np.cos(vv) * np.cos(uu), np.sin(vv), np.cos(vv) * np.sin(uu)], -1)
    i0 = np.arange(nu)[:, None]; j0 = np.arange(nv)[None, :]
    i1 = (i0 + 1) % nu; j1 = (j0 + 1) % nv
    def g(A, i, j): return A[np.broadcast_to(i, (nu, nv)), np.broadcast_to(j, (nu, nv))].reshape(-1, 3)
    tri_p = np.concatenate([np.concatenate([g(P, i0, j0), g(P, i1, j0), g(P, i1, j1)], 1),
                            np.concatenate([g(P, i0, j0), g(P, i1, j1), g(P, i0, j1)], 1)]).astype(np.float32)
    tri_n = np.concatenate([np.concatenate([g(N, i0, j0), g(N, i1, j0), g(N, i1, j1)], 1),
                            np.concatenate([g(N, i0, j0), g(N, i1, j1), g(N, i0, j1)], 1)]).astype(np.float32)
    n = tri_p.shape[0]
    vprim = np.zeros((n, 12), dtype=np.float32); vprim[:, 0:9] = tri_p
    t = np.full(n, wire.PT_TRIANGLE, dtype=np.int32)
    m = (np.arange(n) % 2).astype(np.int32)
    extra = [(wire.PT_SPHERE, (-5000.0, 5000.0, -15000.0) + (0,) * 6 + (1.0, 0, 0), light_id)]
    extra += _ground(-5000.0, 20000.0, 2)
    et = np.array([e[0] for e in extra], dtype=np.int32)
    ev = np.array([e[1] for e in extra], dtype=np.float32).reshape(-1, 12)
    em = np.array([e[2] for e in extra], dtype=np.int32)
    mat_f = np.stack([x[0] for x in mats]).astype(np.float32)
    mat_i = np.stack([x[1] for x in mats]).astype(np.int32)
    return Scene(name, mat_f, mat_i, np.concatenate([t, et]), np.concatenate([vprim, ev]), np.concatenate([m, em]),
                 bulk_normals=tri_n)
