"""vfd_b200/scene_io.py — the data formats on either side of the solver path, host side.

* `read_scene(path)` reads a VFD scene file — the cereal JSON archive of an entt registry snapshot that
  `Scene::Save` writes (reference: VFD/Source/Scene/Scene.cpp:204-260; component serialisers
  Scene/Components.h and Scene/Components/DFSPHSimulationComponent.h:7-39) — into plain Python: the simulation
  description (field for field `DFSPHSimulationDescription`), the fluid objects and the rigid bodies with their mesh
  source, transform and sampling / collision-map parameters.
* `sample_box_volume(...)` restates `ParticleSampler::SampleMeshVolume` (Utility/Sampler/ParticleSampler.cpp:7-91) for
  an axis-aligned box: the same float32 lattice walk in the three sample modes, the inside test against the box's exact
  signed distance instead of the interpolated SDF grid.  `SampleMode::MinDensity` places samples half a diameter inside
  the lattice cell, so the result is identical to the reference's (bit for bit, same order); in the two denser modes
  lattice points lie ON the faces, where the reference's answer is decided by the interpolation error of its 20^3 SDF grid:
  it keeps some points on (or up to 2e-3 outside) the surface that the exact test drops — tests/test_scene_io_cpu.py
  states that tolerance.
* `read_obj(path)` reads the vertices and triangles the reference's loaders take from an .obj file.
* `build_simulation(scene)` turns a scene into a ready `api.DFSPHSimulation` the way the editor's "Bake Objects" does: every
  fluid object's mesh sampled with particles, every rigid body's mesh turned into a volume map — any closed triangle mesh
  under any transform, on the GPU (csrc/volume_map.cu: the reference's mesh distance, sampling lattice and map quadrature).
  `sample_box_volume` stays as a host-side statement of the sampling lattice for boxes (tests).
"""
import json
import os

import numpy as np

MIN_DENSITY, MEDIUM_DENSITY, MAX_DENSITY = 0, 1, 2

# cereal NVP name (DFSPHSimulationComponent.h:11-37) -> DFSPHSimulationDescription field
DESCRIPTION_FIELDS = {
    "timeStepSize": "TimeStepSize", "minTimeStepSize": "MinTimeStepSize", "maxTimeStepSize": "MaxTimeStepSize",
    "frameLength": "FrameLength", "frameCount": "FrameCount",
    "minPressureSolverIterations": "MinPressureSolverIterations", "maxPressureSolverIterations": "MaxPressureSolverIterations",
    "maxPressureSolverError": "MaxPressureSolverError",
    "enableDivergenceSolverError": "EnableDivergenceSolverError",
    "minDivergenceSolverIterations": "MinDivergenceSolverIterations", "maxDivergenceSolverIterations": "MaxDivergenceSolverIterations",
    "maxDivergenceSolverError": "MaxDivergenceSolverError",
    "enableViscositySolver": "EnableViscositySolver",
    "minViscositySolverIterations": "MinViscositySolverIterations", "maxViscositySolverIterations": "MaxViscositySolverIterations",
    "maxViscositySolverError": "MaxViscositySolverError",
    "viscosity": "Viscosity", "boundaryViscosity": "BoundaryViscosity", "tangentialDistanceFactor": "TangentialDistanceFactor",
    "enableSurfaceTensionSolver": "EnableSurfaceTensionSolver", "surfaceTensionSmoothPassCount": "SurfaceTensionSmoothPassCount",
    "surfaceTension": "SurfaceTension", "temporalSmoothing": "TemporalSmoothing", "CSDFix": "CSDFix", "CSD": "CSD",
    "particleRadius": "ParticleRadius", "gravity": "Gravity",
}


def _vec(d, keys="xyz"):
    return tuple(d[k] for k in keys)


def _mat4(t):
    """cereal writes a glm::mat4 as value0..value3 = its four COLUMNS; returned row-major (m[row][col])."""
    cols = [_vec(t["value%d" % c], "xyzw") for c in range(4)]
    return np.array(cols, np.float32).T.copy()


def read_scene(path):
    """-> {"description": {Field: value}, "fluid_objects": [...], "rigid_bodies": [...], "entities": {id: {...}}}

    The archive is a flat sequence value0, value1, ... : the entity list, then one pool per component type, each pool as
    its size followed by (entity, component) pairs.  Components are recognised by their member names."""
    with open(path) as f:
        doc = json.load(f)
    seq, i = [], 0
    while "value%d" % i in doc:
        seq.append(doc["value%d" % i])
        i += 1
    ents = {}
    for k, v in enumerate(seq):
        if not isinstance(v, dict) or k == 0 or not isinstance(seq[k - 1], int):
            continue
        e = ents.setdefault(int(seq[k - 1]), {})
        if "tag" in v:
            e["tag"] = v["tag"]
        elif "transform" in v:
            e["transform"] = _mat4(v["transform"])
        elif "meshSource" in v:
            e["mesh"] = v["meshSource"]
        elif "collisionMapResolution" in v:
            e["rigid_body"] = dict(inverted=bool(v["inverted"]), padding=float(v["padding"]), resolution=_vec(v["collisionMapResolution"]))
        elif "sampleMode" in v:
            e["fluid_object"] = dict(inverted=bool(v["inverted"]), resolution=_vec(v["resolution"]), sample_mode=int(v["sampleMode"]),
                                     velocity=_vec(v["velocity"]))
        elif "description" in v:
            e["description"] = {DESCRIPTION_FIELDS[k2]: (_vec(x) if isinstance(x, dict) else x) for k2, x in v["description"].items() if k2 in DESCRIPTION_FIELDS}
        elif "id" in v:
            e["uuid"] = v["id"].get("UUID32")
    out = {"description": None, "fluid_objects": [], "rigid_bodies": [], "entities": ents}
    for eid, e in sorted(ents.items()):
        T = e.get("transform", np.eye(4, dtype=np.float32))
        if "description" in e:
            out["description"] = e["description"]
        if "fluid_object" in e:
            out["fluid_objects"].append(dict(e["fluid_object"], mesh=e.get("mesh"), transform=T, tag=e.get("tag")))
        if "rigid_body" in e:
            out["rigid_bodies"].append(dict(e["rigid_body"], mesh=e.get("mesh"), transform=T, tag=e.get("tag")))
    return out


# what DFSPHSimulationDescription's member initialisers give (Structures/DFSPHSimulationDescription.h:9-50): the loader reads
# every name-value pair, in this order (DFSPHSimulationComponent.h:11-37), so a scene file carries all of them
DESCRIPTION_DEFAULTS = {
    "TimeStepSize": 0.001, "MinTimeStepSize": 0.0001, "MaxTimeStepSize": 0.005, "FrameLength": 0.0016, "FrameCount": 200,
    "MinPressureSolverIterations": 0, "MaxPressureSolverIterations": 100, "MaxPressureSolverError": 10.0,
    "EnableDivergenceSolverError": True, "MinDivergenceSolverIterations": 0, "MaxDivergenceSolverIterations": 100, "MaxDivergenceSolverError": 10.0,
    "EnableViscositySolver": True, "MinViscositySolverIterations": 0, "MaxViscositySolverIterations": 100, "MaxViscositySolverError": 0.1,
    "Viscosity": 10.0, "BoundaryViscosity": 10.0, "TangentialDistanceFactor": 0.5,
    "EnableSurfaceTensionSolver": False, "SurfaceTensionSmoothPassCount": 1, "SurfaceTension": 1.0, "TemporalSmoothing": False,
    "CSDFix": -1, "CSD": 10000, "ParticleRadius": 0.025, "Gravity": (0.0, -9.81, 0.0),
}
_BOOL_FIELDS = ("EnableDivergenceSolverError", "EnableViscositySolver", "EnableSurfaceTensionSolver", "TemporalSmoothing")

# MaterialComponent of the three kinds of entity (shader + its property block as raw bytes), as the editor saves them
_MATERIALS = {
    "sim": ("Resources/Shaders/Normal/DFSPHParticleSimpleShader.glsl",
            np.array([0.0, 0.843, 0.561, 1.0, 0.0, 0.2, 0.976, 1.0], np.float32)),
    "rb": ("Resources/Shaders/Normal/BasicDiffuseShader.glsl", np.array([0.4, 0.4, 0.4, 1.0], np.float32)),
    "fo": ("Resources/Shaders/Normal/ColorShader.glsl", np.array([1.0, 1.0, 1.0, 1.0], np.float32)),
}

# the component pools of the snapshot, in the order Scene::Save / Scene::Load name them (Scene.cpp:204-260): the loader reads
# them by position, so an absent component type still needs its (empty) pool
POOL_ORDER = ("ID", "Tag", "Relationship", "Transform", "SPHSimulation", "Material", "Mesh", "RigidBody", "FluidObject", "DFSPHSimulation")


def write_scene(path, description, fluid_objects=(), rigid_bodies=()):
    """Writes a scene in the archive layout the editor's `Scene::Load` reads (and `read_scene`): the entity list, then one pool
    per component type in POOL_ORDER — the (empty) SPH pool and the material pool included — with the complete description.
    Entity 0 is the simulation, then the rigid bodies, then the fluid objects."""
    inv = {v: k for k, v in DESCRIPTION_FIELDS.items()}
    ents = [dict(tag="GPU Simulation", sim=description)]
    ents += [dict(tag=b.get("tag", "Rigid Body"), rb=b) for b in rigid_bodies]
    ents += [dict(tag=f.get("tag", "Fluid Object"), fo=f) for f in fluid_objects]
    n = len(ents)

    def xyz(v, keys="xyz"):
        return {k: (int(x) if isinstance(x, (int, np.integer)) else float(x)) for k, x in zip(keys, v)}

    def mat(T):
        T = np.asarray(T, np.float32)
        return {"transform": {"value%d" % c: xyz(T[:, c], "xyzw") for c in range(4)}}

    def material(e):
        shader, props = _MATERIALS["sim" if "sim" in e else ("rb" if "rb" in e else "fo")]
        return {"shaderSource": shader, "properties": [int(b) for b in props.tobytes()]}
    seq = [n + 1, 4294967295] + list(range(n))

    def pool(items):
        seq.append(len(items))
        for e, c in items:
            seq.extend([e, c])
    full = dict(DESCRIPTION_DEFAULTS)
    full.update(description)
    d = {}
    for name, field in DESCRIPTION_FIELDS.items():                 # the reference's order
        v = full[field]
        d[name] = xyz(v) if isinstance(v, (tuple, list, np.ndarray)) else (bool(v) if field in _BOOL_FIELDS else v)
    pools = {
        "ID": [(i, {"id": {"UUID32": 1000 + i}}) for i in range(n)],
        "Tag": [(i, {"tag": e["tag"]}) for i, e in enumerate(ents)],
        "Relationship": [(i, {"parent": {"UUID32": 0}, "children": []}) for i in range(n)],
        "Transform": [(i, mat((e.get("rb") or e.get("fo") or {}).get("transform", np.eye(4)))) for i, e in enumerate(ents)],
        "SPHSimulation": [],
        "Material": [(i, material(e)) for i, e in enumerate(ents)],
        "Mesh": [(i, {"meshSource": (e.get("rb") or e.get("fo"))["mesh"]}) for i, e in enumerate(ents) if "rb" in e or "fo" in e],
        "RigidBody": [(i, {"inverted": bool(e["rb"]["inverted"]), "padding": float(e["rb"]["padding"]), "collisionMapResolution": xyz(e["rb"]["resolution"])})
                      for i, e in enumerate(ents) if "rb" in e],
        "FluidObject": [(i, {"inverted": bool(e["fo"]["inverted"]), "resolution": xyz(e["fo"]["resolution"]), "sampleMode": int(e["fo"]["sample_mode"]),
                             "velocity": xyz(e["fo"].get("velocity", (0.0, 0.0, 0.0)))}) for i, e in enumerate(ents) if "fo" in e],
        "DFSPHSimulation": [(0, {"description": d})],
    }
    for name in POOL_ORDER:
        pool(pools[name])
    doc = {"value%d" % i: v for i, v in enumerate(seq)}
    doc["sceneData"] = {"cameraPosition": xyz((4.0, 4.0, 4.0)), "cameraPivot": xyz((0.0, 0.0, 0.0)), "readMe": ""}
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)


def pool_sequence(path):
    """The archive's shape: entity count, then for every pool (its size, the member names of its first component).  Two files
    the editor's loader treats alike have the same pool sequence up to the sizes."""
    with open(path) as f:
        doc = json.load(f)
    seq, i = [], 0
    while "value%d" % i in doc:
        seq.append(doc["value%d" % i])
        i += 1
    ents = int(seq[0]) - 1
    k, out = 2 + ents, []
    while k < len(seq):
        size = int(seq[k])
        keys = tuple(seq[k + 2].keys()) if size else ()
        if keys == ("description",):
            keys = ("description",) + tuple(seq[k + 2]["description"].keys())
        out.append((size, keys))
        k += 1 + 2 * size
    return ents, out


def unit_cube_box(transform):
    """Axis-aligned box of the unit cube mesh (Resources/Models/Cube.obj spans [-1, 1]^3) under a scale/translation
    transform; None when the transform rotates or shears (then it is not an axis-aligned box)."""
    T = np.asarray(transform, np.float64)
    A = T[:3, :3]
    if np.abs(A - np.diag(np.diag(A))).max() > 1e-6 * max(1.0, np.abs(A).max()):
        # a rotation by a multiple of 90 degrees still maps the cube onto an axis-aligned box
        if not np.allclose(np.sort(np.abs(A), axis=1)[:, :2], 0.0, atol=1e-6 * max(1.0, np.abs(A).max())):
            return None
    corners = np.array([[x, y, z, 1.0] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)]) @ T.T
    lo, hi = corners[:, :3].min(axis=0), corners[:, :3].max(axis=0)
    return lo.astype(np.float32), hi.astype(np.float32)


def sample_box_volume(bmin, bmax, radius, mode=MIN_DENSITY, inverted=False):
    """`ParticleSampler::SampleMeshVolume` (ParticleSampler.cpp:7-91) for an axis-aligned box: float32 lattice walk over
    the bounds (`for (z = min.z; z <= max.z; z += diameter)` ... — the accumulating float loop counters are reproduced),
    position per sample mode, kept when the signed distance to the box is negative."""
    f = np.float32
    bmin, bmax = np.asarray(bmin, f), np.asarray(bmax, f)
    # the reference's BoundingBox starts from min = max = (0, 0, 0) and is only ever extended (Core/Structures/BoundingBox.h:
    # 16-31, 47-48): the lattice (and its SDF grid) always spans the origin as well.  Reproduced: it fixes the lattice phase.
    lat_min, lat_max = np.minimum(bmin, f(0.0)), np.maximum(bmax, f(0.0))
    r = f(radius)
    d = f(2.0) * r
    sx, sy = d, d
    if mode == MEDIUM_DENSITY:
        sy = f(np.sqrt(f(3.0))) * r
    elif mode == MAX_DENSITY:
        sx = f(np.sqrt(f(3.0))) * r
        sy = f(np.sqrt(f(6.0))) * d / f(3.0)

    def axis(lo, hi, step):
        v, out = f(lo), []
        while v <= hi:
            out.append(v)
            v = f(v + step)
        return np.array(out, f)
    xs, ys, zs = axis(lat_min[0], lat_max[0], sx), axis(lat_min[1], lat_max[1], sy), axis(lat_min[2], lat_max[2], d)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")                     # x fastest, as the reference's loop nest
    cy = np.broadcast_to(np.arange(len(ys))[None, :, None], X.shape)
    cx = np.broadcast_to(np.arange(len(xs))[None, None, :], X.shape)
    if mode == MIN_DENSITY:
        P = np.stack([X + r, Y + r, Z + r], axis=-1)
    elif mode == MEDIUM_DENSITY:
        even = (cy % 2 == 0)
        P = np.stack([np.where(even, X, X + r), Y + r, np.where(even, Z + r, Z)], axis=-1)
    else:
        P = np.stack([X, Y + r, Z + r], axis=-1).astype(f)
        shz = np.where(cx % 2 == 1, d / (f(2.0) * np.where(cy % 2 == 1, f(-1.0), f(1.0))), f(0.0)).astype(f)
        shx = np.where(cy % 2 == 1, sx / f(2.0), f(0.0)).astype(f)
        shz = (shz + np.where(cy % 2 == 1, d / f(2.0), f(0.0))).astype(f)
        P = P + np.stack([shx, np.zeros_like(shx), shz], axis=-1)
    P = P.reshape(-1, 3).astype(f)
    # exact signed distance of a box: negative inside
    q = np.maximum(bmin - P, P - bmax)
    outside = np.sqrt((np.maximum(q, 0.0) ** 2).sum(axis=1))
    dist = np.where((q <= 0).all(axis=1), q.max(axis=1), outside)
    if inverted:
        dist = -dist
    return P[dist < 0.0]


def read_obj(path):
    """The (vertices, triangles) the reference's loaders take from an .obj file (TriangleMesh::LoadOBJ, Renderer/Mesh/
    TriangleMesh.cpp:39-107; EdgeMesh.cpp:62-100, both through tinyobjloader): every `v` line in file order, every `f` line as
    vertex indices (1-based, negative = relative to the vertices read so far; texture / normal indices ignored), polygons
    with more than three corners fan-triangulated around their first corner."""
    verts, tris = [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                verts.append([float(t[1]), float(t[2]), float(t[3])])
            elif t[0] == "f":
                ids = []
                for c in t[1:]:
                    k = int(c.split("/")[0])
                    ids.append(k - 1 if k > 0 else len(verts) + k)
                for j in range(1, len(ids) - 1):
                    tris.append([ids[0], ids[j], ids[j + 1]])
    return np.array(verts, np.float32).reshape(-1, 3), np.array(tris, np.uint32).reshape(-1, 3)


def build_simulation(scene, device=0, resources_root=None, meshes=None, **overrides):
    """A ready `DFSPHSimulation` for a scene as `read_scene` returns it — what the editor's "Bake Objects" does
    (SURVEY.md section 3.1): every fluid object's mesh under its transform is sampled with particles, every rigid body's mesh
    becomes a volume map, both on the GPU (csrc/volume_map.cu: vfd_sample_mesh_volume, vfd_volume_map_build_mesh — the
    reference's mesh distance, sampling lattice and map integration).

    Meshes are looked up in `meshes` ({mesh source or its file name: (vertices, triangles)}) and otherwise read from
    `resources_root`/<mesh source> with `read_obj`.  Without either, the unit cube (Cube.obj spans [-1, 1]^3) is the only
    mesh known by name."""
    from . import api
    desc = dict(scene["description"] or {})
    desc.update(overrides)
    for k, v in list(desc.items()):
        if isinstance(v, bool):
            desc[k] = int(v)
    sim = api.DFSPHSimulation(api.DFSPHSimulationDescription(**desc), device=device)
    radius = float(desc.get("ParticleRadius", 0.025))

    def mesh_of(obj):
        src = obj.get("mesh") or ""
        for key in (src, os.path.basename(src)):
            if meshes and key in meshes:
                return meshes[key]
        if resources_root and src and os.path.exists(os.path.join(resources_root, src)):
            return read_obj(os.path.join(resources_root, src))
        if os.path.basename(src).lower() == "cube.obj":
            return UNIT_CUBE
        raise FileNotFoundError("mesh %r: pass it in `meshes` or give `resources_root`" % src)
    fluids = []
    for fo in scene["fluid_objects"]:
        v, t = mesh_of(fo)
        pos = api.sample_mesh_volume(v, t, radius, fo["resolution"], fo["inverted"], fo["sample_mode"], transform=fo["transform"], device=device)
        fluids.append(api.FluidObject(pos, velocity=fo["velocity"]))
    sim.SetFluidObjects(fluids)
    maps = []
    for rb in scene["rigid_bodies"]:
        v, t = mesh_of(rb)
        maps.append(api.VolumeMap.build_mesh(v, t, transform=rb["transform"], inverted=rb["inverted"], padding=rb["padding"],
                                             resolution=rb["resolution"], particle_radius=radius, device=device))
    sim.SetRigidBodies(maps)
    return sim


# the reference's TriangleMesh(AABB) topology (TriangleMesh.cpp:18-37) on [-1, 1]^3
UNIT_CUBE = (np.array([[-1, -1, -1], [1, -1, -1], [1, -1, 1], [-1, -1, 1], [-1, 1, -1], [1, 1, -1], [1, 1, 1], [-1, 1, 1]], np.float32),
             np.array([[0, 1, 2], [0, 2, 3], [4, 7, 6], [4, 6, 5], [0, 3, 7], [0, 7, 4], [1, 5, 6], [1, 6, 2], [0, 4, 5], [0, 5, 1], [3, 2, 6], [3, 6, 7]], np.uint32))
