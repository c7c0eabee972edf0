"""Synthetic scenes for the BASELINE.json configs. Pure numpy data generation: the
same descriptor arrays are fed to the CUDA world and (in tests) to the CPU oracle.

RNG: counter-based splitmix64(seed, stream, index) -> f32 in [0,1) from the top 24 bits, so a scene is
a pure function of (config, seed).
"""
import numpy as np

from . import _abi as A

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform(seed, stream, n, offset=0):
    """n floats in [0,1): f32(top 24 bits of splitmix64) * 2^-24."""
    with np.errstate(over="ignore"):
        idx = np.arange(offset, offset + n, dtype=np.uint64)
        key = np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + np.uint64(stream) * np.uint64(0xAF251AF3B0F025B5)
        z = _splitmix64(idx + key)
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


class Scene:
    """Plain container: descriptors + index-based topology (indices refer to insertion order)."""

    def __init__(self, name, gravity=(0.0, -30.0)):
        self.name = name
        self.gravity = gravity
        self.constraints = []          # [(x, y, r)]
        self.bodies = A.body_descs(0)
        self.colliders = A.collider_descs(0)
        self.col_parent = np.zeros(0, dtype=np.int64)   # body index per collider
        self.springs = []              # [(a_idx, b_idx, rest, k, c)]  (numpy arrays allowed)
        self.joints = []               # [(a_idx, b_idx)]
        self.cell_size = None          # BLOBS_PARAM_CELL_SIZE (reference spatial-hash cell)
        self.interleave = True         # insert body i then its collider(s) (the reference's spawn order)

    @property
    def n_bodies(self):
        return len(self.bodies)

    @property
    def n_colliders(self):
        return len(self.colliders)


def build(world, scene):
    """Insert a Scene into a world (blobs_b200.World or the oracle's OracleWorld). Returns handle arrays."""
    for (x, y, r) in scene.constraints:
        world.constraint_push((x, y), r)
    if scene.cell_size is not None:
        world.set_param(A.PARAM_CELL_SIZE, scene.cell_size)
    bh = world.insert_bodies(scene.bodies)
    ch = world.insert_colliders(scene.colliders, bh[scene.col_parent]) if scene.n_colliders else np.zeros(0, np.uint64)
    sh, jh = _insert_links(world, scene, bh)
    return {"bodies": bh, "colliders": ch, "springs": sh, "joints": jh}


def _insert_links(world, scene, bh):
    """springs and joints of a Scene; uses the bulk entry points when the world has them (the CUDA world does)"""
    springs, joints = scene.springs, scene.joints
    if isinstance(springs, np.ndarray):   # structured: a, b, rest, k, c
        if hasattr(world, "insert_springs") and len(springs):
            sh = world.insert_springs(bh[springs["a"]], bh[springs["b"]], np.stack([springs["rest"], springs["k"], springs["c"]], axis=1))
        else:
            sh = [world.spring_insert(bh[s["a"]], bh[s["b"]], float(s["rest"]), float(s["k"]), float(s["c"])) for s in springs]
    else:
        sh = [world.spring_insert(bh[a], bh[b], rest, k, c) for (a, b, rest, k, c) in springs]
    if isinstance(joints, np.ndarray):    # (n, 2) body indices
        if hasattr(world, "insert_joints") and len(joints):
            jh = world.insert_joints(bh[joints[:, 0]], bh[joints[:, 1]])
        else:
            jh = [world.joint_insert(bh[a], bh[b]) for a, b in joints]
    else:
        jh = [world.joint_insert(bh[a], bh[b]) for (a, b) in joints]
    return sh, jh


SPRING_DTYPE = np.dtype([("a", "<i8"), ("b", "<i8"), ("rest", "<f4"), ("k", "<f4"), ("c", "<f4")])


def _spheres(name, pos, radius, vel_req=None, gravity=(0.0, -30.0)):
    n = len(pos)
    s = Scene(name, gravity)
    b = A.body_descs(n)
    b["position"]["x"] = pos[:, 0]
    b["position"]["y"] = pos[:, 1]
    b["position_old"] = b["position"]          # RigidBodyBuilder::position sets both (rigid_body.rs:320-324)
    if vel_req is not None:
        b["has_velocity_request"] = 1
        b["velocity_request"]["x"] = vel_req[:, 0]
        b["velocity_request"]["y"] = vel_req[:, 1]
    c = A.collider_descs(n)
    c["radius"] = radius
    c["shape_radius"] = radius
    # spawn_rbd_entity passes absolute_transform = from_translation(position) (demo/src/simulation.rs:107)
    c["absolute_transform"]["translation"] = b["position"]
    s.bodies, s.colliders = b, c
    s.col_parent = np.arange(n, dtype=np.int64)
    return s


def lattice_scene(nx, ny, pitch, centre, seed, r_lo, r_hi, jitter=0.0, vel_disc=0.0, constraint_r=None, name="lattice",
                  cell_size=None):
    """nx*ny single-collider dynamic spheres on a jittered lattice (row-major slots)."""
    n = nx * ny
    ix = (np.arange(n) % nx).astype(np.float32)
    iy = (np.arange(n) // nx).astype(np.float32)
    x = (ix - np.float32((nx - 1) / 2.0)) * np.float32(pitch) + np.float32(centre[0])
    y = (iy - np.float32((ny - 1) / 2.0)) * np.float32(pitch) + np.float32(centre[1])
    if jitter:
        x = x + (uniform(seed, 1, n) - np.float32(0.5)) * np.float32(jitter)
        y = y + (uniform(seed, 2, n) - np.float32(0.5)) * np.float32(jitter)
    pos = np.stack([x, y], axis=1).astype(np.float32)
    if r_hi > r_lo:
        radius = (np.float32(r_lo) + uniform(seed, 3, n) * np.float32(r_hi - r_lo)).astype(np.float32)
    else:
        radius = np.full(n, r_lo, dtype=np.float32)
    vel = None
    if vel_disc:
        ang = uniform(seed, 4, n) * np.float32(2 * np.pi)
        rad = np.sqrt(uniform(seed, 5, n)) * np.float32(vel_disc)
        vel = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1).astype(np.float32)
    s = _spheres(name, pos, radius, vel)
    if constraint_r is not None:
        s.constraints.append((0.0, 0.0, float(constraint_r)))
    s.cell_size = cell_size
    return s


def cfg1(seed=1, n_side=32):
    """'benches scene': 1024 spheres r~U[0.05,0.2) falling inside a circle constraint."""
    return lattice_scene(n_side, n_side, 0.45, (0.0, 6.0 if n_side == 32 else 0.0), seed, 0.05, 0.2, jitter=0.05, vel_disc=3.0,
                         constraint_r=8.0 if n_side == 32 else 4.0, name=f"cfg1_{n_side * n_side}")


def cfg2(seed=1, side=1024, varied=False):
    """1M single-collider spheres in one world: r=0.5 (or U[0.25,0.5)), pitch 1.05, circle R=800."""
    r_lo, r_hi = (0.25, 0.5) if varied else (0.5, 0.5)
    return lattice_scene(side, side, 1.05, (0.0, 0.0), seed, r_lo, r_hi, jitter=0.04, vel_disc=1.0,
                         constraint_r=800.0 * side / 1024.0, name=f"cfg2_{side * side}", cell_size=1.0)


def cfg2_dense(seed=1, side=256):
    """Contact-rich cfg2 variant: the lattice starts overlapping (pitch 0.9 < 2r), so every sphere has ~4 contacts per
    substep while the block relaxes; the circle is roomy enough to contain the lattice corners."""
    return lattice_scene(side, side, 0.9, (0.0, 0.0), seed, 0.5, 0.5, jitter=0.08, vel_disc=2.0,
                         constraint_r=0.9 * side * 0.75, name=f"cfg2_dense_{side * side}", cell_size=1.0)


def cfg3(n_worlds=4096, side=16, seed=1):
    """Batched independent worlds: n_worlds x (side*side) bodies, each world = cfg1 at that size
    (constraint radius 4 as in demo/src/main.rs:51-54), per-world seed = seed + world id. Returns one Scene per world."""
    return [cfg1(seed + w, n_side=side) for w in range(n_worlds)]


def build_batch(world, scene_list):
    """Insert a list of per-world Scenes into ONE GPU world as batched independent worlds (BLOBS_PARAM_BATCH_WORLD).
    Gravity/constraints are shared (taken from the first scene). Returns the per-world handle dicts."""
    first = scene_list[0]
    for (x, y, r) in first.constraints:
        world.constraint_push((x, y), r)
    if first.cell_size is not None:
        world.set_param(A.PARAM_CELL_SIZE, first.cell_size)
    out = []
    for w, sc in enumerate(scene_list):
        world.set_param(A.PARAM_BATCH_WORLD, w)
        bh = world.insert_bodies(sc.bodies)
        ch = world.insert_colliders(sc.colliders, bh[sc.col_parent]) if sc.n_colliders else np.zeros(0, np.uint64)
        sh, jh = _insert_links(world, sc, bh)
        out.append({"bodies": bh, "colliders": ch, "springs": sh, "joints": jh})
    world.set_param(A.PARAM_BATCH_WORLD, 0)
    return out


def cfg4(n_blobs=100_000, k=16, seed=1):
    """Soft blobs: rings of k single-collider bodies r=0.1 on a circle of radius 0.5; adjacent bodies
    joined by fixed joints, second neighbours and opposite bodies by springs k=1000 c=50; lattice pitch 1.6."""
    side = int(np.ceil(np.sqrt(n_blobs)))
    n = n_blobs * k
    blob = np.arange(n) // k
    j = np.arange(n) % k
    bx = ((blob % side).astype(np.float32) - np.float32((side - 1) / 2.0)) * np.float32(1.6)
    by = ((blob // side).astype(np.float32) - np.float32((side - 1) / 2.0)) * np.float32(1.6)
    ang = (j.astype(np.float32) * np.float32(2 * np.pi / k)).astype(np.float32)
    jit = (uniform(seed, 7, n) - np.float32(0.5)) * np.float32(0.01)
    pos = np.stack([bx + np.float32(0.5) * np.cos(ang) + jit, by + np.float32(0.5) * np.sin(ang)], axis=1).astype(np.float32)
    s = _spheres(f"cfg4_{n_blobs}x{k}", pos, np.full(n, 0.1, dtype=np.float32))
    # roomy circle: contains the lattice corners (half-diagonal 1.13 * side), R = 400 at the full 317 x 316 lattice (SURVEY cfg4)
    s.constraints.append((0.0, 0.0, float(1.26 * side + 2.0)))
    base = (blob * k).astype(np.int64)
    nxt = base + (j + 1) % k
    s.joints = np.stack([np.arange(n, dtype=np.int64), nxt], axis=1)
    parts = []
    for step in (2, k // 2):
        other = base + (j + step) % k
        a = np.arange(n, dtype=np.int64)
        keep = a < other if step == k // 2 else np.ones(n, dtype=bool)
        d = pos[other] - pos
        rest = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)
        sp = np.zeros(int(keep.sum()), dtype=SPRING_DTYPE)
        sp["a"], sp["b"], sp["rest"], sp["k"], sp["c"] = a[keep], other[keep], rest[keep], 1000.0, 50.0
        parts.append(sp)
    s.springs = np.concatenate(parts)
    return s
