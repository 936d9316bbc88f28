"""Second, independent restatement of the reference shaders in pure Python / numpy float32.

TEST INFRASTRUCTURE ONLY.  Written directly from the GLSL text with vector operations (one
numpy float32 array per GLSL vec3), independently of oracle.c, so that agreement between the
two pins the C oracle's reading of the shader (there is no runnable reference here; SURVEY
§8c).  Slow: use on small cases only.

Follows assets/shaders/map.glsl:21-47,57-60,83-248, primary.comp.glsl:23-68,
secondary.comp.glsl:18-50.
"""
import numpy as np

F = np.float32
EPSILON = F(0.001)
SUN_DIR = np.array([7.52185881e-01, 6.58950984e-01, 7.52185881e-01], dtype=F)
NORMALS = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]  # map.glsl:72-79
ENTITY_POSITIONS = [(256., 21., 256.), (251., 21., 259.), (253., 21., 256.), (251., 21., 256.), (257., 21., 261.)]


def _ivec(v):
    """ivec3(vec3): truncate toward zero, saturate, NaN -> 0."""
    out = []
    for x in v:
        x = float(x)
        if x != x:
            out.append(0)
        else:
            out.append(int(max(min(np.trunc(x), 2147483647), -2147483648)))
    return np.array(out, dtype=np.int64)


def _uvec(v):
    out = []
    for x in v:
        x = float(x)
        out.append(0 if (x != x or x <= 0) else int(min(np.trunc(x), 4294967295)))
    return np.array(out, dtype=np.int64)


def _gmin(a, b):
    return np.where(b < a, b, a).astype(F)


def _gmax(a, b):
    return np.where(a < b, b, a).astype(F)


def intersect_aabb(ro, rd, bmin, bmax):
    """map.glsl:21-29"""
    with np.errstate(divide="ignore", invalid="ignore"):
        t_min = ((bmin - ro).astype(F) / rd).astype(F)
        t_max = ((bmax - ro).astype(F) / rd).astype(F)
    t1, t2 = _gmin(t_min, t_max), _gmax(t_min, t_max)
    near = _gmax(_gmax(t1[0], t1[1]), t1[2])
    far = _gmin(_gmin(t2[0], t2[1]), t2[2])
    return F(near), F(far)


class World:
    def __init__(self, dim, chunks, bricks, atlas):
        self.dim, self.cd = dim, dim // 8
        self.chunks = np.asarray(chunks, dtype=np.uint32).reshape(-1)
        self.bricks = np.asarray(bricks, dtype=np.uint32).reshape(-1)
        self.atlas = np.asarray(atlas, dtype=np.uint32).reshape(256, 256, 256)  # [z][y][x]

    def chunk_flags(self, c):  # map.glsl:31-36
        if (c < 0).any() or (c >= self.cd).any():
            return 0
        return int(self.chunks[c[0] + self.cd * (c[1] + c[2] * self.cd)])

    def get_voxel(self, b):  # map.glsl:38-47
        idx = self.chunk_flags(b >> 3)
        if idx > 0:
            return int(self.bricks[(idx - 1) * 512 + (b[0] % 8) + ((b[2] % 8) * 8 + (b[1] % 8)) * 8]), True
        return 0, False

    def get_sub_voxel(self, mdl, p):  # map.glsl:57-60
        o = np.array([(mdl & 31) * 8, ((mdl // 32) & 31) * 8, ((mdl // 1024) & 31) * 8])
        q = p + o
        return int(self.atlas[q[2], q[1], q[0]])


def trace_map(world, ray_origin, ray_dir, max_steps):
    """map.glsl:83-168.  Returns dict(data, hit_pos, normal, p, face, block, trips, exit_kind, t_in, t_chunk, t_block)."""
    rd = np.array(ray_dir, dtype=F)
    ro = np.array(ray_origin, dtype=F)
    rd[rd == 0] = F(0.001)
    bounds = 8 * world.dim
    sgn = np.sign(rd).astype(np.int64)
    positivity = (1 + sgn) >> 1
    inv = (F(1.0) / rd).astype(F)
    min_idx = 0
    o8 = (ro * F(8.0)).astype(F)
    g = _ivec(o8)
    w = (o8 - g.astype(F)).astype(F)
    step = 0
    res = dict(data=0, hit_pos=(-1.0, -1.0, -1.0), normal=(0.0, 0.0, 0.0), p=(0xFFFFFFFF,) * 3, face=0, block=0,
               trips=0, exit_kind=1, t_in=0, t_chunk=0, t_block=0)
    for trip in range(max_steps):
        if (g >= bounds).any() or (g < 0).any():
            res["exit_kind"] = 2
            res["trips"] = trip
            return res
        res["t_in"] += 1
        p = g + _uvec(w)
        block, chunk_hit = world.get_voxel(p >> 3)
        res["t_chunk"] += int(chunk_hit)
        if block != 0:
            res["t_block"] += 1
            sub = world.get_sub_voxel(block & 0xFFFFFFF, p % 8)
            if sub != 0:
                face = {0: 2 - positivity[0], 1: 4 - positivity[1], 2: 6 - positivity[2]}[min_idx]
                hp = (g.astype(F) + w).astype(F)
                res.update(data=sub, hit_pos=tuple(float(x) for x in hp), normal=tuple(float(x) for x in NORMALS[face - 1]),
                           p=tuple(int(x) for x in p), face=int(face), block=block, trips=trip + 1, exit_kind=0)
                return res
            elif step != 0:
                g = g + _ivec(w)
                w = (w - np.floor(w)).astype(F)
                step = 0
        elif step != 3:
            w = (w + (g & 7).astype(F)).astype(F)
            g = g - (g & 7)
            step = 3
        t = ((((positivity << step).astype(F)) - w).astype(F) * inv).astype(F)
        min_idx = (0 if t[0] < t[2] else 2) if t[0] < t[1] else (1 if t[1] < t[2] else 2)
        g[min_idx] += int(sgn[min_idx] << step) if sgn[min_idx] >= 0 else -int((-sgn[min_idx]) << step)
        w = (w + (rd * t[min_idx]).astype(F)).astype(F)
        w[min_idx] = F(F((1 - positivity[min_idx]) << step) * F(0.999))
    res["trips"] = max_steps
    return res


def trace_entities(ro, rd, max_distance):
    """map.glsl:172-201 (live part).  True when HitInfo.data != 0."""
    ro = np.asarray(ro, dtype=F)
    rd = np.asarray(rd, dtype=F)
    prev_d = F(np.inf)
    chosen = None
    for i, pos in enumerate(ENTITY_POSITIONS):
        pos = np.array(pos, dtype=F)
        diff = (ro - pos).astype(F)
        sq = (diff * diff).astype(F)
        dist = np.sqrt(F(F(sq[0] + sq[1]) + sq[2]), dtype=F)
        if dist >= max_distance:
            continue
        near, far = intersect_aabb(ro, rd, pos, (pos + F(1.0)).astype(F))
        if far >= near and prev_d >= far:
            chosen, prev_d = i, far
    if chosen is None:
        return False
    pos = np.array(ENTITY_POSITIONS[chosen], dtype=F)
    near, far = intersect_aabb(ro, rd, pos, (pos + F(1.0)).astype(F))
    return bool(far >= near)


def trace_entities_models(model, ro, rd, max_distance, positions=ENTITY_POSITIONS, size=8, max_steps=64):
    """map.glsl:172-248 read with the early return of :199 deleted (the sub-model DDA made live).
    `model[z, y, x]` = size^3 texels; box edge = size / 8 blocks (1 for the literal 8^3).
    Returns dict(data, hit_pos, normal, face, p, entity, trips)."""
    ro = np.asarray(ro, dtype=F)
    rd = np.asarray(rd, dtype=F)
    edge = F(F(size) / F(8.0))
    miss = {"data": 0, "hit_pos": (0.0, 0.0, 0.0), "normal": (0.0, 0.0, 0.0), "face": 0, "p": None, "entity": None, "trips": 0}
    prev_d = F(np.inf)
    chosen = None
    for i, pos in enumerate(positions):
        pos = np.array(pos, dtype=F)
        diff = (ro - pos).astype(F)
        sq = (diff * diff).astype(F)
        if np.sqrt(F(F(sq[0] + sq[1]) + sq[2]), dtype=F) >= max_distance:
            continue
        near, far = intersect_aabb(ro, rd, pos, (pos + edge).astype(F))
        if far >= near and prev_d >= far:
            chosen, prev_d = i, far
    if chosen is None:
        return miss
    pos = np.array(positions[chosen], dtype=F)
    near, far = intersect_aabb(ro, rd, pos, (pos + edge).astype(F))
    if not far >= near:
        return miss
    bounds = size
    t0 = near if F(0.0) < near else F(0.0)                       # max(hit.x, 0)
    ro = (ro + (rd * t0).astype(F)).astype(F)                    # :204
    sgn = np.array([1 if c > 0 else (-1 if c < 0 else 0) for c in rd], dtype=np.int64)   # ivec3(sign(rayDir))
    positivity = (1 + sgn) >> 1
    with np.errstate(divide="ignore"):
        inv = (F(1.0) / rd).astype(F)
    min_idx = 0
    g = _ivec((((ro - EPSILON).astype(F) - pos).astype(F) * F(8.0)).astype(F))            # :211
    w = (((ro - pos).astype(F) * F(8.0)).astype(F) - g.astype(F)).astype(F)               # :212
    trips = 0
    for _ in range(max_steps):
        if (g >= bounds).any() or (g < 0).any():
            break
        trips += 1
        p = ((g & 0xFFFFFFFF) + _uvec(w)) & 0xFFFFFFFF
        q = p & (size - 1)
        block = int(model[q[2], q[1], q[0]])
        if block != 0:
            face = [2 - positivity[0], 4 - positivity[1], 6 - positivity[2]][min_idx]
            hp = (pos + ((g.astype(F) + w).astype(F) / F(8.0)).astype(F)).astype(F)       # :230
            return {"data": block, "hit_pos": tuple(float(x) for x in hp), "normal": tuple(float(x) for x in NORMALS[face - 1]),
                    "face": int(face), "p": tuple(int(x) for x in q), "entity": chosen, "trips": trips}
        g = g + _ivec(w)
        with np.errstate(invalid="ignore"):
            w = (w - np.floor(w)).astype(F)
            t = ((positivity.astype(F) - w).astype(F) * inv).astype(F)
        min_idx = (0 if t[0] < t[2] else 2) if t[0] < t[1] else (1 if t[1] < t[2] else 2)
        g[min_idx] += sgn[min_idx]
        with np.errstate(invalid="ignore"):
            w = (w + (rd * t[min_idx]).astype(F)).astype(F)
        w[min_idx] = F(F(1 - positivity[min_idx]) * F(0.999))
    return dict(miss, trips=trips)


def primary_ray(cam_pos, cam_mat, fov, W, H, px, py, map_dim):
    """primary.comp.glsl:31-43 -> (origin, dir, traceMap start)."""
    uv = (np.array([px, py], dtype=F) / np.array([W, H], dtype=F)).astype(F)
    uv = (uv * F(2.0) - F(1.0)).astype(F)
    uv[1] = F(uv[1] * F(F(H) / F(W)))
    uv = (uv * np.tan(F(fov) / F(2.0), dtype=F)).astype(F)
    M = np.asarray(cam_mat, dtype=F).reshape(4, 4)  # row j = GLSL column j
    vin = np.array([uv[0], uv[1], 1.0, 1.0], dtype=F)
    v4 = np.zeros(4, dtype=F)
    for i in range(4):
        acc = F(M[0, i] * vin[0])
        for j in (1, 2, 3):
            acc = F(acc + F(M[j, i] * vin[j]))
        v4[i] = acc
    sq = (v4 * v4).astype(F)
    ln = np.sqrt(F(F(F(sq[0] + sq[1]) + sq[2]) + sq[3]), dtype=F)
    rd = (v4[:3] / ln).astype(F)
    ro = np.asarray(cam_pos[:3], dtype=F)
    D = F(map_dim)
    near, _ = intersect_aabb(ro, rd, np.zeros(3, dtype=F), np.array([D, D, D], dtype=F))
    t0 = near if F(0.0) < near else F(0.0)  # max(intersection.x, 0)
    start = ((ro + (rd * t0).astype(F)).astype(F) - EPSILON).astype(F)
    return ro, rd, start


def shadow_origin(position_xyz, normal_rgba8):
    """secondary.comp.glsl:36-37 from the quantised G-buffer values."""
    n = np.array([(normal_rgba8 >> s) & 255 for s in (0, 8, 16)], dtype=F) / F(255.0)
    return (np.asarray(position_xyz, dtype=F) + (n.astype(F) * F(0.001)).astype(F)).astype(F)
