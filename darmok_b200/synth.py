"""Seeded synthetic assets for the BASELINE.json configs.

Everything here is workload generation: a 67-joint depth-first skeleton and looping clips written as real
ozz-format archives (Skeleton v2 / Animation v6) so that the product loader
(`dmk_skeleton_load_ozz` / `dmk_clip_load_ozz`) and the oracle parse the very same bytes, a 4-influence mesh
in the layout `MeshData::exportData` produces (/root/reference/src/mesh_core.cpp:323-365), an armature
(`ArmatureJoint{name, inverseBindPose}`, include/darmok/skeleton.hpp:363-367) and crowd / camera inputs.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

SQRT2 = 1.4142135623730950488016887242097


# ----------------------------------------------------------------------------------------------
# small quaternion / matrix helpers (float64, generation only)
# ----------------------------------------------------------------------------------------------
def quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
        aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def quat_axis_angle(axis, angle):
    axis = axis / np.linalg.norm(axis, axis=-1, keepdims=True)
    s = np.sin(angle * 0.5)[..., None]
    return np.concatenate([axis * s, np.cos(angle * 0.5)[..., None]], axis=-1)


def trs_matrix(t, q, s):
    """column-major-in-math (M @ v) 4x4 from translation, quaternion (x,y,z,w), scale"""
    x, y, z, w = q
    r = np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    m = np.eye(4)
    m[:3, :3] = r * np.asarray(s)[None, :]
    m[:3, 3] = t
    return m


# ----------------------------------------------------------------------------------------------
# ozz archives
# ----------------------------------------------------------------------------------------------
@dataclass
class SynthSkeleton:
    names: list
    parents: np.ndarray       # int16 [J]
    rest_t: np.ndarray        # f32 [J,3]
    rest_r: np.ndarray        # f32 [J,4]  (x,y,z,w)
    rest_s: np.ndarray        # f32 [J,3]
    archive: bytes

    @property
    def num_joints(self):
        return len(self.names)


def soa_pack(t, r, s):
    """AoS joint transforms -> ozz SoaTransform array (40 floats per 4 joints), identity padded."""
    j = t.shape[0]
    soa = (j + 3) // 4
    tt = np.zeros((soa * 4, 3), np.float32); tt[:j] = t
    rr = np.zeros((soa * 4, 4), np.float32); rr[:, 3] = 1.0; rr[:j] = r
    ss = np.ones((soa * 4, 3), np.float32); ss[:j] = s
    out = np.zeros((soa, 10, 4), np.float32)
    for g in range(soa):
        sl = slice(4 * g, 4 * g + 4)
        out[g, 0:3] = tt[sl].T
        out[g, 3:7] = rr[sl].T
        out[g, 7:10] = ss[sl].T
    return out.reshape(soa, 40)


def write_skeleton_archive(names, parents, rest_t, rest_r, rest_s) -> bytes:
    j = len(names)
    buf = bytearray()
    buf += struct.pack("<B", 1)                       # endianness tag: little
    buf += b"ozz-skeleton\0"
    buf += struct.pack("<I", 2)                       # version
    buf += struct.pack("<i", j)
    if j == 0:
        return bytes(buf)
    chars = b"".join(n.encode() + b"\0" for n in names)
    buf += struct.pack("<i", len(chars))
    buf += chars
    buf += np.asarray(parents, "<i2").tobytes()
    buf += soa_pack(rest_t, rest_r, rest_s).astype("<f4").tobytes()
    return bytes(buf)


def make_skeleton(num_joints=67, seed=1234, max_depth=12) -> SynthSkeleton:
    rng = np.random.default_rng(seed)
    parents = np.full(num_joints, -1, np.int16)
    depth = np.zeros(num_joints, np.int32)
    for i in range(1, num_joints):
        # parent drawn from the ancestor chain of i-1 (incl. i-1) => depth-first order, subtrees contiguous
        chain = []
        a = i - 1
        while a >= 0:
            chain.append(a)
            a = parents[a]
        chain = [c for c in chain if depth[c] + 1 <= max_depth]
        # bias towards deeper ancestors so that limbs form
        p = chain[min(int(rng.geometric(0.55)) - 1, len(chain) - 1)]
        parents[i] = p
        depth[i] = depth[p] + 1
    rest_t = rng.uniform(-0.3, 0.3, (num_joints, 3))
    axis = rng.normal(size=(num_joints, 3))
    angle = rng.uniform(0, np.pi / 6, num_joints)
    rest_r = quat_axis_angle(axis, angle)
    rest_s = np.ones((num_joints, 3))
    names = [f"joint_{i:02d}" for i in range(num_joints)]
    t32, r32, s32 = rest_t.astype(np.float32), rest_r.astype(np.float32), rest_s.astype(np.float32)
    return SynthSkeleton(names, parents, t32, r32, s32, write_skeleton_archive(names, parents, t32, r32, s32))


def compress_quat(q):
    """ozz offline CompressQuat: largest component dropped, the others * 32767*sqrt(2), rounded."""
    q = np.asarray(q, np.float64)
    largest = np.argmax(np.abs(q), axis=-1)
    sign = (np.take_along_axis(q, largest[..., None], -1)[..., 0] < 0).astype(np.uint8)
    mapping = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
    others = np.take_along_axis(q, mapping[largest], -1)
    quant = np.floor(others * (32767.0 * SQRT2) + 0.5)
    quant = np.clip(quant, -32767, 32767).astype(np.int16)
    return largest.astype(np.uint8), sign, quant


def _sorted_stream(ratios_per_track):
    """ozz AnimationBuilder key order: sort by (previous key ratio of the same track, track); the first
    key of a track has previous = -1.  Returns (track, key_in_track) arrays in stream order."""
    tr, ki, prev = [], [], []
    for t, r in enumerate(ratios_per_track):
        n = len(r)
        tr.append(np.full(n, t)); ki.append(np.arange(n))
        prev.append(np.concatenate([[-1.0], r[:-1]]))
    tr, ki, prev = np.concatenate(tr), np.concatenate(ki), np.concatenate(prev)
    order = np.lexsort((ki, tr, prev))
    return tr[order], ki[order]


def compress_quat_v7(q):
    """ozz >= 0.15 QuaternionKey: largest component dropped, the others mapped from [-1/sqrt 2, 1/sqrt 2] to 15 bits."""
    q = np.asarray(q, np.float64)
    largest = np.argmax(np.abs(q), axis=-1)
    sign = (np.take_along_axis(q, largest[..., None], -1)[..., 0] < 0).astype(np.uint8)
    mapping = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
    others = np.take_along_axis(q, mapping[largest], -1)
    quant = np.floor((others + 1.0 / SQRT2) * (32767.0 / SQRT2) + 0.5)
    quant = np.clip(quant, 0, 32767).astype(np.int64)
    return largest.astype(np.uint8), sign, quant


def write_animation_archive_v7(name, duration, num_tracks, t_tracks, r_tracks, s_tracks) -> bytes:
    """The same clip as an ozz-animation >= 0.15 archive (Animation version 7): shared timepoints, per component ratio indices
    (u8 / u16), previouses (distance back to the previous key of the same track), no iframes, 3 x u16 values per key.
    Layout from memory of upstream (SURVEY.md Appendix A.3'; darmok_b200/csrc/ozz_io.hpp): unvalidated against ozz itself."""
    timepoints = np.unique(np.concatenate([np.asarray(t[0], np.float32) for tracks in (t_tracks, r_tracks, s_tracks) for t in tracks]))
    wide = len(timepoints) > 256

    def ctrl(tracks, values_of):
        tr, ki = _sorted_stream([t[0] for t in tracks])
        ratios = bytearray()
        previouses = bytearray()
        values = bytearray()
        last = {}
        for i, (a, b) in enumerate(zip(tr, ki)):
            index = int(np.searchsorted(timepoints, np.float32(tracks[a][0][b])))
            assert timepoints[index] == np.float32(tracks[a][0][b])
            ratios += struct.pack("<H" if wide else "<B", index)
            back = i - last[a] if a in last else 0
            assert 0 <= back <= 0xffff, "a track's keys are further apart in the stream than a u16 offset reaches"
            previouses += struct.pack("<H", back)
            last[a] = i
            values += values_of(a, b)
        return len(tr), bytes(ratios) + bytes(previouses) + struct.pack("<f", 0.0) + bytes(values)   # no iframes, interval 0

    def f3_values(tracks):
        halves = [np.asarray(t[1], np.float32).astype(np.float16).view(np.uint16) for t in tracks]
        return lambda a, b: struct.pack("<3H", int(halves[a][b][0]), int(halves[a][b][1]), int(halves[a][b][2]))

    def q_values(tracks):
        comp = [compress_quat_v7(t[1]) for t in tracks]

        def one(a, b):
            lg, sg, v = comp[a]
            packed = int(lg[b]) | (int(sg[b]) << 2) | (int(v[b][0]) << 3) | (int(v[b][1]) << 18) | (int(v[b][2]) << 33)
            return struct.pack("<3H", packed & 0xffff, (packed >> 16) & 0xffff, (packed >> 32) & 0xffff)
        return one

    nt, bt = ctrl(t_tracks, f3_values(t_tracks))
    nr, br = ctrl(r_tracks, q_values(r_tracks))
    ns, bs = ctrl(s_tracks, f3_values(s_tracks))
    nm = name.encode()
    buf = bytearray()
    buf += struct.pack("<B", 1)
    buf += b"ozz-animation\0"
    buf += struct.pack("<I", 7)
    buf += struct.pack("<fIII", float(duration), int(num_tracks), len(nm), len(timepoints))
    buf += struct.pack("<III", nt, nr, ns)
    buf += struct.pack("<IIIIII", 0, 0, 0, 0, 0, 0)   # iframe entries / descriptors of the three components: none
    buf += nm
    buf += timepoints.astype("<f4").tobytes()
    buf += bt + br + bs
    return bytes(buf)


def write_animation_archive(name, duration, num_tracks, t_tracks, r_tracks, s_tracks, version=6) -> bytes:
    """*_tracks: list (len = padded track count) of (ratios f32[n], values) with values f32[n,3] (T,S) or
    f32[n,4] (R).  Values are quantised here (half / smallest-three 16 bit).  version: ozz Animation archive version (6 or 7)."""
    if version == 7:
        return write_animation_archive_v7(name, duration, num_tracks, t_tracks, r_tracks, s_tracks)
    assert version == 6
    def pack_f3(tracks):
        tr, ki = _sorted_stream([t[0] for t in tracks])
        out = bytearray()
        halves = [np.asarray(t[1], np.float32).astype(np.float16).view(np.uint16) for t in tracks]
        for a, b in zip(tr, ki):
            h = halves[a][b]
            out += struct.pack("<fH3H", float(tracks[a][0][b]), int(a), int(h[0]), int(h[1]), int(h[2]))
        return len(tr), out

    def pack_q(tracks):
        tr, ki = _sorted_stream([t[0] for t in tracks])
        comp = [compress_quat(t[1]) for t in tracks]
        out = bytearray()
        for a, b in zip(tr, ki):
            lg, sg, v = comp[a]
            out += struct.pack("<fHBB3h", float(tracks[a][0][b]), int(a), int(lg[b]), int(sg[b]),
                               int(v[b][0]), int(v[b][1]), int(v[b][2]))
        return len(tr), out

    nt, bt = pack_f3(t_tracks)
    nr, br = pack_q(r_tracks)
    ns, bs = pack_f3(s_tracks)
    nm = name.encode()
    buf = bytearray()
    buf += struct.pack("<B", 1)
    buf += b"ozz-animation\0"
    buf += struct.pack("<I", 6)
    buf += struct.pack("<fiiiii", float(duration), int(num_tracks), len(nm), nt, nr, ns)
    buf += nm
    buf += bt + br + bs
    return bytes(buf)


@dataclass
class SynthClip:
    name: str
    duration: float
    archive: bytes


def make_clip(skel: SynthSkeleton, name="clip", seed=0, duration=1.5, hz=30.0, drop=0.25, version=6) -> SynthClip:
    """Looping clip: uniform `hz` keys with `drop` of the interior keys removed per track and component.
    version: ozz Animation archive version the clip is written as (6 = ozz 0.14, 7 = ozz >= 0.15); same keys either way."""
    rng = np.random.default_rng(seed + 1000003)
    j = skel.num_joints
    padded = (j + 3) // 4 * 4
    nkeys = int(round(duration * hz)) + 1          # 46 for 1.5 s @ 30 Hz
    base = np.arange(nkeys, dtype=np.float64) / (nkeys - 1)
    base_f32 = base.astype(np.float32)
    base_f32[-1] = 1.0

    def keep_mask():
        m = rng.random(nkeys) >= drop
        m[0] = m[-1] = True
        return m

    t_tracks, r_tracks, s_tracks = [], [], []
    for tr in range(padded):
        if tr >= j:   # builder pads with identity keys at ratio 0 and 1
            r01 = np.array([0.0, 1.0], np.float32)
            t_tracks.append((r01, np.zeros((2, 3))))
            r_tracks.append((r01, np.tile([0.0, 0.0, 0.0, 1.0], (2, 1))))
            s_tracks.append((r01, np.ones((2, 3))))
            continue
        phase = rng.uniform(0, 2 * np.pi, 3)
        freq = rng.integers(1, 4, 3)
        amp = rng.uniform(0.0, 0.05, 3)
        tv = skel.rest_t[tr][None, :] + amp[None, :] * np.sin(2 * np.pi * base[:, None] * freq[None, :] + phase[None, :])
        axis = rng.normal(size=3)
        ang = rng.uniform(0.1, 0.6) * np.sin(2 * np.pi * base * rng.integers(1, 3) + rng.uniform(0, 2 * np.pi))
        rv = quat_mul(np.tile(skel.rest_r[tr].astype(np.float64), (nkeys, 1)), quat_axis_angle(np.tile(axis, (nkeys, 1)), ang))
        rv /= np.linalg.norm(rv, axis=-1, keepdims=True)
        sv = 1.0 + rng.uniform(0.0, 0.05) * np.sin(2 * np.pi * base[:, None] + rng.uniform(0, 2 * np.pi, 3)[None, :])
        mt, mr, ms = keep_mask(), keep_mask(), keep_mask()
        rk = rv[mr]
        for i in range(1, len(rk)):   # builder: adjacent keys on the same hemisphere (sampling does plain nlerp)
            if np.dot(rk[i - 1], rk[i]) < 0:
                rk[i] = -rk[i]
        t_tracks.append((base_f32[mt], tv[mt]))
        r_tracks.append((base_f32[mr], rk))
        s_tracks.append((base_f32[ms], sv[ms]))
    return SynthClip(name, float(np.float32(duration)),
                     write_animation_archive(name, duration, j, t_tracks, r_tracks, s_tracks, version=version))


# ----------------------------------------------------------------------------------------------
# armature + mesh
# ----------------------------------------------------------------------------------------------
def rest_models(skel: SynthSkeleton) -> np.ndarray:
    """float64 model-space rest matrices [J,4,4] (M @ v convention), right-handed ozz space"""
    out = np.zeros((skel.num_joints, 4, 4))
    for i in range(skel.num_joints):
        local = trs_matrix(skel.rest_t[i].astype(np.float64), skel.rest_r[i].astype(np.float64), skel.rest_s[i].astype(np.float64))
        out[i] = local if skel.parents[i] < 0 else out[skel.parents[i]] @ local
    return out


@dataclass
class SynthArmature:
    names: list               # armature joint k -> skeleton joint name
    remap: np.ndarray         # int32 [K] skeleton joint index of armature joint k
    inv_bind: np.ndarray      # f32 [K,16] column-major mat4


def make_armature(skel: SynthSkeleton, seed=7) -> SynthArmature:
    rng = np.random.default_rng(seed)
    remap = rng.permutation(skel.num_joints).astype(np.int32)
    z = np.diag([1.0, 1.0, -1.0, 1.0])
    models = rest_models(skel)
    inv = np.zeros((len(remap), 16), np.float32)
    for k, ji in enumerate(remap):
        m = np.linalg.inv(z @ models[ji] @ z)      # inverse of the left-handed bind matrix
        m[3] = [0, 0, 0, 1]
        inv[k] = m.T.reshape(16).astype(np.float32)   # column-major
    return SynthArmature([skel.names[i] for i in remap], remap, inv)


@dataclass
class SynthMesh:
    pos: np.ndarray       # f32 [V,3]
    nrm: np.ndarray
    tan: np.ndarray
    indices: np.ndarray   # f32 [V,4]  bone index into the armature, -1 = unused
    weights: np.ndarray   # f32 [V,4]


def make_mesh(num_vertices=5000, num_bones=67, seed=11, influence_dist=(0.1, 0.2, 0.3, 0.4),
              unskinned_fraction=0.0) -> SynthMesh:
    rng = np.random.default_rng(seed)
    v = num_vertices
    pos = rng.uniform(-1, 1, (v, 3)).astype(np.float32)

    def unit(n):
        a = rng.normal(size=(n, 3))
        return (a / np.linalg.norm(a, axis=1, keepdims=True)).astype(np.float32)

    nrm, tan = unit(v), unit(v)
    counts = rng.choice([1, 2, 3, 4], size=v, p=np.asarray(influence_dist) / np.sum(influence_dist))
    counts = np.minimum(counts, num_bones)
    indices = np.full((v, 4), -1.0, np.float32)
    weights = np.zeros((v, 4), np.float32)
    weights[:, 0] = 1.0                                   # VertexDataWriter default (src/vertex.cpp:228-235)
    for i in range(v):
        c = counts[i]
        b = rng.choice(num_bones, size=c, replace=False)
        w = np.sort(rng.dirichlet(np.ones(c)))[::-1].astype(np.float32)
        indices[i, :c] = b
        weights[i, :] = 0.0
        weights[i, :c] = w
    if unskinned_fraction > 0:
        un = rng.random(v) < unskinned_fraction          # C17: indices (-1,-1,-1,-1), weights (1,0,0,0)
        indices[un] = -1.0
        weights[un] = [1, 0, 0, 0]
    return SynthMesh(pos, nrm, tan, indices, weights)


# ----------------------------------------------------------------------------------------------
# crowd + camera
# ----------------------------------------------------------------------------------------------
@dataclass
class SynthCrowd:
    clip: np.ndarray      # int32 [L,N]
    ratio: np.ndarray     # f32 [L,N]
    weight: np.ndarray    # f32 [L,N]
    speed: np.ndarray     # f32 [L,N]


def make_crowd(n, num_layers, num_clips, seed=42) -> SynthCrowd:
    rng = np.random.default_rng(seed)
    clip = rng.integers(0, num_clips, (num_layers, n)).astype(np.int32)
    if num_layers == 2 and num_clips > 1:   # two distinct clips A, B
        same = clip[0] == clip[1]
        clip[1, same] = (clip[1, same] + 1) % num_clips
    ratio = rng.random((num_layers, n)).astype(np.float32)
    speed = rng.uniform(0.5, 1.5, (num_layers, n)).astype(np.float32)
    if num_layers == 1:
        weight = np.ones((1, n), np.float32)
    else:
        v = rng.random(n).astype(np.float32)          # transition form: (1 - v, v)
        weight = np.zeros((num_layers, n), np.float32)
        weight[0] = np.float32(1.0) - v
        weight[1] = v
    return SynthCrowd(clip, ratio, weight, speed)


def perspective_lh(fovy_deg, aspect, near, far, homogeneous=False) -> np.ndarray:
    """bx::mtxProj left-handed as Math::perspective uses it (src/math.cpp:44-49); column-major float32[16]."""
    h = 1.0 / np.tan(np.radians(fovy_deg) * 0.5)
    w = h / aspect
    diff = far - near
    aa = (far + near) / diff if homogeneous else far / diff
    bb = (2.0 * far * near) / diff if homogeneous else near * aa
    m = np.zeros(16, np.float32)
    m[0], m[5], m[10], m[11], m[14] = w, h, aa, 1.0, -bb
    return m


def look_at_lh(eye, at, up=(0, 1, 0)) -> np.ndarray:
    eye, at, up = (np.asarray(a, np.float64) for a in (eye, at, up))
    f = at - eye; f /= np.linalg.norm(f)
    s = np.cross(up, f); s /= np.linalg.norm(s)
    u = np.cross(f, s)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, f
    m[0, 3], m[1, 3], m[2, 3] = -s @ eye, -u @ eye, -f @ eye
    return m.T.reshape(16).astype(np.float32)   # column-major


def mat4_mul_cm(a, b) -> np.ndarray:
    """column-major float32 product a * b (float64 accumulate; generation only)"""
    am = np.asarray(a, np.float64).reshape(4, 4).T
    bm = np.asarray(b, np.float64).reshape(4, 4).T
    return (am @ bm).T.reshape(16).astype(np.float32)


def sample_camera_view_proj() -> np.ndarray:
    """samples/ozz camera: at (0,2,-2) looking at (0,1,0), perspective 60 deg, near 0.3, far 1000
    (/root/reference/samples/ozz/src/main.cpp:76-81), 16:9."""
    return mat4_mul_cm(perspective_lh(60.0, 16.0 / 9.0, 0.3, 1000.0), look_at_lh((0, 2, -2), (0, 1, 0)))


@dataclass
class SynthInstances:
    world_inverse: np.ndarray   # f32 [N,16] column-major
    bbox: np.ndarray            # f32 [N,6] min.xyz max.xyz


def make_instances(n, seed=5, extent=500.0) -> SynthInstances:
    rng = np.random.default_rng(seed)
    s = rng.uniform(0.5, 2.0, n)
    x = rng.uniform(-extent, extent, n)
    z = rng.uniform(-extent, extent, n)
    yaw = rng.uniform(0, 2 * np.pi, n)
    # world = T * R_y * S ; inverse = S^-1 * R_y^T * T^-1   (closed form, float64 then rounded)
    c, sn = np.cos(yaw), np.sin(yaw)
    inv = np.zeros((n, 4, 4))
    inv[:, 0, 0], inv[:, 0, 2] = c / s, -sn / s
    inv[:, 1, 1] = 1.0 / s
    inv[:, 2, 0], inv[:, 2, 2] = sn / s, c / s
    inv[:, 0, 3] = -(inv[:, 0, 0] * x + inv[:, 0, 2] * z)
    inv[:, 2, 3] = -(inv[:, 2, 0] * x + inv[:, 2, 2] * z)
    inv[:, 3, 3] = 1.0
    wi = inv.transpose(0, 2, 1).reshape(n, 16).astype(np.float32)
    half = np.stack([0.3 * s, 0.9 * s, 0.3 * s], 1)
    # bbox is in the entity's LOCAL space (culling.cpp:274-282 brings the frustum to local space)
    bbox = np.concatenate([-half, half], 1).astype(np.float32)
    return SynthInstances(wi, bbox)


@dataclass
class SynthTransforms:
    position: np.ndarray   # f32 [N,3]
    rotation: np.ndarray   # f32 [N,4] quaternion x, y, z, w
    scale: np.ndarray      # f32 [N,3]
    bbox: np.ndarray       # f32 [N,6]


def make_transforms(n, seed=5, extent=500.0, uniform_scale=False) -> SynthTransforms:
    """What darmok's Transform holds for each entity of the config-5 field: translation on the ground plane, yaw plus a
    small tilt, scale; local-space bounding box of a character."""
    rng = np.random.default_rng(seed)
    s = rng.uniform(0.5, 2.0, (n, 1 if uniform_scale else 3))
    s = np.repeat(s, 3, axis=1) if uniform_scale else s
    pos = np.stack([rng.uniform(-extent, extent, n), rng.uniform(-1, 1, n), rng.uniform(-extent, extent, n)], 1)
    yaw = quat_axis_angle(np.tile([0.0, 1.0, 0.0], (n, 1)), rng.uniform(0, 2 * np.pi, n))
    tilt = quat_axis_angle(rng.normal(size=(n, 3)), rng.uniform(0, 0.2, n))
    q = quat_mul(yaw, tilt)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    half = np.array([0.3, 0.9, 0.3])
    bbox = np.concatenate([np.tile(-half, (n, 1)), np.tile(half, (n, 1))], 1)
    return SynthTransforms(pos.astype(np.float32), q.astype(np.float32), s.astype(np.float32), bbox.astype(np.float32))
