"""Pins the cull oracle on the reference's own known-answer tests (test/src/shape_test.cpp)."""
import numpy as np

from darmok_b200 import synth
from oracle import oracle_py as O

IDENTITY = np.eye(4, dtype=np.float32).reshape(16)


def test_plane_distance_kat():
    # shape_test.cpp:52-59 "Plane distance": Plane((-1,0,0), 2)
    n, d = (-1.0, 0.0, 0.0), 2.0
    assert abs(O.plane_signed_distance(n, d, (1, 0, 0)) - (-3.0)) < 1e-6
    assert abs(O.plane_signed_distance(n, d, (-1, 0, 0)) - (-1.0)) < 1e-6
    assert abs(O.plane_signed_distance(n, d, (-3, 0, 0)) - 1.0) < 1e-6


def test_frustum_planes_kat():
    # shape_test.cpp:61-84 "Frustum planes are correct": Frustum(Math::ortho((-1,-1),(1,1),0,1)).
    # bx::mtxOrtho(-1,1,-1,1,0,1, offset 0, homogeneousDepth=false, left-handed) is the identity matrix and
    # Math::getNormalizedNearDepth() is 0 with the zero-initialised bgfx caps the test runs under.
    corners = O.frustum_corners(IDENTITY, 0.0)
    planes = O.frustum_planes(corners)
    expected = {  # PlaneType order Near, Far, Bottom, Top, Left, Right; REQUIRE(... == ...) is exact
        0: ((0, 0, -1), 0.0), 1: ((0, 0, 1), 1.0), 2: ((0, -1, 0), 1.0),
        3: ((0, 1, 0), 1.0), 4: ((-1, 0, 0), 1.0), 5: ((1, 0, 0), 1.0)}
    for i, (normal, dist) in expected.items():
        assert tuple(planes[i, :3]) == tuple(float(x) for x in normal), (i, planes[i])
        assert planes[i, 3] == dist, (i, planes[i])


def test_frustum_corner_order():
    # include/darmok/shape.hpp:372-383 CornerType: NBL NBR NTL NTR FBL FBR FTL FTR
    c = O.frustum_corners(IDENTITY, 0.0)
    assert np.array_equal(c, np.array([[-1, -1, 0], [1, -1, 0], [-1, 1, 0], [1, 1, 0],
                                       [-1, -1, 1], [1, -1, 1], [-1, 1, 1], [1, 1, 1]], np.float32))


def test_can_see_unit_box_cases():
    corners = O.frustum_corners(IDENTITY, 0.0)   # the box [-1,1]^2 x [0,1]
    inside = np.array([-0.1, -0.1, 0.4, 0.1, 0.1, 0.6], np.float32)
    outside = np.array([2.0, 2.0, 0.4, 3.0, 3.0, 0.6], np.float32)
    straddle = np.array([0.9, -0.1, 0.4, 1.5, 0.1, 0.6], np.float32)
    touching = np.array([1.0, -0.1, 0.4, 1.5, 0.1, 0.6], np.float32)  # a corner with sd == 0 is NOT "in front"
    assert O.cull_visible(corners, IDENTITY, inside)
    assert not O.cull_visible(corners, IDENTITY, outside)
    assert O.cull_visible(corners, IDENTITY, straddle)
    assert O.cull_visible(corners, IDENTITY, touching)
    # a translated entity: world = T(3,0,0) -> worldInverse = T(-3,0,0); the same box seen from x in [2,4] is outside
    wi = np.eye(4, dtype=np.float32); wi[0, 3] = -3.0
    assert not O.cull_visible(corners, wi.T.reshape(16), inside)


def test_mat4_inverse_matches_numpy():
    rng = np.random.default_rng(0)
    for _ in range(20):
        m = rng.normal(size=(4, 4)).astype(np.float32)
        inv = O.mat4_inverse(m.T.reshape(16)).reshape(4, 4).T
        assert np.allclose(inv, np.linalg.inv(m.astype(np.float64)), rtol=2e-3, atol=2e-4)


def test_batch_equals_scalar_and_has_both_outcomes():
    inst = synth.make_instances(4000, seed=3, extent=60.0)
    corners = O.frustum_corners(synth.sample_camera_view_proj(), 0.0)
    bits = O.cull_batch(corners, inst.world_inverse, inst.bbox)
    vis = np.unpackbits(bits.view(np.uint8), bitorder="little")[:4000].astype(bool)
    one = np.array([O.cull_visible(corners, inst.world_inverse[i], inst.bbox[i]) for i in range(4000)])
    assert np.array_equal(vis, one)
    assert 0 < vis.sum() < 4000


def test_transform_update_world_inverse():
    # Transform::update (src/transform.cpp:286-309) without parent: inverse(translate * mat4_cast * scale)
    tr = synth.make_transforms(50, seed=1)
    for i in range(50):
        local = O.transform_local_matrix(tr.position[i], tr.rotation[i], tr.scale[i]).reshape(4, 4).T.astype(np.float64)
        x, y, z, w = tr.rotation[i].astype(np.float64)
        r = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.allclose(local[:3, :3], r * tr.scale[i][None, :], atol=1e-6)
        assert np.array_equal(local[:3, 3].astype(np.float32), tr.position[i]) and np.array_equal(local[3], [0, 0, 0, 1])
        wi = O.transform_world_inverse(tr.position[i], tr.rotation[i], tr.scale[i]).reshape(4, 4).T.astype(np.float64)
        assert np.allclose(wi @ local, np.eye(4), atol=5e-3)   # float32 inverse, positions up to 500
    # identity transform: exact identity
    ident = O.transform_world_inverse([0, 0, 0], [0, 0, 0, 1], [1, 1, 1])
    assert np.array_equal(ident, np.eye(4, dtype=np.float32).reshape(16))


def make_forest(m, seed=0):
    """random transform forest in shuffled order (a parent may come after its child), depth <= 6"""
    rng = np.random.default_rng(seed)
    parent = np.full(m, -1, np.int32)
    depth = np.zeros(m, np.int32)
    order = rng.permutation(m)
    for k, i in enumerate(order):
        if k and rng.random() < 0.75:
            cand = order[rng.integers(0, k)]
            if depth[cand] < 5:
                parent[i], depth[i] = cand, depth[cand] + 1
    q = rng.normal(size=(m, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)
    pos = rng.uniform(-20, 20, (m, 3)).astype(np.float32)
    scl = rng.uniform(0.5, 2.0, (m, 3)).astype(np.float32)
    return pos, q, scl, parent


def test_transform_hierarchy_matches_explicit_products():
    # Transform::update with _parent (src/transform.cpp:300-305): world = parent.world * local,
    # worldInverse = localInverse * parent.worldInverse, parents first
    pos, q, scl, parent = make_forest(60, seed=3)
    world, inv = O.transform_nodes(pos, q, scl, parent)
    assert (parent >= 0).sum() > 30 and (parent == -1).sum() >= 1

    def chain(i):
        local = O.transform_local_matrix(pos[i], q[i], scl[i])
        linv = O.transform_world_inverse(pos[i], q[i], scl[i])
        if parent[i] < 0:
            return local, linv
        pw, pi = chain(parent[i])
        return O.mat4_mul(pw, local), O.mat4_mul(linv, pi)

    for i in range(60):
        w, wi = chain(i)
        assert np.array_equal(world[i], w) and np.array_equal(inv[i], wi), f"node {i}"
        a = world[i].reshape(4, 4).T.astype(np.float64) @ inv[i].reshape(4, 4).T.astype(np.float64)
        assert np.allclose(a, np.eye(4), atol=2e-3)


def test_parented_cull_equals_cull_on_composed_inverse():
    pos, q, scl, parent = make_forest(12, seed=5)
    _, node_inv = O.transform_nodes(pos, q, scl, parent)
    tr = synth.make_transforms(500, seed=9)
    rng = np.random.default_rng(1)
    ep = rng.integers(-1, 12, 500).astype(np.int32)
    corners = O.frustum_corners(synth.sample_camera_view_proj(), 0.0)
    bits, wi = O.cull_batch_trs_parented(corners, tr.position, tr.rotation, tr.scale, tr.bbox, ep, node_inv)
    ref = O.cull_batch(corners, wi, tr.bbox)
    assert np.array_equal(bits, ref)
    free = ep < 0
    bits_free, wi_free = O.cull_batch_trs(corners, tr.position, tr.rotation, tr.scale, tr.bbox, want_world_inverse=True)
    assert np.array_equal(wi[free], wi_free[free])


def test_math_transform_conventions_match_reference_plane_kat():
    """test/src/shape_test.cpp:9-50 ("Plane transform") pins what Transform::update builds its matrices from: the
    translation column of Math::transform, the diagonal of glm::scale, and the rotation sense / component order of
    glm::mat4_cast(glm::quat).  Plane::operator*= (src/shape.cpp:527-537) restated: origin = normal * distance;
    origin' = M * (origin, 1) / w; normal' = M * (normal, 0); distance' = dot(normal', origin')."""
    ident_q, one = [0, 0, 0, 1], [1, 1, 1]

    def apply(plane, m16):
        m = np.asarray(m16, np.float32).reshape(4, 4).T
        normal, distance = plane
        origin = m @ np.append(normal * np.float32(distance), np.float32(1))
        origin = origin[:3] / origin[3]
        normal = (m @ np.append(normal, np.float32(0)))[:3]
        return normal, np.float32(np.dot(normal, origin))

    plane = (np.array([0, 0, 1], np.float32), np.float32(1))
    plane = apply(plane, O.transform_local_matrix([1, 0, 0], ident_q, one))          # Math::transform(vec3(1, 0, 0))
    assert np.array_equal(plane[0], [0, 0, 1]) and plane[1] == 1
    plane = apply(plane, O.transform_local_matrix([0, 0, 3], ident_q, one))
    assert np.array_equal(plane[0], [0, 0, 1]) and plane[1] == 4
    scale = O.transform_local_matrix([0, 0, 0], ident_q, [2, 3, 4])                   # glm::scale(mat4(1), vec3(2, 3, 4))
    plane = apply(plane, scale)
    assert np.array_equal(plane[0], [0, 0, 4]) and plane[1] == 64
    # glm::quat(glm::radians(vec3(0, -90, 0))): a pure rotation about y, (x, y, z, w) = (0, sin(-45 deg), 0, cos(-45 deg))
    half = np.radians(-90.0) / 2
    roty = O.transform_local_matrix([0, 0, 0], [0, np.sin(half), 0, np.cos(half)], one)
    plane = apply((np.array([0, 0, 1], np.float32), np.float32(1)), roty)
    assert np.allclose(plane[0], [-1, 0, 0], atol=1e-6) and np.isclose(plane[1], 1, atol=1e-6)
    plane = apply(plane, scale)
    assert np.allclose(plane[0], [-2, 0, 0], atol=1e-6) and np.isclose(plane[1], 4, atol=1e-5)
    plane = apply((np.array([0, 0, 1], np.float32), np.float32(0)), roty)
    assert np.allclose(plane[0], [-1, 0, 0], atol=1e-6) and np.isclose(plane[1], 0, atol=1e-6)
    plane = apply((np.array([0, 0, -1], np.float32), np.float32(1)), roty)
    assert np.allclose(plane[0], [1, 0, 0], atol=1e-6) and np.isclose(plane[1], 1, atol=1e-6)


def test_product_frustum_corners_equal_the_oracle_on_the_host():
    # dmk_frustum_corners (Frustum::Frustum(mtx), src/shape.cpp:1141-1163) is host code of the product and needs no device:
    # bit-identical to the oracle's restatement for the sample camera and for random view-projection matrices
    import ctypes
    from darmok_b200 import build, capi, synth
    build.build()
    rng = np.random.default_rng(5)
    mats = [synth.sample_camera_view_proj()] + [rng.normal(size=(4, 4)).astype(np.float32) for _ in range(200)]
    for vp in mats:
        for nd in (0.0, -1.0):
            v = np.ascontiguousarray(vp, np.float32)
            out = np.zeros((8, 3), np.float32)
            st = capi.lib().dmk_frustum_corners(v.ctypes.data_as(ctypes.c_void_p), nd, out.ctypes.data_as(ctypes.c_void_p))
            assert st == 0
            ref = O.frustum_corners(v, nd)
            assert np.array_equal(out, ref, equal_nan=True)
