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
