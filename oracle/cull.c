/*
 * oracle/cull.c -- TEST INFRASTRUCTURE (see dmk_oracle.h).
 *
 * Restates FrustumCuller::updateCulled (src/culling.cpp:264-285) and the shape arithmetic it uses
 * (src/shape.cpp) in glm's scalar operation order.  Pinned by the reference's
 * KATs test/src/shape_test.cpp:52-59 ("Plane distance") and :61-84 ("Frustum planes are correct").
 * Must be compiled with -ffp-contract=off: the reference build does not contract a*b+c.
 */
#include "dmk_oracle.h"

#include <math.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z; } v3;

static v3 v3_sub(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
/* glm::dot(vec3): tmp = a * b; (tmp.x + tmp.y) + tmp.z */
static float v3_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
/* glm::cross */
static v3 v3_cross(v3 x, v3 y) {
    v3 r = {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
    return r;
}
/* glm::normalize: v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x) */
static v3 v3_normalize(v3 v) {
    const float inv = 1.f / sqrtf(v3_dot(v, v));
    v3 r = {v.x * inv, v.y * inv, v.z * inv};
    return r;
}
/* glm mat4 * vec4: (m[0] * v.x + m[1] * v.y) + (m[2] * v.z + m[3] * v.w), per component */
static void mat4_mul_vec4(const float* m, const float* v, float* out) {
    for (int i = 0; i < 4; ++i) out[i] = (m[0 + i] * v[0] + m[4 + i] * v[1]) + (m[8 + i] * v[2] + m[12 + i] * v[3]);
}

void orc_frustum_corners(const float* view_proj, float near_depth, float* corners) {
    /* Frustum::Frustum(mtx, inverse = false) (src/shape.cpp:1141-1163); near_depth =
     * Math::getNormalizedNearDepth() (src/math.cpp:32-36): -1 when homogeneousDepth else 0. */
    const float nd = near_depth;
    const float ndc[8][3] = {{-1, -1, nd}, {1, -1, nd}, {-1, 1, nd}, {1, 1, nd},
                             {-1, -1, 1},  {1, -1, 1},  {-1, 1, 1},  {1, 1, 1}};
    float inv[16];
    orc_mat4_inverse(view_proj, inv);
    for (int c = 0; c < 8; ++c) {
        const float v[4] = {ndc[c][0], ndc[c][1], ndc[c][2], 1.0f};
        float t[4];
        mat4_mul_vec4(inv, v, t);
        /* corner = tcorner / tcorner.w  (vec4 / scalar, then truncated to vec3) */
        corners[c * 3 + 0] = t[0] / t[3];
        corners[c * 3 + 1] = t[1] / t[3];
        corners[c * 3 + 2] = t[2] / t[3];
    }
}

/* Frustum::CornerType (include/darmok/shape.hpp:372-383) */
enum { NBL = 0, NBR, NTL, NTR, FBL, FBR, FTL, FTR };
/* Frustum::getPlane corner triples, PlaneType order Near, Far, Bottom, Top, Left, Right
 * (src/shape.cpp:1244-1255, include/darmok/shape.hpp:385-394) */
static const int kPlaneCorners[6][3] = {{NTR, NBR, NBL}, {FTL, FBL, FBR}, {FBL, NBL, NBR},
                                        {FTR, NTR, NTL}, {NTL, NBL, FBL}, {FTR, FBR, NBR}};

typedef struct { v3 normal; float distance; } plane_t;

static plane_t plane_from_triangle(v3 a, v3 b, v3 c) {
    /* Triangle::getNormal (src/shape.cpp:223-229) = cross(v1 - v0, v2 - v0);
     * Plane(Triangle) (src/shape.cpp:423-427): normal = normalize(...), distance = dot(normal, v0) */
    plane_t p;
    p.normal = v3_normalize(v3_cross(v3_sub(b, a), v3_sub(c, a)));
    p.distance = v3_dot(p.normal, a);
    return p;
}

static float plane_signed_distance(const plane_t* p, v3 point) {
    /* Plane::signedDistanceTo (src/shape.cpp:449-452): the stored normal is normalised AGAIN */
    return v3_dot(v3_normalize(p->normal), point) - p->distance;
}

float orc_plane_signed_distance(const float* normal, float distance, const float* point) {
    plane_t p = {{normal[0], normal[1], normal[2]}, distance};
    v3 pt = {point[0], point[1], point[2]};
    return plane_signed_distance(&p, pt);
}

static void frustum_planes(const v3* corners, plane_t* planes) {
    for (int p = 0; p < 6; ++p)
        planes[p] = plane_from_triangle(corners[kPlaneCorners[p][0]], corners[kPlaneCorners[p][1]],
                                        corners[kPlaneCorners[p][2]]);
}

void orc_frustum_planes(const float* corners, float* out) {
    v3 c[8];
    plane_t planes[6];
    for (int i = 0; i < 8; ++i) { c[i].x = corners[i * 3]; c[i].y = corners[i * 3 + 1]; c[i].z = corners[i * 3 + 2]; }
    frustum_planes(c, planes);
    for (int p = 0; p < 6; ++p) {
        out[p * 4 + 0] = planes[p].normal.x;
        out[p * 4 + 1] = planes[p].normal.y;
        out[p * 4 + 2] = planes[p].normal.z;
        out[p * 4 + 3] = planes[p].distance;
    }
}

int orc_cull_visible(const float* corners, const float* world_inverse, const float* bbox) {
    /* src/culling.cpp:274-282: frust = camFrust; frust *= trans->getWorldInverse();
     * culled iff !frust.canSee(bounds) */
    v3 c[8];
    for (int i = 0; i < 8; ++i) {
        /* Frustum::operator*= (src/shape.cpp:1317-1324): corner = trans * vec4(corner, 1) truncated to vec3 */
        const float v[4] = {corners[i * 3], corners[i * 3 + 1], corners[i * 3 + 2], 1.0f};
        float t[4];
        mat4_mul_vec4(world_inverse, v, t);
        c[i].x = t[0]; c[i].y = t[1]; c[i].z = t[2];
    }
    plane_t planes[6];
    frustum_planes(c, planes);
    /* BoundingBox::getCorners order (src/shape.cpp:1057-1069) */
    const float *mn = bbox, *mx = bbox + 3;
    const v3 bc[8] = {{mn[0], mn[1], mn[2]}, {mn[0], mx[1], mn[2]}, {mn[0], mn[1], mx[2]}, {mn[0], mx[1], mx[2]},
                      {mx[0], mn[1], mn[2]}, {mx[0], mx[1], mn[2]}, {mx[0], mn[1], mx[2]}, {mx[0], mx[1], mx[2]}};
    /* Frustum::canSee (src/shape.cpp:1213-1223) / Plane::isInFront(BoundingBox) (:459-469) */
    for (int p = 0; p < 6; ++p) {
        int in_front = 1;
        for (int k = 0; k < 8; ++k) {
            if (plane_signed_distance(&planes[p], bc[k]) <= 0.f) {
                in_front = 0;
                break;
            }
        }
        if (in_front) return 0;
    }
    return 1;
}

void orc_cull_batch(const float* corners, const float* world_inverse, const float* bbox, int64_t n,
                    uint32_t* vis_bits, int num_threads) {
    const int64_t words = (n + 31) / 32;
#ifdef _OPENMP
    if (num_threads < 1) num_threads = 1;
#pragma omp parallel for schedule(static) num_threads(num_threads)
#endif
    for (int64_t w = 0; w < words; ++w) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int64_t i = w * 32 + b;
            if (i >= n) break;
            if (orc_cull_visible(corners, world_inverse + i * 16, bbox + i * 6)) bits |= (1u << b);
        }
        vis_bits[w] = bits;
    }
    (void)num_threads;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * "next" row: Transform::update (src/transform.cpp:286-309) for an entity without parent:
 *   _localMatrix  = Math::transform(position, rotation, scale) = glm::translate(p) * glm::mat4_cast(q) * glm::scale(s)
 *   _localInverse = glm::inverse(_localMatrix);   _worldInverse = _localInverse
 * in glm's operation order (mat3_cast of glm/gtc/quaternion.inl; products as orc_mat4_mul; inverse as orc_mat4_inverse).
 * rotation is (x, y, z, w).
 * ---------------------------------------------------------------------------------------------- */
void orc_transform_local_matrix(const float* p, const float* q, const float* s, float* out) {
    const float qx = q[0], qy = q[1], qz = q[2], qw = q[3];
    const float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz;
    const float qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
    float r[16] = {0}, t[16] = {0}, sc[16] = {0}, tr[16];
    r[0] = 1.f - 2.f * (qyy + qzz); r[1] = 2.f * (qxy + qwz);       r[2] = 2.f * (qxz - qwy);
    r[4] = 2.f * (qxy - qwz);       r[5] = 1.f - 2.f * (qxx + qzz); r[6] = 2.f * (qyz + qwx);
    r[8] = 2.f * (qxz + qwy);       r[9] = 2.f * (qyz - qwx);       r[10] = 1.f - 2.f * (qxx + qyy);
    r[15] = 1.f;
    t[0] = t[5] = t[10] = t[15] = 1.f;
    t[12] = p[0]; t[13] = p[1]; t[14] = p[2];
    sc[0] = s[0]; sc[5] = s[1]; sc[10] = s[2]; sc[15] = 1.f;
    orc_mat4_mul(t, r, tr);
    orc_mat4_mul(tr, sc, out);
}

void orc_transform_world_inverse(const float* p, const float* q, const float* s, float* out) {
    float local[16];
    orc_transform_local_matrix(p, q, s, local);
    orc_mat4_inverse(local, out);
}

void orc_cull_batch_trs(const float* corners, const float* position, const float* rotation, const float* scale,
                        const float* bbox, int64_t n, uint32_t* vis_bits, float* world_inverse_out, int num_threads) {
    const int64_t words = (n + 31) / 32;
    if (num_threads < 1) num_threads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(num_threads)
#endif
    for (int64_t w = 0; w < words; ++w) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int64_t i = w * 32 + b;
            if (i >= n) break;
            float wi[16];
            orc_transform_world_inverse(position + i * 3, rotation + i * 4, scale + i * 3, wi);
            if (world_inverse_out) memcpy(world_inverse_out + i * 16, wi, sizeof wi);
            if (orc_cull_visible(corners, wi, bbox + i * 6)) bits |= (1u << b);
        }
        vis_bits[w] = bits;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Transform::update with a parent (src/transform.cpp:286-309):
 *   _worldMatrix  = _parent->getWorldMatrix() * _localMatrix
 *   _worldInverse = _localInverse * _parent->getWorldInverse()
 * the parent being updated first (recursively).  Nodes may come in any order; parent[i] = -1 for a root.
 * ---------------------------------------------------------------------------------------------- */
static void node_update(int i, const float* pos, const float* rot, const float* scale, const int32_t* parent, unsigned char* done,
                        float* world, float* world_inverse) {
    if (done[i]) return;
    float local[16], local_inverse[16];
    orc_transform_local_matrix(pos + (size_t)i * 3, rot + (size_t)i * 4, scale + (size_t)i * 3, local);
    orc_mat4_inverse(local, local_inverse);
    if (parent[i] >= 0) {
        node_update(parent[i], pos, rot, scale, parent, done, world, world_inverse);
        orc_mat4_mul(world + (size_t)parent[i] * 16, local, world + (size_t)i * 16);
        orc_mat4_mul(local_inverse, world_inverse + (size_t)parent[i] * 16, world_inverse + (size_t)i * 16);
    } else {
        memcpy(world + (size_t)i * 16, local, sizeof local);
        memcpy(world_inverse + (size_t)i * 16, local_inverse, sizeof local_inverse);
    }
    done[i] = 1;
}

void orc_transform_nodes(int num_nodes, const float* position, const float* rotation, const float* scale, const int32_t* parent,
                         float* out_world, float* out_world_inverse) {
    unsigned char done[num_nodes > 0 ? num_nodes : 1];
    memset(done, 0, sizeof done);
    for (int i = 0; i < num_nodes; ++i) node_update(i, position, rotation, scale, parent, done, out_world, out_world_inverse);
}

void orc_cull_batch_trs_parented(const float* corners, const float* position, const float* rotation, const float* scale,
                                 const float* bbox, int64_t n, const int32_t* entity_parent, const float* node_world_inverse,
                                 uint32_t* vis_bits, float* world_inverse_out, int num_threads) {
    const int64_t words = (n + 31) / 32;
    if (num_threads < 1) num_threads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(num_threads)
#endif
    for (int64_t w = 0; w < words; ++w) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int64_t i = w * 32 + b;
            if (i >= n) break;
            float wi[16];
            orc_transform_world_inverse(position + i * 3, rotation + i * 4, scale + i * 3, wi);
            if (entity_parent[i] >= 0) orc_mat4_mul(wi, node_world_inverse + (size_t)entity_parent[i] * 16, wi);
            if (world_inverse_out) memcpy(world_inverse_out + i * 16, wi, sizeof wi);
            if (orc_cull_visible(corners, wi, bbox + i * 6)) bits |= (1u << b);
        }
        vis_bits[w] = bits;
    }
}
