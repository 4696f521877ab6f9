// host_math.hpp -- the few glm operations the host side of the cull needs, in glm's scalar operation order
//  so that the frustum corners handed to the kernel are the ones darmok would compute:
// Frustum::Frustum(mtx) (src/shape.cpp:1141-1163) = inverse(mtx) * ndc corner, divided by w.
#pragma once

namespace dmk {

// glm::inverse(mat4): cofactor expansion of glm/detail/func_matrix.inl, column-major m[c * 4 + r]
inline void mat4_inverse_glm(const float* m, float* out) {
    auto M = [&](int c, int r) { return m[c * 4 + r]; };
    const float s00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), s02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
    const float s03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3), s04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    const float s06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), s07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    const float s08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), s10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
    const float s11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2), s12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
    const float s14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), s15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    const float s16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), s18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
    const float s19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2), s20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    const float s22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), s23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
    const float fac[6][4] = {{s00, s00, s02, s03}, {s04, s04, s06, s07}, {s08, s08, s10, s11},
                             {s12, s12, s14, s15}, {s16, s16, s18, s19}, {s20, s20, s22, s23}};
    const float vec[4][4] = {{M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, {M(1, 1), M(0, 1), M(0, 1), M(0, 1)},
                             {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, {M(1, 3), M(0, 3), M(0, 3), M(0, 3)}};
    float inv[16];
    for (int i = 0; i < 4; ++i) {
        const float sa = (i & 1) ? -1.f : 1.f, sb = -sa;
        inv[0 + i] = ((vec[1][i] * fac[0][i] - vec[2][i] * fac[1][i]) + vec[3][i] * fac[2][i]) * sa;
        inv[4 + i] = ((vec[0][i] * fac[0][i] - vec[2][i] * fac[3][i]) + vec[3][i] * fac[4][i]) * sb;
        inv[8 + i] = ((vec[0][i] * fac[1][i] - vec[1][i] * fac[3][i]) + vec[3][i] * fac[5][i]) * sa;
        inv[12 + i] = ((vec[0][i] * fac[2][i] - vec[1][i] * fac[4][i]) + vec[2][i] * fac[5][i]) * sb;
    }
    const float d0 = M(0, 0) * inv[0], d1 = M(0, 1) * inv[4], d2 = M(0, 2) * inv[8], d3 = M(0, 3) * inv[12];
    const float one_over_det = 1.f / ((d0 + d1) + (d2 + d3));
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * one_over_det;
}

// Frustum::Frustum(mtx): corners in Frustum::CornerType order
inline void frustum_corners_glm(const float* view_proj, float near_depth, float* corners) {
    float inv[16];
    mat4_inverse_glm(view_proj, inv);
    const float nd = near_depth;
    const float ndc[8][3] = {{-1, -1, nd}, {1, -1, nd}, {-1, 1, nd}, {1, 1, nd}, {-1, -1, 1}, {1, -1, 1}, {-1, 1, 1}, {1, 1, 1}};
    for (int k = 0; k < 8; ++k) {
        float t[4];
        for (int r = 0; r < 4; ++r)  // mat4 * vec4: (m0*x + m1*y) + (m2*z + m3*w)
            t[r] = (inv[0 + r] * ndc[k][0] + inv[4 + r] * ndc[k][1]) + (inv[8 + r] * ndc[k][2] + inv[12 + r] * 1.0f);
        corners[k * 3 + 0] = t[0] / t[3];
        corners[k * 3 + 1] = t[1] / t[3];
        corners[k * 3 + 2] = t[2] / t[3];
    }
}

}  // namespace dmk
