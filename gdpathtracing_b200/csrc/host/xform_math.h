// xform_math.h -- the few Godot math types the host side of the path needs,
// without godot-cpp.  When the adapter is built inside the real GDExtension the
// godot::Transform3D / godot::Projection values are converted to these PODs at
// the call site; standalone (tests, bench) they are the scene description.
//
// Arithmetic follows godot-cpp @56571dc (branch 4.3), real_t = float:
//   Basis::invert            godot-cpp/src/variant/basis.cpp:57-72
//   Transform3D::affine_invert  godot-cpp/src/variant/transform3d.cpp:37-46
//   Projection(Transform3D)  godot-cpp/src/variant/projection.cpp:916-936
//   Projection::operator Transform3D  godot-cpp/src/variant/projection.cpp:886-907
//   Projection::operator*    godot-cpp/src/variant/projection.cpp:709-723
//   Projection::invert       godot-cpp/src/variant/projection.cpp:601-698
//   Projection::set_perspective godot-cpp/src/variant/projection.cpp:254-278
//   Math::is_equal_approx    godot-cpp/include/godot_cpp/core/math.hpp:624-635
#ifndef GDPT_XFORM_MATH_H
#define GDPT_XFORM_MATH_H

#include <cmath>

namespace gdpt {

struct Vec3 {
    float x = 0.f, y = 0.f, z = 0.f;
    Vec3() {}
    Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float dot(const Vec3 &o) const { return x * o.x + y * o.y + z * o.z; }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

// Row-major 3x3, like godot::Basis (rows[] holds the transposed axes).
struct Basis3 {
    Vec3 rows[3] = { Vec3(1, 0, 0), Vec3(0, 1, 0), Vec3(0, 0, 1) };
    Vec3 column(int i) const { return Vec3(rows[0][i], rows[1][i], rows[2][i]); }
    Vec3 xform(const Vec3 &v) const { return Vec3(rows[0].dot(v), rows[1].dot(v), rows[2].dot(v)); }
    float minor2(int r1, int c1, int r2, int c2) const { return rows[r1][c1] * rows[r2][c2] - rows[r1][c2] * rows[r2][c1]; }
    Basis3 inverse() const
    {
        const float c0 = minor2(1, 1, 2, 2), c1 = minor2(1, 2, 2, 0), c2 = minor2(1, 0, 2, 1);
        const float det = rows[0][0] * c0 + rows[0][1] * c1 + rows[0][2] * c2;
        const float s = 1.0f / det;
        Basis3 r;
        r.rows[0] = Vec3(c0 * s, minor2(0, 2, 2, 1) * s, minor2(0, 1, 1, 2) * s);
        r.rows[1] = Vec3(c1 * s, minor2(0, 0, 2, 2) * s, minor2(0, 2, 1, 0) * s);
        r.rows[2] = Vec3(c2 * s, minor2(0, 1, 2, 0) * s, minor2(0, 0, 1, 1) * s);
        return r;
    }
};

inline bool approx_equal(float a, float b)
{
    if (a == b) return true;
    float tol = 0.00001f * std::fabs(a);
    if (tol < 0.00001f) tol = 0.00001f;
    return std::fabs(a - b) < tol;
}

struct Xform3 { // godot::Transform3D
    Basis3 basis;
    Vec3 origin;
    Xform3 affine_inverse() const
    {
        Xform3 r;
        r.basis = basis.inverse();
        r.origin = r.basis.xform(Vec3(-origin.x, -origin.y, -origin.z));
        return r;
    }
    bool is_equal_approx(const Xform3 &o) const
    {
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++)
                if (!approx_equal(basis.rows[r][c], o.basis.rows[r][c])) return false;
        return approx_equal(origin.x, o.origin.x) && approx_equal(origin.y, o.origin.y) && approx_equal(origin.z, o.origin.z);
    }
    // Column-major 4x4 exactly as Utils::transform_to_float (src/utils.h:15-37).
    void to_float16(float *m) const
    {
        for (int c = 0; c < 3; c++) {
            const Vec3 col = basis.column(c);
            m[c * 4 + 0] = col.x; m[c * 4 + 1] = col.y; m[c * 4 + 2] = col.z; m[c * 4 + 3] = 0.0f;
        }
        m[12] = origin.x; m[13] = origin.y; m[14] = origin.z; m[15] = 1.0f;
    }
    // 12 floats: basis rows then origin (the layout the C API uses).
    static Xform3 from_rows12(const float *t)
    {
        Xform3 x;
        for (int r = 0; r < 3; r++) x.basis.rows[r] = Vec3(t[r * 3 + 0], t[r * 3 + 1], t[r * 3 + 2]);
        x.origin = Vec3(t[9], t[10], t[11]);
        return x;
    }
};

struct Mat4 { // godot::Projection, m[col][row]
    float m[4][4];
    Mat4()
    {
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++) m[c][r] = (c == r) ? 1.0f : 0.0f;
    }
    explicit Mat4(const Xform3 &t)
    {
        for (int c = 0; c < 3; c++) {
            m[c][0] = t.basis.rows[0][c]; m[c][1] = t.basis.rows[1][c]; m[c][2] = t.basis.rows[2][c]; m[c][3] = 0.0f;
        }
        m[3][0] = t.origin.x; m[3][1] = t.origin.y; m[3][2] = t.origin.z; m[3][3] = 1.0f;
    }
    static Mat4 perspective(float fovy_degrees, float aspect, float z_near, float z_far)
    {
        Mat4 p;
        const float radians = (fovy_degrees / 2.0f) * 3.14159265358979323846f / 180.f;
        const float dz = z_far - z_near;
        const float sine = std::sin(radians);
        if (dz == 0 || sine == 0 || aspect == 0) return p;
        const float cot = std::cos(radians) / sine;
        p.m[0][0] = cot / aspect;
        p.m[1][1] = cot;
        p.m[2][2] = -(z_far + z_near) / dz;
        p.m[2][3] = -1;
        p.m[3][2] = -2 * z_near * z_far / dz;
        p.m[3][3] = 0;
        return p;
    }
    Mat4 operator*(const Mat4 &o) const
    {
        Mat4 r;
        for (int j = 0; j < 4; j++)
            for (int i = 0; i < 4; i++) {
                float ab = 0;
                for (int k = 0; k < 4; k++) ab += m[k][i] * o.m[j][k];
                r.m[j][i] = ab;
            }
        return r;
    }
    // Full-pivot Gauss-Jordan, same pivoting and operation order as Projection::invert.
    Mat4 inverse() const
    {
        Mat4 a = *this;
        int pi[4], pj[4];
        float det = 1.0f;
        for (int k = 0; k < 4; k++) {
            float pv = a.m[k][k];
            pi[k] = k; pj[k] = k;
            for (int i = k; i < 4; i++)
                for (int j = k; j < 4; j++)
                    if (std::fabs(a.m[i][j]) > std::fabs(pv)) { pi[k] = i; pj[k] = j; pv = a.m[i][j]; }
            det *= pv;
            if (std::fabs(det) < 0.00001f) return a; // singular: left half-reduced, like upstream
            int i = pi[k];
            if (i != k)
                for (int j = 0; j < 4; j++) { float h = -a.m[k][j]; a.m[k][j] = a.m[i][j]; a.m[i][j] = h; }
            int j = pj[k];
            if (j != k)
                for (i = 0; i < 4; i++) { float h = -a.m[i][k]; a.m[i][k] = a.m[i][j]; a.m[i][j] = h; }
            for (i = 0; i < 4; i++)
                if (i != k) a.m[i][k] /= (-pv);
            for (i = 0; i < 4; i++) {
                const float h = a.m[i][k];
                for (j = 0; j < 4; j++)
                    if (i != k && j != k) a.m[i][j] += h * a.m[k][j];
            }
            for (j = 0; j < 4; j++)
                if (j != k) a.m[k][j] /= pv;
            a.m[k][k] = 1.0f / pv;
        }
        for (int k = 2; k >= 0; k--) {
            int i = pj[k];
            if (i != k)
                for (int j = 0; j < 4; j++) { float h = a.m[k][j]; a.m[k][j] = -a.m[i][j]; a.m[i][j] = h; }
            int j = pi[k];
            if (j != k)
                for (i = 0; i < 4; i++) { float h = a.m[i][k]; a.m[i][k] = -a.m[i][j]; a.m[i][j] = h; }
        }
        return a;
    }
    // Projection -> Transform3D -> Projection (godot-cpp/src/variant/projection.cpp:886-907, 916-936): what an
    // assignment of a Projection to a Transform3D keeps -- the projective row becomes (0, 0, 0, 1).
    Mat4 affine_part() const
    {
        Mat4 r = *this;
        r.m[0][3] = 0.0f; r.m[1][3] = 0.0f; r.m[2][3] = 0.0f; r.m[3][3] = 1.0f;
        return r;
    }
    // Utils::projection_to_float (src/utils.h:39-49)
    void to_float16(float *out) const
    {
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++) out[c * 4 + r] = m[c][r];
    }
};

} // namespace gdpt
#endif
