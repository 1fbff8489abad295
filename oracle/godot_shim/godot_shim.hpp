// godot_shim.hpp -- TEST INFRASTRUCTURE ONLY (oracle/).
//
// The handful of godot-cpp types that /root/reference/src/bvh/{bvh.h,bvh.cpp,
// vec.h} and src/utils.h touch, so that those reference sources can be compiled
// VERBATIM, in place, into oracle/_ref/ (see oracle/Makefile).  Nothing under
// gdpathtracing_b200/ includes or links this.
//
// godot-cpp itself cannot be built here (it needs generated bindings and a Godot
// engine), so the few arithmetic routines the reference host code calls on these
// types are restated from the pinned submodule (godot-cpp @56571dc, branch 4.3):
//   Basis::invert          godot-cpp/src/variant/basis.cpp:57-72 (cofac :35-36)
//   Basis::xform           godot-cpp/include/godot_cpp/variant/basis.hpp:297-302
//   Transform3D::affine_invert  godot-cpp/src/variant/transform3d.cpp:37-46
#pragma once
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <limits>
#include <string>
#include <vector>

namespace godot {

typedef float real_t; // precision=single (godot-cpp/tools/godotcpp.py:254-256)

class String {
    std::string s_;
public:
    String() {}
    String(const char *p) : s_(p ? p : "") {}
    String(const std::string &p) : s_(p) {}
    String operator+(const String &o) const { return String(s_ + o.s_); }
    friend String operator+(const char *a, const String &b) { return String(std::string(a) + b.s_); }
    const std::string &str() const { return s_; }
};

struct Vector2 { real_t x = 0, y = 0; Vector2() {} Vector2(real_t px, real_t py) : x(px), y(py) {} };

struct Vector3 {
    real_t x = 0, y = 0, z = 0;
    Vector3() {}
    Vector3(real_t px, real_t py, real_t pz) : x(px), y(py), z(pz) {}
    real_t dot(const Vector3 &v) const { return x * v.x + y * v.y + z * v.z; }
    Vector3 operator-() const { return Vector3(-x, -y, -z); }
    real_t &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const real_t &operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

struct Vector4 { real_t x = 0, y = 0, z = 0, w = 0; };

struct Basis {
    Vector3 rows[3] = { Vector3(1, 0, 0), Vector3(0, 1, 0), Vector3(0, 0, 1) };
    Vector3 get_column(int i) const { return Vector3(rows[0][i], rows[1][i], rows[2][i]); }
    Vector3 xform(const Vector3 &v) const { return Vector3(rows[0].dot(v), rows[1].dot(v), rows[2].dot(v)); }
    void set(real_t xx, real_t xy, real_t xz, real_t yx, real_t yy, real_t yz, real_t zx, real_t zy, real_t zz) {
        rows[0] = Vector3(xx, xy, xz); rows[1] = Vector3(yx, yy, yz); rows[2] = Vector3(zx, zy, zz);
    }
    void invert() {
#define GDSHIM_COFAC(r1, c1, r2, c2) (rows[r1][c1] * rows[r2][c2] - rows[r1][c2] * rows[r2][c1])
        real_t co[3] = { GDSHIM_COFAC(1, 1, 2, 2), GDSHIM_COFAC(1, 2, 2, 0), GDSHIM_COFAC(1, 0, 2, 1) };
        real_t det = rows[0][0] * co[0] + rows[0][1] * co[1] + rows[0][2] * co[2];
        real_t s = 1.0f / det;
        set(co[0] * s, GDSHIM_COFAC(0, 2, 2, 1) * s, GDSHIM_COFAC(0, 1, 1, 2) * s,
            co[1] * s, GDSHIM_COFAC(0, 0, 2, 2) * s, GDSHIM_COFAC(0, 2, 1, 0) * s,
            co[2] * s, GDSHIM_COFAC(0, 1, 2, 0) * s, GDSHIM_COFAC(0, 0, 1, 1) * s);
#undef GDSHIM_COFAC
    }
};

struct Transform3D {
    Basis basis;
    Vector3 origin;
    void affine_invert() { basis.invert(); origin = basis.xform(-origin); }
    Transform3D affine_inverse() const { Transform3D r = *this; r.affine_invert(); return r; }
};

struct Projection { Vector4 columns[4]; };

template <typename T> struct PackedArrayShim {
    const T *ptr = nullptr;
    int64_t n = 0;
    int64_t size() const { return n; }
    const T &operator[](int64_t i) const { return ptr[i]; }
};
typedef PackedArrayShim<int32_t> PackedInt32Array;
typedef PackedArrayShim<Vector3> PackedVector3Array;
typedef PackedArrayShim<Vector2> PackedVector2Array;

// One mesh surface's arrays, indexed by Mesh::ARRAY_* like a godot::Array.
struct Variant {
    PackedInt32Array i32; PackedVector3Array v3; PackedVector2Array v2;
    operator PackedInt32Array() const { return i32; }
    operator PackedVector3Array() const { return v3; }
    operator PackedVector2Array() const { return v2; }
};
struct Array {
    Variant slots[13];
    const Variant &operator[](int i) const { return slots[i]; }
};

struct Mesh { enum ArrayType { ARRAY_VERTEX = 0, ARRAY_NORMAL = 1, ARRAY_TANGENT = 2, ARRAY_COLOR = 3, ARRAY_TEX_UV = 4, ARRAY_INDEX = 12 }; };

struct ArrayMesh : Mesh {
    std::vector<Array> surfaces;
    int get_surface_count() const { return (int)surfaces.size(); }
    Array surface_get_arrays(int i) const { return surfaces[i]; }
};

template <typename T> struct Ref {
    T *p = nullptr;
    Ref() {}
    explicit Ref(T *q) : p(q) {}
    T *operator->() const { return p; }
};

struct UtilityFunctions {
    static void print(const String &s) { std::fputs(s.str().c_str(), stderr); std::fputc('\n', stderr); }
    template <typename T> static void print(const T &) {}
};

} // namespace godot
