// ORACLE — TEST INFRASTRUCTURE ONLY. Never imported, linked or executed by the product path
// (rasterize_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it, and only as the checker / CPU baseline.
//
// CPU restatement (f64, no FMA contraction, no fast-math) of the geometry layer of
// aslpavel/rasterize v0.6.7.  Every function cites the reference file:line it follows
// (paths relative to /root/reference).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <limits>
#include <optional>
#include <vector>
#include <array>
#include <string>
#include <algorithm>

namespace orc {

using Scalar = double;                                            // src/geometry.rs:12
constexpr Scalar EPSILON = std::numeric_limits<double>::epsilon(); // src/geometry.rs:14
constexpr Scalar EPSILON_SQRT = 1.4901161193847656e-8;             // src/geometry.rs:16
constexpr Scalar PI = 3.14159265358979323846264338327950288;       // src/geometry.rs:18

// ---- Rust numeric semantics helpers -------------------------------------------------------
// `f64::max`/`f64::min`: if one operand is NaN the other is returned.
inline Scalar rmax(Scalar a, Scalar b) { return std::isnan(a) ? b : (std::isnan(b) ? a : (a > b ? a : b)); }
inline Scalar rmin(Scalar a, Scalar b) { return std::isnan(a) ? b : (std::isnan(b) ? a : (a < b ? a : b)); }
// `as usize` / `as i32` / `as u8`: saturating, NaN -> 0.
inline size_t as_usize(Scalar v) {
    if (std::isnan(v) || v <= 0.0) return 0;
    if (v >= 18446744073709551615.0) return SIZE_MAX;
    return (size_t)v;
}
inline int32_t as_i32(Scalar v) {
    if (std::isnan(v)) return 0;
    if (v <= -2147483648.0) return INT32_MIN;
    if (v >= 2147483647.0) return INT32_MAX;
    return (int32_t)v;
}
inline uint8_t f32_as_u8(float v) {
    if (std::isnan(v) || v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
// `f64::rem_euclid`
inline Scalar rem_euclid(Scalar x, Scalar rhs) {
    Scalar r = std::fmod(x, rhs);
    return r < 0.0 ? r + std::fabs(rhs) : r;
}
// `f64::powi` with a run-time exponent lowers to compiler-rt `__powidf2` (square-and-multiply,
// reciprocal at the end for negative exponents) — used by the scalar parser, src/svg.rs:233-234.
inline Scalar powi(Scalar a, int b) {
    const bool recip = b < 0;
    Scalar r = 1.0;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

// src/utils.rs:6-18
template <class T> inline T clamp(T val, T min, T max) { return val < min ? min : (val > max ? max : val); }

// ---- Point: src/geometry.rs:103-294 ------------------------------------------------------
struct Point {
    Scalar x = 0.0, y = 0.0;
    Point() = default;
    Point(Scalar x_, Scalar y_) : x(x_), y(y_) {}
    Scalar length() const { return std::hypot(x, y); }                     // :144-147
    Scalar dist(Point o) const { return Point(x - o.x, y - o.y).length(); } // :150-152
    Scalar dot(Point o) const { return x * o.x + y * o.y; }                 // :155-159
    Scalar cross(Point o) const { return x * o.y - y * o.x; }               // :162-166
    Point normal() const { return Point(y, -x); }                           // :169-172
    std::optional<Point> normalize() const {                                // :175-183
        Scalar len = length();
        if (len < EPSILON) return std::nullopt;
        return Point(x / len, y / len);
    }
    std::optional<Scalar> cos_between(Point o) const {                      // :196-203
        Scalar lengths = length() * o.length();
        if (lengths < EPSILON) return std::nullopt;
        return dot(o) / lengths;
    }
    std::optional<Scalar> angle_between(Point o) const {                    // :186-193
        auto c = cos_between(o);
        if (!c) return std::nullopt;
        Scalar angle = std::acos(clamp(*c, -1.0, 1.0));
        return cross(o) < 0.0 ? -angle : angle;
    }
    bool is_close_to(Point o) const {                                       // :212-216
        return std::fabs(x - o.x) < EPSILON && std::fabs(y - o.y) < EPSILON;
    }
};
inline Point operator+(Point a, Point b) { return Point(a.x + b.x, a.y + b.y); }
inline Point operator-(Point a, Point b) { return Point(a.x - b.x, a.y - b.y); }
inline Point operator*(Scalar s, Point p) { return Point(s * p.x, s * p.y); }
inline Point operator/(Point p, Scalar s) { return Point(p.x / s, p.y / s); }
inline bool operator==(Point a, Point b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(Point a, Point b) { return !(a == b); }

enum class Align { Min, Mid, Max };  // src/geometry.rs:298-305

struct BBox;

// ---- Transform: src/geometry.rs:317-539 --------------------------------------------------
struct Transform {
    Scalar m[6] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0};  // [m00, m01, m02, m10, m11, m12]
    Transform() = default;
    Transform(Scalar m00, Scalar m01, Scalar m02, Scalar m10, Scalar m11, Scalar m12) {
        m[0] = m00; m[1] = m01; m[2] = m02; m[3] = m10; m[4] = m11; m[5] = m12;
    }
    static Transform identity() { return Transform(); }
    Point apply(Point p) const {  // :363-367 — two rounded products and two rounded sums per coordinate
        return Point(p.x * m[0] + p.y * m[1] + m[2], p.x * m[3] + p.y * m[4] + m[5]);
    }
    std::optional<Transform> invert() const {  // :370-384
        Scalar det = m[0] * m[4] - m[3] * m[1];
        if (std::fabs(det) <= EPSILON) return std::nullopt;
        Scalar o00 = m[4] / det;
        Scalar o01 = -m[1] / det;
        Scalar o10 = -m[3] / det;
        Scalar o11 = m[0] / det;
        Scalar o02 = -o00 * m[2] - o01 * m[5];
        Scalar o12 = -o10 * m[2] - o11 * m[5];
        return Transform(o00, o01, o02, o10, o11, o12);
    }
    static Transform new_translate(Scalar tx, Scalar ty) { return Transform(1.0, 0.0, tx, 0.0, 1.0, ty); }  // :391
    static Transform new_scale(Scalar sx, Scalar sy) { return Transform(sx, 0.0, 0.0, 0.0, sy, 0.0); }      // :400
    static Transform new_rotate(Scalar a) {                                                               // :409-412
        Scalar s = std::sin(a), c = std::cos(a);
        return Transform(c, -s, 0.0, s, c, 0.0);
    }
    static Transform new_skew(Scalar ax, Scalar ay) { return Transform(1.0, std::tan(ax), 0.0, std::tan(ay), 1.0, 0.0); }  // :427
    Transform mul(const Transform& o) const {  // :519-539
        const Scalar* s = m;
        return Transform(s[0] * o.m[0] + s[1] * o.m[3], s[0] * o.m[1] + s[1] * o.m[4],
                         s[0] * o.m[2] + s[1] * o.m[5] + s[2], s[3] * o.m[0] + s[4] * o.m[3],
                         s[3] * o.m[1] + s[4] * o.m[4], s[3] * o.m[2] + s[4] * o.m[5] + s[5]);
    }
    Transform pre_concat(const Transform& o) const { return mul(o); }                                    // :432
    Transform pre_translate(Scalar tx, Scalar ty) const { return pre_concat(new_translate(tx, ty)); }    // :387
    Transform pre_scale(Scalar sx, Scalar sy) const { return pre_concat(new_scale(sx, sy)); }            // :396
    Transform pre_rotate(Scalar a) const { return pre_concat(new_rotate(a)); }                           // :405
    static Transform fit_bbox(const BBox& src, const BBox& dst, Align align);
};
inline Transform operator*(const Transform& a, const Transform& b) { return a.mul(b); }

// ---- BBox: src/geometry.rs:543-688 -------------------------------------------------------
struct BBox {
    Point min, max;
    BBox() = default;
    BBox(Point p0, Point p1) {  // :552-561
        Scalar x0 = p0.x, x1 = p1.x, y0 = p0.y, y1 = p1.y;
        if (!(x0 <= x1)) std::swap(x0, x1);
        if (!(y0 <= y1)) std::swap(y0, y1);
        min = Point(x0, y0);
        max = Point(x1, y1);
    }
    Scalar x() const { return min.x; }
    Scalar y() const { return min.y; }
    Scalar width() const { return max.x - min.x; }
    Scalar height() const { return max.y - min.y; }
    bool contains(Point p) const {  // :605-608
        return min.x <= p.x && p.x <= max.x && min.y <= p.y && p.y <= max.y;
    }
    BBox extend(Point p) const {  // :611-633
        Scalar x0 = min.x, y0 = min.y, x1 = max.x, y1 = max.y;
        if (p.x < x0) x0 = p.x; else if (p.x > x1) x1 = p.x;
        if (p.y < y0) y0 = p.y; else if (p.y > y1) y1 = p.y;
        BBox r;
        r.min = Point(x0, y0);
        r.max = Point(x1, y1);
        return r;
    }
    BBox union_(const BBox& o) const { return extend(o.min).extend(o.max); }  // :636-638
    BBox union_opt(const std::optional<BBox>& o) const { return o ? union_(*o) : *this; }  // :640-645
    std::optional<BBox> intersect(const BBox& o) const {  // :648-657, 677-688
        if (min.x > o.max.x || o.min.x > max.x) return std::nullopt;
        if (min.y > o.max.y || o.min.y > max.y) return std::nullopt;
        return BBox(Point(rmax(min.x, o.min.x), rmax(min.y, o.min.y)),
                    Point(rmin(max.x, o.max.x), rmin(max.y, o.max.y)));
    }
    Transform unit_transform() const {  // :662-664
        return Transform::new_translate(x(), y()).pre_scale(width(), height());
    }
};

inline Transform Transform::fit_bbox(const BBox& src, const BBox& dst, Align align) {  // :470-487
    Scalar scale = rmin(dst.height() / src.height(), dst.width() / src.width());
    Transform base = new_translate(dst.x(), dst.y()).pre_scale(scale, scale).pre_translate(-src.x(), -src.y());
    Transform al;
    switch (align) {
        case Align::Min: al = identity(); break;
        case Align::Mid:
            al = new_translate((dst.width() - src.width() * scale) / 2.0, (dst.height() - src.height() * scale) / 2.0);
            break;
        case Align::Max:
            al = new_translate(dst.width() - src.width() * scale, dst.height() - src.height() * scale);
            break;
    }
    return al * base;
}

struct Size { size_t width = 0, height = 0; };  // src/rasterize.rs:38-41

// src/geometry.rs:490-516
inline std::pair<Size, Transform> fit_size(BBox src, Size size, Align align) {
    src = BBox(Point(std::floor(src.min.x), std::floor(src.min.y)), Point(std::ceil(src.max.x), std::ceil(src.max.y)));
    Scalar height, width;
    if (size.height == 0 && size.width == 0) {
        height = src.height(); width = src.width();
    } else if (size.width == 0) {
        height = (Scalar)size.height;
        width = std::ceil(src.width() * height / src.height());
    } else if (size.height == 0) {
        width = (Scalar)size.width;
        height = std::ceil(src.height() * width / src.width());
    } else {
        height = (Scalar)size.height; width = (Scalar)size.width;
    }
    BBox dst(Point(0.0, 0.0), Point(width, height));
    Size out;
    out.height = as_usize(height);
    out.width = as_usize(width);
    return {out, Transform::fit_bbox(src, dst, align)};
}

// src/utils.rs:205-231 — returns roots in push order; n in {0,1,2}
struct Roots2 { int n = 0; Scalar v[2] = {0, 0}; void push(Scalar t) { v[n++] = t; } };
inline Roots2 quadratic_solve(Scalar a, Scalar b, Scalar c) {
    Roots2 result;
    if (std::fabs(a) < EPSILON) {
        if (std::fabs(b) > EPSILON) result.push(-c / b);
        return result;
    }
    Scalar disc = b * b - 4.0 * a * c;
    if (std::fabs(disc) < EPSILON) {
        result.push(-b / (2.0 * a));
    } else if (disc > 0.0) {
        Scalar sq = std::sqrt(disc);
        if (b >= 0.0) {
            Scalar mul = -b - sq;
            result.push(mul / (2.0 * a));
            result.push(2.0 * c / mul);
        } else {
            Scalar mul = -b + sq;
            result.push(2.0 * c / mul);
            result.push(mul / (2.0 * a));
        }
    }
    return result;
}

}  // namespace orc
