// Device-side stroking primitives: `Path::stroke` of the reference (src/path.rs:374-415, 692-732) and the curve
// helpers it stands on (src/curve.rs:195-224, 377-388, 647-667 `ends`; :270-273, 505-513, 810-818 `bbox`;
// :978-1078 `line_join` / `line_cap`; :1283-1433 offsets; src/ellipse.rs:40-96, 167-214 for the round joins).
//
// Everything is f64 in the reference's expression order; the translation unit that includes this header is compiled
// with --fmad=false, so no product is fused into a following sum.  `Point::length` is `f64::hypot`, i.e. the C
// library's: `hypot_libm` restates glibc (>= 2.35) `__hypot` (sysdeps/ieee754/dbl-64/e_hypot.c, non-FMA kernel), which
// is what the reference calls on x86-64 Linux.  sqrt and division are IEEE on the device.  sin / cos / tan / acos of the
// round joins are CUDA's (<= 2 ulp), the C library's are < 1 ulp: arcs agree to a few ulp, everything else bit for bit.
#pragma once
#include <cmath>
#include <cstdint>
#include "stroke_units.hpp"

#ifndef SD_FN
#define SD_FN __device__
#endif

namespace rgpu {
namespace sk {

constexpr double kEps = 2.220446049250313e-16;               // src/geometry.rs:14
constexpr double kPi = 3.14159265358979323846264338327950288;  // src/geometry.rs:18

enum { kJoinMiter = 0, kJoinBevel = 1, kJoinRound = 2 };  // `LineJoin`, src/path.rs:78-87
enum { kCapButt = 0, kCapSquare = 1, kCapRound = 2 };     // `LineCap`, src/path.rs:103-110

struct Style {  // `StrokeStyle`, src/path.rs:121-136
    double width;
    double miter_limit;
    int join, cap;
};

struct P2 {
    double x, y;
};
SD_FN inline P2 mk(double x, double y) {
    P2 r;
    r.x = x;
    r.y = y;
    return r;
}
SD_FN inline P2 operator+(P2 a, P2 b) { return mk(a.x + b.x, a.y + b.y); }
SD_FN inline P2 operator-(P2 a, P2 b) { return mk(a.x - b.x, a.y - b.y); }
SD_FN inline P2 operator*(double s, P2 p) { return mk(s * p.x, s * p.y); }
SD_FN inline P2 operator/(P2 p, double s) { return mk(p.x / s, p.y / s); }

SD_FN inline double hypot_kernel(double ax, double ay) {
    double t1, t2;
    double h = sqrt(ax * ax + ay * ay);
    if (h <= 2.0 * ay) {
        const double delta = h - ay;
        t1 = ax * (2.0 * delta - ax);
        t2 = (delta - 2.0 * (ax - ay)) * delta;
    } else {
        const double delta = h - ax;
        t1 = 2.0 * delta * (ax - 2.0 * ay);
        t2 = (4.0 * delta - ay) * ay + delta * delta;
    }
    h -= (t1 + t2) / (2.0 * h);
    return h;
}
SD_FN inline double hypot_libm(double x, double y) {
    const double dbl_max = 1.7976931348623157e308;
    if (!(fabs(x) <= dbl_max) || !(fabs(y) <= dbl_max)) return (fabs(x) > dbl_max || fabs(y) > dbl_max) ? INFINITY : x + y;
    x = fabs(x);
    y = fabs(y);
    const double ax = x < y ? y : x, ay = x < y ? x : y;
    const double scale = 0x1p-600, large = 0x1p+511, tiny = 0x1p-459, eps = 0x1p-54;
    if (ax > large) {
        if (ay <= ax * eps) return ax + ay;
        return hypot_kernel(ax * scale, ay * scale) / scale;
    }
    if (ay < tiny) {
        if (ax >= ay / eps) return ax + ay;
        return hypot_kernel(ax / scale, ay / scale) * scale;
    }
    if (ax >= ay / eps) return ax + ay;
    return hypot_kernel(ax, ay);
}

// ---- Point, src/geometry.rs:144-216 ----
SD_FN inline double length(P2 p) { return hypot_libm(p.x, p.y); }
SD_FN inline double dot(P2 a, P2 b) { return a.x * b.x + a.y * b.y; }
SD_FN inline double cross(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }
SD_FN inline P2 normal(P2 p) { return mk(p.y, -p.x); }
SD_FN inline bool normalize(P2 p, P2& out) {
    const double len = length(p);
    if (len < kEps) return false;
    out = mk(p.x / len, p.y / len);
    return true;
}
SD_FN inline bool cos_between(P2 a, P2 b, double& c) {
    const double lengths = length(a) * length(b);
    if (lengths < kEps) return false;
    c = dot(a, b) / lengths;
    return true;
}
SD_FN inline bool angle_between(P2 a, P2 b, double& angle) {
    double c;
    if (!cos_between(a, b, c)) return false;
    c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);  // src/utils.rs:6-18
    const double ang = acos(c);
    angle = cross(a, b) < 0.0 ? -ang : ang;
    return true;
}
SD_FN inline bool close_to(P2 a, P2 b) { return fabs(a.x - b.x) < kEps && fabs(a.y - b.y) < kEps; }

// ---- Line, src/curve.rs:163-253 ----
struct Ln {
    P2 a, b;
};
SD_FN inline Ln mkln(P2 a, P2 b) {
    Ln l;
    l.a = a;
    l.b = b;
    return l;
}
SD_FN inline P2 ln_at(const Ln& l, double t) { return (1.0 - t) * l.a + t * l.b; }
SD_FN inline P2 ln_dir(const Ln& l) { return l.b - l.a; }
SD_FN inline double ln_length(const Ln& l) { return length(l.a - l.b); }
SD_FN inline bool ln_intersect(const Ln& s, const Ln& o, double& t0, double& t1) {  // :204-214
    const double x1 = s.a.x, y1 = s.a.y, x2 = s.b.x, y2 = s.b.y;
    const double x3 = o.a.x, y3 = o.a.y, x4 = o.b.x, y4 = o.b.y;
    const double det = (x4 - x3) * (y1 - y2) - (x1 - x2) * (y4 - y3);
    if (fabs(det) < kEps) return false;
    t0 = ((y3 - y4) * (x1 - x3) + (x4 - x3) * (y1 - y3)) / det;
    t1 = ((y1 - y2) * (x1 - x3) + (x2 - x1) * (y1 - y3)) / det;
    return true;
}
SD_FN inline bool ln_intersect_point(const Ln& s, const Ln& o) {  // :217-224 (only the Some / None answer is used)
    double t0, t1;
    if (!ln_intersect(s, o, t0, t1)) return false;
    return t0 >= 0.0 && t0 <= 1.0 && t1 >= 0.0 && t1 <= 1.0;
}
SD_FN inline bool ln_offset(const Ln& l, double dist, Ln& out) {  // :1283-1287
    P2 n;
    if (!normalize(normal(l.b - l.a), n)) return false;
    const P2 off = dist * n;
    out = mkln(l.a + off, l.b + off);
    return true;
}

// ---- Segment (2 = Line, 3 = Quad, 4 = Cubic control points) ----
struct Seg {
    P2 p[4];
    int kind;
};
SD_FN inline Seg seg_line(P2 a, P2 b) {
    Seg s;
    s.kind = 2;
    s.p[0] = a;
    s.p[1] = b;
    s.p[2] = b;
    s.p[3] = b;
    return s;
}
SD_FN inline Seg seg_quad(P2 a, P2 b, P2 c) {
    Seg s;
    s.kind = 3;
    s.p[0] = a;
    s.p[1] = b;
    s.p[2] = c;
    s.p[3] = c;
    return s;
}
SD_FN inline Seg seg_cubic(P2 a, P2 b, P2 c, P2 d) {
    Seg s;
    s.kind = 4;
    s.p[0] = a;
    s.p[1] = b;
    s.p[2] = c;
    s.p[3] = d;
    return s;
}
SD_FN inline P2 seg_start(const Seg& s) { return s.p[0]; }
SD_FN inline P2 seg_end(const Seg& s) { return s.kind == 4 ? s.p[3] : (s.kind == 3 ? s.p[2] : s.p[1]); }
SD_FN inline Seg seg_reverse(const Seg& s) {  // src/curve.rs:279-282, 519-522, 829-832
    if (s.kind == 2) return seg_line(s.p[1], s.p[0]);
    if (s.kind == 3) return seg_quad(s.p[2], s.p[1], s.p[0]);
    return seg_cubic(s.p[3], s.p[2], s.p[1], s.p[0]);
}
SD_FN inline P2 seg_at(const Seg& s, double t) {  // src/curve.rs:250-253, 432-441, 712-723
    if (s.kind == 2) return (1.0 - t) * s.p[0] + t * s.p[1];
    const double t1 = t, t_1 = 1.0 - t;
    const double t2 = t1 * t1, t_2 = t_1 * t_1;
    if (s.kind == 3) return t_2 * s.p[0] + 2.0 * t1 * t_1 * s.p[1] + t2 * s.p[2];
    const double t3 = t2 * t1, t_3 = t_2 * t_1;
    return t_3 * s.p[0] + 3.0 * t1 * t_2 * s.p[1] + 3.0 * t2 * t_1 * s.p[2] + t3 * s.p[3];
}
// `Quad::split_at(0.5)`, src/curve.rs:458-468 (what `quad_offset_rec` calls)
SD_FN inline void quad_split_half(const Seg& q, Seg& c0, Seg& c1) {
    const double t = 0.5, t1 = t, t_1 = 1.0 - t;
    const double t2 = t1 * t1, t_2 = t_1 * t_1;
    const P2 mid = t_2 * q.p[0] + 2.0 * t1 * t_1 * q.p[1] + t2 * q.p[2];
    c0 = seg_quad(q.p[0], t_1 * q.p[0] + t * q.p[1], mid);
    c1 = seg_quad(mid, t_1 * q.p[1] + t * q.p[2], q.p[2]);
}
// `Cubic::split`, src/curve.rs:731-747
SD_FN inline void cubic_split(const Seg& c, Seg& c0, Seg& c1) {
    const P2 mid = 0.125 * c.p[0] + 0.375 * c.p[1] + 0.375 * c.p[2] + 0.125 * c.p[3];
    c0 = seg_cubic(c.p[0], 0.5 * c.p[0] + 0.5 * c.p[1], 0.25 * c.p[0] + 0.5 * c.p[1] + 0.25 * c.p[2], mid);
    c1 = seg_cubic(mid, 0.25 * c.p[1] + 0.5 * c.p[2] + 0.25 * c.p[3], 0.5 * c.p[2] + 0.5 * c.p[3], c.p[3]);
}
// tangent lines at the two ends: src/curve.rs:195-197 (Line), :377-388 (Quad), :647-667 (Cubic)
SD_FN inline void seg_ends(const Seg& s, Ln& first, Ln& second) {
    if (s.kind == 2) {
        first = second = mkln(s.p[0], s.p[1]);
        return;
    }
    if (s.kind == 3) {
        const Ln a = mkln(s.p[0], s.p[1]), b = mkln(s.p[1], s.p[2]);
        if (close_to(s.p[0], s.p[1])) {
            first = second = b;
        } else if (close_to(s.p[1], s.p[2])) {
            first = second = a;
        } else {
            first = a;
            second = b;
        }
        return;
    }
    int si = 0, ei = 0;
    for (int i = 0; i < 3; i++)
        if (!close_to(s.p[i], s.p[i + 1])) {
            si = i;
            break;
        }
    for (int i = 3; i >= 1; i--)
        if (!close_to(s.p[i], s.p[i - 1])) {
            ei = i;
            break;
        }
    if (ei == 0) ei = 1;  // four coincident points: the reference panics on ps[end - 1] (index underflow)
    first = mkln(s.p[si], s.p[si + 1]);
    second = mkln(s.p[ei - 1], s.p[ei]);
}

// ---- BBox, src/geometry.rs:552-633 ----
struct Box {
    P2 lo, hi;
};
SD_FN inline Box box_new(P2 p0, P2 p1) {
    double x0 = p0.x, x1 = p1.x, y0 = p0.y, y1 = p1.y;
    if (!(x0 <= x1)) {
        const double t = x0;
        x0 = x1;
        x1 = t;
    }
    if (!(y0 <= y1)) {
        const double t = y0;
        y0 = y1;
        y1 = t;
    }
    Box b;
    b.lo = mk(x0, y0);
    b.hi = mk(x1, y1);
    return b;
}
SD_FN inline bool box_contains(const Box& b, P2 p) { return b.lo.x <= p.x && p.x <= b.hi.x && b.lo.y <= p.y && p.y <= b.hi.y; }
SD_FN inline Box box_extend(const Box& b, P2 p) {
    double x0 = b.lo.x, y0 = b.lo.y, x1 = b.hi.x, y1 = b.hi.y;
    if (p.x < x0) x0 = p.x;
    else if (p.x > x1) x1 = p.x;
    if (p.y < y0) y0 = p.y;
    else if (p.y > y1) y1 = p.y;
    Box r;
    r.lo = mk(x0, y0);
    r.hi = mk(x1, y1);
    return r;
}
// src/utils.rs:205-231 (roots in push order)
SD_FN inline int quadratic_solve(double a, double b, double c, double* r) {
    int n = 0;
    if (fabs(a) < kEps) {
        if (fabs(b) > kEps) r[n++] = -c / b;
        return n;
    }
    const double disc = b * b - 4.0 * a * c;
    if (fabs(disc) < kEps) {
        r[n++] = -b / (2.0 * a);
    } else if (disc > 0.0) {
        const double sq = sqrt(disc);
        if (b >= 0.0) {
            const double mul = -b - sq;
            r[n++] = mul / (2.0 * a);
            r[n++] = 2.0 * c / mul;
        } else {
            const double mul = -b + sq;
            r[n++] = 2.0 * c / mul;
            r[n++] = mul / (2.0 * a);
        }
    }
    return n;
}
// `Curve::bbox(init)`: src/curve.rs:270-273 (Line), :505-513 + extremities :535-555 (Quad), :810-818 + :853-864 (Cubic);
// `BBox::union_opt` src/geometry.rs:636-645
SD_FN inline Box seg_bbox_acc(const Seg& s, bool has_init, const Box& init) {
    Box bb = box_new(seg_start(s), seg_end(s));
    if (has_init) bb = box_extend(box_extend(bb, init.lo), init.hi);
    if (s.kind == 2) return bb;
    if (s.kind == 3) {
        if (box_contains(bb, s.p[1])) return bb;
        const P2 a = s.p[2] - 2.0 * s.p[1] + s.p[0];
        const P2 b = s.p[1] - s.p[0];
        if (fabs(a.x) > kEps) {
            const double t0 = -b.x / a.x;
            if (t0 >= 0.0 && t0 <= 1.0) bb = box_extend(bb, seg_at(s, t0));
        }
        if (fabs(a.y) > kEps) {
            const double t1 = -b.y / a.y;
            if (t1 >= 0.0 && t1 <= 1.0) bb = box_extend(bb, seg_at(s, t1));
        }
        return bb;
    }
    if (box_contains(bb, s.p[1]) && box_contains(bb, s.p[2])) return bb;
    const P2 a = -1.0 * s.p[0] + 3.0 * s.p[1] - 3.0 * s.p[2] + 1.0 * s.p[3];
    const P2 b = 2.0 * s.p[0] - 4.0 * s.p[1] + 2.0 * s.p[2];
    const P2 c = -1.0 * s.p[0] + s.p[1];
    double rx[2], ry[2];
    const int nx = quadratic_solve(a.x, b.x, c.x, rx);
    const int ny = quadratic_solve(a.y, b.y, c.y, ry);
    for (int i = 0; i < nx; i++)
        if (rx[i] >= 0.0 && rx[i] <= 1.0) bb = box_extend(bb, seg_at(s, rx[i]));
    for (int i = 0; i < ny; i++)
        if (ry[i] >= 0.0 && ry[i] <= 1.0) bb = box_extend(bb, seg_at(s, ry[i]));
    return bb;
}
SD_FN inline Box seg_bbox(const Seg& s) { return seg_bbox_acc(s, false, Box()); }

// ---- polyline offset, src/curve.rs:1294-1346 ----
SD_FN inline bool polyline_offset(P2* ps, int len, double dist) {
    bool has_prev = false;
    Ln prev = mkln(mk(0.0, 0.0), mk(0.0, 0.0));
    int index = 0;
    while (true) {
        int repeats = 1;
        for (int i = index; i + 1 < len; i++) {
            if (!close_to(ps[i], ps[i + 1])) break;
            repeats += 1;
        }
        if (index + repeats >= len) break;
        index += repeats;
        Ln next;
        if (!ln_offset(mkln(ps[index - 1], ps[index]), dist, next)) return false;
        P2 point;
        if (!has_prev) {
            point = next.a;
        } else {
            double t0, t1;
            point = ln_intersect(prev, next, t0, t1) ? ln_at(prev, t0) : next.a;
        }
        for (int i = index - repeats; i < index; i++) ps[i] = point;
        prev = next;
        has_prev = true;
    }
    if (!has_prev) return false;
    for (int i = index; i < len; i++) ps[i] = prev.b;
    return true;
}

// ---- elliptic arc of a round join: src/ellipse.rs:40-96 (`new_param`), :167-214 (`to_cubics`) ----
struct Rot {  // `Transform::new_rotate`, src/geometry.rs:409-412, applied as in :363-367
    double c, s;
};
SD_FN inline Rot rot_new(double a) {
    Rot r;
    r.s = sin(a);
    r.c = cos(a);
    return r;
}
SD_FN inline P2 rot_apply(const Rot& r, P2 p) { return mk(p.x * r.c + p.y * -r.s + 0.0, p.x * r.s + p.y * r.c + 0.0); }
SD_FN inline double rem_euclid(double x, double rhs) {
    const double r = fmod(x, rhs);
    return r < 0.0 ? r + fabs(rhs) : r;
}

// The arc from `src` to `dst` with radii (rx, ry), as cubics into `sink`; false when the parametrisation fails
// (the caller then falls back to a line).  NaN angles (e.g. an arc whose end points coincide: 0 / 0 in the centre) make the
// reference's iterator spin for ever; here, as in the host builders (api.py / .hpp `arc_to`), they count as a failed
// parametrisation.
template <class Sink>
SD_FN inline bool arc_cubics(P2 src, P2 dst, double rx, double ry, double x_axis_rot, bool large_flag, bool sweep_flag, Sink& sink) {
    rx = fabs(rx);
    ry = fabs(ry);
    const double phi = x_axis_rot * kPi / 180.0;
    const P2 p1 = rot_apply(rot_new(-phi), 0.5 * (src - dst));
    const double x1 = p1.x, y1 = p1.y;
    const double ax = x1 / rx, ay = y1 / ry;
    const double s = ax * ax + ay * ay;
    if (s > 1.0) {
        const double sq = sqrt(s);
        rx = rx * sq;
        ry = ry * sq;
    }
    const double rxry = rx * ry, rxy1 = rx * y1, ryx1 = ry * x1;
    double sq = sqrt(fmax(rxry * rxry / (rxy1 * rxy1 + ryx1 * ryx1) - 1.0, 0.0));
    sq = (large_flag == sweep_flag) ? -sq : sq;
    P2 center = sq * mk(rx * y1 / ry, -ry * x1 / rx);
    const double cx = center.x, cy = center.y;
    const Rot phi_tr = rot_new(phi);
    center = rot_apply(phi_tr, center) + 0.5 * (dst + src);
    const P2 v0 = mk(1.0, 0.0);
    const P2 v1 = mk((x1 - cx) / rx, (y1 - cy) / ry);
    const P2 v2 = mk((-x1 - cx) / rx, (-y1 - cy) / ry);
    double eta, eta_delta;
    if (!angle_between(v0, v1, eta)) return false;
    if (!angle_between(v1, v2, eta_delta)) return false;
    if (eta != eta || eta_delta != eta_delta) return false;
    eta_delta = rem_euclid(eta_delta, 2.0 * kPi);
    if (!sweep_flag && eta_delta > 0.0) eta_delta = eta_delta - 2.0 * kPi;
    else if (sweep_flag && eta_delta < 0.0) eta_delta = eta_delta + 2.0 * kPi;

    const double segment_max_angle = kPi / 2.0;
    double segment_count = ceil(fabs(eta_delta) / segment_max_angle);
    const double segment_delta = eta_delta / segment_count;
    double segment_index = 0.0;
    segment_count = segment_count - 1.0;
    while (!(segment_index > segment_count)) {
        const double eta_1 = eta + segment_delta * segment_index;
        const double eta_2 = eta_1 + segment_delta;
        segment_index += 1.0;
        const double tn = tan((eta_2 - eta_1) / 2.0);
        const double sq2 = sqrt(4.0 + 3.0 * (tn * tn));
        const double alpha = sin(eta_2 - eta_1) * (sq2 - 1.0) / 3.0;
        const double sn1 = sin(eta_1), cs1 = cos(eta_1);
        const P2 a0 = rot_apply(phi_tr, mk(rx * cs1, ry * sn1)) + center;
        const P2 d0 = rot_apply(phi_tr, mk(-rx * sn1, ry * cs1));
        const double sn2 = sin(eta_2), cs2 = cos(eta_2);
        const P2 a3 = rot_apply(phi_tr, mk(rx * cs2, ry * sn2)) + center;
        const P2 d3 = rot_apply(phi_tr, mk(-rx * sn2, ry * cs2));
        sink.push(seg_cubic(a0, a0 + alpha * d0, a3 - alpha * d3, a3));
    }
    return true;
}

// ---- joins and caps, src/curve.rs:978-1078 ----
template <class Sink>
SD_FN inline void line_join(const Seg& self, const Seg& other, const Style& style, Sink& sink) {
    if (close_to(seg_end(self), seg_start(other))) return;
    const Ln bevel = mkln(seg_end(self), seg_start(other));
    if (style.join == kJoinBevel) {
        sink.push(seg_line(bevel.a, bevel.b));
        return;
    }
    Ln unused, start, end;
    seg_ends(self, unused, start);
    seg_ends(other, end, unused);
    if (style.join == kJoinMiter) {
        double t0, t1;
        if (!ln_intersect(start, end, t0, t1)) {
            sink.push(seg_line(bevel.a, bevel.b));
        } else if (t0 >= 0.0 && t0 <= 1.0 && t1 >= 0.0 && t1 <= 1.0) {
            sink.push(seg_line(bevel.a, bevel.b));
        } else {
            const P2 p0 = start.b - start.a;
            const P2 p1 = end.a - end.b;
            double c;
            bool done = false;
            if (cos_between(p0, p1, c)) {
                const double miter_length = style.width / sqrt((1.0 - c) / 2.0);
                if (miter_length < style.miter_limit) {
                    const P2 p = ln_at(start, t0);
                    sink.push(seg_line(start.b, p));
                    sink.push(seg_line(p, end.a));
                    done = true;
                }
            }
            if (!done) sink.push(seg_line(bevel.a, bevel.b));
        }
        return;
    }
    // Round
    if (ln_intersect_point(start, end)) {
        sink.push(seg_line(bevel.a, bevel.b));
    } else {
        const bool sweep_flag = cross(ln_dir(start), ln_dir(bevel)) >= 0.0;
        const double radius = style.width / 2.0;
        if (!arc_cubics(start.b, end.a, radius, radius, 0.0, false, sweep_flag, sink)) sink.push(seg_line(bevel.a, bevel.b));
    }
}

template <class Sink>
SD_FN inline void line_cap(const Seg& self, const Seg& other, const Style& style, Sink& sink) {
    if (close_to(seg_end(self), seg_start(other))) return;
    const Ln butt = mkln(seg_end(self), seg_start(other));
    if (style.cap == kCapButt) {
        sink.push(seg_line(butt.a, butt.b));
    } else if (style.cap == kCapSquare) {
        Ln unused, from;
        seg_ends(self, unused, from);
        P2 tang;
        if (normalize(ln_dir(from), tang)) {
            const P2 e = seg_end(self);
            const P2 l0b = e + style.width / 2.0 * tang;
            sink.push(seg_line(e, l0b));
            const P2 l1b = l0b + ln_dir(butt);
            sink.push(seg_line(l0b, l1b));
            sink.push(seg_line(l1b, seg_start(other)));
        }
    } else {
        Style st = style;
        st.join = kJoinRound;
        line_join(self, other, st, sink);
    }
}

// ---- curve offsets, src/curve.rs:1349-1433 ----
SD_FN inline bool quad_offset_should_split(const Seg& q) {
    const P2 p0 = q.p[0], p1 = q.p[1], p2 = q.p[2];
    if (dot(p0 - p1, p2 - p1) > 0.0) return true;
    const P2 c_mass = (p0 + p1 + p2) / 3.0;
    const P2 c_mid = seg_at(q, 0.5);
    const double dist = length(c_mass - c_mid);
    const Box bb = seg_bbox(q);
    const double bbox_diag = ln_length(mkln(bb.lo, bb.hi));
    return bbox_diag * 0.1 < dist;
}
SD_FN inline bool cubic_offset_should_split(const Seg& c) {
    const P2 p0 = c.p[0], p1 = c.p[1], p2 = c.p[2], p3 = c.p[3];
    if (dot(p3 - p0, p2 - p1) < 0.0) return true;
    const double a0 = cross(p3 - p0, p1 - p0);
    const double a1 = cross(p3 - p0, p2 - p0);
    if (a0 * a1 < 0.0) return true;
    const P2 c_mass = (p0 + p1 + p2 + p3) / 4.0;
    const P2 c_mid = seg_at(c, 0.5);
    const double dist = length(c_mass - c_mid);
    const Box bb = seg_bbox(c);
    const double bbox_diag = ln_length(mkln(bb.lo, bb.hi));
    return bbox_diag * 0.1 < dist;
}

// `Curve::offset` (src/curve.rs:275-277, 515-517, 825-827): the recursions of `quad_offset_rec` / `cubic_offset_rec`
// (depth < 3) as a depth-first walk with an explicit stack (one pending sibling per level).
template <class Sink>
SD_FN inline void segment_offset(const Seg& s, double dist, Sink& sink) {
    if (s.kind == 2) {
        Ln l;
        if (ln_offset(mkln(s.p[0], s.p[1]), dist, l)) sink.push(seg_line(l.a, l.b));
        return;
    }
    Seg stack[4];
    int depth[4];
    int sp = 0;
    stack[sp] = s;
    depth[sp++] = 0;
    bool has_last = false;
    Seg last = s;
    while (sp) {
        const Seg cur = stack[--sp];
        const int d = depth[sp];
        const bool split = d < 3 && (s.kind == 3 ? quad_offset_should_split(cur) : cubic_offset_should_split(cur));
        if (split) {
            Seg c0, c1;
            if (s.kind == 3) quad_split_half(cur, c0, c1);
            else cubic_split(cur, c0, c1);
            stack[sp] = c1;
            depth[sp++] = d + 1;
            stack[sp] = c0;
            depth[sp++] = d + 1;
            continue;
        }
        P2 pts[4] = {cur.p[0], cur.p[1], cur.p[2], cur.p[3]};
        if (s.kind == 3) {
            if (polyline_offset(pts, 3, dist)) sink.push(seg_quad(pts[0], pts[1], pts[2]));
        } else if (polyline_offset(pts, 4, dist)) {
            const Seg result = seg_cubic(pts[0], pts[1], pts[2], pts[3]);
            if (has_last && !close_to(seg_end(last), seg_start(result))) {
                Style st;
                st.width = dist * 2.0;
                st.miter_limit = 4.0;
                st.join = kJoinRound;
                st.cap = kCapRound;
                line_join(last, result, st, sink);
            }
            sink.push(result);
            last = result;
            has_last = true;
        }
    }
}

// ---- the unit passes (see stroke.cu) ----
struct PieceRec {  // first / last piece of a unit
    double p[8];
    int kind, pad;
};
SD_FN inline void store_piece(PieceRec& r, const Seg& s) {
    for (int i = 0; i < 4; i++) {
        r.p[2 * i] = s.p[i].x;
        r.p[2 * i + 1] = s.p[i].y;
    }
    r.kind = s.kind;
}
SD_FN inline Seg load_piece(const PieceRec& r) {
    Seg s;
    for (int i = 0; i < 4; i++) s.p[i] = mk(r.p[2 * i], r.p[2 * i + 1]);
    s.kind = r.kind;
    return s;
}
// the source segment of an ordinary unit (reversed when the unit walks backwards); pts = (x, y) pairs
SD_FN inline Seg load_src(const StrokeUnit& u, const double* pts) {
    const int kind = (int)(u.b & 7u);
    const bool rev = (u.b & kUnitReversed) != 0;
    Seg s;
    s.kind = kind;
    for (int i = 0; i < 4; i++) {
        const int j = i < kind ? i : kind - 1;
        const double* v = pts + 2 * (size_t)(u.a + (uint32_t)(rev ? kind - 1 - j : j));
        s.p[i] = mk(v[0], v[1]);
    }
    return s;
}

struct CountSink {
    uint32_t seg = 0, pts = 0, curves = 0;
    Seg first, last;
    SD_FN void push(const Seg& s) {
        if (seg == 0) first = s;
        last = s;
        seg++;
        pts += (uint32_t)s.kind;
        curves += s.kind != 2;
    }
};
// Writes segments as a device path: points, the item list in the reference's order and the curves-first copy of it
// (curves at [0, total_curves), lines and closing items behind them, each group in order).  I2 = uint2.
template <class I2>
struct EmitSink {
    double* pts;
    I2* items;
    I2* packed;
    uint32_t pt, item, curve, total_curves;  // running global indices
    SD_FN void put(uint32_t x, uint32_t y, bool is_curve) {
        I2 it;
        it.x = x;
        it.y = y;
        items[item] = it;
        if (is_curve) packed[curve] = it;
        else packed[total_curves + (item - curve)] = it;
        curve += is_curve;
        item++;
    }
    SD_FN void push(const Seg& s) {
        for (int i = 0; i < s.kind; i++) {
            pts[2 * (size_t)(pt + i)] = s.p[i].x;
            pts[2 * (size_t)(pt + i) + 1] = s.p[i].y;
        }
        put(pt, (uint32_t)s.kind, s.kind != 2);
        pt += (uint32_t)s.kind;
    }
};

// what a closer appends: `stroke_close` (src/path.rs:708-732) or the final cap of an open subpath (:403-411)
template <class Sink>
SD_FN inline void closer_segments(const StrokeUnit& u, const double* pts, const Seg& first, const Seg& last, const Style& style, Sink& sink) {
    const uint32_t mode = (u.b >> 16) & 3u;
    if (mode == kCloserOpen) {
        line_cap(last, first, style, sink);
        return;
    }
    const P2 sp_start = mk(pts[2 * (size_t)u.a], pts[2 * (size_t)u.a + 1]), sp_end = mk(pts[2 * (size_t)u.d], pts[2 * (size_t)u.d + 1]);
    const Ln close = mode == kCloserForward ? mkln(sp_end, sp_start) : mkln(sp_start, sp_end);
    Ln off;
    if (ln_offset(close, style.width / 2.0, off) && ln_length(off) * 100.0 > style.width) {
        const Seg c = seg_line(off.a, off.b);
        line_join(last, c, style, sink);
        sink.push(c);
        line_join(c, first, style, sink);
    } else {
        line_join(last, first, style, sink);
    }
}

// Nearest unit of [lo, hi) with pieces, searching downwards from hi - 1 (upwards from lo); `none` when there is none.
// cnt_seg[j] > 0 exactly when unit j has pieces, before and after pass 2 adds its join to it.
SD_FN inline uint32_t prev_nonempty(const uint32_t* cnt_seg, uint32_t lo, uint32_t hi, uint32_t none) {
    for (uint32_t j = hi; j > lo; j--)
        if (cnt_seg[j - 1]) return j - 1;
    return none;
}
SD_FN inline uint32_t next_nonempty(const uint32_t* cnt_seg, uint32_t lo, uint32_t hi, uint32_t none) {
    for (uint32_t j = lo; j < hi; j++)
        if (cnt_seg[j]) return j;
    return none;
}

// pass 1: the pieces of ordinary unit i (counts, first and last piece)
SD_FN inline void unit_pieces(uint32_t i, const StrokeUnit* units, const double* pts, const Style& style, uint32_t* cnt_seg, uint32_t* cnt_pts,
                              uint32_t* cnt_curves, PieceRec* first, PieceRec* last) {
    const StrokeUnit u = units[i];
    if (u.b & kUnitCloser) return;  // counted in pass 2
    CountSink sink;
    segment_offset(load_src(u, pts), style.width / 2.0, sink);
    cnt_seg[i] = sink.seg;
    cnt_pts[i] = sink.pts;
    cnt_curves[i] = sink.curves;
    if (sink.seg) {
        store_piece(first[i], sink.first);
        store_piece(last[i], sink.last);
    }
}

// The join (or cap) in front of the pieces of ordinary unit i, or everything closer unit i appends, into `sink`.
// Returns false when the unit emits nothing at all (no pieces / an empty contour).
template <class Sink>
SD_FN inline bool unit_lead(uint32_t i, const StrokeUnit& u, uint32_t n_units, const double* pts, const Style& style, const uint32_t* cnt_seg,
                            const PieceRec* first, const PieceRec* last, Sink& sink) {
    const uint32_t u0 = u.c;
    if (u.b & kUnitCloser) {
        const uint32_t f = next_nonempty(cnt_seg, u0, i, n_units);
        if (f == n_units) return false;
        const uint32_t l = prev_nonempty(cnt_seg, u0, i, n_units);
        closer_segments(u, pts, load_piece(first[f]), load_piece(last[l]), style, sink);
        return true;
    }
    if (!cnt_seg[i]) return false;
    const uint32_t p = prev_nonempty(cnt_seg, u0, i, n_units);
    if (p != n_units) {
        const Seg src = load_piece(last[p]), dst = load_piece(first[i]);
        if (u.b & kUnitCap) line_cap(src, dst, style, sink);
        else line_join(src, dst, style, sink);
    }
    return true;
}

// pass 2: add what unit_lead emits to the counts of unit i; a closer of a contour with output counts one subpath
SD_FN inline void unit_count(uint32_t i, const StrokeUnit* units, uint32_t n_units, const double* pts, const Style& style, uint32_t* cnt_seg,
                             uint32_t* cnt_pts, uint32_t* cnt_curves, uint32_t* cnt_close, const PieceRec* first, const PieceRec* last) {
    const StrokeUnit u = units[i];
    CountSink sink;
    const bool any = unit_lead(i, u, n_units, pts, style, cnt_seg, first, last, sink);
    if (u.b & kUnitCloser) {
        cnt_seg[i] = sink.seg;
        cnt_pts[i] = sink.pts;
        cnt_curves[i] = sink.curves;
        cnt_close[i] = any ? 1u : 0u;
    } else {
        cnt_close[i] = 0;
        if (sink.seg) {
            cnt_seg[i] += sink.seg;
            cnt_pts[i] += sink.pts;
            cnt_curves[i] += sink.curves;
        }
    }
}

// pass 3: write unit i at its scanned offsets (`sink` starts there).  `closing_flags` = the item flags of a closing item,
// `contour_first_pt` = off_pts[u.c], the first output point of the unit's contour.
template <class Sink>
SD_FN inline void unit_emit(uint32_t i, const StrokeUnit* units, uint32_t n_units, const double* pts, const Style& style, const uint32_t* cnt_seg,
                            const PieceRec* first, const PieceRec* last, uint32_t contour_first_pt, uint32_t closing_flags, Sink& sink) {
    const StrokeUnit u = units[i];
    if (!unit_lead(i, u, n_units, pts, style, cnt_seg, first, last, sink)) return;
    if (u.b & kUnitCloser) sink.put(sink.pt - 1, closing_flags | contour_first_pt, false);
    else segment_offset(load_src(u, pts), style.width / 2.0, sink);
}

}  // namespace sk
}  // namespace rgpu
