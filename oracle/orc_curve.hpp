// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_geom.hpp header).
// Restatement of src/curve.rs (Line/Quad/Cubic/Segment, flatness, split, bbox, offset, joins)
// and src/ellipse.rs (SVG arc -> cubics).
#pragma once
#include "orc_geom.hpp"

namespace orc {

enum class LineJoin { Miter, Bevel, Round };  // src/path.rs:78-87
enum class LineCap { Butt, Square, Round };   // src/path.rs:103-110
struct StrokeStyle {                          // src/path.rs:121-136
    Scalar width = 0.0;
    LineJoin line_join = LineJoin::Miter;
    Scalar miter_limit = 4.0;  // LineJoin::default() = Miter(4.0), src/path.rs:89-93
    LineCap line_cap = LineCap::Butt;
};

struct Line {  // src/curve.rs:163
    Point p[2];
    Line() = default;
    Line(Point a, Point b) { p[0] = a; p[1] = b; }
    Point start() const { return p[0]; }
    Point end() const { return p[1]; }
    Scalar length() const { return p[0].dist(p[1]); }                      // :185-188
    Point at(Scalar t) const { return (1.0 - t) * p[0] + t * p[1]; }        // :250-253
    Point direction() const { return end() - start(); }                     // :227-229
    Line transform(const Transform& tr) const { return Line(tr.apply(p[0]), tr.apply(p[1])); }  // :237-240
    // :204-214
    std::optional<std::pair<Scalar, Scalar>> intersect(const Line& o) const {
        Scalar x1 = p[0].x, y1 = p[0].y, x2 = p[1].x, y2 = p[1].y;
        Scalar x3 = o.p[0].x, y3 = o.p[0].y, x4 = o.p[1].x, y4 = o.p[1].y;
        Scalar det = (x4 - x3) * (y1 - y2) - (x1 - x2) * (y4 - y3);
        if (std::fabs(det) < EPSILON) return std::nullopt;
        Scalar t0 = ((y3 - y4) * (x1 - x3) + (x4 - x3) * (y1 - y3)) / det;
        Scalar t1 = ((y1 - y2) * (x1 - x3) + (x2 - x1) * (y1 - y3)) / det;
        return std::make_pair(t0, t1);
    }
    // :217-224
    std::optional<Point> intersect_point(const Line& o) const {
        auto t = intersect(o);
        if (!t) return std::nullopt;
        if (t->first >= 0.0 && t->first <= 1.0 && t->second >= 0.0 && t->second <= 1.0) return at(t->first);
        return std::nullopt;
    }
};

enum class SegKind : uint8_t { Line = 2, Quad = 3, Cubic = 4 };  // value = number of control points

// `Segment` enum of src/curve.rs:905-909 stored as a tagged fixed array.
struct Segment {
    SegKind kind = SegKind::Line;
    Point p[4];

    static Segment line(Point a, Point b) { Segment s; s.kind = SegKind::Line; s.p[0] = a; s.p[1] = b; return s; }
    static Segment line(const Line& l) { return line(l.p[0], l.p[1]); }
    static Segment quad(Point a, Point b, Point c) { Segment s; s.kind = SegKind::Quad; s.p[0] = a; s.p[1] = b; s.p[2] = c; return s; }
    static Segment cubic(Point a, Point b, Point c, Point d) {
        Segment s; s.kind = SegKind::Cubic; s.p[0] = a; s.p[1] = b; s.p[2] = c; s.p[3] = d; return s;
    }
    int npts() const { return (int)kind; }
    Point start() const { return p[0]; }
    Point end() const { return p[npts() - 1]; }

    // src/curve.rs:233-235 (Line), :413-417 (Quad), :692-697 (Cubic) — returns 16*f^2
    Scalar flatness() const {
        switch (kind) {
            case SegKind::Line: return 0.0;
            case SegKind::Quad: {
                Point d = 2.0 * p[1] - p[0] - p[2];
                return d.x * d.x + d.y * d.y;
            }
            default: {
                Point u = 3.0 * p[1] - 2.0 * p[0] - p[3];
                Point v = 3.0 * p[2] - p[0] - 2.0 * p[3];
                return rmax(u.x * u.x, v.x * v.x) + rmax(u.y * u.y, v.y * v.y);
            }
        }
    }
    // src/curve.rs:237-240, 419-422, 699-702
    Segment transform(const Transform& tr) const {
        Segment s;
        s.kind = kind;
        for (int i = 0; i < npts(); i++) s.p[i] = tr.apply(p[i]);
        return s;
    }
    // src/curve.rs:250-253, 432-441, 712-723
    Point at(Scalar t) const {
        switch (kind) {
            case SegKind::Line: return (1.0 - t) * p[0] + t * p[1];
            case SegKind::Quad: {
                Scalar t1 = t, t_1 = 1.0 - t;
                Scalar t2 = t1 * t1, t_2 = t_1 * t_1;
                return t_2 * p[0] + 2.0 * t1 * t_1 * p[1] + t2 * p[2];
            }
            default: {
                Scalar t1 = t, t_1 = 1.0 - t;
                Scalar t2 = t1 * t1, t_2 = t_1 * t_1;
                Scalar t3 = t2 * t1, t_3 = t_2 * t_1;
                return t_3 * p[0] + 3.0 * t1 * t_2 * p[1] + 3.0 * t2 * t_1 * p[2] + t3 * p[3];
            }
        }
    }
    // `Curve::split`: Line uses the default split_at(0.5) (src/curve.rs:43-45, 260-264);
    // Quad :449-456; Cubic :731-747.
    std::pair<Segment, Segment> split() const {
        switch (kind) {
            case SegKind::Line: {
                Point mid = at(0.5);
                return {line(p[0], mid), line(mid, p[1])};
            }
            case SegKind::Quad: {
                Point mid = 0.25 * (p[0] + 2.0 * p[1] + p[2]);
                return {quad(p[0], 0.5 * (p[0] + p[1]), mid), quad(mid, 0.5 * (p[1] + p[2]), p[2])};
            }
            default: {
                Point mid = 0.125 * p[0] + 0.375 * p[1] + 0.375 * p[2] + 0.125 * p[3];
                Segment c0 = cubic(p[0], 0.5 * p[0] + 0.5 * p[1], 0.25 * p[0] + 0.5 * p[1] + 0.25 * p[2], mid);
                Segment c1 = cubic(mid, 0.25 * p[1] + 0.5 * p[2] + 0.25 * p[3], 0.5 * p[2] + 0.5 * p[3], p[3]);
                return {c0, c1};
            }
        }
    }
    // Quad::split_at, src/curve.rs:458-468 (only the quad variant is needed: quad_offset_rec)
    std::pair<Segment, Segment> quad_split_at(Scalar t) const {
        Scalar t1 = t, t_1 = 1.0 - t;
        Scalar t2 = t1 * t1, t_2 = t_1 * t_1;
        Point mid = t_2 * p[0] + 2.0 * t1 * t_1 * p[1] + t2 * p[2];
        return {quad(p[0], t_1 * p[0] + t * p[1], mid), quad(mid, t_1 * p[1] + t * p[2], p[2])};
    }
    // src/curve.rs:279-282, 519-522, 829-832
    Segment reverse() const {
        Segment s;
        s.kind = kind;
        int n = npts();
        for (int i = 0; i < n; i++) s.p[i] = p[n - 1 - i];
        return s;
    }
    // src/curve.rs:1081-1089
    bool has_nans() const {
        for (int i = 0; i < npts(); i++)
            if (std::isnan(p[i].x) || std::isnan(p[i].y)) return true;
        return false;
    }
    // src/curve.rs:296-298 (Line: none), :535-555 (Quad), :853-864 (Cubic); order = x roots then y roots
    int extremities(Scalar out[6]) const {
        int n = 0;
        if (kind == SegKind::Quad) {
            Point a = p[2] - 2.0 * p[1] + p[0];
            Point b = p[1] - p[0];
            if (std::fabs(a.x) > EPSILON) {
                Scalar t0 = -b.x / a.x;
                if (t0 >= 0.0 && t0 <= 1.0) out[n++] = t0;
            }
            if (std::fabs(a.y) > EPSILON) {
                Scalar t1 = -b.y / a.y;
                if (t1 >= 0.0 && t1 <= 1.0) out[n++] = t1;
            }
        } else if (kind == SegKind::Cubic) {
            Point a = -1.0 * p[0] + 3.0 * p[1] - 3.0 * p[2] + 1.0 * p[3];
            Point b = 2.0 * p[0] - 4.0 * p[1] + 2.0 * p[2];
            Point c = -1.0 * p[0] + p[1];
            Roots2 rx = quadratic_solve(a.x, b.x, c.x);
            Roots2 ry = quadratic_solve(a.y, b.y, c.y);
            for (int i = 0; i < rx.n; i++) if (rx.v[i] >= 0.0 && rx.v[i] <= 1.0) out[n++] = rx.v[i];
            for (int i = 0; i < ry.n; i++) if (ry.v[i] >= 0.0 && ry.v[i] <= 1.0) out[n++] = ry.v[i];
        }
        return n;
    }
    // src/curve.rs:270-273 (Line), :505-513 (Quad), :810-818 (Cubic)
    BBox bbox(const std::optional<BBox>& init) const {
        BBox bb = BBox(start(), end()).union_opt(init);
        if (kind == SegKind::Line) return bb;
        if (kind == SegKind::Quad) {
            if (bb.contains(p[1])) return bb;
        } else {
            if (bb.contains(p[1]) && bb.contains(p[2])) return bb;
        }
        Scalar ts[6];
        int n = extremities(ts);
        for (int i = 0; i < n; i++) bb = bb.extend(at(ts[i]));
        return bb;
    }
    // src/curve.rs:195-197 (Line), :377-388 (Quad), :647-667 (Cubic): tangent lines at the ends
    std::pair<Line, Line> ends() const {
        switch (kind) {
            case SegKind::Line: return {Line(p[0], p[1]), Line(p[0], p[1])};
            case SegKind::Quad: {
                Line s(p[0], p[1]), e(p[1], p[2]);
                if (p[0].is_close_to(p[1])) return {e, e};
                if (p[1].is_close_to(p[2])) return {s, s};
                return {s, e};
            }
            default: {
                int s = 0;
                for (int i = 0; i < 3; i++) if (!p[i].is_close_to(p[i + 1])) { s = i; break; }
                int e = 0;
                for (int i = 3; i >= 1; i--) if (!p[i].is_close_to(p[i - 1])) { e = i; break; }
                // NOTE: with all four points coincident the reference indexes ps[end-1] with end == 0 and panics.
                if (e == 0) e = 1;
                return {Line(p[s], p[s + 1]), Line(p[e - 1], p[e])};
            }
        }
    }
};

// ---- Elliptic arc: src/ellipse.rs -------------------------------------------------------
struct EllipArc {
    Point center;
    Scalar rx, ry, phi, eta, eta_delta;

    // src/ellipse.rs:40-96
    static std::optional<EllipArc> new_param(Point src, Point dst, Scalar rx, Scalar ry, Scalar x_axis_rot, bool large_flag,
                                             bool sweep_flag) {
        rx = std::fabs(rx);
        ry = std::fabs(ry);
        Scalar phi = x_axis_rot * PI / 180.0;
        Point p1 = Transform::new_rotate(-phi).apply(0.5 * (src - dst));
        Scalar x1 = p1.x, y1 = p1.y;
        Scalar ax = x1 / rx, ay = y1 / ry;
        Scalar s = ax * ax + ay * ay;
        if (s > 1.0) {
            Scalar sq = std::sqrt(s);
            rx = rx * sq;
            ry = ry * sq;
        }
        Scalar rxry = rx * ry, rxy1 = rx * y1, ryx1 = ry * x1;
        Scalar sq = std::sqrt(rmax(rxry * rxry / (rxy1 * rxy1 + ryx1 * ryx1) - 1.0, 0.0));
        sq = (large_flag == sweep_flag) ? -sq : sq;
        Point center = sq * Point(rx * y1 / ry, -ry * x1 / rx);
        Scalar cx = center.x, cy = center.y;
        center = Transform::new_rotate(phi).apply(center) + 0.5 * (dst + src);
        Point v0(1.0, 0.0);
        Point v1((x1 - cx) / rx, (y1 - cy) / ry);
        Point v2((-x1 - cx) / rx, (-y1 - cy) / ry);
        auto eta = v0.angle_between(v1);
        if (!eta) return std::nullopt;
        auto ed = v1.angle_between(v2);
        if (!ed) return std::nullopt;
        Scalar eta_delta = rem_euclid(*ed, 2.0 * PI);
        if (!sweep_flag && eta_delta > 0.0) eta_delta = eta_delta - 2.0 * PI;
        else if (sweep_flag && eta_delta < 0.0) eta_delta = eta_delta + 2.0 * PI;
        EllipArc a;
        a.center = center; a.rx = rx; a.ry = ry; a.phi = phi; a.eta = *eta; a.eta_delta = eta_delta;
        return a;
    }

    // src/ellipse.rs:167-214 — EllipArcCubicIter collected into a vector.
    // Float division by zero is kept as in Rust: a zero sweep gives segment_count = 0 - 1 = -1 and
    // segment_delta = NaN, and the iterator yields nothing.
    std::vector<Segment> to_cubics() const {
        std::vector<Segment> out;
        Transform phi_tr = Transform::new_rotate(phi);
        Scalar segment_max_angle = PI / 2.0;
        Scalar segment_count = std::ceil(std::fabs(eta_delta) / segment_max_angle);
        Scalar segment_delta = eta_delta / segment_count;
        Scalar segment_index = 0.0;
        segment_count = segment_count - 1.0;
        auto at = [&](Scalar alpha) -> std::pair<Point, Point> {
            Scalar sn = std::sin(alpha), cs = std::cos(alpha);
            Point a = phi_tr.apply(Point(rx * cs, ry * sn)) + center;
            Point d = phi_tr.apply(Point(-rx * sn, ry * cs));
            return {a, d};
        };
        while (!(segment_index > segment_count)) {
            Scalar eta_1 = eta + segment_delta * segment_index;
            Scalar eta_2 = eta_1 + segment_delta;
            segment_index += 1.0;
            Scalar tn = std::tan((eta_2 - eta_1) / 2.0);
            Scalar sq = std::sqrt(4.0 + 3.0 * (tn * tn));
            Scalar alpha = std::sin(eta_2 - eta_1) * (sq - 1.0) / 3.0;
            auto [p0, d0] = at(eta_1);
            auto [p3, d3] = at(eta_2);
            Point p1 = p0 + alpha * d0;
            Point p2 = p3 - alpha * d3;
            out.push_back(Segment::cubic(p0, p1, p2, p3));
            if (out.size() > 64) break;  // guard against NaN loops; the reference cannot exceed 4 here
        }
        return out;
    }
};

// ---- stroke helpers: src/curve.rs:978-1078, 1283-1433 -----------------------------------

// :1283-1287
inline std::optional<Line> line_offset(const Line& line, Scalar dist) {
    auto n = (line.p[1] - line.p[0]).normal().normalize();
    if (!n) return std::nullopt;
    Point offset = dist * *n;
    return Line(line.p[0] + offset, line.p[1] + offset);
}

// :1294-1346
inline bool polyline_offset(Point* ps, size_t len, Scalar dist) {
    if (len == 0) return true;
    std::optional<Line> prev;
    size_t index = 0;
    while (true) {
        size_t repeats = 1;
        for (size_t i = index; i + 1 < len; i++) {
            if (!ps[i].is_close_to(ps[i + 1])) break;
            repeats += 1;
        }
        if (index + repeats >= len) break;
        index += repeats;
        auto next = line_offset(Line(ps[index - 1], ps[index]), dist);
        if (!next) return false;  // reference: expect("polyline implementation error")
        Point point;
        if (!prev) {
            point = next->start();
        } else {
            auto t = prev->intersect(*next);
            point = t ? prev->at(t->first) : next->start();
        }
        for (size_t i = index - repeats; i < index; i++) ps[i] = point;
        prev = next;
    }
    if (!prev) return false;
    for (size_t i = index; i < len; i++) ps[i] = prev->end();
    return true;
}

std::vector<Segment> line_join(const Segment& self, const Segment& other, const StrokeStyle& style);

// :1349-1362
inline bool quad_offset_should_split(const Segment& q) {
    Point p0 = q.p[0], p1 = q.p[1], p2 = q.p[2];
    if ((p0 - p1).dot(p2 - p1) > 0.0) return true;
    Point c_mass = (p0 + p1 + p2) / 3.0;
    Point c_mid = q.at(0.5);
    Scalar dist = (c_mass - c_mid).length();
    BBox bb = q.bbox(std::nullopt);
    Scalar bbox_diag = Line(bb.min, bb.max).length();
    return bbox_diag * 0.1 < dist;
}
// :1365-1376
inline void quad_offset_rec(const Segment& q, Scalar dist, std::vector<Segment>& out, size_t depth) {
    if (quad_offset_should_split(q) && depth < 3) {
        auto [c0, c1] = q.quad_split_at(0.5);
        quad_offset_rec(c0, dist, out, depth + 1);
        quad_offset_rec(c1, dist, out, depth + 1);
    } else {
        Point pts[3] = {q.p[0], q.p[1], q.p[2]};
        if (polyline_offset(pts, 3, dist)) out.push_back(Segment::quad(pts[0], pts[1], pts[2]));
    }
}
// :1414-1433
inline bool cubic_offset_should_split(const Segment& c) {
    Point p0 = c.p[0], p1 = c.p[1], p2 = c.p[2], p3 = c.p[3];
    if ((p3 - p0).dot(p2 - p1) < 0.0) return true;
    Scalar a0 = (p3 - p0).cross(p1 - p0);
    Scalar a1 = (p3 - p0).cross(p2 - p0);
    if (a0 * a1 < 0.0) return true;
    Point c_mass = (p0 + p1 + p2 + p3) / 4.0;
    Point c_mid = c.at(0.5);
    Scalar dist = (c_mass - c_mid).length();
    BBox bb = c.bbox(std::nullopt);
    Scalar bbox_diag = Line(bb.min, bb.max).length();
    return bbox_diag * 0.1 < dist;
}
// :1379-1411
inline std::optional<Segment> cubic_offset_rec(const Segment& c, std::optional<Segment> last, Scalar dist,
                                               std::vector<Segment>& out, size_t depth) {
    if (cubic_offset_should_split(c) && depth < 3) {
        auto [c0, c1] = c.split();
        last = cubic_offset_rec(c0, last, dist, out, depth + 1);
        return cubic_offset_rec(c1, last, dist, out, depth + 1);
    }
    Point pts[4] = {c.p[0], c.p[1], c.p[2], c.p[3]};
    if (polyline_offset(pts, 4, dist)) {
        Segment result = Segment::cubic(pts[0], pts[1], pts[2], pts[3]);
        if (last) {
            if (!last->end().is_close_to(result.start())) {
                StrokeStyle st;
                st.width = dist * 2.0;
                st.line_join = LineJoin::Round;
                st.line_cap = LineCap::Round;
                auto j = line_join(*last, result, st);
                out.insert(out.end(), j.begin(), j.end());
            }
        }
        out.push_back(result);
        return result;
    }
    return last;
}
// `Curve::offset`: src/curve.rs:275-277, 515-517, 825-827
inline void segment_offset(const Segment& s, Scalar dist, std::vector<Segment>& out) {
    switch (s.kind) {
        case SegKind::Line: {
            auto l = line_offset(Line(s.p[0], s.p[1]), dist);
            if (l) out.push_back(Segment::line(*l));
            break;
        }
        case SegKind::Quad: quad_offset_rec(s, dist, out, 0); break;
        default: cubic_offset_rec(s, std::nullopt, dist, out, 0); break;
    }
}

// src/curve.rs:978-1047
inline std::vector<Segment> line_join(const Segment& self, const Segment& other, const StrokeStyle& style) {
    std::vector<Segment> result;
    if (self.end().is_close_to(other.start())) return result;
    Line bevel(self.end(), other.start());
    switch (style.line_join) {
        case LineJoin::Bevel: result.push_back(Segment::line(bevel)); break;
        case LineJoin::Miter: {
            Line start = self.ends().second;
            Line end = other.ends().first;
            auto t = start.intersect(end);
            if (!t) {
                result.push_back(Segment::line(bevel));
            } else if (t->first >= 0.0 && t->first <= 1.0 && t->second >= 0.0 && t->second <= 1.0) {
                result.push_back(Segment::line(bevel));
            } else {
                Point p0 = start.end() - start.start();
                Point p1 = end.start() - end.end();
                auto c = p0.cos_between(p1);
                bool done = false;
                if (c) {
                    Scalar miter_length = style.width / std::sqrt((1.0 - *c) / 2.0);
                    if (miter_length < style.miter_limit) {
                        Point p = start.at(t->first);
                        result.push_back(Segment::line(start.end(), p));
                        result.push_back(Segment::line(p, end.start()));
                        done = true;
                    }
                }
                if (!done) result.push_back(Segment::line(bevel));
            }
            break;
        }
        case LineJoin::Round: {
            Line start = self.ends().second;
            Line end = other.ends().first;
            if (start.intersect_point(end)) {
                result.push_back(Segment::line(bevel));
            } else {
                bool sweep_flag = start.direction().cross(bevel.direction()) >= 0.0;
                Scalar radius = style.width / 2.0;
                auto arc = EllipArc::new_param(start.end(), end.start(), radius, radius, 0.0, false, sweep_flag);
                if (arc) {
                    auto cs = arc->to_cubics();
                    result.insert(result.end(), cs.begin(), cs.end());
                } else {
                    result.push_back(Segment::line(bevel));
                }
            }
            break;
        }
    }
    return result;
}

// src/curve.rs:1050-1078
inline std::vector<Segment> line_cap(const Segment& self, const Segment& other, const StrokeStyle& style) {
    std::vector<Segment> result;
    if (self.end().is_close_to(other.start())) return result;
    Line butt(self.end(), other.start());
    switch (style.line_cap) {
        case LineCap::Butt: result.push_back(Segment::line(butt)); break;
        case LineCap::Square: {
            Line from = self.ends().second;
            auto tang = from.direction().normalize();
            if (tang) {
                Line l0(self.end(), self.end() + style.width / 2.0 * *tang);
                result.push_back(Segment::line(l0));
                Line l1(l0.end(), l0.end() + butt.direction());
                result.push_back(Segment::line(l1));
                Line l2(l1.end(), other.start());
                result.push_back(Segment::line(l2));
            }
            break;
        }
        case LineCap::Round: {
            StrokeStyle st = style;
            st.line_join = LineJoin::Round;
            result = line_join(self, other, st);
            break;
        }
    }
    return result;
}

}  // namespace orc
