// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_geom.hpp header).
// Restatement of src/path.rs (Path, PathBuilder, flatten, stroke, bbox, size) and
// src/svg.rs (scalar / path / transform parsers).
#pragma once
#include "orc_curve.hpp"
#include <stdexcept>
#include <cstring>

namespace orc {

constexpr Scalar DEFAULT_FLATNESS = 0.05;  // src/path.rs:16

enum class FillRule : int { NonZero = 0, EvenOdd = 1 };  // src/path.rs:21-29

// src/path.rs:32-46
inline Scalar alpha_from_winding(FillRule rule, Scalar winding) {
    if (rule == FillRule::EvenOdd) return std::fabs(rem_euclid(winding + 1.0, 2.0) - 1.0);
    Scalar value = std::fabs(winding);
    if (value >= 1.0) return 1.0;
    if (value < 1e-6) return 0.0;
    return value;
}

struct ParseError : std::runtime_error { using std::runtime_error::runtime_error; };

// src/path.rs:227-233
struct Path {
    std::vector<Segment> segments;
    std::vector<size_t> subpaths;  // segments[subpaths[i]..subpaths[i+1]] is the i-th subpath
    std::vector<uint8_t> closed;

    size_t len() const { return subpaths.size() < 2 ? 0 : subpaths.size() - 1; }  // :259-265
    bool is_empty() const { return subpaths.empty(); }

    // src/path.rs:336-346
    void push(const std::vector<Segment>& segs, bool is_closed) {
        if (segs.empty()) return;
        if (subpaths.empty()) subpaths.push_back(0);
        segments.insert(segments.end(), segs.begin(), segs.end());
        subpaths.push_back(segments.size());
        closed.push_back(is_closed);
    }
    // src/path.rs:359-363
    void transform(const Transform& tr) {
        for (auto& s : segments) s = s.transform(tr);
    }
    // src/path.rs:428-431 + SubPath::bbox :217-222
    std::optional<BBox> bbox(const Transform& tr) const {
        std::optional<BBox> bb;
        for (size_t i = 0; i < len(); i++)
            for (size_t s = subpaths[i]; s < subpaths[i + 1]; s++) bb = segments[s].transform(tr).bbox(bb);
        return bb;
    }
    // src/path.rs:439-451
    struct SizeResult { Size size; Transform tr; Point min; };
    std::optional<SizeResult> size(const Transform& tr) const {
        auto bb = bbox(tr);
        if (!bb) return std::nullopt;
        Scalar min_x = bb->min.x, min_y = bb->min.y, max_x = bb->max.x, max_y = bb->max.y;
        Point mn(std::floor(min_x) - 1.0, std::floor(min_y) - 1.0);
        Point mx(std::ceil(max_x) + 1.0, std::ceil(max_y) + 1.0);
        SizeResult r;
        r.size.width = as_usize(std::round(mx.x - mn.x));
        r.size.height = as_usize(std::round(mx.y - mn.y));
        Transform shift = Transform::new_translate(1.0 - min_x, 1.0 - min_y);
        r.tr = shift * tr;
        r.min = mn;
        return r;
    }

    // `Path::flatten` = PathFlattenIter, src/path.rs:418-425, 744-795. Lines are appended as
    // (x0,y0,x1,y1). Throws on NaN control points (reference panics, :765-767).
    void flatten(const Transform& tr, Scalar flatness, bool close, std::vector<Line>& out) const {
        Scalar thr = 16.0 * flatness * flatness;  // :749
        std::vector<Segment> stack;
        for (size_t sp = 0; sp < len(); sp++) {
            size_t b = subpaths[sp], e = subpaths[sp + 1];
            for (size_t si = b; si < e; si++) {
                stack.push_back(segments[si].transform(tr));
                while (!stack.empty()) {
                    Segment seg = stack.back();
                    stack.pop_back();
                    if (seg.has_nans()) throw std::runtime_error("cannot flatten segment with NaN");
                    if (seg.flatness() < thr) {
                        out.push_back(Line(seg.start(), seg.end()));
                        continue;
                    }
                    auto [s0, s1] = seg.split();
                    stack.push_back(s1);
                    stack.push_back(s0);
                }
            }
            if (closed[sp] || close) {  // :781-785 — emitted even when zero length
                out.push_back(Line(segments[e - 1].end(), segments[b].start()).transform(tr));
            }
        }
    }

    Path stroke(const StrokeStyle& style) const;
};

// src/path.rs:692-706
template <class JoinFn>
inline void stroke_segment(std::vector<Segment>& segments, const Segment& segment, const StrokeStyle& style, JoinFn join) {
    size_t offset = segments.size();
    segment_offset(segment, style.width / 2.0, segments);
    if (offset != 0) {
        if (offset - 1 < segments.size() && offset < segments.size()) {
            Segment src = segments[offset - 1];
            Segment dst = segments[offset];
            auto j = join(src, dst, style);
            segments.insert(segments.begin() + (ptrdiff_t)offset, j.begin(), j.end());
        }
    }
}

// src/path.rs:708-732
inline void stroke_close(const Path& path, size_t sp, std::vector<Segment>& segments, const StrokeStyle& style, bool forward) {
    if (segments.empty()) return;
    Segment first = segments.front(), last = segments.back();
    Point sp_start = path.segments[path.subpaths[sp]].start();
    Point sp_end = path.segments[path.subpaths[sp + 1] - 1].end();
    Line close = forward ? Line(sp_end, sp_start) : Line(sp_start, sp_end);
    auto off = line_offset(close, style.width / 2.0);
    if (off && off->length() * 100.0 > style.width) {
        Segment c = Segment::line(*off);
        auto j0 = line_join(last, c, style);
        segments.insert(segments.end(), j0.begin(), j0.end());
        segments.push_back(c);
        auto j1 = line_join(c, first, style);
        segments.insert(segments.end(), j1.begin(), j1.end());
    } else {
        auto j = line_join(last, first, style);
        segments.insert(segments.end(), j.begin(), j.end());
    }
}

// src/path.rs:374-415
inline Path Path::stroke(const StrokeStyle& style) const {
    Path result;
    std::vector<Segment> segs;
    for (size_t sp = 0; sp < len(); sp++) {
        size_t b = subpaths[sp], e = subpaths[sp + 1];
        bool is_closed = closed[sp] != 0;
        for (size_t i = b; i < e; i++) stroke_segment(segs, segments[i], style, line_join);
        // backward = segments reversed, each reversed
        size_t back = e;  // next backward index is back-1
        if (is_closed) {
            stroke_close(*this, sp, segs, style, true);
            result.push(segs, true);
            segs.clear();
        } else {
            if (back > b) {
                back--;
                stroke_segment(segs, segments[back].reverse(), style, line_cap);
            }
        }
        while (back > b) {
            back--;
            stroke_segment(segs, segments[back].reverse(), style, line_join);
        }
        if (is_closed) {
            stroke_close(*this, sp, segs, style, false);
            result.push(segs, true);
            segs.clear();
        } else {
            if (!segs.empty()) {
                Segment last = segs.back(), first = segs.front();
                auto c = line_cap(last, first, style);
                segs.insert(segs.end(), c.begin(), c.end());
            }
            result.push(segs, true);
            segs.clear();
        }
    }
    return result;
}

// ---- PathBuilder: src/path.rs:800-1056 ---------------------------------------------------
struct PathBuilder {
    Point position{0.0, 0.0};
    std::vector<Segment> segments;
    std::vector<size_t> subpaths;
    std::vector<uint8_t> closed;

    // :849-868
    void subpath_finish(bool close) {
        if (segments.empty() || (!subpaths.empty() && subpaths.back() == segments.size())) return;
        if (subpaths.empty()) subpaths.push_back(0);
        if (close) {
            size_t first_index = subpaths.back();
            if (first_index < segments.size()) position = segments[first_index].start();
        }
        subpaths.push_back(segments.size());
        closed.push_back(close);
    }
    Path build() {  // :832-841
        subpath_finish(false);
        Path p;
        p.segments = std::move(segments);
        p.subpaths = std::move(subpaths);
        p.closed = std::move(closed);
        *this = PathBuilder();
        return p;
    }
    PathBuilder& move_to(Point p) { subpath_finish(false); position = p; return *this; }  // :882-886
    PathBuilder& close() { subpath_finish(true); return *this; }                           // :889-892
    PathBuilder& line_to(Point p) {                                                        // :895-903
        if (!position.is_close_to(p)) {
            segments.push_back(Segment::line(position, p));
            position = p;
        }
        return *this;
    }
    PathBuilder& quad_to(Point p1, Point p2) {  // :906-911
        segments.push_back(Segment::quad(position, p1, p2));
        position = p2;
        return *this;
    }
    PathBuilder& cubic_to(Point p1, Point p2, Point p3) {  // :923-933
        segments.push_back(Segment::cubic(position, p1, p2, p3));
        position = p3;
        return *this;
    }
    PathBuilder& arc_to(Point radii, Scalar x_axis_rot, bool large, bool sweep, Point p) {  // :945-972
        auto arc = EllipArc::new_param(position, p, radii.x, radii.y, x_axis_rot, large, sweep);
        if (!arc) return line_to(p);
        auto cs = arc->to_cubics();
        segments.insert(segments.end(), cs.begin(), cs.end());
        position = p;
        return *this;
    }
    PathBuilder& circle(Scalar radius) {  // :977-996
        Scalar offset = 0.5522847498307935 * radius;
        Point x_offset(offset, 0.0), y_offset(0.0, offset);
        Point center = position;
        Point p0 = center - Point(radius, 0.0);
        Point p1 = center - Point(0.0, radius);
        Point p2 = center + Point(radius, 0.0);
        Point p3 = center + Point(0.0, radius);
        move_to(p0);
        cubic_to(p0 - y_offset, p1 - x_offset, p1);
        cubic_to(p1 + x_offset, p2 - y_offset, p2);
        cubic_to(p2 + y_offset, p3 + x_offset, p3);
        cubic_to(p3 - x_offset, p0 + y_offset, p0);
        close();
        return move_to(center);
    }
    PathBuilder& checkerboard(const BBox& bbox, Scalar cell_size) {  // :1031-1050
        Scalar x = bbox.x(), y = bbox.y();
        while (y < bbox.max.y) {
            while (x < bbox.max.x) {
                Point offset(x, y);
                move_to(offset);
                line_to(offset + Point(cell_size, 0.0));
                line_to(offset + Point(cell_size, 2.0 * cell_size));
                line_to(offset + Point(2.0 * cell_size, 2.0 * cell_size));
                line_to(offset + Point(2.0 * cell_size, cell_size));
                line_to(offset + Point(0.0, cell_size));
                close();
                x += 2.0 * cell_size;
            }
            x = bbox.x();
            y += 2.0 * cell_size;
        }
        return move_to(bbox.min);
    }
};

// ---- byte parser: src/svg.rs:62-236 ------------------------------------------------------
struct ByteParser {
    const uint8_t* data;
    size_t len, pos = 0;
    ByteParser(const char* s, size_t n) : data((const uint8_t*)s), len(n) {}
    int peek() const { return pos < len ? data[pos] : -1; }
    int next() { return pos < len ? data[pos++] : -1; }
    void separators() {  // :150-162
        while (pos < len) {
            uint8_t b = data[pos];
            if (b == ' ' || b == '\t' || b == '\r' || b == '\n' || b == ',') pos++; else break;
        }
    }
    // :165-235 — value = (i64 mantissa as f64) * powi(10, exponent); NOT strtod.
    bool try_scalar(Scalar& out) {
        separators();
        uint64_t mantissa = 0;  // wrapping arithmetic as in the reference (i64 wrapping_mul/add)
        int64_t exponent = 0;
        int64_t sign = 1;
        int c = peek();
        if (c == '-' || c == '+') { if (c == '-') sign = -1; pos++; }
        size_t whole = 0, frac = 0;
        while (pos < len && data[pos] >= '0' && data[pos] <= '9') { mantissa = mantissa * 10 + (uint64_t)(data[pos] - '0'); pos++; whole++; }
        if (peek() == '.') {
            pos++;
            while (pos < len && data[pos] >= '0' && data[pos] <= '9') {
                mantissa = mantissa * 10 + (uint64_t)(data[pos] - '0'); pos++; frac++; exponent -= 1;
            }
        }
        int64_t m = (int64_t)mantissa * sign;
        if (whole + frac == 0) return false;
        c = peek();
        if (c == 'e' || c == 'E') {
            pos++;
            int64_t sci = 0, sci_sign = 1;
            c = peek();
            if (c == '-' || c == '+') { if (c == '-') sci_sign = -1; pos++; }
            size_t nd = 0;
            while (pos < len && data[pos] >= '0' && data[pos] <= '9') { sci = sci * 10 + (data[pos] - '0'); pos++; nd++; }
            if (nd == 0) return false;
            exponent = exponent + sci_sign * sci;
        }
        out = (Scalar)m * powi(10.0, (int)(int32_t)exponent);
        return true;
    }
    Scalar scalar() {
        Scalar v;
        if (!try_scalar(v)) throw ParseError("InvalidScalar at offset " + std::to_string(pos));
        return v;
    }
};

// ---- SvgPathParser: src/svg.rs:241-421, applied straight to a PathBuilder (SvgPathCmd::apply :43-59)
inline void parse_svg_path(const char* text, size_t n, PathBuilder& builder) {
    ByteParser ps(text, n);
    int prev_op = -1;
    enum PrevKind { None, QuadTo, CubicTo, Other } prev_kind = None;
    Point prev_c1, prev_c2;  // for QuadTo: (p1,p2); for CubicTo: (p2,p3)
    Point position(0.0, 0.0), subpath_start(0.0, 0.0);

    auto parse_point = [&]() -> Point {  // :265-271
        Scalar x = ps.scalar();
        Scalar y = ps.scalar();
        Point p(x, y);
        if (prev_op >= 0 && prev_op >= 'a' && prev_op <= 'z') return p + position;
        return p;
    };
    auto parse_flag = [&]() -> bool {  // :274-289
        ps.separators();
        int b = ps.peek();
        if (b == '0') { ps.pos++; return false; }
        if (b == '1') { ps.pos++; return true; }
        throw ParseError("InvalidFlag at offset " + std::to_string(ps.pos));
    };

    while (true) {
        ps.separators();
        int op = ps.next();
        if (op < 0) break;
        if (std::strchr("MmLlVvHhCcSsQqTtAaZz", op) != nullptr && op != 0) {  // :297-310
            if (op == 'm') prev_op = 'l';
            else if (op == 'M') prev_op = 'L';
            else if (op == 'Z' || op == 'z') prev_op = -1;
            else prev_op = op;
        } else {  // :311-320 implicit repeat of the previous op
            ps.pos--;
            if (prev_op < 0) throw ParseError("InvalidCmd at offset " + std::to_string(ps.pos));
            op = prev_op;
        }
        Point dst;
        PrevKind kind = Other;
        Point k1, k2;
        switch (op) {
            case 'M': case 'm': {
                dst = parse_point();
                subpath_start = dst;
                builder.move_to(dst);
                break;
            }
            case 'L': case 'l': dst = parse_point(); builder.line_to(dst); break;
            case 'V': case 'v': {
                Scalar y = ps.scalar();
                dst = (op == 'v') ? Point(position.x, position.y + y) : Point(position.x, y);
                builder.line_to(dst);
                break;
            }
            case 'H': case 'h': {
                Scalar x = ps.scalar();
                dst = (op == 'h') ? Point(position.x + x, position.y) : Point(x, position.y);
                builder.line_to(dst);
                break;
            }
            case 'Q': case 'q': {
                Point p1 = parse_point();
                Point p2 = parse_point();
                builder.quad_to(p1, p2);
                dst = p2; kind = QuadTo; k1 = p1; k2 = p2;
                break;
            }
            case 'T': case 't': {
                Point p1 = (prev_kind == QuadTo) ? (2.0 * prev_c2 - prev_c1) : position;
                Point p2 = parse_point();
                builder.quad_to(p1, p2);
                dst = p2; kind = QuadTo; k1 = p1; k2 = p2;
                break;
            }
            case 'C': case 'c': {
                Point p1 = parse_point();
                Point p2 = parse_point();
                Point p3 = parse_point();
                builder.cubic_to(p1, p2, p3);
                dst = p3; kind = CubicTo; k1 = p2; k2 = p3;
                break;
            }
            case 'S': case 's': {
                Point p1 = (prev_kind == CubicTo) ? (2.0 * prev_c2 - prev_c1) : position;
                Point p2 = parse_point();
                Point p3 = parse_point();
                builder.cubic_to(p1, p2, p3);
                dst = p3; kind = CubicTo; k1 = p2; k2 = p3;
                break;
            }
            case 'A': case 'a': {
                Scalar rx = ps.scalar();
                Scalar ry = ps.scalar();
                Scalar rot = ps.scalar();
                bool large = parse_flag();
                bool sweep = parse_flag();
                dst = parse_point();
                builder.arc_to(Point(rx, ry), rot, large, sweep, dst);
                break;
            }
            case 'Z': case 'z': dst = subpath_start; builder.close(); break;
            default: throw ParseError("unreachable");
        }
        position = dst;  // :409
        prev_kind = kind; prev_c1 = k1; prev_c2 = k2;
    }
}

inline Path path_from_svg(const char* text, size_t n) {  // Path::read_svg_path, src/path.rs:528-534
    PathBuilder b;
    parse_svg_path(text, n, b);
    return b.build();
}

// ---- transform strings: src/svg.rs:423-602 ----------------------------------------------
inline Transform parse_transform(const char* text, size_t n) {
    ByteParser ps(text, n);
    Transform tr = Transform::identity();
    auto ident = [&]() -> std::string {
        std::string s;
        while (ps.pos < ps.len && std::isalpha(ps.data[ps.pos])) s.push_back((char)ps.data[ps.pos++]);
        return s;
    };
    auto angle = [&]() -> Scalar {  // :455-467
        Scalar v = ps.scalar();
        std::string u = ident();
        if (u.empty() || u == "deg") return v * PI / 180.0;
        if (u == "rad") return v;
        throw ParseError("InvalidUnits " + u);
    };
    auto try_length = [&](Scalar& out) -> bool {  // :469-474
        if (!ps.try_scalar(out)) return false;
        ident();
        return true;
    };
    auto length = [&]() -> Scalar {
        Scalar v;
        if (!try_length(v)) throw ParseError("InvalidScalar at offset " + std::to_string(ps.pos));
        return v;
    };
    while (true) {
        ps.separators();
        if (ps.peek() < 0) break;
        std::string op = ident();
        ps.separators();
        if (ps.next() != '(') throw ParseError("BracketExpected");
        Transform t;
        if (op == "matrix") {
            Scalar m00 = ps.scalar(), m10 = ps.scalar(), m01 = ps.scalar(), m11 = ps.scalar(), m02 = ps.scalar(), m12 = ps.scalar();
            t = Transform(m00, m01, m02, m10, m11, m12);
        } else if (op == "rotate") {
            t = Transform::new_rotate(angle());
            Scalar tx;
            if (try_length(tx)) {
                Scalar ty = length();
                t = Transform::new_translate(tx, ty).pre_concat(t).pre_translate(-tx, -ty);
            }
        } else if (op == "translate") {
            Scalar tx = length();
            Scalar ty;
            if (!try_length(ty)) ty = 0.0;
            t = Transform::new_translate(tx, ty);
        } else if (op == "translateX") {
            t = Transform::new_translate(length(), 0.0);
        } else if (op == "translateY") {
            t = Transform::new_translate(0.0, length());
        } else if (op == "scale") {
            Scalar sx = ps.scalar();
            Scalar sy;
            if (!ps.try_scalar(sy)) sy = sx;
            t = Transform::new_scale(sx, sy);
        } else if (op == "scaleX") {
            t = Transform::new_scale(ps.scalar(), 1.0);
        } else if (op == "scaleY") {
            t = Transform::new_scale(1.0, ps.scalar());
        } else if (op == "skewX") {
            t = Transform::new_skew(angle(), 0.0);
        } else if (op == "skewY") {
            t = Transform::new_skew(0.0, angle());
        } else {
            throw ParseError("InvalidTransformOp " + op);
        }
        ps.separators();
        if (ps.next() != ')') throw ParseError("BracketExpected");
        tr = tr * t;  // :597-599
    }
    return tr;
}

}  // namespace orc
