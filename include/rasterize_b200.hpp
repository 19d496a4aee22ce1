// rasterize_b200.hpp — C++17 host-side mirror of the reference's interface for the fill path, header-only, on
// top of the C ABI in rasterize_b200.h.  The reference is compiled code (Rust); with no Rust toolchain in this
// image the same seam is stated in C++: same names, argument meaning and error behaviour
// (aslpavel/rasterize v0.6.7, paths relative to the crate):
//
//   Point, Transform            src/geometry.rs:103, 317-539
//   FillRule, Size              src/path.rs:21-29, src/rasterize.rs:38-41
//   Path, PathBuilder           src/path.rs:227-233, 800-1056   (flat storage: this is what crosses the FFI)
//   Pixel, Rasterizer           src/rasterize.rs:44-101          (name, mask, mask_iter, fill)
//   LinColor, GradLinear/Radial src/color.rs:268-374, src/grad.rs:150-226, 307-426 (as paint descriptions)
//   GpuRasterizer               the new implementor of Rasterizer (north star)
//
// Errors: a non-zero status from the C ABI throws rasterize::Error (the Rust shim panics, as the reference does
// on NaN input, src/path.rs:765-767).  There is no CPU fallback.
#pragma once
#include "rasterize_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

namespace rasterize {

using Scalar = double;
constexpr Scalar EPSILON = std::numeric_limits<double>::epsilon();
constexpr Scalar DEFAULT_FLATNESS = 0.05;  // src/path.rs:16

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("rasterize_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};

struct Point {
    Scalar x = 0, y = 0;
};

struct Size {
    size_t width = 0, height = 0;
};

enum class FillRule : int { NonZero = RGPU_NONZERO, EvenOdd = RGPU_EVENODD };

struct Pixel {
    size_t x, y;
    Scalar alpha;
};

// 2x3 affine [m00, m01, m02, m10, m11, m12]
struct Transform {
    Scalar m[6] = {1, 0, 0, 0, 1, 0};
    static Transform identity() { return {}; }
    static Transform new_translate(Scalar tx, Scalar ty) { Transform t; t.m[2] = tx; t.m[5] = ty; return t; }
    static Transform new_scale(Scalar sx, Scalar sy) { Transform t; t.m[0] = sx; t.m[4] = sy; return t; }
    static Transform new_rotate(Scalar a) {
        Transform t;
        const Scalar s = std::sin(a), c = std::cos(a);
        t.m[0] = c; t.m[1] = -s; t.m[3] = s; t.m[4] = c;
        return t;
    }
    Point apply(Point p) const { return {p.x * m[0] + p.y * m[1] + m[2], p.x * m[3] + p.y * m[4] + m[5]}; }
    Transform operator*(const Transform& o) const {  // src/geometry.rs:519-539
        Transform r;
        r.m[0] = m[0] * o.m[0] + m[1] * o.m[3];
        r.m[1] = m[0] * o.m[1] + m[1] * o.m[4];
        r.m[2] = m[0] * o.m[2] + m[1] * o.m[5] + m[2];
        r.m[3] = m[3] * o.m[0] + m[4] * o.m[3];
        r.m[4] = m[3] * o.m[1] + m[4] * o.m[4];
        r.m[5] = m[3] * o.m[2] + m[4] * o.m[5] + m[5];
        return r;
    }
    Transform pre_translate(Scalar tx, Scalar ty) const { return *this * new_translate(tx, ty); }
    Transform pre_scale(Scalar sx, Scalar sy) const { return *this * new_scale(sx, sy); }
    Transform pre_rotate(Scalar a) const { return *this * new_rotate(a); }
};

// `Shape`, src/image.rs:6-17, over caller-owned memory
template <class P>
struct ImageMut {
    P* data;
    rgpu_shape shape;
    static ImageMut dense(P* data, size_t height, size_t width) { return {data, rgpu_shape{0, width, height, width, 1}}; }
    size_t width() const { return shape.width; }
    size_t height() const { return shape.height; }
    P& at(size_t row, size_t col) { return data[shape.start + row * shape.row_stride + col * shape.col_stride]; }
};

class PathBuilder;

// Flat `Path { segments, subpaths, closed }`
class Path {
public:
    std::vector<Scalar> points;          // x,y pairs, segments back to back
    std::vector<uint8_t> kinds;          // 2 = Line, 3 = Quad, 4 = Cubic
    std::vector<uint32_t> subpath_offsets;
    std::vector<uint8_t> closed;

    static PathBuilder builder();
    bool is_empty() const { return closed.empty(); }
    size_t segments_count() const { return kinds.size(); }
    rgpu_path ffi() const {
        rgpu_path p;
        p.points = points.data();
        p.kinds = kinds.data();
        p.subpath_offsets = closed.empty() ? nullptr : subpath_offsets.data();
        p.closed = closed.data();
        p.n_points = (uint32_t)(points.size() / 2);
        p.n_segments = (uint32_t)kinds.size();
        p.n_subpaths = (uint32_t)closed.size();
        return p;
    }
};

// `LineJoin` (src/path.rs:78-93), `LineCap` (:103-110), `StrokeStyle` (:121-136)
enum class LineJoin : int { Miter = RGPU_JOIN_MITER, Bevel = RGPU_JOIN_BEVEL, Round = RGPU_JOIN_ROUND };
enum class LineCap : int { Butt = RGPU_CAP_BUTT, Square = RGPU_CAP_SQUARE, Round = RGPU_CAP_ROUND };
struct StrokeStyle {
    Scalar width = 1.0;
    LineJoin line_join = LineJoin::Miter;
    Scalar miter_limit = 4.0;  // `LineJoin::Miter(limit)`, default 4.0
    LineCap line_cap = LineCap::Butt;
    rgpu_stroke_style ffi() const {
        rgpu_stroke_style s;
        s.width = width;
        s.miter_limit = miter_limit;
        s.line_join = (int32_t)line_join;
        s.line_cap = (int32_t)line_cap;
        return s;
    }
};

// Elliptic arc in centre form and its conversion into cubics of at most a quarter turn: `EllipArc::new_param`
// (src/ellipse.rs:40-96), `EllipArcCubicIter` (src/ellipse.rs:167-214), `Point::angle_between` (src/geometry.rs:186-203).
// A `Path` only ever stores lines, quads and cubics (src/curve.rs:905-909): arcs are converted here, on the host, when the
// path is built — the flatten kernel never sees one.  IEEE semantics are kept as in Rust: a zero sweep gives a
// NaN step and no cubic at all.
struct EllipArc {
    Point center;
    Scalar rx, ry, phi, eta, eta_delta;
    static bool angle_between(Point a, Point b, Scalar& out) {
        const Scalar lengths = std::hypot(a.x, a.y) * std::hypot(b.x, b.y);
        if (lengths < EPSILON) return false;
        Scalar c = (a.x * b.x + a.y * b.y) / lengths;
        c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
        const Scalar angle = std::acos(c);
        out = (a.x * b.y - a.y * b.x < 0.0) ? -angle : angle;
        return true;
    }
    static Scalar rem_euclid(Scalar x, Scalar rhs) {
        const Scalar r = std::fmod(x, rhs);
        return r < 0.0 ? r + std::fabs(rhs) : r;
    }
    static bool new_param(Point src, Point dst, Scalar rx, Scalar ry, Scalar x_axis_rot, bool large_flag, bool sweep_flag, EllipArc& out) {
        constexpr Scalar PI = 3.14159265358979323846264338327950288;
        rx = std::fabs(rx);
        ry = std::fabs(ry);
        const Scalar phi = x_axis_rot * PI / 180.0;
        const Point p1 = Transform::new_rotate(-phi).apply({0.5 * (src.x - dst.x), 0.5 * (src.y - dst.y)});
        const Scalar x1 = p1.x, y1 = p1.y;
        const Scalar ax = x1 / rx, ay = y1 / ry;
        const Scalar s = ax * ax + ay * ay;
        if (s > 1.0) {
            const Scalar sq = std::sqrt(s);
            rx = rx * sq;
            ry = ry * sq;
        }
        const Scalar rxry = rx * ry, rxy1 = rx * y1, ryx1 = ry * x1;
        Scalar q = rxry * rxry / (rxy1 * rxy1 + ryx1 * ryx1) - 1.0;
        q = (q > 0.0 || q != q) ? q : 0.0;  // f64::max(0.0): NaN.max(0.0) is 0.0 in Rust
        if (q != q) q = 0.0;
        Scalar sq = std::sqrt(q);
        sq = (large_flag == sweep_flag) ? -sq : sq;
        const Scalar cx = sq * (rx * y1 / ry), cy = sq * (-ry * x1 / rx);
        const Point rc = Transform::new_rotate(phi).apply({cx, cy});
        const Point center{rc.x + 0.5 * (dst.x + src.x), rc.y + 0.5 * (dst.y + src.y)};
        Scalar eta, ed;
        if (!angle_between({1.0, 0.0}, {(x1 - cx) / rx, (y1 - cy) / ry}, eta)) return false;
        if (!angle_between({(x1 - cx) / rx, (y1 - cy) / ry}, {(-x1 - cx) / rx, (-y1 - cy) / ry}, ed)) return false;
        // NaN angles (coincident end points, a zero radius): the reference's cubic iterator then never terminates
        // (`segment_index > NaN` is false for ever, src/ellipse.rs:198-201); treated like the degenerate arcs it does detect
        if (eta != eta || ed != ed) return false;
        Scalar eta_delta = rem_euclid(ed, 2.0 * PI);
        if (!sweep_flag && eta_delta > 0.0) eta_delta = eta_delta - 2.0 * PI;
        else if (sweep_flag && eta_delta < 0.0) eta_delta = eta_delta + 2.0 * PI;
        out = EllipArc{center, rx, ry, phi, eta, eta_delta};
        return true;
    }
    template <class F>  // f(p0, p1, p2, p3) for every cubic
    void to_cubics(F f) const {
        constexpr Scalar PI = 3.14159265358979323846264338327950288;
        const Transform phi_tr = Transform::new_rotate(phi);
        Scalar segment_count = std::ceil(std::fabs(eta_delta) / (PI / 2.0));
        const Scalar segment_delta = eta_delta / segment_count;
        Scalar segment_index = 0.0;
        segment_count = segment_count - 1.0;
        auto at = [&](Scalar alpha, Point& a, Point& d) {
            const Scalar sn = std::sin(alpha), cs = std::cos(alpha);
            const Point t = phi_tr.apply({rx * cs, ry * sn});
            a = {t.x + center.x, t.y + center.y};
            d = phi_tr.apply({-rx * sn, ry * cs});
        };
        while (!(segment_index > segment_count)) {
            const Scalar eta_1 = eta + segment_delta * segment_index;
            const Scalar eta_2 = eta_1 + segment_delta;
            segment_index += 1.0;
            const Scalar tn = std::tan((eta_2 - eta_1) / 2.0);
            const Scalar sq = std::sqrt(4.0 + 3.0 * (tn * tn));
            const Scalar alpha = std::sin(eta_2 - eta_1) * (sq - 1.0) / 3.0;
            Point p0, d0, p3, d3;
            at(eta_1, p0, d0);
            at(eta_2, p3, d3);
            f(p0, Point{p0.x + alpha * d0.x, p0.y + alpha * d0.y}, Point{p3.x - alpha * d3.x, p3.y - alpha * d3.y}, p3);
        }
    }
};

// `PathBuilder` (move_to / line_to / quad_to / cubic_to / arc_to / close / build), src/path.rs:813-972.
class PathBuilder {
public:
    // `PathBuilder::arc_to` (src/path.rs:945-972): SVG endpoint arc -> cubics stored in the path; a degenerate arc is a line
    PathBuilder& arc_to(Point radii, Scalar x_axis_rot, bool large, bool sweep, Point p) {
        EllipArc arc;
        if (!EllipArc::new_param(pos_, p, radii.x, radii.y, x_axis_rot, large, sweep, arc)) return line_to(p);
        arc.to_cubics([&](Point p0, Point p1, Point p2, Point p3) {
            push(p0); push(p1); push(p2); push(p3);
            path_.kinds.push_back(4);
        });
        pos_ = p;
        return *this;
    }
    PathBuilder& move_to(Point p) { finish(false); pos_ = p; return *this; }
    PathBuilder& close() { finish(true); return *this; }
    PathBuilder& line_to(Point p) {
        if (!(std::fabs(pos_.x - p.x) < EPSILON && std::fabs(pos_.y - p.y) < EPSILON)) {  // src/path.rs:895-903
            push(pos_); push(p);
            path_.kinds.push_back(2);
            pos_ = p;
        }
        return *this;
    }
    PathBuilder& quad_to(Point p1, Point p2) { push(pos_); push(p1); push(p2); path_.kinds.push_back(3); pos_ = p2; return *this; }
    PathBuilder& cubic_to(Point p1, Point p2, Point p3) {
        push(pos_); push(p1); push(p2); push(p3);
        path_.kinds.push_back(4);
        pos_ = p3;
        return *this;
    }
    Path build() {
        finish(false);
        Path out = std::move(path_);
        *this = PathBuilder();
        return out;
    }

private:
    void push(Point p) { path_.points.push_back(p.x); path_.points.push_back(p.y); }
    void finish(bool close) {  // subpath_finish, src/path.rs:849-868
        const uint32_t n = (uint32_t)path_.kinds.size();
        if (n == 0 || (!path_.subpath_offsets.empty() && path_.subpath_offsets.back() == n)) return;
        if (path_.subpath_offsets.empty()) path_.subpath_offsets.push_back(0);
        if (close) {
            size_t first = path_.subpath_offsets.back(), off = 0;
            for (size_t i = 0; i < first; i++) off += path_.kinds[i];
            pos_ = {path_.points[2 * off], path_.points[2 * off + 1]};
        }
        path_.subpath_offsets.push_back(n);
        path_.closed.push_back(close ? 1 : 0);
    }
    Path path_;
    Point pos_{0, 0};
};
inline PathBuilder Path::builder() { return PathBuilder(); }

// Paint description handed to `Rasterizer::fill` (the `gpu_desc` hook of INTEGRATION.md §3)
struct Paint {
    rgpu_paint desc{};
    std::vector<double> stop_pos;
    std::vector<float> stop_colors;
    static Paint solid(float r, float g, float b, float a) {
        Paint p;
        p.desc.kind = RGPU_PAINT_SOLID;
        p.desc.tr[0] = p.desc.tr[4] = 1.0;
        p.desc.solid[0] = r; p.desc.solid[1] = g; p.desc.solid[2] = b; p.desc.solid[3] = a;
        return p;
    }
    const rgpu_paint* ffi() {
        desc.n_stops = (uint32_t)stop_pos.size();
        desc.stop_pos = stop_pos.data();
        desc.stop_colors = stop_colors.data();
        return &desc;
    }
};

// `pub trait Rasterizer`, src/rasterize.rs:44-101
class Rasterizer {
public:
    virtual ~Rasterizer() = default;
    virtual const char* name() const = 0;
    virtual void mask(const Path& path, Transform tr, ImageMut<Scalar> img, FillRule fill_rule) = 0;
    virtual std::vector<Pixel> mask_iter(const Path& path, Transform tr, Size size, FillRule fill_rule) = 0;
    virtual void fill(const Path& path, Transform tr, FillRule fill_rule, Paint& paint, ImageMut<float> img, const double* path_bbox = nullptr) = 0;
};

class GpuRasterizer final : public Rasterizer {
public:
    explicit GpuRasterizer(Scalar flatness = DEFAULT_FLATNESS, int device = 0) {
        const int rc = rgpu_create(device, flatness, &ctx_);
        if (rc != RGPU_OK) throw Error(rc, rgpu_last_error(nullptr));
    }
    ~GpuRasterizer() override { rgpu_destroy(ctx_); }
    GpuRasterizer(const GpuRasterizer&) = delete;
    GpuRasterizer& operator=(const GpuRasterizer&) = delete;

    const char* name() const override { return rgpu_name(); }

    void mask(const Path& path, Transform tr, ImageMut<Scalar> img, FillRule fill_rule) override {
        const rgpu_path p = path.ffi();
        check(rgpu_mask(ctx_, &p, tr.m, (int)fill_rule, img.data, img.shape));
    }
    std::vector<Pixel> mask_iter(const Path& path, Transform tr, Size size, FillRule fill_rule) override {
        const rgpu_path p = path.ffi();
        std::vector<rgpu_pixel> buf(size.width * size.height ? size.width * size.height : 1);
        size_t n = 0;
        check(rgpu_mask_iter(ctx_, &p, tr.m, size.width, size.height, (int)fill_rule, buf.data(), buf.size(), &n));
        std::vector<Pixel> out(n);
        for (size_t i = 0; i < n; i++) out[i] = {buf[i].x, buf[i].y, buf[i].alpha};
        return out;
    }
    void fill(const Path& path, Transform tr, FillRule fill_rule, Paint& paint, ImageMut<float> img, const double* path_bbox = nullptr) override {
        const rgpu_path p = path.ffi();
        check(rgpu_fill(ctx_, &p, tr.m, (int)fill_rule, paint.ffi(), path_bbox, img.data, img.shape));
    }
    // `Path::flatten(tr, flatness, close)`: (x0,y0,x1,y1) per line, reference order
    std::vector<double> flatten(const Path& path, Transform tr, bool close = true) {
        const rgpu_path p = path.ffi();
        std::vector<double> lines(4 * (path.segments_count() * 32 + 64));
        size_t n = 0;
        int rc = rgpu_flatten(ctx_, &p, tr.m, close, lines.data(), lines.size() / 4, &n);
        if (rc == RGPU_ERR_CAPACITY) {
            lines.resize(4 * n);
            rc = rgpu_flatten(ctx_, &p, tr.m, close, lines.data(), n, &n);
        }
        check(rc);
        lines.resize(4 * n);
        return lines;
    }
    // `Path::stroke(style)` (src/path.rs:374-415), computed on the device and returned as a host `Path`.  Callers that go
    // on to rasterize the outline keep it on the device instead: rgpu_path_stroke hands back an `rgpu_dpath`.
    Path stroke(const Path& path, const StrokeStyle& style) {
        const rgpu_path p = path.ffi();
        const rgpu_stroke_style st = style.ffi();
        rgpu_dpath* dp = nullptr;
        check(rgpu_path_stroke(ctx_, &p, &st, &dp));
        uint32_t n_pts = 0, n_seg = 0, n_sub = 0;
        rgpu_dpath_info(dp, &n_pts, &n_seg, &n_sub);
        Path out;
        out.points.resize(2 * (size_t)n_pts);
        out.kinds.resize(n_seg);
        out.subpath_offsets.resize((size_t)n_sub + 1);
        out.closed.resize(n_sub);
        const int rc = rgpu_dpath_download(ctx_, dp, out.points.data(), out.kinds.data(), out.subpath_offsets.data(), out.closed.data());
        rgpu_path_free(ctx_, dp);
        check(rc);
        return out;
    }
    // Batch `str::parse::<Path>()` + `Path::bbox` + `fit_size` on the device (src/svg.rs:241-421, src/path.rs:428-451,
    // src/geometry.rs:490-516): one host call for n strings.  `info[i]` carries the bbox, the parse status and, when
    // fit_align >= 0, the fitted size and transform; a string that does not parse gives an empty path.
    struct ParsedBatch {
        std::vector<Path> paths;
        std::vector<rgpu_parse_info> info;
    };
    ParsedBatch parse_svg_batch(const std::vector<std::string>& strings, uint32_t fit_width = 0, uint32_t fit_height = 0, int fit_align = -1) {
        std::string text;
        std::vector<uint32_t> off(1, 0u);
        for (const std::string& s : strings) {
            text += s;
            off.push_back((uint32_t)text.size());
        }
        ParsedBatch out;
        out.info.resize(strings.size());
        rgpu_parse_options opt{fit_width, fit_height, fit_align};
        rgpu_dpath_batch* b = nullptr;
        check(rgpu_parse_svg_batch(ctx_, text.data(), off.data(), strings.size(), &opt, &b, out.info.data()));
        size_t n = 0;
        uint32_t n_pts = 0, n_seg = 0, n_sub = 0;
        rgpu_path_batch_info(b, &n, &n_pts, &n_seg, &n_sub);
        std::vector<double> pts(2 * (size_t)n_pts);
        std::vector<uint8_t> kinds(n_seg), closed(n_sub);
        std::vector<uint32_t> sp((size_t)n_sub + 1), psp(n + 1);
        const int rc = rgpu_path_batch_download(ctx_, b, pts.data(), kinds.data(), sp.data(), closed.data(), psp.data());
        rgpu_path_batch_free(ctx_, b);
        check(rc);
        size_t pt = 0;
        out.paths.resize(n);
        for (size_t i = 0; i < n; i++) {
            Path& p = out.paths[i];
            const uint32_t s0 = psp[i], s1 = psp[i + 1];
            if (s1 == s0) continue;
            const uint32_t k0 = sp[s0], k1 = sp[s1];
            size_t np = 0;
            for (uint32_t k = k0; k < k1; k++) np += kinds[k];
            p.points.assign(pts.begin() + 2 * pt, pts.begin() + 2 * (pt + np));
            p.kinds.assign(kinds.begin() + k0, kinds.begin() + k1);
            for (uint32_t s = s0; s <= s1; s++) p.subpath_offsets.push_back(sp[s] - k0);
            p.closed.assign(closed.begin() + s0, closed.begin() + s1);
            pt += np;
        }
        return out;
    }
    // One Fill node as the Fill arm of `Pipeline::render_rec` sets it up (src/scene.rs:407-430): `tr` = align * node
    // transform, (x, y, width, height) = the `view_mut` window of the layer.
    struct SceneFill {
        const Path* path;
        Transform tr;
        FillRule fill_rule;
        Paint* paint;
        uint32_t x, y, width, height;
        const double* path_bbox = nullptr;
    };
    // `Scene::render` of a Fill-only pipeline (src/scene.rs:186-199, 384-435) + the CLI's RGBA8 export in ONE call: a
    // width x height layer created with `bg` (nullptr = transparent), fills blended in order by the scene compositor.
    // Returns RGBA8; `lin` (optional) receives the LinColor layer (4 floats per pixel).
    std::vector<uint8_t> render_scene(const std::vector<SceneFill>& fills, size_t width, size_t height, const float* bg = nullptr,
                                      std::vector<float>* lin = nullptr) {
        std::vector<rgpu_path> paths(fills.size());
        std::vector<rgpu_scene_fill> ffi(fills.size());
        for (size_t i = 0; i < fills.size(); i++) {
            paths[i] = fills[i].path->ffi();
            ffi[i] = rgpu_scene_fill{};
            ffi[i].path = &paths[i];
            for (int k = 0; k < 6; k++) ffi[i].tr[k] = fills[i].tr.m[k];
            ffi[i].fill_rule = (int32_t)fills[i].fill_rule;
            ffi[i].paint = fills[i].paint->ffi();
            ffi[i].path_bbox = fills[i].path_bbox;
            ffi[i].x = fills[i].x; ffi[i].y = fills[i].y; ffi[i].width = fills[i].width; ffi[i].height = fills[i].height;
        }
        std::vector<uint8_t> rgba(width * height * 4);
        if (lin) lin->assign(width * height * 4, 0.0f);
        check(rgpu_render_scene_host(ctx_, ffi.data(), ffi.size(), width, height, bg, lin ? lin->data() : nullptr, rgba.data()));
        return rgba;
    }
    rgpu_ctx* raw() { return ctx_; }
    void check(int rc) const {
        if (rc != RGPU_OK) throw Error(rc, rgpu_last_error(ctx_));
    }

private:
    rgpu_ctx* ctx_ = nullptr;
};

// `Layer<C>` (src/scene.rs:464-501) in device memory: integer origin, dense row-major pixels; C = LinColor (4 floats) or
// Scalar coverage (1 float, the clip mask of the Clip arm).  The methods are the arms of `Pipeline::render_rec`
// (src/scene.rs:397-459) with every pixel operation on the GPU; only `download*` crosses PCIe.
class DeviceLayer {
public:
    DeviceLayer(GpuRasterizer& r, int32_t x, int32_t y, size_t width, size_t height, int channels, const float* color = nullptr)
        : r_(r), x_(x), y_(y), w_(width), h_(height), ch_(channels) {
        r_.check(rgpu_device_alloc(r_.raw(), std::max<size_t>(1, w_ * h_ * ch_ * sizeof(float)), &ptr_));
        if (color && ch_ == 4) r_.check(rgpu_fill_color_dev(r_.raw(), data(), w_ * h_, color));
        else r_.check(rgpu_device_zero(r_.raw(), ptr_, w_ * h_ * ch_ * sizeof(float)));
    }
    ~DeviceLayer() { rgpu_device_free(r_.raw(), ptr_); }
    DeviceLayer(const DeviceLayer&) = delete;
    DeviceLayer& operator=(const DeviceLayer&) = delete;
    float* data() const { return static_cast<float*>(ptr_); }
    size_t width() const { return w_; }
    size_t height() const { return h_; }

    // Fill arm: `path.fill(rasterizer, align * tr, fill_rule, paint, layer.view_mut(..))` on the whole layer
    void fill(const Path& path, Transform tr, FillRule rule, Paint& paint) { job(path, tr, rule, RGPU_JOB_FILL, &paint); }
    // the clip mask of the Clip arm: `path.mask(rasterizer, align * clip_tr, fill_rule, mask_layer)`
    void mask(const Path& path, Transform tr, FillRule rule) { job(path, tr, rule, RGPU_JOB_MASK, nullptr); }
    // `child_layer.compose(mask_layer, |dst, src| dst * src)` on the intersection rectangle (src/scene.rs:453-455)
    void scale_by_mask(const DeviceLayer& m) {
        Rect q;
        if (!intersect(m, q)) return;
        r_.check(rgpu_layer_scale_by_mask_dev(r_.raw(), data(), q.a, w_, m.data(), q.b, m.w_, q.w, q.h));
    }
    // `layer.compose(child_layer, |dst, src| dst.blend_over(src [* opacity]))` (src/scene.rs:436-457)
    void blend_over(const DeviceLayer& src, const float* opacity = nullptr) {
        Rect q;
        if (!intersect(src, q)) return;
        r_.check(rgpu_layer_blend_over_dev(r_.raw(), data(), q.a, w_, src.data(), q.b, src.w_, q.w, q.h, opacity != nullptr, opacity ? *opacity : 1.0f));
    }
    std::vector<float> download() const {
        std::vector<float> out(w_ * h_ * ch_);
        r_.check(rgpu_memcpy_d2h(r_.raw(), out.data(), ptr_, out.size() * sizeof(float)));
        return out;
    }
    std::vector<uint8_t> download_rgba8() const {  // `From<LinColor> for RGBA` on the device, 4 B/pixel over PCIe
        std::vector<uint8_t> out(w_ * h_ * 4);
        r_.check(rgpu_download_rgba8(r_.raw(), data(), w_ * h_, out.data()));
        return out;
    }

private:
    struct Rect { size_t a, b, w, h; };
    bool intersect(const DeviceLayer& o, Rect& q) const {  // src/scene.rs:540-549
        const int64_t x0 = std::max<int64_t>(x_, o.x_), x1 = std::min<int64_t>(x_ + (int64_t)w_, o.x_ + (int64_t)o.w_);
        const int64_t y0 = std::max<int64_t>(y_, o.y_), y1 = std::min<int64_t>(y_ + (int64_t)h_, o.y_ + (int64_t)o.h_);
        if (x1 <= x0 || y1 <= y0) return false;
        q = Rect{(size_t)((y0 - y_) * (int64_t)w_ + (x0 - x_)), (size_t)((y0 - o.y_) * (int64_t)o.w_ + (x0 - o.x_)), (size_t)(x1 - x0), (size_t)(y1 - y0)};
        return true;
    }
    void job(const Path& path, Transform tr, FillRule rule, int mode, Paint* paint) {
        if (w_ == 0 || h_ == 0) return;
        const rgpu_path p = path.ffi();
        rgpu_dpath* dp = nullptr;
        r_.check(rgpu_path_upload(r_.raw(), &p, &dp));
        rgpu_job j{};
        j.path = dp;
        // layer-local coordinates: `Transform::new_translate(-layer.x, -layer.y) * tr`
        const Transform t = Transform::new_translate(-(double)x_, -(double)y_) * tr;
        for (int i = 0; i < 6; i++) j.tr[i] = t.m[i];
        j.fill_rule = (int)rule;
        j.mode = mode;
        j.paint = paint ? paint->ffi() : nullptr;
        j.canvas = ptr_;
        j.row_stride = w_;
        j.width = (uint32_t)w_;
        j.height = (uint32_t)h_;
        const int rc = rgpu_render_batch_sync(r_.raw(), &j, 1, RGPU_BATCH_ORDERED);
        rgpu_path_free(r_.raw(), dp);
        r_.check(rc);
    }
    GpuRasterizer& r_;
    int32_t x_, y_;
    size_t w_, h_;
    int ch_;
    void* ptr_ = nullptr;
};

}  // namespace rasterize
