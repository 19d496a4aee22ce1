// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h).  C API glue over the orc_*.hpp restatement.
#include "oracle.h"
#include "orc_scene.hpp"
#include <thread>
#include <functional>
#include <cstring>

using namespace orc;

struct orc_path { Path p; };
struct orc_paint { Paint p; };
struct orc_scene {
    ScenePtr s;
    // keep-alive storage for orc_scene_fill_jobs
    std::vector<std::unique_ptr<orc_path>> job_paths;
    std::vector<std::unique_ptr<orc_paint>> job_paints;
};
struct orc_layer { Layer<LinColor> l; };

static thread_local std::string g_err;
static Transform TR(const double t[6]) { return Transform(t[0], t[1], t[2], t[3], t[4], t[5]); }
static void TR_out(const Transform& t, double out[6]) { for (int i = 0; i < 6; i++) out[i] = t.m[i]; }
static Shape SH(const orc_shape& s) { Shape r; r.start = s.start; r.width = s.width; r.height = s.height; r.row_stride = s.row_stride; r.col_stride = s.col_stride; return r; }
static BBox BB(const double mm[4]) { return BBox(Point(mm[0], mm[1]), Point(mm[2], mm[3])); }

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }
void orc_set_simd_x86(int on) { simd_x86() = on != 0; }

orc_path* orc_path_parse(const char* svg, size_t len) {
    try {
        auto* r = new orc_path();
        r->p = path_from_svg(svg, len);
        return r;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
orc_path* orc_path_from_flat(const double* pts, const uint8_t* kinds, size_t n_segs, const uint32_t* sub_off, size_t n_sub,
                             const uint8_t* closed) {
    auto* r = new orc_path();
    size_t pi = 0;
    for (size_t i = 0; i < n_segs; i++) {
        Segment s;
        s.kind = (SegKind)kinds[i];
        for (int k = 0; k < (int)kinds[i]; k++, pi++) s.p[k] = Point(pts[2 * pi], pts[2 * pi + 1]);
        r->p.segments.push_back(s);
    }
    if (n_sub > 0) {
        for (size_t i = 0; i <= n_sub; i++) r->p.subpaths.push_back(sub_off[i]);
        for (size_t i = 0; i < n_sub; i++) r->p.closed.push_back(closed[i]);
    }
    return r;
}
void orc_path_free(orc_path* p) { delete p; }
void orc_path_counts(const orc_path* p, size_t* n_segs, size_t* n_points, size_t* n_subpaths) {
    size_t np = 0;
    for (auto& s : p->p.segments) np += (size_t)s.npts();
    if (n_segs) *n_segs = p->p.segments.size();
    if (n_points) *n_points = np;
    if (n_subpaths) *n_subpaths = p->p.len();
}
void orc_path_export(const orc_path* p, double* pts, uint8_t* kinds, uint32_t* sub_off, uint8_t* closed) {
    size_t pi = 0;
    for (size_t i = 0; i < p->p.segments.size(); i++) {
        const Segment& s = p->p.segments[i];
        kinds[i] = (uint8_t)s.kind;
        for (int k = 0; k < s.npts(); k++, pi++) { pts[2 * pi] = s.p[k].x; pts[2 * pi + 1] = s.p[k].y; }
    }
    for (size_t i = 0; i < p->p.subpaths.size(); i++) sub_off[i] = (uint32_t)p->p.subpaths[i];
    for (size_t i = 0; i < p->p.closed.size(); i++) closed[i] = p->p.closed[i];
}
int orc_path_bbox(const orc_path* p, const double tr[6], double out[4]) {
    auto bb = p->p.bbox(TR(tr));
    if (!bb) return 0;
    out[0] = bb->min.x; out[1] = bb->min.y; out[2] = bb->max.x; out[3] = bb->max.y;
    return 1;
}
int orc_path_size(const orc_path* p, const double tr[6], size_t* w, size_t* h, double tr_out[6], double min_out[2]) {
    auto r = p->p.size(TR(tr));
    if (!r) return 0;
    *w = r->size.width; *h = r->size.height;
    TR_out(r->tr, tr_out);
    if (min_out) { min_out[0] = r->min.x; min_out[1] = r->min.y; }
    return 1;
}
orc_path* orc_path_stroke(const orc_path* p, double width, int join, double miter_limit, int cap) {
    StrokeStyle st;
    st.width = width;
    st.line_join = join == 0 ? LineJoin::Miter : (join == 1 ? LineJoin::Bevel : LineJoin::Round);
    st.miter_limit = miter_limit;
    st.line_cap = cap == 0 ? LineCap::Butt : (cap == 1 ? LineCap::Square : LineCap::Round);
    auto* r = new orc_path();
    r->p = p->p.stroke(st);
    return r;
}
orc_path* orc_path_transformed(const orc_path* p, const double tr[6]) {
    auto* r = new orc_path();
    r->p = p->p;
    r->p.transform(TR(tr));
    return r;
}
orc_path* orc_path_checkerboard(const double mm[4], double cell) {
    auto* r = new orc_path();
    PathBuilder b;
    b.checkerboard(BB(mm), cell);
    r->p = b.build();
    return r;
}
orc_path* orc_path_circle(double cx, double cy, double rad) {
    auto* r = new orc_path();
    PathBuilder b;
    b.move_to(Point(cx, cy)).circle(rad);
    r->p = b.build();
    return r;
}

void orc_fit_size(const double mm[4], size_t w, size_t h, int align, size_t* ow, size_t* oh, double tr_out[6]) {
    Size s; s.width = w; s.height = h;
    auto r = fit_size(BB(mm), s, align == 0 ? Align::Min : (align == 1 ? Align::Mid : Align::Max));
    *ow = r.first.width; *oh = r.first.height;
    TR_out(r.second, tr_out);
}
int orc_transform_parse(const char* text, double out[6]) {
    try { TR_out(parse_transform(text, std::strlen(text)), out); return 1; }
    catch (const std::exception& e) { g_err = e.what(); return 0; }
}
void orc_transform_mul(const double a[6], const double b[6], double out[6]) { TR_out(TR(a) * TR(b), out); }
int orc_transform_invert(const double a[6], double out[6]) {
    auto r = TR(a).invert();
    if (!r) return 0;
    TR_out(*r, out);
    return 1;
}
double orc_parse_scalar(const char* text, size_t len, size_t* consumed) {
    ByteParser ps(text, len);
    Scalar v;
    bool ok = ps.try_scalar(v);
    if (consumed) *consumed = ps.pos;
    return ok ? v : std::nan("");
}

long orc_flatten(const orc_path* p, const double tr[6], double flatness, int close, double* lines_out, size_t cap) {
    std::vector<Line> lines;
    try { p->p.flatten(TR(tr), flatness, close != 0, lines); }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
    size_t n = std::min(cap, lines.size());
    for (size_t i = 0; i < n; i++) {
        lines_out[4 * i] = lines[i].p[0].x; lines_out[4 * i + 1] = lines[i].p[0].y;
        lines_out[4 * i + 2] = lines[i].p[1].x; lines_out[4 * i + 3] = lines[i].p[1].y;
    }
    return (long)lines.size();
}
void orc_signed_difference_line(double* data, size_t data_len, orc_shape shape, const double l[4]) {
    signed_difference_line(data, data_len, SH(shape), Line(Point(l[0], l[1]), Point(l[2], l[3])));
}
void orc_signed_difference_to_mask(double* data, orc_shape shape, int rule) {
    signed_difference_to_mask(data, SH(shape), (FillRule)rule);
}
int orc_mask(const orc_path* p, const double tr[6], double flatness, int rule, double* data, size_t data_len, orc_shape shape) {
    try { mask(p->p, TR(tr), flatness, (FillRule)rule, data, data_len, SH(shape)); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int orc_mask_threads(const orc_path* p, const double tr[6], double flatness, int rule, double* data, size_t w, size_t h, int threads) {
    std::vector<Line> lines;
    try { p->p.flatten(TR(tr), flatness, true, lines); }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
    if (threads < 1) threads = 1;
    if ((size_t)threads > h) threads = (int)std::max<size_t>(h, 1);
    auto work = [&](int k) {
        size_t r0 = h * (size_t)k / (size_t)threads, r1 = h * (size_t)(k + 1) / (size_t)threads;
        if (r1 <= r0) return;
        Shape sh = Shape::simple(r1 - r0, w);
        Scalar* band = data + r0 * w;
        Scalar dy = (Scalar)r0;
        for (const Line& l : lines) {
            // band-local translate(0, -r0): the y<0 / y>=H clipping of signed_difference_line crops exactly
            Scalar lo = rmin(l.p[0].y, l.p[1].y), hi = rmax(l.p[0].y, l.p[1].y);
            if (hi < (Scalar)r0 || lo > (Scalar)r1) continue;
            signed_difference_line(band, (r1 - r0) * w, sh, Line(Point(l.p[0].x, l.p[0].y - dy), Point(l.p[1].x, l.p[1].y - dy)));
        }
        signed_difference_to_mask(band, sh, (FillRule)rule);
    };
    if (threads == 1) { work(0); return 0; }
    std::vector<std::thread> ts;
    for (int k = 0; k < threads; k++) ts.emplace_back(work, k);
    for (auto& t : ts) t.join();
    return 0;
}
// A batch of independent paths on `threads` host threads (bench.py --impl reference, config 4): thread k takes paths
// k, k + threads, ... and renders each into its own w x h scratch image — `img.clear()` + `Rasterizer::mask` when
// paint is NULL, `ImageOwned::new_default` + `Rasterizer::fill` otherwise — exactly what a caller of the
// single-threaded reference would do with one rasterizer per thread.
int orc_batch_threads(const orc_path* const* paths, size_t n, const double tr[6], double flatness, int rule, const orc_paint* paint,
                      size_t w, size_t h, int threads) {
    if (threads < 1) threads = 1;
    std::vector<int> rc((size_t)threads, 0);
    auto work = [&](int k) {
        try {
            Shape sh = Shape::simple(h, w);
            std::vector<Scalar> m(paint ? 0 : w * h);
            std::vector<LinColor> c(paint ? w * h : 0);
            for (size_t i = (size_t)k; i < n; i += (size_t)threads) {
                if (paint) {
                    std::fill(c.begin(), c.end(), LinColor{});
                    fill(paths[i]->p, TR(tr), flatness, (FillRule)rule, paint->p, c.data(), sh);
                } else {
                    std::fill(m.begin(), m.end(), 0.0);
                    mask(paths[i]->p, TR(tr), flatness, (FillRule)rule, m.data(), m.size(), sh);
                }
            }
        } catch (const std::exception&) { rc[(size_t)k] = -1; }
    };
    if (threads == 1) { work(0); return rc[0]; }
    std::vector<std::thread> ts;
    for (int k = 0; k < threads; k++) ts.emplace_back(work, k);
    for (auto& t : ts) t.join();
    for (int r : rc) if (r) { g_err = "orc_batch_threads: a path failed"; return -1; }
    return 0;
}
long orc_mask_iter(const orc_path* p, const double tr[6], double flatness, size_t w, size_t h, int rule, orc_pixel* out, size_t cap) {
    size_t n = 0;
    try {
        Size s; s.width = w; s.height = h;
        mask_iter(p->p, TR(tr), flatness, s, (FillRule)rule, [&](const Pixel& px) {
            if (n < cap) { out[n].x = px.x; out[n].y = px.y; out[n].alpha = px.alpha; }
            n++;
        });
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    return (long)n;
}
int orc_fill(const orc_path* p, const double tr[6], double flatness, int rule, const orc_paint* paint, float* data, orc_shape shape) {
    static_assert(sizeof(LinColor) == 16, "LinColor must be 4 x f32");
    try { fill(p->p, TR(tr), flatness, (FillRule)rule, paint->p, reinterpret_cast<LinColor*>(data), SH(shape)); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
}

void orc_rgba_to_lin(const uint8_t rgba[4], float out[4]) {
    RGBA c; std::memcpy(c.v, rgba, 4);
    LinColor l = rgba_to_lin(c);
    std::memcpy(out, l.c, 16);
}
void orc_lin_to_rgba(const float lin[4], uint8_t out[4]) {
    RGBA c = lin_to_rgba(LinColor(lin[0], lin[1], lin[2], lin[3]));
    std::memcpy(out, c.v, 4);
}
void orc_lin_to_rgba_image(const float* lin, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; i++) orc_lin_to_rgba(lin + 4 * i, out + 4 * i);
}
int orc_parse_color(const char* text, float out[4]) {
    RGBA c;
    if (!parse_rgba(text, c)) return 0;
    LinColor l = rgba_to_lin(c);
    std::memcpy(out, l.c, 16);
    return 1;
}
void orc_l2s(const float in[4], float out[4]) { f32x4 x; std::memcpy(x.v, in, 16); x = l2s(x); std::memcpy(out, x.v, 16); }
void orc_s2l(const float in[4], float out[4]) { f32x4 x; std::memcpy(x.v, in, 16); x = s2l(x); std::memcpy(out, x.v, 16); }
float orc_linear_to_srgb(float v) { return linear_to_srgb(v); }
float orc_srgb_to_linear(float v) { return srgb_to_linear(v); }
double orc_spread_at(int spread, double t) { return spread_at((GradSpread)spread, t); }

orc_paint* orc_paint_solid(const float lin[4]) {
    auto* r = new orc_paint();
    r->p = Paint::make_solid(LinColor(lin[0], lin[1], lin[2], lin[3]));
    return r;
}
static GradStops make_stops(const double* pos, const float* colors, size_t n, int sort_stops) {
    std::vector<GradStop> v;
    for (size_t i = 0; i < n; i++) v.push_back({pos[i], LinColor(colors[4 * i], colors[4 * i + 1], colors[4 * i + 2], colors[4 * i + 3])});
    if (sort_stops) return GradStops(v);
    GradStops gs;
    gs.stops = v;
    return gs;
}
orc_paint* orc_paint_linear(const double* pos, const float* colors, size_t n, int sort_stops, int units, int linear_colors, int spread,
                            const double tr[6], const double start[2], const double end[2]) {
    auto* r = new orc_paint();
    r->p = Paint::make_linear(make_stops(pos, colors, n, sort_stops), (Units)units, linear_colors != 0, (GradSpread)spread, TR(tr),
                              Point(start[0], start[1]), Point(end[0], end[1]));
    return r;
}
orc_paint* orc_paint_radial(const double* pos, const float* colors, size_t n, int sort_stops, int units, int linear_colors, int spread,
                            const double tr[6], const double center[2], double radius, const double fcenter[2], double fradius) {
    auto* r = new orc_paint();
    r->p = Paint::make_radial(make_stops(pos, colors, n, sort_stops), (Units)units, linear_colors != 0, (GradSpread)spread, TR(tr),
                              Point(center[0], center[1]), radius, Point(fcenter[0], fcenter[1]), fradius);
    return r;
}
orc_paint* orc_paint_linear_stored(const double* pos, const float* colors, size_t n, int units, int linear_colors, int spread,
                                   const double tr[6], const double start[2], const double end[2]) {
    // construct as linear_colors (no convert_to_srgb), then restore the flag: `at` applies into_linear iff !linear_colors
    auto* r = orc_paint_linear(pos, colors, n, 0, units, 1, spread, tr, start, end);
    r->p.linear_colors = linear_colors != 0;
    return r;
}
orc_paint* orc_paint_radial_stored(const double* pos, const float* colors, size_t n, int units, int linear_colors, int spread,
                                   const double tr[6], const double center[2], double radius, const double fcenter[2], double fradius) {
    auto* r = orc_paint_radial(pos, colors, n, 0, units, 1, spread, tr, center, radius, fcenter, fradius);
    r->p.linear_colors = linear_colors != 0;
    return r;
}
void orc_paint_free(orc_paint* p) { delete p; }
void orc_paint_at(const orc_paint* p, double x, double y, float out[4]) {
    LinColor c = p->p.at(Point(x, y));
    std::memcpy(out, c.c, 16);
}
int orc_paint_radial_offset(const orc_paint* p, double x, double y, double* out) {
    auto o = p->p.radial_offset(Point(x, y));
    if (!o) return 0;
    *out = *o;
    return 1;
}
void orc_paint_describe(const orc_paint* pp, orc_paint_desc* d) {
    const Paint& p = pp->p;
    std::memset(d, 0, sizeof(*d));
    d->kind = (int)p.kind; d->units = (int)p.units; d->linear_colors = p.linear_colors; d->spread = (int)p.spread;
    TR_out(p.transform(), d->tr);
    if (p.kind == PaintKind::Linear) {
        d->p0[0] = p.start.x; d->p0[1] = p.start.y; d->p1[0] = p.end.x; d->p1[1] = p.end.y; d->dir[0] = p.dir.x; d->dir[1] = p.dir.y;
    } else if (p.kind == PaintKind::Radial) {
        d->p0[0] = p.center.x; d->p0[1] = p.center.y; d->p1[0] = p.fcenter.x; d->p1[1] = p.fcenter.y; d->r0 = p.radius; d->r1 = p.fradius;
    }
    std::memcpy(d->solid, p.solid.c, 16);
    d->n_stops = p.stops.stops.size();
}
void orc_paint_stops(const orc_paint* pp, double* pos, float* colors) {
    const auto& st = pp->p.stops.stops;
    for (size_t i = 0; i < st.size(); i++) { pos[i] = st[i].position; std::memcpy(colors + 4 * i, st[i].color.c, 16); }
}

orc_scene* orc_scene_load_json(const char* text, size_t len) {
    try {
        auto* r = new orc_scene();
        r->s = scene_from_json(text, len);
        return r;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
// examples/rasterize.rs:277-308 with default flags: checkerboard(16, EvenOdd, #d0d0d0) + black NonZero fill
orc_scene* orc_scene_cli_rasterize(const orc_path* path, const double tr[6], size_t w, size_t h) {
    auto* r = new orc_scene();
    BBox bbox(Point(0.0, 0.0), Point((Scalar)w, (Scalar)h));
    auto fg = std::make_shared<Paint>(Paint::make_solid(LinColor(0.f, 0.f, 0.f, 1.f)));
    ScenePtr fillp = Scene::transform(Scene::fill(std::make_shared<Path>(path->p), fg, FillRule::NonZero), TR(tr));
    ScenePtr scene = Scene::group({fillp});
    PathBuilder b;
    b.checkerboard(bbox, 16.0);
    auto cb = std::make_shared<Path>(b.build());
    auto cbp = std::make_shared<Paint>(Paint::make_solid(parse_lin_color("#d0d0d0")));
    r->s = Scene::group({Scene::fill(cb, cbp, FillRule::EvenOdd), scene});
    return r;
}

// LCG: benches/scene_bench.rs:53-88
static uint32_t lcg_step(uint32_t* state) {
    *state = ((*state) * 214013u + 2531011u) & 0x7fffffffu;
    return *state >> 16;
}
static uint32_t lcg_u32(uint32_t* state) {
    uint32_t hi = lcg_step(state) & 0xffffu;
    uint32_t lo = lcg_step(state) & 0xffffu;
    return (hi << 16) | lo;
}
static uint64_t lcg_u64(uint32_t* state) {
    uint64_t hi = lcg_u32(state);
    uint64_t lo = lcg_u32(state);
    return (hi << 32) | lo;
}
double orc_lcg_uniform(uint32_t* state) {
    const double bpr_recip = 1.1102230246251565e-16;  // 2^-53
    return (double)(lcg_u64(state) >> 10) * bpr_recip;
}
orc_scene* orc_scene_many_circles(uint32_t seed, size_t count, size_t size) {
    auto* r = new orc_scene();
    uint32_t st = seed;
    Scalar fs = (Scalar)size;
    std::vector<ScenePtr> group;
    for (size_t i = 0; i < count; i++) {
        Scalar px = orc_lcg_uniform(&st);
        Scalar py = orc_lcg_uniform(&st);
        PathBuilder b;
        b.move_to(Point(px * fs, py * fs));
        b.circle(orc_lcg_uniform(&st) * 10.0 + 30.0);
        auto path = std::make_shared<Path>(b.build());
        RGBA c;
        c.v[0] = (uint8_t)(lcg_u32(&st) % 256); c.v[1] = (uint8_t)(lcg_u32(&st) % 256); c.v[2] = (uint8_t)(lcg_u32(&st) % 256); c.v[3] = 255;
        auto paint = std::make_shared<Paint>(Paint::make_solid(rgba_to_lin(c)));
        group.push_back(Scene::fill(path, paint, FillRule::NonZero));
    }
    r->s = Scene::group(group);
    return r;
}
orc_path* orc_glyph(uint32_t seed) {
    auto* r = new orc_path();
    uint32_t st = seed;
    PathBuilder b;
    auto coord = [&]() { return orc_lcg_uniform(&st) * 56.0 + 4.0; };
    auto pt = [&]() { Scalar x = coord(); Scalar y = coord(); return Point(x, y); };
    for (int c = 0; c < 3; c++) {
        b.move_to(pt());
        for (int k = 0; k < 6; k++) { Point q1 = pt(); Point q2 = pt(); Point q3 = pt(); b.cubic_to(q1, q2, q3); }
        b.close();
    }
    r->p = b.build();
    return r;
}
void orc_scene_free(orc_scene* s) { delete s; }
int orc_scene_bbox(const orc_scene* s, const double tr[6], double out[4]) {
    auto bb = s->s->bbox(TR(tr));
    if (!bb) return 0;
    out[0] = bb->min.x; out[1] = bb->min.y; out[2] = bb->max.x; out[3] = bb->max.y;
    return 1;
}
orc_layer* orc_scene_render(const orc_scene* s, double flatness, const double tr[6], const double* view, const float* bg) {
    try {
        std::optional<BBox> v;
        if (view) v = BB(view);
        std::optional<LinColor> b;
        if (bg) b = LinColor(bg[0], bg[1], bg[2], bg[3]);
        auto* r = new orc_layer();
        r->l = scene_render(*s->s, flatness, TR(tr), v, b);
        return r;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void orc_layer_info(const orc_layer* l, int32_t* x, int32_t* y, size_t* w, size_t* h) {
    *x = l->l.x; *y = l->l.y; *w = l->l.width(); *h = l->l.height();
}
const float* orc_layer_data(const orc_layer* l) { return reinterpret_cast<const float*>(l->l.data.data()); }
void orc_layer_free(orc_layer* l) { delete l; }

long orc_scene_fill_jobs(const orc_scene* cs, const double tr[6], const double* view, orc_fill_job* out, size_t cap) {
    auto* s = const_cast<orc_scene*>(cs);
    Pipeline p;
    std::optional<BBox> v;
    if (view) v = BB(view);
    p.build_rec(*s->s, v, TR(tr));
    // render order == depth-first order of Fill nodes under the root
    std::vector<size_t> order;
    bool ok = true;
    std::vector<size_t> stack;
    if (!p.nodes.empty()) {
        std::function<void(size_t)> walk = [&](size_t id) {
            const PipelineNode& n = p.nodes[id];
            if (n.kind == PipelineNode::Fill) order.push_back(id);
            else if (n.kind == PipelineNode::Group) { for (size_t c : n.children) walk(c); }
            else ok = false;
        };
        walk(p.nodes.size() - 1);
    }
    if (!ok) return -1;
    size_t n = 0;
    for (size_t id : order) {
        const PipelineNode& node = p.nodes[id];
        if (n < cap) {
            auto pp = std::make_unique<orc_path>(); pp->p = *node.path;
            auto pa = std::make_unique<orc_paint>(); pa->p = *node.paint;
            out[n].path = pp.get(); out[n].paint = pa.get(); out[n].fill_rule = (int)node.fill_rule;
            TR_out(node.tr, out[n].tr);
            out[n].bbox[0] = node.bbox.min.x; out[n].bbox[1] = node.bbox.min.y; out[n].bbox[2] = node.bbox.max.x; out[n].bbox[3] = node.bbox.max.y;
            s->job_paths.push_back(std::move(pp));
            s->job_paints.push_back(std::move(pa));
        }
        n++;
    }
    return (long)n;
}

long orc_scene_pipeline(const orc_scene* cs, const double tr[6], const double* view, orc_pipe_node* out, size_t cap, long* children,
                        size_t children_cap, size_t* n_children_out) {
    auto* s = const_cast<orc_scene*>(cs);
    Pipeline p;
    std::optional<BBox> v;
    if (view) v = BB(view);
    p.build_rec(*s->s, v, TR(tr));
    size_t nc = 0;
    for (size_t id = 0; id < p.nodes.size(); id++) {
        const PipelineNode& node = p.nodes[id];
        if (out && id < cap) {
            orc_pipe_node& o = out[id];
            o.kind = (int)node.kind;
            o.path = nullptr;
            o.paint = nullptr;
            o.fill_rule = (int)node.fill_rule;
            o.opacity = node.opacity;
            o.child = (long)node.child;
            o.child_begin = (long)nc;
            o.child_count = (long)node.children.size();
            TR_out(node.kind == PipelineNode::Clip ? node.clip_tr : node.tr, o.tr);
            o.bbox[0] = node.bbox.min.x; o.bbox[1] = node.bbox.min.y; o.bbox[2] = node.bbox.max.x; o.bbox[3] = node.bbox.max.y;
            if (node.path) {
                auto pp = std::make_unique<orc_path>(); pp->p = *node.path;
                o.path = pp.get();
                s->job_paths.push_back(std::move(pp));
            }
            if (node.paint) {
                auto pa = std::make_unique<orc_paint>(); pa->p = *node.paint;
                o.paint = pa.get();
                s->job_paints.push_back(std::move(pa));
            }
        }
        for (size_t c : node.children) {
            if (children && nc < children_cap) children[nc] = (long)c;
            nc++;
        }
    }
    if (n_children_out) *n_children_out = nc;
    return (long)p.nodes.size();
}

}  // extern "C"
