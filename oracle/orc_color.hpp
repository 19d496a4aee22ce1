// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_geom.hpp header).
// Restatement of src/color.rs (LinColor / RGBA), src/simd/x86.rs:197-244 (polynomial l2s / s2l on
// all four lanes — the DEFAULT on x86-64 with the `simd` feature) with src/simd/fallback.rs:157-166
// behind a switch, and src/grad.rs (GradStops, GradLinear, GradRadial, GradSpread).
#pragma once
#include "orc_geom.hpp"
#include <cstdio>

namespace orc {

// true  -> src/simd/x86.rs variant (reference default on x86-64)
// false -> src/simd/fallback.rs variant
inline bool& simd_x86() { static bool v = true; return v; }

// src/color.rs:465-477
inline float linear_to_srgb(float x0) {
    if (x0 <= 0.0031308f) return x0 * 12.92f;
    float x1 = std::sqrt(x0);
    float x2 = std::sqrt(x1);
    float x3 = std::sqrt(x2);
    return -0.01848558f * x0 + 0.6445592f * x1 + 0.70994765f * x2 - 0.33605254f * x3;
}
// src/color.rs:480-486
inline float srgb_to_linear(float value) {
    if (value <= 0.04045f) return value / 12.92f;
    return std::pow((value + 0.055f) / 1.055f, 2.4f);
}

struct f32x4 { float v[4]; };

// src/simd/x86.rs:197-214 (per lane; _mm_sqrt_ps is IEEE-exact, no FMA is used) / fallback.rs:157-160
inline f32x4 l2s(f32x4 x) {
    f32x4 r;
    if (simd_x86()) {
        for (int i = 0; i < 4; i++) {
            float x0 = x.v[i];
            float x1 = std::sqrt(x0);
            float x2 = std::sqrt(x1);
            float x3 = std::sqrt(x2);
            float high = -0.01848558f * x0 + 0.6445592f * x1 + 0.70994765f * x2 - 0.33605254f * x3;
            // _mm_blendv_ps(high, x0*12.92, cmple(x0, 0.0031308)): NaN compares false -> high
            r.v[i] = (x0 <= 0.0031308f) ? x0 * 12.92f : high;
        }
    } else {
        for (int i = 0; i < 3; i++) r.v[i] = linear_to_srgb(x.v[i]);
        r.v[3] = x.v[3];
    }
    return r;
}
// src/simd/x86.rs:217-244 / fallback.rs:163-166
inline f32x4 s2l(f32x4 vs) {
    f32x4 r;
    if (simd_x86()) {
        for (int i = 0; i < 4; i++) {
            float v = vs.v[i];
            float x1 = 2.0843103538116825f * v - 1.0843103538116827f;
            float x2 = x1 * x1;
            float x3 = x2 * x1;
            float high = 0.23361048543711943f + 0.4665843122387033f * x1 + 0.26901741378006355f * x2 +
                         0.031661580753065945f * x3;
            r.v[i] = (v <= 0.04045f) ? v * 0.07739938080495357f : high;
        }
    } else {
        for (int i = 0; i < 3; i++) r.v[i] = srgb_to_linear(vs.v[i]);
        r.v[3] = vs.v[3];
    }
    return r;
}

// src/color.rs:268-355 — premultiplied linear RGBA, f32x4
struct LinColor {
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    LinColor() = default;
    LinColor(float r, float g, float b, float a) { c[0] = r; c[1] = g; c[2] = b; c[3] = a; }
    float alpha() const { return c[3]; }
    LinColor scale(float s) const { return LinColor(c[0] * s, c[1] * s, c[2] * s, c[3] * s); }  // Mul<f32> :385-392
    LinColor add(const LinColor& o) const { return LinColor(c[0] + o.c[0], c[1] + o.c[1], c[2] + o.c[2], c[3] + o.c[3]); }
    f32x4 unmultiply() const {  // :308-317
        float a = alpha();
        f32x4 r;
        if (a <= 1e-6f) { r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.f; return r; }
        for (int i = 0; i < 4; i++) r.v[i] = c[i] / a;
        return r;
    }
    LinColor into_srgb() const {  // :322-324
        f32x4 s = l2s(unmultiply());
        float a = alpha();
        return LinColor(s.v[0] * a, s.v[1] * a, s.v[2] * a, s.v[3] * a);
    }
    LinColor into_linear() const {  // :330-332
        f32x4 s = s2l(unmultiply());
        float a = alpha();
        return LinColor(s.v[0] * a, s.v[1] * a, s.v[2] * a, s.v[3] * a);
    }
    LinColor blend_over(const LinColor& other) const {  // :342-344  other + self * (1 - other.alpha)
        return other.add(scale(1.0f - other.alpha()));
    }
    LinColor with_alpha(Scalar alpha) const { return scale((float)alpha); }  // :347-349
    LinColor lerp(const LinColor& other, float t) const {                    // :352-354  other*t + self*(1-t)
        return other.scale(t).add(scale(1.0f - t));
    }
    bool operator==(const LinColor& o) const { return c[0] == o.c[0] && c[1] == o.c[1] && c[2] == o.c[2] && c[3] == o.c[3]; }
};

struct RGBA { uint8_t v[4] = {0, 0, 0, 0}; };

// src/color.rs:164-175
inline RGBA lin_to_rgba(const LinColor& lin) {
    f32x4 s = l2s(lin.unmultiply());
    RGBA o;
    o.v[0] = f32_as_u8(s.v[0] * 255.0f + 0.5f);
    o.v[1] = f32_as_u8(s.v[1] * 255.0f + 0.5f);
    o.v[2] = f32_as_u8(s.v[2] * 255.0f + 0.5f);
    o.v[3] = f32_as_u8(lin.alpha() * 255.0f + 0.5f);
    return o;
}
// src/color.rs:394-406 — scalar exact srgb_to_linear, NOT the polynomial
inline LinColor rgba_to_lin(const RGBA& color) {
    float a = (float)color.v[3] / 255.0f;
    float r = srgb_to_linear((float)color.v[0] / 255.0f) * a;
    float g = srgb_to_linear((float)color.v[1] / 255.0f) * a;
    float b = srgb_to_linear((float)color.v[2] / 255.0f) * a;
    return LinColor(r, g, b, a);
}

// src/color.rs:96-141 — `#rrggbb`, `#rrggbbaa`, optional `/alpha` suffix. SVG colour names are
// not restated (out of scope: no data/ asset uses them); a few basic names are kept for tests.
inline bool parse_rgba(const std::string& text, RGBA& out) {
    std::string color = text;
    bool has_alpha = false;
    float alpha = 1.0f;
    size_t slash = color.rfind('/');
    if (slash != std::string::npos) {
        char* end = nullptr;
        std::string a = color.substr(slash + 1);
        alpha = std::strtof(a.c_str(), &end);
        if (end == a.c_str()) return false;
        has_alpha = true;
        color = color.substr(0, slash);
    }
    RGBA rgba;
    if (!color.empty() && color[0] == '#' && (color.size() == 7 || color.size() == 9)) {
        auto digit = [](char b, int& ok) -> int {
            if (b >= 'A' && b <= 'F') return b - 'A' + 10;
            if (b >= 'a' && b <= 'f') return b - 'a' + 10;
            if (b >= '0' && b <= '9') return b - '0';
            ok = 0;
            return 0;
        };
        int ok = 1;
        size_t n = (color.size() - 1) / 2;
        uint8_t vals[4] = {0, 0, 0, 255};
        for (size_t i = 0; i < n; i++) vals[i] = (uint8_t)((digit(color[1 + 2 * i], ok) << 4) | digit(color[2 + 2 * i], ok));
        if (!ok) return false;
        for (int i = 0; i < 4; i++) rgba.v[i] = vals[i];
    } else if (color == "black") {
        rgba.v[3] = 255;
    } else if (color == "white") {
        rgba.v[0] = rgba.v[1] = rgba.v[2] = rgba.v[3] = 255;
    } else if (color == "red") {
        rgba.v[0] = 255; rgba.v[3] = 255;
    } else {
        return false;
    }
    if (has_alpha) rgba.v[3] = f32_as_u8((float)rgba.v[3] * alpha);  // `as u8` truncation, :138
    out = rgba;
    return true;
}

// ---- paints -----------------------------------------------------------------------------
enum class Units : int { UserSpaceOnUse = 0, BoundingBox = 1 };  // src/rasterize.rs:171-176
enum class GradSpread : int { Pad = 0, Repeat = 1, Reflect = 2 };  // src/grad.rs:13-20

// src/grad.rs:24-30
inline Scalar spread_at(GradSpread s, Scalar t) {
    switch (s) {
        case GradSpread::Pad: return t;
        case GradSpread::Repeat: return rem_euclid(t, 1.0);
        default: return std::fabs(rem_euclid(t + 1.0, 2.0) - 1.0);
    }
}

struct GradStop { Scalar position; LinColor color; };  // src/grad.rs:41-44

// src/grad.rs:81-140
struct GradStops {
    std::vector<GradStop> stops;
    GradStops() = default;
    explicit GradStops(std::vector<GradStop> s) : stops(std::move(s)) {  // :86-99 (stable sort)
        std::stable_sort(stops.begin(), stops.end(), [](const GradStop& a, const GradStop& b) { return a.position < b.position; });
        if (stops.empty()) stops.push_back({0.0, LinColor(0.f, 0.f, 0.f, 1.f)});
    }
    void convert_to_srgb() { for (auto& s : stops) s.color = s.color.into_srgb(); }  // :101-105
    // :116-139 — binary search with a Less/Greater-only comparator == partition point of `position < t`
    LinColor at(Scalar t) const {
        size_t lo = 0, hi = stops.size();
        while (lo < hi) {
            size_t mid = lo + (hi - lo) / 2;
            if (stops[mid].position < t) lo = mid + 1; else hi = mid;
        }
        size_t index = lo, size = stops.size();
        if (index == 0) return stops[0].color;
        if (index == size) return stops[size - 1].color;
        const GradStop& p0 = stops[index - 1];
        const GradStop& p1 = stops[index];
        Scalar ratio = (t - p0.position) / (p1.position - p0.position);
        return p0.color.lerp(p1.color, (float)ratio);
    }
};

enum class PaintKind : int { Solid = 0, Linear = 1, Radial = 2 };

// One struct for the three `Paint` implementors: LinColor (src/color.rs:357-374),
// GradLinear (src/grad.rs:150-226), GradRadial (src/grad.rs:307-426).
struct Paint {
    PaintKind kind = PaintKind::Solid;
    LinColor solid;
    Units units = Units::UserSpaceOnUse;
    bool linear_colors = false;
    GradSpread spread = GradSpread::Pad;
    Transform tr;
    Point start, end, dir;           // linear
    Point center, fcenter;           // radial
    Scalar radius = 0.0, fradius = 0.0;
    GradStops stops;

    static Paint make_solid(const LinColor& c) { Paint p; p.kind = PaintKind::Solid; p.solid = c; return p; }
    // GradLinear::new, src/grad.rs:165-191
    static Paint make_linear(GradStops stops, Units units, bool linear_colors, GradSpread spread, Transform tr, Point start, Point end) {
        Paint p;
        p.kind = PaintKind::Linear;
        if (!linear_colors) stops.convert_to_srgb();
        Point dir = end - start;
        p.stops = std::move(stops); p.units = units; p.linear_colors = linear_colors; p.spread = spread; p.tr = tr;
        p.start = start; p.end = end; p.dir = dir / dir.dot(dir);
        return p;
    }
    // GradRadial::new, src/grad.rs:323-351
    static Paint make_radial(GradStops stops, Units units, bool linear_colors, GradSpread spread, Transform tr, Point center,
                             Scalar radius, Point fcenter, Scalar fradius) {
        Paint p;
        p.kind = PaintKind::Radial;
        if (!linear_colors) stops.convert_to_srgb();
        p.stops = std::move(stops); p.units = units; p.linear_colors = linear_colors; p.spread = spread; p.tr = tr;
        p.center = center; p.radius = radius; p.fcenter = fcenter; p.fradius = fradius;
        return p;
    }
    bool has_units() const { return kind != PaintKind::Solid; }  // `units()` is None for LinColor
    Transform transform() const { return kind == PaintKind::Solid ? Transform::identity() : tr; }

    // GradRadial::offset, src/grad.rs:361-396
    std::optional<Scalar> radial_offset(Point point) const {
        Point cd = center - fcenter;
        Point pd = point - fcenter;
        Scalar rd = radius - fradius;
        Scalar a = cd.dot(cd) - rd * rd;
        Scalar b = -2.0 * (cd.dot(pd) + fradius * rd);
        Scalar c = pd.dot(pd) - fradius * fradius;
        Roots2 r = quadratic_solve(a, b, c);
        if (r.n == 2) return rmax(r.v[0], r.v[1]);
        if (r.n == 1) return r.v[0];
        return std::nullopt;
    }
    // `Paint::at`: src/color.rs:358-360, src/grad.rs:202-211, 400-411
    LinColor at(Point point) const {
        switch (kind) {
            case PaintKind::Solid: return solid;
            case PaintKind::Linear: {
                Scalar t = (point - start).dot(dir);
                LinColor color = stops.at(spread_at(spread, t));
                return linear_colors ? color : color.into_linear();
            }
            default: {
                auto off = radial_offset(point);
                if (!off) return LinColor(0.f, 0.f, 0.f, 0.f);
                LinColor color = stops.at(spread_at(spread, *off));
                return linear_colors ? color : color.into_linear();
            }
        }
    }
};

}  // namespace orc
