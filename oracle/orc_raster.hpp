// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_geom.hpp header).
// Restatement of the SignedDifferenceRasterizer of src/rasterize.rs:299-507, 923-937 and the default
// `Rasterizer::fill` (src/rasterize.rs:70-115), over strided images (src/image.rs:6-35).
#pragma once
#include "orc_path.hpp"
#include "orc_color.hpp"

namespace orc {

// src/image.rs:6-35
struct Shape {
    size_t start = 0, width = 0, height = 0, row_stride = 0, col_stride = 1;
    static Shape simple(size_t height, size_t width) { Shape s; s.start = 0; s.width = width; s.height = height; s.row_stride = width; s.col_stride = 1; return s; }
    size_t offset(size_t row, size_t col) const { return start + row * row_stride + col * col_stride; }
};
// src/image.rs:588-605
inline Shape view_shape(const Shape& shape, size_t row_min, size_t row_max, size_t col_min, size_t col_max) {
    row_min = std::min(row_min, shape.height);
    row_max = std::min(std::max(row_max, row_min), shape.height);
    col_min = std::min(col_min, shape.width);
    col_max = std::min(std::max(col_max, col_min), shape.width);
    Shape s = shape;
    s.start = shape.offset(row_min, col_min);
    s.width = col_max - col_min;
    s.height = row_max - row_min;
    return s;
}

// src/rasterize.rs:923-937
inline bool split_at_zero_x(const Line& line, Line& keep, Line& rest) {
    Point p0 = line.p[0], p1 = line.p[1];
    if (p0.x >= 0.0 && p1.x >= 0.0) { keep = line; return false; }
    if (p0.x <= 0.0 && p1.x <= 0.0) { keep = Line(Point(0.0, p0.y), Point(0.0, p1.y)); return false; }
    Point mid = line.at(p0.x / (p0.x - p1.x));
    if (p0.x < 0.0) {
        keep = Line(mid, p1);
        rest = Line(Point(0.0, p0.y), mid);
    } else {
        keep = Line(p0, mid);
        rest = Line(mid, Point(0.0, p1.y));
    }
    return true;
}

// src/rasterize.rs:365-470.  `data`/`data_len` is the WHOLE backing slice (the reference bounds-checks
// against `data.len()`, :442); `shape` is the (possibly strided sub-)view.
inline void signed_difference_line(Scalar* data, size_t data_len, const Shape& shape, Line line) {
    Point p0 = line.p[0], p1 = line.p[1];
    // :370-387 right edge
    Scalar width = (Scalar)shape.width - 1.0;
    if (p0.x > width || p1.x > width) {
        if (p0.x > width && p1.x > width) {
            line = Line(Point(width - 0.001, p0.y), Point(width - 0.001, p1.y));
        } else {
            Scalar t = (p0.x - width) / (p0.x - p1.x);
            Point mid(width, (1.0 - t) * p0.y + t * p1.y);
            line = (p0.x < width) ? Line(p0, mid) : Line(mid, p1);
        }
    }
    // :389-393 left edge
    Line keep, rest;
    if (split_at_zero_x(line, keep, rest)) signed_difference_line(data, data_len, shape, rest);
    line = keep;
    p0 = line.p[0];
    p1 = line.p[1];
    size_t stride = shape.col_stride;

    if (std::fabs(p0.y - p1.y) < EPSILON) return;  // :400-403
    Scalar dir;
    if (p0.y < p1.y) { dir = 1.0; } else { dir = -1.0; std::swap(p0, p1); }  // :405-409
    Scalar dxdy = (p1.x - p0.x) / (p1.y - p0.y);
    size_t y_begin = as_usize(rmax(p0.y, 0.0));                             // :414
    Scalar x = (p0.y < 0.0) ? p0.x - p0.y * dxdy : p0.x;                    // :415-419
    Scalar x_next = x;
    size_t y_end = std::min(shape.height, as_usize(rmax(std::ceil(p1.y), 0.0)));  // :421
    for (size_t y = y_begin; y < y_end; y++) {
        x = x_next;
        size_t row_offset = shape.offset(y, 0);
        Scalar dy = rmin((Scalar)(y + 1), p1.y) - rmax((Scalar)y, p0.y);
        Scalar d = dir * dy;
        x_next = x + dxdy * dy;
        Scalar x0, x1;
        if (x < x_next) { x0 = x; x1 = x_next; } else { x0 = x_next; x1 = x; }
        Scalar x0_floor = rmax(std::floor(x0), 0.0);
        int32_t x0i = as_i32(x0_floor);
        Scalar x1_ceil = rmin(std::ceil(x1), width);
        int32_t x1i = as_i32(x1_ceil);
        if (x1i <= x0i + 1) {
            Scalar xmf = 0.5 * (x + x_next) - x0_floor;
            data[row_offset + (size_t)x0i * stride] += d * (1.0 - xmf);
            size_t offset = row_offset + (size_t)(x0i + 1) * stride;
            if (offset < data_len) data[offset] += d * xmf;
        } else {
            Scalar s = 1.0 / (x1 - x0);
            Scalar x0f = x0 - x0_floor;
            Scalar x1f = x1 - x1_ceil + 1.0;
            Scalar a0 = 0.5 * s * (1.0 - x0f) * (1.0 - x0f);
            Scalar am = 0.5 * s * x1f * x1f;
            data[row_offset + (size_t)x0i * stride] += d * a0;
            if (x1i == x0i + 2) {
                data[row_offset + (size_t)(x0i + 1) * stride] += d * (1.0 - a0 - am);
            } else {
                Scalar a1 = s * (1.5 - x0f);
                data[row_offset + (size_t)(x0i + 1) * stride] += d * (a1 - a0);
                for (int32_t xi = x0i + 2; xi < x1i - 1; xi++) data[row_offset + (size_t)xi * stride] += d * s;
                Scalar a2 = a1 + (Scalar)(x1i - x0i - 3) * s;
                data[row_offset + (size_t)(x1i - 1) * stride] += d * (1.0 - a2 - am);
            }
            data[row_offset + (size_t)x1i * stride] += d * am;
        }
    }
}

// src/rasterize.rs:473-507
inline void signed_difference_to_mask(Scalar* data, const Shape& shape, FillRule rule) {
    if (rule == FillRule::NonZero) {
        for (size_t y = 0; y < shape.height; y++) {
            Scalar acc = 0.0;
            for (size_t x = 0; x < shape.width; x++) {
                size_t off = shape.offset(y, x);
                acc += data[off];
                Scalar value = std::fabs(acc);
                data[off] = value > 1.0 ? 1.0 : (value < 1e-6 ? 0.0 : value);
            }
        }
    } else {
        for (size_t y = 0; y < shape.height; y++) {
            Scalar acc = 0.0;
            for (size_t x = 0; x < shape.width; x++) {
                size_t off = shape.offset(y, x);
                acc += data[off];
                data[off] = std::fabs(rem_euclid(acc + 1.0, 2.0) - 1.0);
            }
        }
    }
}

// `SignedDifferenceRasterizer::mask`, src/rasterize.rs:299-311
inline void mask(const Path& path, const Transform& tr, Scalar flatness, FillRule rule, Scalar* data, size_t data_len, const Shape& shape) {
    std::vector<Line> lines;
    path.flatten(tr, flatness, true, lines);
    for (const Line& l : lines) signed_difference_line(data, data_len, shape, l);
    signed_difference_to_mask(data, shape, rule);
}

struct Pixel { size_t x, y; Scalar alpha; };  // src/rasterize.rs `Pixel`

// `SignedDifferenceRasterizer::mask_iter`, src/rasterize.rs:313-355 — calls `emit(Pixel)` in row-major order
template <class F>
inline void mask_iter(const Path& path, const Transform& tr, Scalar flatness, Size size, FillRule rule, F emit) {
    if (size.width == 0 || size.height == 0) return;
    size_t width = size.width;
    Shape shape = Shape::simple(size.height, width + 1);
    std::vector<Scalar> img((width + 1) * size.height, 0.0);
    std::vector<Line> lines;
    path.flatten(tr, flatness, true, lines);
    for (const Line& l : lines) signed_difference_line(img.data(), img.size(), shape, l);
    Scalar winding = 0.0;
    for (size_t index = 0; index < img.size(); index++) {
        size_t y = index / shape.width;
        size_t x = index - y * shape.width;
        if (x == 0) winding = 0.0;
        else if (x >= width) continue;
        winding += img[index];
        Scalar alpha = alpha_from_winding(rule, winding);
        if (std::fabs(alpha) < 1e-6) continue;
        emit(Pixel{x, y, alpha});
    }
}

// Default `Rasterizer::fill` + `fill_impl`, src/rasterize.rs:70-115
inline void fill(const Path& path, const Transform& tr, Scalar flatness, FillRule rule, const Paint& paint, LinColor* data,
                 const Shape& shape) {
    Size size{shape.width, shape.height};
    if (!paint.has_units()) {
        LinColor color = paint.at(Point(0.0, 0.0));
        mask_iter(path, tr, flatness, size, rule, [&](const Pixel& px) {
            LinColor& dst = data[shape.offset(px.y, px.x)];
            dst = dst.blend_over(color.with_alpha(px.alpha));
        });
        return;
    }
    Transform units_tr;
    if (paint.units == Units::UserSpaceOnUse) {
        units_tr = tr * paint.transform();
    } else {
        auto bb = path.bbox(Transform::identity());
        if (!bb) return;
        units_tr = tr * bb->unit_transform() * paint.transform();
    }
    auto pixel_tr = units_tr.invert();
    if (!pixel_tr) return;
    // NOTE: in the reference the mask_iter is constructed (lines rasterized) before the early returns;
    // there is no observable difference.
    mask_iter(path, tr, flatness, size, rule, [&](const Pixel& px) {
        Point point((Scalar)px.x + 0.5, (Scalar)px.y + 0.5);
        LinColor color = paint.at(pixel_tr->apply(point));
        LinColor& dst = data[shape.offset(px.y, px.x)];
        dst = dst.blend_over(color.with_alpha(px.alpha));
    });
}

}  // namespace orc
