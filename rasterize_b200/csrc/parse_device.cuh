// Device-side SVG path parser: `SvgPathParser` (src/svg.rs:241-421) over the byte scanner of src/svg.rs:62-236, applied to
// `PathBuilder` semantics (src/path.rs:832-972), with `Path::bbox` (src/path.rs:428-431, src/curve.rs bbox) and `fit_size`
// (src/geometry.rs:470-516) behind it.  One thread parses one path string; the same code runs twice (count, emit) with a
// different output policy.  f64 in the reference's expression order, translation unit compiled with --fmad=false.
//
// The scalar scanner is the reference's own, not strtod: value = (i64 mantissa as f64) * powi(10, exponent), with the
// mantissa accumulated in wrapping 64-bit arithmetic and `powi` = compiler-rt's square-and-multiply (__powidf2).
#pragma once
#include "stroke_device.cuh"

namespace rgpu {
namespace sv {

using namespace sk;

enum { kParseOk = 0, kParseInvalidCmd = 1, kParseInvalidScalar = 2, kParseInvalidFlag = 3 };  // `SvgParserError`, src/svg.rs:604-640

SD_FN inline double powi(double a, int b) {
    const bool recip = b < 0;
    double r = 1.0;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

struct Bytes {  // src/svg.rs:62-162
    const uint8_t* data;
    uint32_t len, pos;
    SD_FN void init(const uint8_t* d, uint32_t n) {
        data = d;
        len = n;
        pos = 0;
    }
    // Plain byte loads.  (Reading through an aligned 16-byte register window was measured: slower, see DESIGN.md.)
    SD_FN uint8_t at(uint32_t p) const { return data[p]; }
    SD_FN int peek() const { return pos < len ? (int)data[pos] : -1; }
    SD_FN int next() { return pos < len ? (int)data[pos++] : -1; }
    SD_FN void separators() {
        while (pos < len) {
            const uint8_t b = data[pos];
            if (b == ' ' || b == '\t' || b == '\r' || b == '\n' || b == ',') pos++;
            else break;
        }
    }
    // src/svg.rs:165-235
    SD_FN bool try_scalar(double& out) {
        separators();
        uint64_t mantissa = 0;
        int64_t exponent = 0, sign = 1;
        int c = peek();
        if (c == '-' || c == '+') {
            if (c == '-') sign = -1;
            pos++;
        }
        uint32_t whole = 0, frac = 0;
        while (pos < len && at(pos) >= '0' && at(pos) <= '9') {
            mantissa = mantissa * 10u + (uint64_t)(at(pos) - '0');
            pos++;
            whole++;
        }
        if (peek() == '.') {
            pos++;
            while (pos < len && at(pos) >= '0' && at(pos) <= '9') {
                mantissa = mantissa * 10u + (uint64_t)(at(pos) - '0');
                pos++;
                frac++;
                exponent -= 1;
            }
        }
        const int64_t m = (int64_t)(mantissa * (uint64_t)sign);
        if (whole + frac == 0) return false;
        c = peek();
        if (c == 'e' || c == 'E') {
            pos++;
            int64_t sci = 0, sci_sign = 1;
            c = peek();
            if (c == '-' || c == '+') {
                if (c == '-') sci_sign = -1;
                pos++;
            }
            uint32_t nd = 0;
            while (pos < len && at(pos) >= '0' && at(pos) <= '9') {
                sci = (int64_t)((uint64_t)sci * 10u + (uint64_t)(at(pos) - '0'));
                pos++;
                nd++;
            }
            if (nd == 0) return false;
            exponent = exponent + sci_sign * sci;
        }
        out = (double)m * powi(10.0, (int)(int32_t)exponent);
        return true;
    }
};

// ---- `PathBuilder`, src/path.rs:832-972: `Out` receives the segments and the subpath ends ----
template <class Out>
struct PathBuild {
    Out& out;
    P2 position;
    uint32_t n_seg, sub_first_seg;
    P2 sub_first_start;
    bool has_box;
    Box box;  // `Path::bbox(identity)`: src/path.rs:428-431 folds `Curve::bbox` over the segments
    SD_FN explicit PathBuild(Out& o) : out(o), position(mk(0.0, 0.0)), n_seg(0), sub_first_seg(0), sub_first_start(mk(0.0, 0.0)), has_box(false) {
        box.lo = box.hi = mk(0.0, 0.0);
    }
    SD_FN void push(const Seg& s) {
        if (n_seg == sub_first_seg) {
            sub_first_start = s.p[0];
            out.begin_subpath();
        }
        out.segment(s);
        // `segments[s].transform(identity).bbox(bb)`: the identity is applied as any transform (src/geometry.rs:363-367), which
        // turns -0.0 into 0.0 and an infinite coordinate into a NaN in the other one
        Seg t = s;
        for (int i = 0; i < 4; i++) t.p[i] = mk(s.p[i].x * 1.0 + s.p[i].y * 0.0 + 0.0, s.p[i].x * 0.0 + s.p[i].y * 1.0 + 0.0);
        box = seg_bbox_acc(t, has_box, box);
        has_box = true;
        n_seg++;
    }
    SD_FN void subpath_finish(bool close) {  // :849-868
        if (n_seg == sub_first_seg) return;
        if (close) position = sub_first_start;
        out.end_subpath(close);
        sub_first_seg = n_seg;
    }
    SD_FN void move_to(P2 p) {
        subpath_finish(false);
        position = p;
    }
    SD_FN void close() { subpath_finish(true); }
    SD_FN void line_to(P2 p) {  // :895-903
        if (!close_to(position, p)) {
            push(seg_line(position, p));
            position = p;
        }
    }
    SD_FN void quad_to(P2 p1, P2 p2) {
        push(seg_quad(position, p1, p2));
        position = p2;
    }
    SD_FN void cubic_to(P2 p1, P2 p2, P2 p3) {
        push(seg_cubic(position, p1, p2, p3));
        position = p3;
    }
    SD_FN void arc_to(P2 radii, double x_axis_rot, bool large, bool sweep, P2 p) {  // :945-972
        if (!arc_cubics(position, p, radii.x, radii.y, x_axis_rot, large, sweep, *this)) {
            line_to(p);
            return;
        }
        position = p;
    }
    SD_FN void finish() { subpath_finish(false); }  // `build`, :832-841
};

// ---- `SvgPathParser`, src/svg.rs:241-421, applied straight to the builder (`SvgPathCmd::apply`, :43-59) ----
// Returns kParseOk or the error kind; `err_pos` = byte offset of the error.
template <class Out>
SD_FN inline int parse_svg_path(const uint8_t* text, uint32_t len, PathBuild<Out>& builder, uint32_t& err_pos) {
    Bytes ps;
    ps.init(text, len);
    int prev_op = -1;
    int prev_kind = 0;  // 0 none / other, 1 QuadTo, 2 CubicTo
    P2 prev_c1 = mk(0.0, 0.0), prev_c2 = mk(0.0, 0.0);
    P2 position = mk(0.0, 0.0), subpath_start = mk(0.0, 0.0);
    int err = kParseOk;

    auto scalar = [&](double& v) -> bool {
        if (ps.try_scalar(v)) return true;
        err = kParseInvalidScalar;
        err_pos = ps.pos;
        return false;
    };
    auto point = [&](P2& p) -> bool {  // :265-271
        double x, y;
        if (!scalar(x) || !scalar(y)) return false;
        p = mk(x, y);
        if (prev_op >= 'a' && prev_op <= 'z') p = p + position;
        return true;
    };
    auto flag = [&](bool& f) -> bool {  // :274-289
        ps.separators();
        const int b = ps.peek();
        if (b == '0' || b == '1') {
            ps.pos++;
            f = b == '1';
            return true;
        }
        err = kParseInvalidFlag;
        err_pos = ps.pos;
        return false;
    };

    while (true) {
        ps.separators();
        int op = ps.next();
        if (op < 0) break;
        bool is_cmd = false;
        switch (op) {
            case 'M': case 'm': case 'L': case 'l': case 'V': case 'v': case 'H': case 'h': case 'C': case 'c': case 'S': case 's':
            case 'Q': case 'q': case 'T': case 't': case 'A': case 'a': case 'Z': case 'z':
                is_cmd = true;
                break;
            default:
                break;
        }
        if (is_cmd) {  // :297-310
            if (op == 'm') prev_op = 'l';
            else if (op == 'M') prev_op = 'L';
            else if (op == 'Z' || op == 'z') prev_op = -1;
            else prev_op = op;
        } else {  // :311-320 implicit repeat of the previous command
            ps.pos--;
            if (prev_op < 0) {
                err_pos = ps.pos;
                return kParseInvalidCmd;
            }
            op = prev_op;
        }
        P2 dst = position;
        int kind = 0;
        P2 k1 = mk(0.0, 0.0), k2 = mk(0.0, 0.0);
        switch (op) {
            case 'M': case 'm': {
                if (!point(dst)) return err;
                subpath_start = dst;
                builder.move_to(dst);
                break;
            }
            case 'L': case 'l': {
                if (!point(dst)) return err;
                builder.line_to(dst);
                break;
            }
            case 'V': case 'v': {
                double y;
                if (!scalar(y)) return err;
                dst = op == 'v' ? mk(position.x, position.y + y) : mk(position.x, y);
                builder.line_to(dst);
                break;
            }
            case 'H': case 'h': {
                double x;
                if (!scalar(x)) return err;
                dst = op == 'h' ? mk(position.x + x, position.y) : mk(x, position.y);
                builder.line_to(dst);
                break;
            }
            case 'Q': case 'q': {
                P2 p1, p2;
                if (!point(p1) || !point(p2)) return err;
                builder.quad_to(p1, p2);
                dst = p2;
                kind = 1;
                k1 = p1;
                k2 = p2;
                break;
            }
            case 'T': case 't': {
                const P2 p1 = prev_kind == 1 ? 2.0 * prev_c2 - prev_c1 : position;
                P2 p2;
                if (!point(p2)) return err;
                builder.quad_to(p1, p2);
                dst = p2;
                kind = 1;
                k1 = p1;
                k2 = p2;
                break;
            }
            case 'C': case 'c': {
                P2 p1, p2, p3;
                if (!point(p1) || !point(p2) || !point(p3)) return err;
                builder.cubic_to(p1, p2, p3);
                dst = p3;
                kind = 2;
                k1 = p2;
                k2 = p3;
                break;
            }
            case 'S': case 's': {
                const P2 p1 = prev_kind == 2 ? 2.0 * prev_c2 - prev_c1 : position;
                P2 p2, p3;
                if (!point(p2) || !point(p3)) return err;
                builder.cubic_to(p1, p2, p3);
                dst = p3;
                kind = 2;
                k1 = p2;
                k2 = p3;
                break;
            }
            case 'A': case 'a': {
                double rx, ry, rot;
                bool large, sweep;
                if (!scalar(rx) || !scalar(ry) || !scalar(rot) || !flag(large) || !flag(sweep) || !point(dst)) return err;
                builder.arc_to(mk(rx, ry), rot, large, sweep, dst);
                break;
            }
            default: {  // 'Z' / 'z'
                dst = subpath_start;
                builder.close();
                break;
            }
        }
        position = dst;  // :409
        prev_kind = kind;
        prev_c1 = k1;
        prev_c2 = k2;
    }
    builder.finish();
    return kParseOk;
}

// ---- `fit_size`, src/geometry.rs:490-516, over `Transform::fit_bbox` :470-487 and `Transform` products :519-539 ----
struct Xf {
    double m[6];  // m00 m01 m02 m10 m11 m12
};
SD_FN inline Xf xf(double m00, double m01, double m02, double m10, double m11, double m12) {
    Xf t;
    t.m[0] = m00; t.m[1] = m01; t.m[2] = m02; t.m[3] = m10; t.m[4] = m11; t.m[5] = m12;
    return t;
}
SD_FN inline Xf xf_mul(const Xf& a, const Xf& o) {
    const double* s = a.m;
    return xf(s[0] * o.m[0] + s[1] * o.m[3], s[0] * o.m[1] + s[1] * o.m[4], s[0] * o.m[2] + s[1] * o.m[5] + s[2],
              s[3] * o.m[0] + s[4] * o.m[3], s[3] * o.m[1] + s[4] * o.m[4], s[3] * o.m[2] + s[4] * o.m[5] + s[5]);
}
SD_FN inline Xf xf_translate(double tx, double ty) { return xf(1.0, 0.0, tx, 0.0, 1.0, ty); }
SD_FN inline Xf xf_scale(double sx, double sy) { return xf(sx, 0.0, 0.0, 0.0, sy, 0.0); }
SD_FN inline uint32_t as_size(double v) {  // `as usize`, saturating; sizes are 32-bit here
    if (!(v > 0.0)) return 0;
    if (v >= 4294967295.0) return 0xffffffffu;
    return (uint32_t)v;
}
enum { kAlignMin = 0, kAlignMid = 1, kAlignMax = 2 };  // `Align`, src/geometry.rs:298-305
SD_FN inline void fit_size(Box src, uint32_t want_w, uint32_t want_h, int align, double* tr_out, uint32_t& out_w, uint32_t& out_h) {
    src = box_new(mk(floor(src.lo.x), floor(src.lo.y)), mk(ceil(src.hi.x), ceil(src.hi.y)));
    const double sw = src.hi.x - src.lo.x, sh = src.hi.y - src.lo.y;
    double height, width;
    if (want_h == 0 && want_w == 0) {
        height = sh;
        width = sw;
    } else if (want_w == 0) {
        height = (double)want_h;
        width = ceil(sw * height / sh);
    } else if (want_h == 0) {
        width = (double)want_w;
        height = ceil(sh * width / sw);
    } else {
        height = (double)want_h;
        width = (double)want_w;
    }
    const Box dst = box_new(mk(0.0, 0.0), mk(width, height));
    out_h = as_size(height);
    out_w = as_size(width);
    // Transform::fit_bbox(src, dst, align)
    const double dw = dst.hi.x - dst.lo.x, dh = dst.hi.y - dst.lo.y;
    const double scale = fmin(dh / sh, dw / sw);
    const Xf base = xf_mul(xf_mul(xf_translate(dst.lo.x, dst.lo.y), xf_scale(scale, scale)), xf_translate(-src.lo.x, -src.lo.y));
    Xf al = xf(1.0, 0.0, 0.0, 0.0, 1.0, 0.0);
    if (align == kAlignMid) al = xf_translate((dw - sw * scale) / 2.0, (dh - sh * scale) / 2.0);
    else if (align == kAlignMax) al = xf_translate(dw - sw * scale, dh - sh * scale);
    const Xf r = xf_mul(al, base);
    for (int i = 0; i < 6; i++) tr_out[i] = r.m[i];
}

// ---- output policies ----
struct CountOut {
    uint32_t seg = 0, pts = 0, curves = 0, sub = 0;
    SD_FN void begin_subpath() {}
    SD_FN void segment(const Seg& s) {
        seg++;
        pts += (uint32_t)s.kind;
        curves += s.kind != 2;
    }
    SD_FN void end_subpath(bool) { sub++; }
};
// Writes one path of a device path batch: control points at [pt, ...), its items at [item, ...) of the reference-order list
// and of the curves-first list (this path's curves first, then its lines and closing items).  I2 = uint2.
template <class I2>
struct EmitOut {
    double* pts;
    I2* items;
    I2* packed;
    uint32_t pt, item, curve, rest, sub_first_pt;
    uint32_t closing_flag, closed_flag;
    SD_FN void begin_subpath() { sub_first_pt = pt; }
    SD_FN void segment(const Seg& s) {
        for (int i = 0; i < s.kind; i++) {
            pts[2 * (size_t)(pt + i)] = s.p[i].x;
            pts[2 * (size_t)(pt + i) + 1] = s.p[i].y;
        }
        I2 it;
        it.x = pt;
        it.y = (uint32_t)s.kind;
        items[item++] = it;
        if (s.kind != 2) packed[curve++] = it;
        else packed[rest++] = it;
        pt += (uint32_t)s.kind;
    }
    SD_FN void end_subpath(bool closed) {
        I2 it;
        it.x = pt - 1;
        it.y = closing_flag | (closed ? closed_flag : 0u) | sub_first_pt;
        items[item++] = it;
        packed[rest++] = it;
    }
};

}  // namespace sv
}  // namespace rgpu
