// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_geom.hpp header).
// Restatement of src/scene.rs: Scene tree, Pipeline::build / render (:262-459), Layer (:464-566),
// plus a minimal JSON reader for the serde shapes of src/scene.rs:13-63, 587-640 and
// src/grad.rs:249-269, 449-469.  JSON numbers go through strtod (serde_json 1.x, version unpinned by
// the reference: "parity unpinned" at that boundary, irrelevant at 1e-4 / 1 LSB).
#pragma once
#include "orc_raster.hpp"
#include <map>
#include <memory>

namespace orc {

// ---- minimal JSON ------------------------------------------------------------------------
struct Json {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;
    const Json* get(const std::string& key) const {
        for (auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
};
struct JsonParser {
    const char* s; size_t n, pos = 0;
    JsonParser(const char* s_, size_t n_) : s(s_), n(n_) {}
    void ws() { while (pos < n && (s[pos] == ' ' || s[pos] == '\t' || s[pos] == '\n' || s[pos] == '\r')) pos++; }
    [[noreturn]] void fail(const char* what) { throw ParseError(std::string("json: ") + what + " at " + std::to_string(pos)); }
    Json parse() { Json v = value(); ws(); if (pos != n) fail("trailing data"); return v; }
    Json value() {
        ws();
        if (pos >= n) fail("eof");
        char c = s[pos];
        Json v;
        if (c == '{') {
            v.type = Json::Object; pos++; ws();
            if (pos < n && s[pos] == '}') { pos++; return v; }
            while (true) {
                ws();
                Json k = string_();
                ws();
                if (pos >= n || s[pos] != ':') fail("':' expected");
                pos++;
                v.obj.emplace_back(k.str, value());
                ws();
                if (pos < n && s[pos] == ',') { pos++; continue; }
                if (pos < n && s[pos] == '}') { pos++; break; }
                fail("',' or '}' expected");
            }
        } else if (c == '[') {
            v.type = Json::Array; pos++; ws();
            if (pos < n && s[pos] == ']') { pos++; return v; }
            while (true) {
                v.arr.push_back(value());
                ws();
                if (pos < n && s[pos] == ',') { pos++; continue; }
                if (pos < n && s[pos] == ']') { pos++; break; }
                fail("',' or ']' expected");
            }
        } else if (c == '"') {
            v = string_();
        } else if (c == 't' && n - pos >= 4 && !std::strncmp(s + pos, "true", 4)) { v.type = Json::Bool; v.b = true; pos += 4; }
        else if (c == 'f' && n - pos >= 5 && !std::strncmp(s + pos, "false", 5)) { v.type = Json::Bool; v.b = false; pos += 5; }
        else if (c == 'n' && n - pos >= 4 && !std::strncmp(s + pos, "null", 4)) { v.type = Json::Null; pos += 4; }
        else {
            std::string tmp;
            size_t p = pos;
            while (p < n && (std::isdigit((unsigned char)s[p]) || s[p] == '-' || s[p] == '+' || s[p] == '.' || s[p] == 'e' || s[p] == 'E')) tmp.push_back(s[p++]);
            if (tmp.empty()) fail("value expected");
            char* end = nullptr;
            v.type = Json::Number;
            v.num = std::strtod(tmp.c_str(), &end);
            pos = p;
        }
        return v;
    }
    Json string_() {
        if (pos >= n || s[pos] != '"') fail("string expected");
        pos++;
        Json v; v.type = Json::String;
        while (pos < n && s[pos] != '"') {
            if (s[pos] == '\\' && pos + 1 < n) {
                char e = s[pos + 1];
                pos += 2;
                switch (e) {
                    case 'n': v.str.push_back('\n'); break;
                    case 't': v.str.push_back('\t'); break;
                    case 'r': v.str.push_back('\r'); break;
                    default: v.str.push_back(e); break;
                }
            } else {
                v.str.push_back(s[pos++]);
            }
        }
        if (pos >= n) fail("unterminated string");
        pos++;
        return v;
    }
};

// ---- Scene tree: src/scene.rs:19-63 -----------------------------------------------------
struct Scene;
using ScenePtr = std::shared_ptr<const Scene>;
struct Scene {
    enum Kind { Fill, Stroke, Group, TransformK, Opacity, Clip } kind = Group;
    FillRule fill_rule = FillRule::NonZero;
    std::shared_ptr<const Paint> paint;
    std::shared_ptr<const Path> path;  // Fill/Stroke path or Clip path
    StrokeStyle style;
    std::vector<ScenePtr> children;
    Transform tr;
    ScenePtr child;
    Scalar opacity = 1.0;
    Units units = Units::UserSpaceOnUse;

    static ScenePtr fill(std::shared_ptr<const Path> path, std::shared_ptr<const Paint> paint, FillRule rule) {
        auto s = std::make_shared<Scene>(); s->kind = Fill; s->path = path; s->paint = paint; s->fill_rule = rule; return s;
    }
    // :104-109 — single-child groups collapse
    static ScenePtr group(std::vector<ScenePtr> children) {
        if (children.size() == 1) return children[0];
        auto s = std::make_shared<Scene>(); s->kind = Group; s->children = std::move(children); return s;
    }
    // :139-151 — nested transforms collapse
    static ScenePtr transform(const ScenePtr& self, const Transform& tr) {
        if (self->kind == TransformK) return transform(self->child, tr * self->tr);
        auto s = std::make_shared<Scene>(); s->kind = TransformK; s->child = self; s->tr = tr; return s;
    }
    // :154-183
    std::optional<BBox> bbox(const Transform& t) const {
        switch (kind) {
            case Fill: case Stroke: return path->bbox(t);
            case Group: {
                std::optional<BBox> bb;
                for (auto& c : children) {
                    if (bb) bb = bb->union_opt(c->bbox(t)); else bb = c->bbox(t);
                }
                return bb;
            }
            case TransformK: return child->bbox(t * tr);
            case Opacity: return child->bbox(t);
            default: {
                Transform clip_tr = t;
                if (units == Units::BoundingBox) {
                    auto bb = child->bbox(Transform::identity());
                    if (!bb) return std::nullopt;
                    clip_tr = t * bb->unit_transform();
                }
                auto cb = child->bbox(t);
                if (!cb) return std::nullopt;
                auto pb = path->bbox(clip_tr);
                if (!pb) return std::nullopt;
                return cb->intersect(*pb);
            }
        }
    }
};

// ---- Layer: src/scene.rs:464-566 --------------------------------------------------------
template <class C>
struct Layer {
    Shape shape;
    std::vector<C> data;
    int32_t x = 0, y = 0;
    Layer() = default;
    Layer(const BBox& bbox, const std::optional<C>& color) {  // :483-501
        int32_t x0 = as_i32(std::floor(bbox.min.x));
        int32_t x1 = as_i32(std::ceil(bbox.max.x));
        int32_t y0 = as_i32(std::floor(bbox.min.y));
        int32_t y1 = as_i32(std::ceil(bbox.max.y));
        size_t w = (size_t)(int64_t)(x1 - x0), h = (size_t)(int64_t)(y1 - y0);
        shape = Shape::simple(h, w);
        data.assign(w * h, color ? *color : C());
        x = x0; y = y0;
    }
    size_t width() const { return shape.width; }
    size_t height() const { return shape.height; }
    // :532-565 (loop order does not matter: pixels are independent)
    template <class CO, class F>
    void compose(const Layer<CO>& other, F f) {
        int32_t x0 = std::max(x, other.x);
        int32_t x1 = std::min(x + (int32_t)width(), other.x + (int32_t)other.width());
        int32_t y0 = std::max(y, other.y);
        int32_t y1 = std::min(y + (int32_t)height(), other.y + (int32_t)other.height());
        for (int32_t yy = y0; yy < y1; yy++)
            for (int32_t xx = x0; xx < x1; xx++) {
                C& dst = data[shape.offset((size_t)(yy - y), (size_t)(xx - x))];
                const CO& src = other.data[other.shape.offset((size_t)(yy - other.y), (size_t)(xx - other.x))];
                dst = f(dst, src);
            }
    }
};

// ---- Pipeline: src/scene.rs:225-459 -----------------------------------------------------
struct PipelineNode {
    enum Kind { Fill, Group, Opacity, Clip } kind;
    std::shared_ptr<const Path> path;     // Fill path / Clip path
    std::shared_ptr<const Paint> paint;
    FillRule fill_rule = FillRule::NonZero;
    std::vector<size_t> children;
    size_t child = 0;
    Scalar opacity = 1.0;
    Transform clip_tr;
    BBox bbox;
    Transform tr;
};

struct Pipeline {
    std::vector<PipelineNode> nodes;
    Scalar flatness = DEFAULT_FLATNESS;

    static std::optional<BBox> view_apply(const std::optional<BBox>& view, const std::optional<BBox>& bbox) {  // :277-282
        if (!view) return bbox;
        if (!bbox) return std::nullopt;
        return view->intersect(*bbox);
    }
    size_t alloc(PipelineNode n) { nodes.push_back(std::move(n)); return nodes.size() - 1; }

    // :268-357
    std::optional<size_t> build_rec(const Scene& scene, const std::optional<BBox>& view, const Transform& tr) {
        switch (scene.kind) {
            case Scene::Fill: {
                auto bb = view_apply(view, scene.path->bbox(tr));
                if (!bb) return std::nullopt;
                PipelineNode n; n.kind = PipelineNode::Fill; n.path = scene.path; n.paint = scene.paint; n.fill_rule = scene.fill_rule;
                n.bbox = *bb; n.tr = tr;
                return alloc(std::move(n));
            }
            case Scene::Stroke: {
                auto stroked = std::make_shared<Path>(scene.path->stroke(scene.style));
                auto bb = view_apply(view, stroked->bbox(tr));
                if (!bb) return std::nullopt;
                PipelineNode n; n.kind = PipelineNode::Fill; n.path = stroked; n.paint = scene.paint; n.fill_rule = FillRule::NonZero;
                n.bbox = *bb; n.tr = tr;
                return alloc(std::move(n));
            }
            case Scene::Group: {
                std::vector<size_t> children;
                for (auto& c : scene.children) {
                    auto id = build_rec(*c, view, tr);
                    if (id) children.push_back(*id);
                }
                std::optional<BBox> bb;
                for (size_t id : children) bb = nodes[id].bbox.union_opt(bb);
                if (!bb) return std::nullopt;
                PipelineNode n; n.kind = PipelineNode::Group; n.children = std::move(children); n.bbox = *bb; n.tr = tr;
                return alloc(std::move(n));
            }
            case Scene::Clip: {
                Transform clip_tr = tr;
                if (scene.units == Units::BoundingBox) {
                    auto bb = scene.child->bbox(Transform::identity());
                    if (!bb) return std::nullopt;
                    clip_tr = tr * bb->unit_transform();
                }
                auto clip_bbox = view_apply(view, scene.path->bbox(clip_tr));
                if (!clip_bbox) return std::nullopt;
                auto child_id = build_rec(*scene.child, clip_bbox, tr);
                if (!child_id) return std::nullopt;
                auto bb = clip_bbox->intersect(nodes[*child_id].bbox);
                if (!bb) return std::nullopt;
                PipelineNode n; n.kind = PipelineNode::Clip; n.child = *child_id; n.path = scene.path; n.clip_tr = clip_tr;
                n.fill_rule = scene.fill_rule; n.bbox = *bb; n.tr = tr;
                return alloc(std::move(n));
            }
            case Scene::Opacity: {
                auto child_id = build_rec(*scene.child, view, tr);
                if (!child_id) return std::nullopt;
                PipelineNode n; n.kind = PipelineNode::Opacity; n.child = *child_id; n.opacity = scene.opacity;
                n.bbox = nodes[*child_id].bbox; n.tr = tr;
                return alloc(std::move(n));
            }
            default: return build_rec(*scene.child, view, tr * scene.tr);
        }
    }

    // :384-395
    Layer<LinColor> render(size_t node_id, const std::optional<BBox>& view, const std::optional<LinColor>& bg) const {
        Layer<LinColor> layer(view ? *view : nodes[node_id].bbox, bg);
        render_rec(node_id, layer);
        return layer;
    }

    // :397-459
    void render_rec(size_t node_id, Layer<LinColor>& layer) const {
        const PipelineNode& node = nodes[node_id];
        switch (node.kind) {
            case PipelineNode::Fill: {
                int32_t col_min = as_i32(std::floor(node.bbox.min.x)) - layer.x;
                int32_t col_max = as_i32(std::ceil(node.bbox.max.x)) - layer.x + 1;
                int32_t row_min = as_i32(std::floor(node.bbox.min.y)) - layer.y;
                int32_t row_max = as_i32(std::ceil(node.bbox.max.y)) - layer.y + 1;
                // `as usize` of a negative i32 sign-extends to a huge value which view_shape clamps
                Shape view = view_shape(layer.shape, (size_t)(int64_t)row_min, (size_t)(int64_t)row_max, (size_t)(int64_t)col_min,
                                        (size_t)(int64_t)col_max);
                Transform align = Transform::new_translate(-std::floor(node.bbox.min.x), -std::floor(node.bbox.min.y));
                fill(*node.path, align * node.tr, flatness, node.fill_rule, *node.paint, layer.data.data(), view);
                break;
            }
            case PipelineNode::Group:
                for (size_t c : node.children) render_rec(c, layer);
                break;
            case PipelineNode::Opacity: {
                Layer<LinColor> child_layer = render(node.child, std::nullopt, std::nullopt);
                float opacity = (float)node.opacity;
                layer.compose(child_layer, [&](const LinColor& dst, const LinColor& src) { return dst.blend_over(src.scale(opacity)); });
                break;
            }
            case PipelineNode::Clip: {
                Layer<Scalar> mask_layer(node.bbox, std::nullopt);
                Transform align = Transform::new_translate(-(Scalar)mask_layer.x, -(Scalar)mask_layer.y);
                Layer<LinColor> child_layer = render(node.child, std::nullopt, std::nullopt);
                mask(*node.path, align * node.clip_tr, flatness, node.fill_rule, mask_layer.data.data(), mask_layer.data.size(), mask_layer.shape);
                child_layer.compose(mask_layer, [](const LinColor& dst, const Scalar& src) { return dst.scale((float)src); });
                layer.compose(child_layer, [](const LinColor& dst, const LinColor& src) { return dst.blend_over(src); });
                break;
            }
        }
    }
};

// `Scene::render`, src/scene.rs:186-199
inline Layer<LinColor> scene_render(const Scene& scene, Scalar flatness, const Transform& tr, const std::optional<BBox>& view,
                                    const std::optional<LinColor>& bg) {
    Pipeline p;
    p.flatness = flatness;
    p.build_rec(scene, view, tr);
    if (p.nodes.empty()) return Layer<LinColor>();
    return p.render(p.nodes.size() - 1, view, bg);
}

// ---- JSON -> Scene ----------------------------------------------------------------------
inline Point json_point(const Json& j) {
    if (j.type != Json::Array || j.arr.size() != 2) throw ParseError("point expected");
    return Point(j.arr[0].num, j.arr[1].num);
}
inline LinColor parse_lin_color(const std::string& s) {  // LinColor::from_str, src/color.rs:414-420
    RGBA c;
    if (!parse_rgba(s, c)) throw ParseError("bad color: " + s);
    return rgba_to_lin(c);
}
inline std::shared_ptr<const Paint> json_paint(const Json& j) {  // src/scene.rs:604-639
    if (j.type == Json::String) return std::make_shared<Paint>(Paint::make_solid(parse_lin_color(j.str)));
    if (j.type != Json::Object) throw ParseError("failed to parse paint");
    const Json* type = j.get("type");
    if (!type) throw ParseError("paint: missing type");
    Units units = Units::UserSpaceOnUse;
    if (auto u = j.get("units")) {
        if (u->str == "userSpaceOnUse") units = Units::UserSpaceOnUse;
        else if (u->str == "objectBoundingBox") units = Units::BoundingBox;
        else throw ParseError("bad units");
    }
    bool linear_colors = false;
    if (auto l = j.get("linear_colors")) linear_colors = l->b;
    GradSpread spread = GradSpread::Pad;
    if (auto s = j.get("spread")) {
        if (s->str == "pad") spread = GradSpread::Pad;
        else if (s->str == "repeat") spread = GradSpread::Repeat;
        else if (s->str == "reflect") spread = GradSpread::Reflect;
        else throw ParseError("bad spread");
    }
    Transform tr;
    if (auto t = j.get("tr")) tr = parse_transform(t->str.data(), t->str.size());
    std::vector<GradStop> stops;
    if (auto st = j.get("stops")) {
        for (auto& e : st->arr) {
            if (e.type != Json::Array || e.arr.size() != 2) throw ParseError("bad stop");
            stops.push_back({e.arr[0].num, parse_lin_color(e.arr[1].str)});
        }
    } else throw ParseError("missing stops");
    // GradStops deserializes `transparent` as the raw vector — no sort on the serde path
    // (src/grad.rs:79-83); GradLinear::new / GradRadial::new take `impl Into<GradStops>` = identity here.
    GradStops gs;
    gs.stops = stops;
    if (type->str == "linear-gradient") {
        return std::make_shared<Paint>(Paint::make_linear(gs, units, linear_colors, spread, tr, json_point(*j.get("start")), json_point(*j.get("end"))));
    } else if (type->str == "radial-gradient") {
        Point center = json_point(*j.get("center"));
        Scalar radius = j.get("radius")->num;
        Point fcenter = center;
        if (auto f = j.get("fcenter")) if (f->type == Json::Array) fcenter = json_point(*f);
        Scalar fradius = 0.0;
        if (auto f = j.get("fradius")) fradius = f->num;
        return std::make_shared<Paint>(Paint::make_radial(gs, units, linear_colors, spread, tr, center, radius, fcenter, fradius));
    }
    throw ParseError("unknown paint type: " + type->str);
}
inline FillRule json_fill_rule(const Json* j) {
    if (!j) return FillRule::NonZero;
    if (j->str == "nonzero") return FillRule::NonZero;
    if (j->str == "evenodd") return FillRule::EvenOdd;
    throw ParseError("bad fill_rule");
}
inline ScenePtr json_scene(const Json& j) {
    if (j.type != Json::Object) throw ParseError("scene: object expected");
    const Json* type = j.get("type");
    if (!type) throw ParseError("scene: missing type");
    auto s = std::make_shared<Scene>();
    auto get_path = [&](const char* key) {
        const Json* p = j.get(key);
        if (!p || p->type != Json::String) throw ParseError(std::string("scene: missing ") + key);
        return std::make_shared<const Path>(path_from_svg(p->str.data(), p->str.size()));
    };
    const std::string& t = type->str;
    if (t == "fill") {
        s->kind = Scene::Fill; s->fill_rule = json_fill_rule(j.get("fill_rule")); s->paint = json_paint(*j.get("paint")); s->path = get_path("path");
    } else if (t == "stroke") {
        s->kind = Scene::Stroke; s->paint = json_paint(*j.get("paint")); s->path = get_path("path");
        s->style.width = j.get("width")->num;
        if (auto lj = j.get("line_join")) {
            if (lj->type == Json::String) {
                if (lj->str == "bevel") s->style.line_join = LineJoin::Bevel;
                else if (lj->str == "round") s->style.line_join = LineJoin::Round;
            } else if (lj->type == Json::Object) {
                if (auto m = lj->get("miter")) { s->style.line_join = LineJoin::Miter; s->style.miter_limit = m->num; }
            }
        }
        if (auto lc = j.get("line_cap")) {
            if (lc->str == "butt") s->style.line_cap = LineCap::Butt;
            else if (lc->str == "square") s->style.line_cap = LineCap::Square;
            else if (lc->str == "round") s->style.line_cap = LineCap::Round;
        }
    } else if (t == "group") {
        s->kind = Scene::Group;
        for (auto& c : j.get("children")->arr) s->children.push_back(json_scene(c));
    } else if (t == "transform") {
        s->kind = Scene::TransformK;
        const Json* tr = j.get("tr");
        s->tr = parse_transform(tr->str.data(), tr->str.size());
        s->child = json_scene(*j.get("child"));
    } else if (t == "opacity") {
        s->kind = Scene::Opacity; s->opacity = j.get("opacity")->num; s->child = json_scene(*j.get("child"));
    } else if (t == "clip") {
        s->kind = Scene::Clip; s->fill_rule = json_fill_rule(j.get("fill_rule"));
        if (auto u = j.get("units")) s->units = (u->str == "objectBoundingBox") ? Units::BoundingBox : Units::UserSpaceOnUse;
        s->path = get_path("clip"); s->child = json_scene(*j.get("child"));
    } else {
        throw ParseError("scene: unknown type " + t);
    }
    return s;
}
inline ScenePtr scene_from_json(const char* text, size_t n) {
    JsonParser jp(text, n);
    return json_scene(jp.parse());
}

}  // namespace orc
