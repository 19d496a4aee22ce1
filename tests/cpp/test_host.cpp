// C++ restatement of the reference's own rasterizer tests (src/rasterize.rs:1065-1161: test_rasterizer and
// test_fill_rule) with the GpuRasterizer in the loop, through include/rasterize_b200.hpp -> the C ABI.
// Built and run by tests/test_gpu_cpp_host.py on the GPU box.  Exit code 0 = all assertions hold.
#include "rasterize_b200.hpp"

#include <cstdio>
#include <cstdlib>

using namespace rasterize;

#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                  \
        }                                                                  \
    } while (0)

static void test_rasterizer(Rasterizer& rasterizer) {
    const double expected[] = {
        0.0, 1.0, 1.0,  1.0,  1.0, 1.0,  1.0,  1.0, 0.0,
        0.0, 0.5, 1.0,  1.0,  1.0, 1.0,  1.0,  0.5, 0.0,
        0.0, 0.0, 0.25, 0.75, 1.0, 0.75, 0.25, 0.0, 0.0,
        0.0, 0.0, 0.0,  0.0,  0.0, 0.0,  0.0,  0.0, 0.0,
    };
    Path path = Path::builder()
                    .move_to({1.0, 0.0}).line_to({1.0, 1.0}).line_to({2.0, 2.0}).line_to({4.0, 3.0}).line_to({5.0, 3.0})
                    .line_to({7.0, 2.0}).line_to({8.0, 1.0}).line_to({8.0, 0.0}).close().build();
    std::vector<double> img(9 * 4, 0.0);
    rasterizer.mask(path, Transform::identity(), ImageMut<double>::dense(img.data(), 4, 9), FillRule::EvenOdd);
    for (int i = 0; i < 36; i++) CHECK(std::fabs(img[i] - expected[i]) <= 1e-6);
    // square that goes exactly through the sides of the image
    Path sq = Path::builder().move_to({1, 1}).line_to({1, 3}).line_to({2, 3}).line_to({2, 1}).close().build();
    std::vector<double> img2(9, 0.0);
    rasterizer.mask(sq, Transform::identity(), ImageMut<double>::dense(img2.data(), 3, 3), FillRule::EvenOdd);
}

static void test_fill_rule(Rasterizer& rasterizer) {
    // M50,0 21,90 98,35 2,35 79,90z M110,0 h90 v90 h-90z M130,20 h50 v50 h-50 z M210,0 h90 v90 h-90 z M230,20 v50 h50 v-50 z
    Path path = Path::builder()
                    .move_to({50, 0}).line_to({21, 90}).line_to({98, 35}).line_to({2, 35}).line_to({79, 90}).close()
                    .move_to({110, 0}).line_to({200, 0}).line_to({200, 90}).line_to({110, 90}).close()
                    .move_to({130, 20}).line_to({180, 20}).line_to({180, 70}).line_to({130, 70}).close()
                    .move_to({210, 0}).line_to({300, 0}).line_to({300, 90}).line_to({210, 90}).close()
                    .move_to({230, 20}).line_to({230, 70}).line_to({280, 70}).line_to({280, 20}).close()
                    .build();
    // Path::size: bbox (2,0)-(300,90) -> 300 x 92 image, shift = translate(1 - 2, 1 - 0)
    const size_t w = 300, h = 92;
    const Transform tr = Transform::new_translate(-1.0, 1.0);
    std::vector<double> img(w * h, 0.0);
    auto view = ImageMut<double>::dense(img.data(), h, w);
    const size_t y = 50, x0 = 50, x1 = 150, x2 = 250;
    rasterizer.mask(path, tr, view, FillRule::EvenOdd);
    CHECK(std::fabs(view.at(y, x0)) < 1e-6 && std::fabs(view.at(y, x1)) < 1e-6 && std::fabs(view.at(y, x2)) < 1e-6);
    double area = 0;
    for (double v : img) area += v;
    CHECK(std::fabs(area - 13130.0) < 1.0);
    std::fill(img.begin(), img.end(), 0.0);
    rasterizer.mask(path, tr, view, FillRule::NonZero);
    CHECK(std::fabs(view.at(y, x0) - 1.0) < 1e-6 && std::fabs(view.at(y, x1) - 1.0) < 1e-6 && std::fabs(view.at(y, x2)) < 1e-6);
    area = 0;
    for (double v : img) area += v;
    CHECK(std::fabs(area - 16492.5) < 1.0);
    // mask_iter never yields pixels outside the size and agrees with mask away from the overflow column
    auto px = rasterizer.mask_iter(path, tr, Size{w, h}, FillRule::NonZero);
    CHECK(!px.empty());
    for (const Pixel& p : px) {
        CHECK(p.x < w && p.y < h && std::fabs(p.alpha) >= 1e-6);
        if (p.x + 1 < w) CHECK(std::fabs(p.alpha - view.at(p.y, p.x)) <= 1e-4);
    }
    // fill: solid paint over a transparent LinColor image
    std::vector<float> lin(w * h * 4, 0.f);
    Paint red = Paint::solid(0.5f, 0.0f, 0.0f, 0.5f);
    rasterizer.fill(path, tr, FillRule::NonZero, red, ImageMut<float>{lin.data(), rgpu_shape{0, w, h, w, 1}});
    CHECK(std::fabs(lin[(y * w + x0) * 4 + 3] - 0.5f) < 1e-4f && std::fabs(lin[(y * w + x2) * 4 + 3]) < 1e-6f);
}

// The Opacity and Clip arms of Pipeline::render_rec (src/scene.rs:436-457) on device layers, with solid colours so that
// the expected pixels follow from `blend_over` / `Mul<f32>` by hand.
static void test_layers(GpuRasterizer& r) {
    const float bg[4] = {0.2f, 0.2f, 0.2f, 1.0f};
    DeviceLayer layer(r, 0, 0, 40, 30, 4, bg);
    // child layer at (10, 5), 20 x 10, covered entirely by a rectangle path in scene coordinates
    DeviceLayer child(r, 10, 5, 20, 10, 4);
    Path rect = Path::builder().move_to({5, 0}).line_to({35, 0}).line_to({35, 25}).line_to({5, 25}).close().build();
    Paint red = Paint::solid(0.5f, 0.0f, 0.0f, 0.5f);
    child.fill(rect, Transform::identity(), FillRule::NonZero, red);
    // clip mask layer at (15, 0), 10 x 30: left half of it (x < 20 in scene coordinates) is inside the clip path
    DeviceLayer mask(r, 15, 0, 10, 30, 1);
    Path clip = Path::builder().move_to({0, 0}).line_to({20, 0}).line_to({20, 30}).line_to({0, 30}).close().build();
    mask.mask(clip, Transform::identity(), FillRule::NonZero);
    child.scale_by_mask(mask);  // only columns 15..24 of the scene are touched: 15..19 keep the colour, 20..24 become 0
    const float opacity = 0.5f;
    layer.blend_over(child, &opacity);
    const std::vector<float> px = layer.download();
    auto at = [&](int x, int y, int c) { return px[(size_t)(y * 40 + x) * 4 + c]; };
    // outside the child layer: background
    CHECK(at(2, 2, 0) == 0.2f && at(35, 20, 3) == 1.0f);
    // child pixel left of the mask layer (x = 12): never multiplied by the mask (compose touches the intersection only)
    // src = (0.5,0,0,0.5) * 0.5 = (0.25,0,0,0.25); dst = src + bg * 0.75
    CHECK(std::fabs(at(12, 8, 0) - (0.25f + 0.2f * 0.75f)) < 1e-6f && std::fabs(at(12, 8, 1) - 0.15f) < 1e-6f && std::fabs(at(12, 8, 3) - 1.0f) < 1e-6f);
    // inside the clip (x = 17): mask 1 -> same as above
    CHECK(std::fabs(at(17, 8, 0) - 0.4f) < 1e-6f);
    // masked out (x = 22): child * 0 -> background unchanged
    CHECK(std::fabs(at(22, 8, 0) - 0.2f) < 1e-6f && std::fabs(at(22, 8, 3) - 1.0f) < 1e-6f);
    // RGBA8 export: background 0.2 linear -> l2s(0.2) * 255 + 0.5 = 124 (x86 polynomial), alpha 255
    const std::vector<uint8_t> rgba = layer.download_rgba8();
    CHECK(rgba[3] == 255 && rgba[0] >= 123 && rgba[0] <= 125);
}

// `Scene::render` of a Fill-only pipeline through the scene compositor (one raster launch) against the same fills applied
// one by one with the trait's `fill` on a host image: identical bits, and the RGBA8 export of that image.
static void test_scene(GpuRasterizer& r) {
    const size_t W = 700, H = 40;
    const float bg[4] = {0.1f, 0.1f, 0.1f, 1.0f};
    Path star = Path::builder().move_to({50, 0}).line_to({21, 90}).line_to({98, 35}).line_to({2, 35}).line_to({79, 90}).close().build();
    Path blob = Path::builder().move_to({10, 10}).cubic_to({200, -30}, {400, 80}, {650, 5}).quad_to({300, 60}, {10, 30}).close().build();
    Paint red = Paint::solid(0.5f, 0.0f, 0.0f, 0.5f), green = Paint::solid(0.0f, 0.3f, 0.0f, 0.3f);
    // windows at unaligned offsets inside the layer; paths in window-local coordinates
    std::vector<GpuRasterizer::SceneFill> fills = {
        {&blob, Transform::identity(), FillRule::NonZero, &red, 13, 3, 680, 36},
        {&star, Transform::new_translate(-3.0, -20.0), FillRule::EvenOdd, &green, 301, 7, 95, 30},
        {&blob, Transform::new_translate(-100.0, 0.0), FillRule::EvenOdd, &green, 0, 0, 700, 40},
    };
    std::vector<float> lin;
    const std::vector<uint8_t> rgba = r.render_scene(fills, W, H, bg, &lin);
    std::vector<float> ref(W * H * 4);
    for (size_t i = 0; i < W * H; i++) for (int c = 0; c < 4; c++) ref[i * 4 + c] = bg[c];
    for (auto& f : fills) {
        ImageMut<float> view{ref.data(), rgpu_shape{(size_t)f.y * W + f.x, f.width, f.height, W, 1}};
        r.fill(*f.path, f.tr, f.fill_rule, *f.paint, view);
    }
    bool drew = false;
    for (size_t i = 0; i < ref.size(); i++) {
        CHECK(ref[i] == lin[i]);
        drew = drew || (i % 4 == 0 && ref[i] > 0.3f);
    }
    CHECK(drew);
    CHECK(rgba[3] == 255 && rgba.size() == W * H * 4);
}

// src/path.rs:1225-1269 test_stroke, first case: "M2,2L8,2C11,2 11,8 8,8L5,4" stroked with width 1 (miter / butt)
static void test_stroke(GpuRasterizer& r) {
    Path path = Path::builder().move_to({2, 2}).line_to({8, 2}).cubic_to({11, 2}, {11, 8}, {8, 8}).line_to({5, 4}).build();
    StrokeStyle style;
    style.width = 1.0;
    Path s = r.stroke(path, style);
    const uint8_t kinds[] = {2, 4, 4, 2, 2, 2, 2, 2, 2, 4, 4, 2, 2};
    const double pts[] = {2, 1.5, 8, 1.5, 8, 1.5, 9.80902, 1.5, 10.75, 3.38197, 10.75, 5, 10.75, 5, 10.75, 6.61803, 9.80902, 8.5, 8, 8.5,
                          8, 8.5, 7.75, 8.5, 7.75, 8.5, 7.6, 8.3, 7.6, 8.3, 4.6, 4.3, 4.6, 4.3, 5.4, 3.7, 5.4, 3.7, 8.4, 7.7, 8.4, 7.7, 8, 7.5,
                          8, 7.5, 9.19098, 7.5, 9.75, 6.38197, 9.75, 5, 9.75, 5, 9.75, 3.61803, 9.19098, 2.5, 8, 2.5, 8, 2.5, 2, 2.5, 2, 2.5, 2, 1.5};
    CHECK(s.kinds.size() == sizeof(kinds));
    for (size_t i = 0; i < sizeof(kinds); i++) CHECK(s.kinds[i] == kinds[i]);
    CHECK(s.closed.size() == 1 && s.closed[0] == 1 && s.subpath_offsets[1] == sizeof(kinds));
    CHECK(s.points.size() == sizeof(pts) / sizeof(double));
    for (size_t i = 0; i < s.points.size(); i++) CHECK(std::fabs(s.points[i] - pts[i]) < 1e-4);
}

// src/path.rs:1155-1176 (test_path_parse, the small cases) and :1098-1105 (test_bbox) through the batch parser
static void test_parse(GpuRasterizer& r) {
    auto b = r.parse_svg_batch({" M0,0L1-1L1,0ZL0,1 L1,1Z ", "M.5-3-11-.11", " m.5,-3 -11.5\n2.89 ", "M0,0 L", "M12 1C9.79 1 8 2.31 8 3.92"}, 64, 64, RGPU_ALIGN_MID);
    CHECK(b.paths.size() == 5);
    const Path& p0 = b.paths[0];
    CHECK(p0.kinds.size() == 4 && p0.closed.size() == 2 && p0.closed[0] == 1 && p0.closed[1] == 1);
    const double want0[] = {0, 0, 1, -1, 1, -1, 1, 0, 0, 0, 0, 1, 0, 1, 1, 1};
    CHECK(p0.points.size() == 16);
    for (int i = 0; i < 16; i++) CHECK(p0.points[i] == want0[i]);
    for (int k = 1; k <= 2; k++) {
        const Path& p = b.paths[k];
        CHECK(p.kinds.size() == 1 && p.kinds[0] == 2 && p.closed.size() == 1 && p.closed[0] == 0);
        const double want[] = {0.5, -3.0, -11.0, -0.11};
        for (int i = 0; i < 4; i++) CHECK(std::fabs(p.points[i] - want[i]) < 1e-12);
    }
    CHECK(b.info[3].status == RGPU_PARSE_INVALID_SCALAR && b.info[3].error_offset == 6 && b.paths[3].is_empty());
    CHECK(b.info[4].has_bbox && b.info[4].bbox[0] == 8.0 && b.info[4].bbox[2] == 12.0 && b.info[4].fit_width == 64);
}

int main() {
    GpuRasterizer r;
    CHECK(std::string(r.name()) == "gpu-signed-difference");
    test_rasterizer(r);
    test_fill_rule(r);
    test_layers(r);
    test_scene(r);
    test_stroke(r);
    test_parse(r);
    // NaN control point -> error (reference panics, src/path.rs:765-767)
    Path bad = Path::builder().move_to({0, 0}).quad_to({std::nan(""), 1}, {2, 2}).build();
    bool threw = false;
    try { r.flatten(bad, Transform::identity()); } catch (const Error& e) { threw = e.code == RGPU_ERR_NAN; }
    CHECK(threw);
    std::puts("cpp host tests ok");
    return 0;
}
