/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * C API of the CPU restatement of aslpavel/rasterize's fill pipeline (flatten -> signed-difference
 * raster -> paint/composite).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker / CPU baseline; the product
 * (rasterize_b200/) never links, imports or executes it.
 *
 * Parity pin: the restatement is checked against every known-answer test the reference holds for
 * this path (tests/test_oracle_kat.py, vectors from src/rasterize.rs:946-1161, src/path.rs:1098-1269,
 * src/grad.rs:515-602, src/color.rs:539-555, src/scene.rs:669-693, src/svg.rs:668-695).  The Rust
 * reference itself cannot be built here (no cargo/rustc), so there is no oracle/_ref.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_path orc_path;
typedef struct orc_paint orc_paint;
typedef struct orc_scene orc_scene;
typedef struct orc_layer orc_layer;

typedef struct { size_t start, width, height, row_stride, col_stride; } orc_shape; /* src/image.rs:6-17 */
typedef struct { size_t x, y; double alpha; } orc_pixel;                            /* rasterize::Pixel */

const char* orc_last_error(void);
void orc_set_simd_x86(int on); /* 1 = src/simd/x86.rs colour maths (default), 0 = src/simd/fallback.rs */

/* ---- paths ---- */
orc_path* orc_path_parse(const char* svg, size_t len);
orc_path* orc_path_from_flat(const double* pts, const uint8_t* kinds, size_t n_segs, const uint32_t* sub_off, size_t n_sub,
                             const uint8_t* closed);
void orc_path_free(orc_path*);
void orc_path_counts(const orc_path*, size_t* n_segs, size_t* n_points, size_t* n_subpaths);
/* pts: 2*n_points doubles; kinds: n_segs (2,3,4 = control points); sub_off: n_subpaths+1; closed: n_subpaths */
void orc_path_export(const orc_path*, double* pts, uint8_t* kinds, uint32_t* sub_off, uint8_t* closed);
int orc_path_bbox(const orc_path*, const double tr[6], double out_minmax[4]); /* 0 = empty */
int orc_path_size(const orc_path*, const double tr[6], size_t* w, size_t* h, double tr_out[6], double min_out[2]);
orc_path* orc_path_stroke(const orc_path*, double width, int join /*0 miter,1 bevel,2 round*/, double miter_limit,
                          int cap /*0 butt,1 square,2 round*/);
orc_path* orc_path_transformed(const orc_path*, const double tr[6]);
orc_path* orc_path_checkerboard(const double bbox_minmax[4], double cell);
orc_path* orc_path_circle(double cx, double cy, double r);

/* ---- geometry ---- */
void orc_fit_size(const double bbox_minmax[4], size_t w, size_t h, int align /*0 min,1 mid,2 max*/, size_t* ow, size_t* oh,
                  double tr_out[6]);
int orc_transform_parse(const char* text, double out[6]);
void orc_transform_mul(const double a[6], const double b[6], double out[6]);
int orc_transform_invert(const double a[6], double out[6]);
double orc_parse_scalar(const char* text, size_t len, size_t* consumed); /* NaN on error */

/* ---- hot path ---- */
/* returns the number of lines; writes min(count, cap) lines as (x0,y0,x1,y1). -1 on NaN input. */
long orc_flatten(const orc_path*, const double tr[6], double flatness, int close, double* lines_out, size_t cap);
void orc_signed_difference_line(double* data, size_t data_len, orc_shape shape, const double line[4]);
void orc_signed_difference_to_mask(double* data, orc_shape shape, int fill_rule);
int orc_mask(const orc_path*, const double tr[6], double flatness, int fill_rule, double* data, size_t data_len, orc_shape shape);
/* band-parallel variant of orc_mask over a dense row-major image (rows are independent) */
int orc_mask_threads(const orc_path*, const double tr[6], double flatness, int fill_rule, double* data, size_t w, size_t h,
                     int threads);
long orc_mask_iter(const orc_path*, const double tr[6], double flatness, size_t w, size_t h, int fill_rule, orc_pixel* out,
                   size_t cap);
int orc_fill(const orc_path*, const double tr[6], double flatness, int fill_rule, const orc_paint*, float* data /*LinColor*/,
             orc_shape shape);

/* ---- colours & paints ---- */
void orc_rgba_to_lin(const uint8_t rgba[4], float out[4]);
void orc_lin_to_rgba(const float lin[4], uint8_t out[4]);
void orc_lin_to_rgba_image(const float* lin, size_t n_pixels, uint8_t* out);
int orc_parse_color(const char* text, float out_lin[4]);
void orc_l2s(const float in[4], float out[4]);
void orc_s2l(const float in[4], float out[4]);
float orc_linear_to_srgb(float v);
float orc_srgb_to_linear(float v);
double orc_spread_at(int spread, double t);

orc_paint* orc_paint_solid(const float lin[4]);
/* stops: n x (pos f64), colors: n x 4 f32 LINEAR premultiplied colours as given to GradLinear::new */
orc_paint* orc_paint_linear(const double* stop_pos, const float* stop_colors, size_t n, int sort_stops, int units,
                            int linear_colors, int spread, const double tr[6], const double start[2], const double end[2]);
orc_paint* orc_paint_radial(const double* stop_pos, const float* stop_colors, size_t n, int sort_stops, int units,
                            int linear_colors, int spread, const double tr[6], const double center[2], double radius,
                            const double fcenter[2], double fradius);
/* same, but stop colours are given in STORED space (already converted when !linear_colors) and kept verbatim */
orc_paint* orc_paint_linear_stored(const double* stop_pos, const float* stop_colors, size_t n, int units, int linear_colors, int spread,
                                   const double tr[6], const double start[2], const double end[2]);
orc_paint* orc_paint_radial_stored(const double* stop_pos, const float* stop_colors, size_t n, int units, int linear_colors, int spread,
                                   const double tr[6], const double center[2], double radius, const double fcenter[2], double fradius);
void orc_paint_free(orc_paint*);
void orc_paint_at(const orc_paint*, double x, double y, float out[4]);
int orc_paint_radial_offset(const orc_paint*, double x, double y, double* out); /* 0 = None */
/* flat description of a paint for the GPU boundary (stored-space stop colours, i.e. post convert_to_srgb) */
typedef struct {
    int kind, units, linear_colors, spread;
    double tr[6];
    double p0[2], p1[2]; /* linear: start,end ; radial: center,fcenter */
    double dir[2];       /* linear: precomputed dir */
    double r0, r1;       /* radial: radius, fradius */
    float solid[4];
    size_t n_stops;
} orc_paint_desc;
void orc_paint_describe(const orc_paint*, orc_paint_desc* out);
void orc_paint_stops(const orc_paint*, double* pos, float* colors);

/* ---- scenes ---- */
orc_scene* orc_scene_load_json(const char* text, size_t len);
orc_scene* orc_scene_cli_rasterize(const orc_path* path, const double tr[6], size_t w, size_t h); /* examples/rasterize.rs:277-308 */
orc_scene* orc_scene_many_circles(uint32_t seed, size_t count, size_t size);                     /* benches/scene_bench.rs:6-32 */
void orc_scene_free(orc_scene*);
int orc_scene_bbox(const orc_scene*, const double tr[6], double out_minmax[4]);
orc_layer* orc_scene_render(const orc_scene*, double flatness, const double tr[6], const double* view_minmax /*nullable*/,
                            const float* bg_lin /*nullable*/);
void orc_layer_info(const orc_layer*, int32_t* x, int32_t* y, size_t* w, size_t* h);
const float* orc_layer_data(const orc_layer*);
void orc_layer_free(orc_layer*);
/* Flattened list of the Fill nodes Pipeline::build produces, in render order (only valid for scenes made
 * of fill/stroke/group/transform nodes; returns -1 if the scene has clip/opacity nodes). */
typedef struct {
    orc_path* path;   /* borrowed */
    orc_paint* paint; /* borrowed */
    int fill_rule;
    double tr[6];     /* node.tr (before the view alignment translate) */
    double bbox[4];   /* node.bbox (already intersected with the view) */
} orc_fill_job;
long orc_scene_fill_jobs(const orc_scene*, const double tr[6], const double* view_minmax, orc_fill_job* out, size_t cap);

/* The whole node table Pipeline::build produces (src/scene.rs:268-357), in allocation order: children come before their
 * parents and the root is the last node.  kind: 0 Fill, 1 Group, 2 Opacity, 3 Clip.  Group children are
 * children[child_begin .. child_begin + child_count); Opacity / Clip have one child in `child`. */
typedef struct {
    int kind;
    orc_path* path;   /* borrowed: Fill path / Clip path, else NULL */
    orc_paint* paint; /* borrowed: Fill paint, else NULL */
    int fill_rule;
    double tr[6];     /* Fill: node.tr ; Clip: node.clip_tr */
    double bbox[4];   /* node.bbox */
    double opacity;
    long child;
    long child_begin, child_count;
} orc_pipe_node;
/* returns the node count (call with out == NULL to size); *n_children_out = total entries of `children` */
long orc_scene_pipeline(const orc_scene*, const double tr[6], const double* view_minmax, orc_pipe_node* out, size_t cap,
                        long* children, size_t children_cap, size_t* n_children_out);

/* LCG of benches/scene_bench.rs:53-88 */
double orc_lcg_uniform(uint32_t* state);
/* synthetic glyph of SURVEY §8d C4: 3 closed contours x 6 cubics, coords uniform()*56+4, seed = index+1 */
orc_path* orc_glyph(uint32_t seed);
/* n independent paths on `threads` host threads, each into a private w x h image: clear + mask (paint NULL) or clear + fill */
int orc_batch_threads(const orc_path* const* paths, size_t n, const double tr[6], double flatness, int fill_rule, const orc_paint* paint,
                      size_t w, size_t h, int threads);

#ifdef __cplusplus
}
#endif
#endif
