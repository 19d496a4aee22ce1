"""Generates the committed fixtures under tests/golden/ from the reference's data/ assets.

Run in the dev container only (needs /root/reference):  python tests/golden/make_golden.py
The SVG / JSON text is parsed by the CPU oracle (oracle/, a restatement of the reference's own parser and
builder, pinned in tests/test_oracle_kat.py) and stored in the flat encoding that crosses the C ABI, so the
GPU box needs neither /root/reference nor a parser.

  paths/<name>.npz      points[n,2] f64, kinds[n] u8, subpath_offsets[s+1] u32, closed[s] u8
  scenes/<name>.npz     the Fill jobs `Pipeline::build` produces for one render transform + view, in render
                        order: per job path arrays, node transform, fill rule, bbox, paint description
  expected.json         oracle-derived known answers (line counts, coverage sums, RGBA digests) that pin the
                        fixtures and the oracle against silent drift
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle as O  # noqa: E402

DATA = "/root/reference/data"


def save_path(name, p):
    pts, kinds, sub, closed = p.export()
    np.savez_compressed(os.path.join(HERE, "paths", name + ".npz"), points=pts, kinds=kinds, subpath_offsets=sub, closed=closed)


def jobs_arrays(jobs):
    out = {"n_jobs": np.array(len(jobs))}
    for i, j in enumerate(jobs):
        pts, kinds, sub, closed = j["path"].export()
        d = j["paint"].describe()
        bb = j["path"].bbox()
        out.update({
            f"j{i}_points": pts, f"j{i}_kinds": kinds, f"j{i}_sub": sub, f"j{i}_closed": closed,
            f"j{i}_tr": j["tr"], f"j{i}_rule": np.array(j["fill_rule"]), f"j{i}_bbox": j["bbox"],
            f"j{i}_path_bbox": bb if bb is not None else np.zeros(4),
            f"j{i}_paint_kind": np.array(d["kind"]), f"j{i}_paint_units": np.array(d["units"]),
            f"j{i}_paint_linear_colors": np.array(d["linear_colors"]), f"j{i}_paint_spread": np.array(d["spread"]),
            f"j{i}_paint_tr": d["tr"], f"j{i}_paint_p0": d["p0"], f"j{i}_paint_p1": d["p1"], f"j{i}_paint_r0": np.array(d["r0"]),
            f"j{i}_paint_r1": np.array(d["r1"]), f"j{i}_paint_solid": d["solid"], f"j{i}_paint_stop_pos": d["stop_pos"],
            f"j{i}_paint_stop_colors": d["stop_colors"],
        })
    return out


def paint_arrays(prefix, paint):
    d = paint.describe()
    return {
        f"{prefix}_paint_kind": np.array(d["kind"]), f"{prefix}_paint_units": np.array(d["units"]),
        f"{prefix}_paint_linear_colors": np.array(d["linear_colors"]), f"{prefix}_paint_spread": np.array(d["spread"]),
        f"{prefix}_paint_tr": d["tr"], f"{prefix}_paint_p0": d["p0"], f"{prefix}_paint_p1": d["p1"], f"{prefix}_paint_r0": np.array(d["r0"]),
        f"{prefix}_paint_r1": np.array(d["r1"]), f"{prefix}_paint_solid": d["solid"], f"{prefix}_paint_stop_pos": d["stop_pos"],
        f"{prefix}_paint_stop_colors": d["stop_colors"],
    }


def pipeline_arrays(nodes):
    """Node table of Pipeline::build (kinds 0 Fill, 1 Group, 2 Opacity, 3 Clip; children before parents, root last)."""
    out = {"n_nodes": np.array(len(nodes))}
    for i, n in enumerate(nodes):
        out.update({f"n{i}_kind": np.array(n["kind"]), f"n{i}_rule": np.array(n["fill_rule"]), f"n{i}_tr": n["tr"], f"n{i}_bbox": n["bbox"],
                    f"n{i}_opacity": np.array(n["opacity"]), f"n{i}_child": np.array(n["child"]),
                    f"n{i}_children": np.asarray(n["children"], dtype=np.int64)})
        if n["path"] is not None:
            pts, kinds, sub, closed = n["path"].export()
            bb = n["path"].bbox()
            out.update({f"n{i}_points": pts, f"n{i}_kinds": kinds, f"n{i}_sub": sub, f"n{i}_closed": closed,
                        f"n{i}_path_bbox": bb if bb is not None else np.zeros(4)})
        if n["paint"] is not None:
            out.update(paint_arrays(f"n{i}", n["paint"]))
    return out


# A scene that exercises every Pipeline node kind (src/scene.rs:13-63): groups, transforms, an opacity group holding a
# clip, a clip with bounding-box units, a stroke, linear / radial gradients and a nested opacity.
# (The fractional translate keeps flattened end points off x == layer width: there the reference's right-edge clip keeps
# the OUTSIDE half of the line (src/rasterize.rs:377-383 tests `p0.x() < width`) and its cells wrap into the next row — a
# defect the oracle reproduces and the GPU path does not; see tests/test_oracle_kat.py::test_right_edge_wrap_quirk.)
NESTED_SCENE = """
{"type": "group", "children": [
  {"type": "fill", "paint": "#204060", "path": "M5,5 h190 v150 h-190 z"},
  {"type": "opacity", "opacity": 0.6, "child": {"type": "group", "children": [
     {"type": "fill", "fill_rule": "evenodd",
      "paint": {"type": "linear-gradient", "start": [20, 20], "end": [150, 120], "stops": [[0, "#ff0000"], [0.5, "#00ff00aa"], [1, "#0000ff"]]},
      "path": "M20,20 C80,-10 160,40 150,90 S60,170 30,110 Q0,70 20,20 Z M60,50 h40 v40 h-40 z"},
     {"type": "clip", "clip": "M100,40 C140,40 170,70 170,100 C170,140 130,150 100,150 C60,150 40,120 50,90 C60,60 80,40 100,40 Z",
      "child": {"type": "transform", "tr": "rotate(12) translate(8 -6)", "child": {"type": "group", "children": [
         {"type": "fill", "paint": {"type": "radial-gradient", "center": [110, 95], "radius": 55, "fcenter": [95, 80],
                                    "spread": "repeat", "stops": [[0, "#ffffff"], [0.6, "#ff8800"], [1, "#40004080"]]},
          "path": "M40,30 h150 v130 h-150 z"},
         {"type": "stroke", "paint": "#000000c0", "width": 3.5, "path": "M50,60 C90,20 130,140 180,70"}]}}}]}},
  {"type": "transform", "tr": "translate(120.3 10.2) scale(0.5)", "child":
     {"type": "clip", "units": "objectBoundingBox", "fill_rule": "evenodd", "clip": "M0.1,0.1 h0.8 v0.8 h-0.8 z M0.3,0.3 h0.4 v0.4 h-0.4 z",
      "child": {"type": "opacity", "opacity": 0.85, "child":
         {"type": "fill", "paint": {"type": "linear-gradient", "units": "objectBoundingBox", "start": [0, 0], "end": [1, 1],
                                    "linear_colors": true, "stops": [[0, "#ffff00"], [1, "#00ffff"]]},
          "path": "M10,10 C60,0 110,0 150,20 C170,70 170,110 150,150 C100,170 60,170 10,150 C0,100 0,60 10,10 Z"}}}}
]}
"""


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    os.makedirs(os.path.join(HERE, "paths"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "scenes"), exist_ok=True)
    expected = {"paths": {}, "scenes": {}, "glyphs": {}}
    paths = {}
    for name in ["squirrel", "tv", "rust", "material", "ava", "huyak"]:
        p = O.OraclePath.parse(open(f"{DATA}/{name}.path", "rb").read())
        paths[name] = p
        save_path(name, p)
    # config-5 input: tv.path stroked with the CLI's `-s 0.5` style (examples/rasterize.rs:255-262)
    paths["tv_stroked"] = paths["tv"].stroke(0.5, "round", 4.0, "round")
    save_path("tv_stroked", paths["tv_stroked"])
    paths["squirrel_stroked"] = paths["squirrel"].stroke(1.0, "round", 4.0, "round")
    save_path("squirrel_stroked", paths["squirrel_stroked"])
    for name, p in paths.items():
        (w, h), tr, _ = p.size()
        lines = p.flatten(tr)
        e = {"counts": list(p.counts()), "bbox": list(map(float, p.bbox())), "size": [w, h], "size_tr": list(map(float, tr)),
             "lines_at_size": len(lines), "lines_digest": digest(lines)}
        if w * h <= 4_000_000:
            for rule, rn in ((O.NONZERO, "nonzero"), (O.EVENODD, "evenodd")):
                img = np.zeros((h, w))
                p.mask(tr, rule, img)
                e["mask_sum_" + rn] = float(img.sum())
        expected["paths"][name] = e
    # fit transforms of the BASELINE configs
    (sz, tr) = O.fit_size(paths["squirrel"].bbox(), 512, 0)
    expected["paths"]["squirrel"]["c1"] = {"size": list(sz), "tr": list(map(float, tr)), "lines": len(paths["squirrel"].flatten(tr))}
    (sz, tr) = O.fit_size(paths["material"].bbox(), 4096, 4096)
    expected["paths"]["material"]["c2"] = {"size": list(sz), "tr": list(map(float, tr)), "lines": len(paths["material"].flatten(tr))}
    (sz, tr) = O.fit_size(paths["tv_stroked"].bbox(), 32768, 32768)
    expected["paths"]["tv_stroked"]["c5"] = {"size": list(sz), "tr": list(map(float, tr)), "lines": len(paths["tv_stroked"].flatten(tr))}

    # scenes: resolved Fill jobs for fixed render transforms
    def scene_fixture(name, scene, tr, view, bg, small=True):
        jobs = scene.fill_jobs(tr, view)
        arrs = jobs_arrays(jobs)
        arrs["render_tr"] = np.asarray(tr, dtype=np.float64)
        arrs["view"] = np.asarray(view, dtype=np.float64)
        arrs["bg"] = np.asarray(bg if bg is not None else [0, 0, 0, 0], dtype=np.float32)
        arrs["has_bg"] = np.array(bg is not None)
        np.savez_compressed(os.path.join(HERE, "scenes", name + ".npz"), **arrs)
        x, y, img = scene.render(tr, view, bg)
        rgba = O.lin_to_rgba(img)
        expected["scenes"][name] = {"n_jobs": len(jobs), "layer": [x, y, img.shape[1], img.shape[0]],
                                    "lines": int(sum(len(j["path"].flatten(j["tr"])) for j in jobs)),
                                    "rgba_digest": digest(rgba), "lin_sum": [float(v) for v in img.reshape(-1, 4).sum(0, dtype=np.float64)]}

    ff = O.OracleScene.load_json(open(f"{DATA}/firefox.scene", "rb").read())
    for s in (256, 2048):
        (sz, tr) = O.fit_size(ff.bbox(), s, s)
        scene_fixture(f"firefox_{s}", ff, tr, (0, 0, sz[0], sz[1]), None)
    lc = O.OracleScene.load_json(open(f"{DATA}/linear-colors.scene", "rb").read())
    bb = lc.bbox()
    (sz, tr) = O.fit_size(bb, 520, 0)
    scene_fixture("linear_colors", lc, tr, (0, 0, sz[0], sz[1]), None)
    # config 1: examples/rasterize.rs default scene for squirrel.path -w 512
    (sz, tr) = O.fit_size(paths["squirrel"].bbox(), 512, 0)
    cli = O.OracleScene.cli_rasterize(paths["squirrel"], tr, sz[0], sz[1])
    scene_fixture("squirrel_cli_512", cli, O.IDENTITY, (0, 0, sz[0], sz[1]), O.parse_color("#f0f0f0"))
    # many-circles bench scene (benches/scene_bench.rs) at reduced count for tests
    mc = O.OracleScene.many_circles(0, 64, 1024)
    scene_fixture("many_circles_64", mc, O.IDENTITY, (0, 0, 1024, 1024), None)

    # full pipelines (clip / opacity nodes included): the node table Pipeline::build produces for a render transform + view
    os.makedirs(os.path.join(HERE, "pipelines"), exist_ok=True)
    expected["pipelines"] = {}

    def pipeline_fixture(name, scene, tr, view, bg):
        nodes = scene.pipeline(tr, view)
        arrs = pipeline_arrays(nodes)
        arrs["render_tr"] = np.asarray(tr, dtype=np.float64)
        arrs["view"] = np.asarray(view if view is not None else [0, 0, 0, 0], dtype=np.float64)
        arrs["has_view"] = np.array(view is not None)
        arrs["bg"] = np.asarray(bg if bg is not None else [0, 0, 0, 0], dtype=np.float32)
        arrs["has_bg"] = np.array(bg is not None)
        np.savez_compressed(os.path.join(HERE, "pipelines", name + ".npz"), **arrs)
        x, y, img = scene.render(tr, view, bg)
        expected["pipelines"][name] = {"n_nodes": len(nodes), "kinds": [int(n["kind"]) for n in nodes], "layer": [x, y, img.shape[1], img.shape[0]],
                                       "rgba_digest": digest(O.lin_to_rgba(img)),
                                       "lin_sum": [float(v) for v in img.reshape(-1, 4).sum(0, dtype=np.float64)]}

    gs = O.OracleScene.load_json(open(f"{DATA}/grad.scene", "rb").read())
    pipeline_fixture("grad", gs, O.IDENTITY, None, None)
    (sz, tr) = O.fit_size(gs.bbox(), 1024, 0)
    pipeline_fixture("grad_1024", gs, tr, (0, 0, sz[0], sz[1]), O.parse_color("#ffffff"))
    ns = O.OracleScene.load_json(NESTED_SCENE)
    pipeline_fixture("nested", ns, O.IDENTITY, None, None)
    (sz, tr) = O.fit_size(ns.bbox(), 900, 0)
    pipeline_fixture("nested_900", ns, tr, (0, 0, sz[0], sz[1]), O.parse_color("#101010"))
    # also as a pipeline: firefox.scene (Fill nodes only) so that both code paths render the same thing
    (sz, tr) = O.fit_size(ff.bbox(), 512, 512)
    pipeline_fixture("firefox_512", ff, tr, (0, 0, sz[0], sz[1]), None)

    # synthetic glyphs: pin the generator
    for seed in (1, 2, 3, 100):
        g = O.OraclePath.glyph(seed)
        lines = g.flatten()
        img = np.zeros((64, 64))
        g.mask(O.IDENTITY, O.NONZERO, img)
        expected["glyphs"][str(seed)] = {"lines": len(lines), "points_digest": digest(g.export()[0]), "mask_sum": float(img.sum())}

    json.dump(expected, open(os.path.join(HERE, "expected.json"), "w"), indent=1, sort_keys=True)
    print("wrote fixtures:", sorted(os.listdir(os.path.join(HERE, "paths"))), sorted(os.listdir(os.path.join(HERE, "scenes"))))


if __name__ == "__main__":
    main()
