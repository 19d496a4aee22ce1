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
