"""CPU: the Fill arm's window arithmetic (`rasterize_b200.scene.fill_window`) against a literal restatement of the
reference's casts — `floor() as i32 - layer.x`, then `as usize`, then `view_shape`'s clamps (src/scene.rs:412-423,
src/image.rs:588-605) — including the wrap of negative bounds to an empty window (the round-1 code clamped
negative bounds to 0 and drew the node shifted)."""
import ctypes
import math
import random

from rasterize_b200.scene import fill_window


def reference_window(bbox, lx, ly, W, H):
    def as_i32(v):  # Rust `f64 as i32`: saturating truncation
        return max(-2 ** 31, min(2 ** 31 - 1, int(math.trunc(v))))

    def as_usize(v):  # Rust `i32 as usize` on a 64-bit target: sign extension
        return ctypes.c_uint64(ctypes.c_int64(ctypes.c_int32(v).value).value).value

    col_min = as_usize(as_i32(math.floor(bbox[0])) - lx)
    col_max = as_usize(as_i32(math.ceil(bbox[2])) - lx + 1)
    row_min = as_usize(as_i32(math.floor(bbox[1])) - ly)
    row_max = as_usize(as_i32(math.ceil(bbox[3])) - ly + 1)
    row_min = min(row_min, H)
    row_max = min(max(row_max, row_min), H)  # usize::clamp(row_min, height)
    col_min = min(col_min, W)
    col_max = min(max(col_max, col_min), W)
    return col_min, row_min, col_max - col_min, row_max - row_min


def test_inside_and_clamped():
    assert fill_window([3.2, 4.1, 10.5, 12.2], 0, 0, 20, 20) == (3, 4, 9, 10)
    assert fill_window([3.2, 4.1, 30.5, 32.2], 0, 0, 20, 20) == (3, 4, 17, 16)
    assert fill_window([25.0, 4.0, 30.0, 9.0], 0, 0, 20, 20)[2] == 0  # right of the layer


def test_negative_bounds_wrap_to_an_empty_window():
    assert fill_window([-3.2, 4.1, 10.5, 12.2], 0, 0, 20, 20) == (20, 4, 0, 10)
    assert fill_window([3.2, -0.1, 10.5, 12.2], 2, 0, 20, 20) == (1, 20, 9, 0)
    assert fill_window([5.0, 5.0, 9.0, 9.0], 7, 0, 20, 20)[2:] == (0, 5)  # layer origin right of the node's left edge


def test_matches_the_casts_on_random_boxes():
    rnd = random.Random(7)
    for _ in range(5000):
        x0, y0 = rnd.uniform(-60, 90), rnd.uniform(-60, 90)
        bbox = [x0, y0, x0 + rnd.uniform(0, 80), y0 + rnd.uniform(0, 80)]
        lx, ly, W, H = rnd.randint(-20, 40), rnd.randint(-20, 40), rnd.randint(0, 70), rnd.randint(0, 70)
        assert fill_window(bbox, lx, ly, W, H) == reference_window(bbox, lx, ly, W, H), (bbox, lx, ly, W, H)
