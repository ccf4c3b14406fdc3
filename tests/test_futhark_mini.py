"""The interpreter behind the golden vectors (tests/golden/futhark_mini.py) on the points where Futhark's semantics differ
from Python's or C's, and on every language construct the reference sources use -- small programs with known answers."""
import sys

import numpy as np
import pytest

from util import ROOT

sys.path.insert(0, str(ROOT / "tests" / "golden"))
import futhark_mini as F  # noqa: E402


def run(tmp_path, src, name="main", *args, files=None):
    for fn, text in (files or {}).items():
        (tmp_path / fn).write_text(text)
    p = tmp_path / "prog.fut"
    p.write_text(src)
    val = F.Program({}).load(p)[name]
    return F.call(val, *args) if args else val


def test_integer_division_and_modulo_round_towards_negative_infinity(tmp_path):
    assert run(tmp_path, "def f (a: i64) (b: i64) = (a / b, a % b)", "f", -7, 2) == (-4, 1)
    assert run(tmp_path, "def f (a: i64) (b: i64) = (a / b, a % b)", "f", 7, -2) == (-4, -1)
    assert run(tmp_path, "def f (n: i64) = (0 - 1) % n", "f", 14) == 13              # explorer.fut:149 (scenario_id - 1) % n


def test_f32_arithmetic_and_truncation(tmp_path):
    src = """
def f (td: f32) (unused: f32) (sps: i32) =
  let time_total = td + unused
  let steps_new = time_total * f32.i32 sps
  let steps_new' = i64.f32 steps_new
  in (steps_new', time_total - f32.i64 steps_new' / f32.i32 sps)
"""
    n, rest = run(tmp_path, src, "f", np.float32(0.05), np.float32(0.016), 32)
    tt = np.float32(0.05) + np.float32(0.016)
    assert n == int(tt * np.float32(32)) == 2
    assert isinstance(rest, np.float32) and rest == tt - np.float32(2) / np.float32(32)
    assert run(tmp_path, "def g (x: f32) = (i64.f32 x, t32 x)", "g", np.float32(-2.7)) == (-2, -2)   # towards zero
    assert run(tmp_path, "def h = 1.0f32 / 3.0", "h") == np.float32(1) / np.float32(3)


def test_booleans_short_circuit(tmp_path):
    src = "def f (xs: []i64) (i: i64) = i >= 0 && i < length xs && xs[i] != 0"
    assert run(tmp_path, src, "f", [1, 0], 5) is False and run(tmp_path, src, "f", [1, 0], 0) is True
    with pytest.raises(F.FutharkError, match="out of bounds"):
        run(tmp_path, "def f (xs: []i64) (i: i64) = xs[i]", "f", [1, 0], 2)


def test_definitions_shadow_and_are_not_recursive(tmp_path):
    src = """
def step (x: i64) (n: i64) = x + n
local def step (td: i64) (x: i64) = if td > 0 then step x td else x
def event (td: i64) (x: i64) = step td x
"""
    assert run(tmp_path, src, "event", 3, 10) == 13                              # designer.fut defines `step` twice like this
    assert run(tmp_path, "def f (x: i64) = let x = x + 1 let x = x * 2 in x", "f", 3) == 8


def test_currying_sections_and_lambdas(tmp_path):
    src = """
def in_bounds (h: i64) (w: i64) (y: i64) (x: i64): bool = y >= 0 && y < h && x >= 0 && x < w
def f (gh: i64) (gw: i64) =
  let inb = in_bounds gh gw
  in (inb 1 1, inb 5 0, map (.1) [(1, 2), (3, 4)], map (\\(a, b) -> a * b) [(1, 2), (3, 4)], reduce (+) 0 [1, 2, 3])
"""
    assert run(tmp_path, src, "f", 3, 3) == (True, False, [2, 4], [2, 12], 6)


def test_records_with_updates_and_equality(tmp_path):
    src = """
def f (c: {a: {y: i8, x: i8}, b: bool}) =
  let c' = c with a.y = 5 with b = !c.b
  in (c'.a.y, c'.a.x, c'.b, c.a.y, c'.a == {y=5, x=c.a.x}, c' != c)
"""
    assert run(tmp_path, src, "f", {"a": {"y": 1, "x": -1}, "b": False}) == (5, -1, True, 1, True, True)


def test_arrays_with_updates_slices_and_in_place_sugar(tmp_path):
    src = """
def f (g: [][]i64) =
  let g2 = copy g with [1, 0] = 9
  let xs = [10, 20, 30] ++ [40]
  let xs[0] = 11
  in (g2[1, 0], g[1, 0], xs[:2], length xs, flatten g2, 0..<3)
"""
    assert run(tmp_path, src, "f", [[1, 2], [3, 4]]) == (9, 3, [11, 20], 4, [1, 2, 9, 4], [0, 1, 2])


def test_loops(tmp_path):
    src = """
def f (n: i64) =
  let a = loop acc = 0 for i < n do acc + i
  let (b, _) = loop (x, going) = (1, true) while going do (x * 2, x * 2 < 100)
  let acc = 5
  let c = loop acc for i < 3 do acc + 1
  in (a, b, c)
"""
    assert run(tmp_path, src, "f", 5) == (10, 128, 8)


def test_match_on_tuples_constructors_and_literals(tmp_path):
    src = """
type option 'a = #some a | #none
def g (a: bool) (b: bool) (p: bool) =
  match (a, b, p)
  case (_, true, true) -> 1
  case (true, _, false) -> 2
  case (true, false, true) -> 3
  case (false, true, false) -> 4
  case (false, false, q) -> if q then 5 else 6
def h (o: option (i64, i64)) = match o case #some (y, _) -> y case #none -> -1
def k (e: i64) = match e case 00 -> 10 case 08 -> 18 case _ -> 0
"""
    assert [run(tmp_path, src, "g", *v) for v in [(False, True, True), (True, True, False), (True, False, True), (False, True, False),
                                                   (False, False, True), (False, False, False)]] == [1, 2, 3, 4, 5, 6]
    assert run(tmp_path, src, "h", ("#some", (7, 8))) == 7 and run(tmp_path, src, "h", ("#none",)) == -1
    assert run(tmp_path, src, "k", 8) == 18


def test_size_parameters_come_from_the_arguments(tmp_path):
    src = """
type state [n] [gh] [gw] = {grid: [gh][gw]i64, cycles: [n][gh][gw]i64}
def f [gh][gw] (g: *[gh][gw]i64) (extra: i64) = (gh, gw, unflatten (flatten g))
def s [gh][gw] (st: state [] [gh] [gw]) = gh * 100 + gw
"""
    assert run(tmp_path, src, "f", [[1, 2, 3], [4, 5, 6]], 0) == (2, 3, [[1, 2, 3], [4, 5, 6]])
    assert run(tmp_path, src, "s", {"grid": [[0] * 7] * 4, "cycles": []}) == 407


def test_modules_functors_imports_and_strings(tmp_path):
    lib = """
module type shape = { val name: () -> string [] val area: i64 -> i64 }
module square: shape = { def name () = "square" def area (s: i64) = s * s }
module mk_scaled (input: { val k: i64 }) (sh: shape): shape = {
  def k = input.k
  def name () = "scaled " ++ sh.name () ++ " x" ++ ['0' + u8.i64 k]
  def area (s: i64) = k * sh.area s
}
"""
    src = """
module lib = import "lib"
module user = {
  open lib
  module S2 = mk_scaled { def k = 2i64 } square
  def describe (s: i64) = (S2.name (), S2.area s, length (S2.name ()))
}
def main (s: i64) = user.describe s
"""
    assert run(tmp_path, src, "main", 3, files={"lib.fut": lib}) == ("scaled square x2", 18, 16)


def test_failed_tuple_components_only_fail_when_used(tmp_path):
    """explorer_scenario_helper.fut:26-31 computes (init on a dummy 1 x 1 grid, ..., name ()) and projects the name out"""
    src = "def f (xs: []i64) = (xs[5], 42).1\ndef g (xs: []i64) = (xs[5], 42).0 + 1"
    assert run(tmp_path, src, "f", [1]) == 42
    with pytest.raises(F.FutharkError):
        run(tmp_path, src, "g", [1])
