"""The -DHB_DEBUG build (bounds asserts on every shared-memory / global scratch index, cgraphs_b200/build.py --debug) runs
the small configurations of every kernel family without tripping an assert and with the same results - the stand-in for
compute-sanitizer, which is closed on this GPU pool."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r)
from cgraphs_b200 import synth
from cgraphs_b200.universe import Universe
from cgraphs_b200.mdhbond import HbondAnalysis, WireAnalysis, HbondBatch, HydrationAnalysis, _native
from oracle import hbond_port, xtc_port
assert "debug" in _native.LIB_PATH
def same(a, b):
    assert {k: np.asarray(v).tobytes() for k, v in a.items()} == {k: np.asarray(v).tobytes() for k, v in b.items()}
s = synth.make_system(3000, 30, 7, hydrogens=True, water="TIP3")
u = s.universe(40)
ref = hbond_port.run(u, "protein", "water_around").initial_results
for opts in ({}, {"fused": 0}, {"fused": 0, "frame_pairs": 0}, {"fused_lw_cap": 8}, {"table_capacity": 1024, "max_batch_frames": 7}):
    h = HbondAnalysis("protein", u, native_options=opts)
    h.set_hbonds_in_selection_and_water_around(3.5)
    same(h.initial_results, ref)
h = HbondAnalysis("protein or resname TIP3", u)
h.set_hbonds_in_selection()
same(h.initial_results, hbond_port.run(u, "protein or resname TIP3", "in_selection").initial_results)
h = HbondAnalysis("protein", u, pbc=True)
h.set_hbonds_in_selection_and_water_around(3.5)
same(h.initial_results, ref)
w = WireAnalysis("protein", Universe(s.topology, s.frames(0, 6)))
for hull in (False, True):
    w.set_water_wires(max_water=3, water_in_convex_hull=hull)
    r = hbond_port.run_wires(Universe(s.topology, s.frames(0, 6)), "protein", max_water=3, water_in_convex_hull=hull)
    assert {k: np.nan_to_num(v, posinf=-1).tolist() for k, v in w.wire_lengths.items()} == \
           {k: np.nan_to_num(v, posinf=-1).tolist() for k, v in r.wire_lengths.items()}
structs = [synth.make_system(2500 + 100 * i, 100, 50 + i, hydrogens=False, water="HOH").universe(1) for i in range(5)]
b = HbondBatch(structs, "protein", check_angle=False, add_donors_without_hydrogen=True)
b.set_hbonds_in_selection_and_water_around(3.5)
for st, res in zip(structs, b.results):
    same(res.initial_results, hbond_port.run(st, "protein", "water_around", check_angle=False, add_donors_without_hydrogen=True).initial_results)
xyz = s.frames(0, 12)
buf, off = xtc_port.encode((xyz * np.float32(0.1)).astype(np.float32))
dec, _ = xtc_port.decode(buf, off, 3000, scale=10.0)
assert np.array_equal(_native.xtc_decode_device(buf, off, 3000).view(np.uint32), dec.view(np.uint32))
print("debug build ok")
'''


def test_small_configs_under_the_debug_build():
    from cgraphs_b200 import build as hb_build
    lib = hb_build.build(debug=True)
    env = dict(os.environ, HBGPU_LIB=lib)
    res = subprocess.run([sys.executable, "-c", SCRIPT % dict(root=ROOT)], env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "debug build ok" in res.stdout, (res.stdout[-1500:], res.stderr[-3000:])
    assert "HB_DASSERT failed" not in res.stdout + res.stderr
