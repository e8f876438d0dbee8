"""Timing of the XTC path on the C3 system (device decode alone, end-to-end pass from pinned XTC bytes)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cgraphs_b200 import synth
from cgraphs_b200.mdhbond import HbondAnalysis, _native
from cgraphs_b200.mdhbond.calib import cos_threshold
from cgraphs_b200.mdhbond.topology import pack_systems

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
system, cfg = synth.config("C3")
na = system.n_atoms
dev = torch.device("cuda", 0)
d_xyz = torch.empty((nf, na, 3), dtype=torch.float32, device=dev)
system.frames_torch(0, nf, dev, out=d_xyz)
xyz = d_xyz.cpu().numpy()
t0 = time.perf_counter()
buf, off = _native.xtc_encode(xyz, in_scale=0.1)
print("encode: %.2f s, %.3f bytes/atom, %d threads" % (time.perf_counter() - t0, len(buf) / nf / na, os.cpu_count()))
ctx = _native.Context(0)
out = torch.empty((nf, na, 3), dtype=torch.float32, device=dev)
ctx.xtc_decode_to_device(buf, off, na, nf, out.data_ptr())
torch.cuda.synchronize()
t0 = time.perf_counter()
ctx.xtc_decode_to_device(buf, off, na, nf, out.data_ptr())
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("decode_to_device (pageable copy + kernel): %.1f ms = %.0f frames/s" % (dt * 1e3, nf / dt))
print("max |decoded - original| = %.5f A" % float((out - d_xyz).abs().max()))
hba = HbondAnalysis("protein", system.universe(1), device=0)
top = hba._topology
tables, names = pack_systems([top], [top.key_ids("water_around")], [na])
c = hba._context()
c.set_topology(tables)
c.set_params(3.5, 3.5, cos_threshold(60.0), True, True, _native.METHOD_WATER_AROUND)
pin = _native.PinnedBuffer((len(buf),), dtype=np.uint8)
host = pin.array
host[:] = buf
c.set_option("profile", 1)
if len(sys.argv) > 2:
    c.set_option("batch_bytes", int(float(sys.argv[2]) * (1 << 30)))
for label, src in (("pageable", buf), ("pinned", host)):
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c.begin(nf)
        c.submit_xtc(src, off, na, nf)
        keys, words = c.results()
        dt = time.perf_counter() - t0
    tm = c.timers()
    print("%s XTC e2e: %.1f ms = %.0f frames/s; h2d %.1f MB, ms_h2d %.1f, decode %.1f ms in %d launches, search %.1f ms, %d bonds" % (
        label, dt * 1e3, nf / dt, tm["h2d_bytes"] / 1e6, tm["ms_h2d"], tm["ms_kernel"]["xtc_decode"], tm["launches"]["xtc_decode"],
        tm["ms_total"], len(keys)))
# float32 frames for comparison
pf = _native.PinnedBuffer((nf, na, 3))
pf.array[...] = out.cpu().numpy()
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c.begin(nf)
    c.submit_raw(pf.ptr, na, nf)
    keys2, words2 = c.results()
    dt = time.perf_counter() - t0
print("pinned float32 e2e: %.1f ms = %.0f frames/s" % (dt * 1e3, nf / dt))
print("same result from XTC and from its decoded frames:", np.array_equal(keys, keys2) and np.array_equal(words, words2))
