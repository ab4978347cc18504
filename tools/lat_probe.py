"""Single-pair latency probe: p50/p95 of vslam_process over 200 calls (pinned and pageable inputs) + kernel time."""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "orb-slam3-rust_b200"))
import ctypes as C, numpy as np, torch
from vslam_b200 import CameraModel, StereoProcessor, synth
L, R = synth.make_stereo_batch(range(8))
Lp, Rp = L.pin_memory(), R.pin_memory()
Ln, Rn = L.numpy().copy(), R.numpy().copy()
sp = StereoProcessor(CameraModel(**synth.EUROC_CAMERA), 1200, max_batch=2)
ob = sp.new_output_buffers(1)
def run(lp, rp, n=200):
    lat, dev = [], []
    for i in range(n + 10):
        j = i % 8
        t = time.perf_counter()
        rc = sp._lib.vslam_process(sp._h, C.c_void_p(lp(j)), 752, C.c_void_p(rp(j)), 752, 0, ob.outs)
        lat.append(1e3 * (time.perf_counter() - t)); dev.append(sp.last_batch_ms())
        assert rc == 0
    lat, dev = sorted(lat[10:]), sorted(dev[10:])
    return lat[len(lat)//2], lat[int(len(lat)*0.95)], dev[len(dev)//2]
print("graph=%s" % os.environ.get("VSLAM_GRAPH", "1"))
print("pinned   p50 %.3f p95 %.3f ms, device %.3f ms" % run(lambda j: Lp.data_ptr() + j*752*480, lambda j: Rp.data_ptr() + j*752*480))
print("pageable p50 %.3f p95 %.3f ms, device %.3f ms" % run(lambda j: Ln[j].ctypes.data, lambda j: Rn[j].ctypes.data))
