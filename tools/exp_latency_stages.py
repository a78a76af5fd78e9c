"""GPU experiment: where a single-frame extraction spends its time (stage events at batch 1, then the whole call)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from multiagent_orb_slam2_b200 import ORBextractor, _lib, synth

L = _lib.lib()
img = synth.image("blocks", 640, 480, 0)
ex = ORBextractor(1000, 1.2, 8, 20, 7)
for _ in range(20):
    ex(img)
t = []
for _ in range(300):
    t0 = time.perf_counter(); k, d = ex(img); t.append(time.perf_counter() - t0)
print("graph path: median %.1f us  (%d keypoints)" % (np.median(t) * 1e6, len(k)))
_lib.check(L.orbx_set_stage_timing(ex._h, 1))
acc = []
for _ in range(100):
    ex(img)
    ms = (C.c_float * 8)()
    _lib.check(L.orbx_stage_times(ex._h, ms))
    acc.append(list(ms))
acc = np.median(np.array(acc), axis=0) * 1e3
print("stage events (us, no graph):", dict(zip(["resize", "blur", "fast", "quadtree", "describe"], [round(float(x), 1) for x in acc[:5]])), "sum %.1f" % acc[:5].sum())
_lib.check(L.orbx_set_stage_timing(ex._h, 0))

# device-resident input, graph replay, timed with events on the launching stream: the kernels' share of the call
import torch
st = torch.cuda.Stream()
buf = torch.from_numpy(img).cuda()
with torch.cuda.stream(st):
    for _ in range(5):
        ex.extract_device(buf.data_ptr(), 640, 640 * 480, 1, st.cuda_stream)
    st.synchronize()
    ts = []
    for _ in range(200):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ex.extract_device(buf.data_ptr(), 640, 640 * 480, 1, st.cuda_stream)
        e1.record(st)
        st.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
print("kernel sequence alone (graph replay, device events): median %.1f us, p10 %.1f" % (np.median(ts), np.percentile(ts, 10)))

# host timeline of one call through the C ABI (pageable caller buffers, as a cv::Mat / std::vector caller has them)
from multiagent_orb_slam2_b200.extractor import KP_DTYPE
cap = ex.cap
kps = np.empty(cap, KP_DTYPE); desc = np.empty((cap, 32), np.uint8); cnt = np.zeros(1, np.int32)
p = lambda a: a.ctypes.data_as(C.c_void_p)
st2 = torch.cuda.Stream()
hs = C.c_void_p(st2.cuda_stream)
acc = []
for _ in range(300):
    t0 = time.perf_counter()
    _lib.check(L.orbx_upload_frames(ex._h, p(img), 640, 640 * 480, 1, hs))
    t1 = time.perf_counter()
    _lib.check(L.orbx_extract_staged(ex._h, 1, hs))
    t2 = time.perf_counter()
    _lib.check(L.orbx_download_results(ex._h, 1, p(kps), p(desc), cap, p(cnt), hs))
    t3 = time.perf_counter()
    st2.synchronize()
    t4 = time.perf_counter()
    acc.append([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0])
acc = np.median(np.array(acc), axis=0) * 1e6
print("host timeline (us): upload call %.1f | graph launch call %.1f | download calls %.1f | final sync %.1f | total %.1f" % tuple(acc))
