"""GPU experiment: time the matching calls of AgentFrontend in isolation for several batch sizes."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multiagent_orb_slam2_b200 import _lib
from multiagent_orb_slam2_b200.frontend import AgentFrontend
import bench

L = _lib.lib()
for B in (64, 256):
    frames = bench.make_frames(B, 0)
    fe = AgentFrontend(640, 480, max_batch=B)
    d = torch.from_numpy(frames).cuda()
    fe.process_device(d); torch.cuda.synchronize()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    cap = fe.cap
    def batched():
        _lib.check(L.orbm_knn2_batched_device(C.c_void_p(fe.d_desc), C.c_void_p(fe.d_counts), cap, C.c_void_p(fe.d_desc + cap * 32),
                   C.c_void_p(fe.d_counts + 4), cap, B - 1, C.c_void_p(fe.idx.data_ptr()), C.c_void_p(fe.d1.data_ptr()), C.c_void_p(fe.d2.data_ptr()), st))
    def wrap():
        last = B - 1
        _lib.check(L.orbm_knn2_batched_device(C.c_void_p(fe.d_desc + last * cap * 32), C.c_void_p(fe.d_counts + 4 * last), cap,
                   C.c_void_p(fe.d_desc), C.c_void_p(fe.d_counts), cap, 1, C.c_void_p(fe.idx.data_ptr() + 4 * last * cap),
                   C.c_void_p(fe.d1.data_ptr() + 4 * last * cap), C.c_void_p(fe.d2.data_ptr() + 4 * last * cap), st))
    def ratio():
        _lib.check(L.orbm_ratio_filter_device(C.c_void_p(fe.idx.data_ptr()), C.c_void_p(fe.d1.data_ptr()), C.c_void_p(fe.d2.data_ptr()),
                   B * cap, 50, 1, 0.9, C.c_void_p(fe.match.data_ptr()), st))
    for name, f in (("batched", batched), ("wrap", wrap), ("ratio", ratio), ("all", lambda: fe.match_consecutive(B))):
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(10): f()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print("B=%d %-8s device %.3f ms/call, host enqueue %.3f ms/call" % (B, name, e0.elapsed_time(e1) / 10, (t1 - t0) * 100))
