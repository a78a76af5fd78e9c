"""Where does the end-to-end step lose its overlap? Times the pieces of AgentFrontend.process_async on 3 streams."""
import ctypes as C, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from multiagent_orb_slam2_b200 import _lib
from multiagent_orb_slam2_b200.frontend import AgentFrontend

B = 512
L = _lib.lib()
frames = bench.make_frames(B, 0)
h = torch.from_numpy(frames).pin_memory()
DEPTH = int(sys.argv[1]) if len(sys.argv) > 1 else 3
fes = [AgentFrontend(bench.W, bench.H, bench.NFEAT, bench.SCALE, bench.NLEVELS, bench.INI_TH, bench.MIN_TH, device=0, max_batch=B,
                     nnratio=bench.NNRATIO, th=bench.TH) for _ in range(DEPTH)]
streams = [torch.cuda.Stream() for _ in range(DEPTH)]
outs = [f.pinned_outputs() for f in fes]


def step(fe, out, up, ex, ma, dn):
    st = C.c_void_p(fe._stream())
    hh = fe.ex._h
    if up:
        _lib.check(L.orbx_upload_frames(hh, C.c_void_p(h.data_ptr()), h.stride(1), h.stride(0), B, st))
    if ex:
        _lib.check(L.orbx_extract_staged(hh, B, st))
    if ma:
        fe.match_consecutive(B)
    if dn:
        _lib.check(L.orbx_download_results(hh, B, C.c_void_p(out["kps"].data_ptr()), C.c_void_p(out["desc"].data_ptr()), fe.cap,
                                           C.c_void_p(out["counts"].data_ptr()), st))
        out["match"][:B].copy_(fe.match[:B], non_blocking=True)


def run(name, up, ex, ma, dn, steps=24, api=False):
    for i in range(2 * DEPTH):
        with torch.cuda.stream(streams[i % DEPTH]):
            step(fes[i % DEPTH], outs[i % DEPTH], True, True, True, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tl = 0.0
    for i in range(steps):
        k = i % DEPTH
        t1 = time.perf_counter()
        with torch.cuda.stream(streams[k]):
            if api:
                fes[k].process_async(h, outs[k])
            else:
                step(fes[k], outs[k], up, ex, ma, dn)
        tl += time.perf_counter() - t1
        if i >= DEPTH - 1:
            streams[(i - (DEPTH - 1)) % DEPTH].synchronize()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print("%-28s %.3f ms/step  (host enqueue %.3f ms/step)" % (name, dt * 1e3, tl / steps * 1e3))


run("upload only", 1, 0, 0, 0)
run("extract only", 0, 1, 0, 0)
run("extract+match", 0, 1, 1, 0)
run("download only", 0, 0, 0, 1)
run("upload+extract+match", 1, 1, 1, 0)
run("extract+match+download", 0, 1, 1, 1)
run("all", 1, 1, 1, 1)
run("process_async (API)", 1, 1, 1, 1, api=True)
run("process_async (API) 48 steps", 1, 1, 1, 1, steps=48, api=True)
