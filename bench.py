#!/usr/bin/env python
"""Headline benchmark: frames/s of ORB extract + frame-to-frame match at 640x480 / 1000 keypoints
(BASELINE.json metric; TUM mono settings, 8 levels, scale 1.2) on N GPUs of one node, one agent
stream per GPU (no data-path collective: "weak" scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]

A step = one pass of the hot path over a batch of B synthetic frames per GPU:
  pyramid -> per-cell FAST -> quadtree distribution -> orientation -> blur -> rBRIEF, then
  brute-force Hamming kNN-2 + ratio test of every frame against the next one.
`value`  : frames already resident in HBM (B * 307 KB > L2 so every step streams from HBM).
`e2e`    : the same work through the host-buffer API: pinned host frames -> H2D -> kernels -> D2H of
           keypoints, descriptors, counts and matches, every step, copies inside the timed region.
`--impl reference` times the reference's own CPU extractor (oracle/_ref, its ORBextractor.cc
compiled unmodified) + the oracle's Hamming loop on the host cores, on a bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH = 640, 480, 1000, 1.2, 8, 20, 7
NNRATIO, TH = 0.9, 50
METRIC = "frames/s ORB extract+match @640x480,1000kp"
WORKLOAD = "640x480 gray, 1000 kp, 8 levels, scale 1.2, FAST 20/7; frame-to-frame brute-force match (ratio 0.9, TH_LOW 50)"
STAGES = ["pyramid_resize", "gaussian_blur", "fast_cells", "quadtree", "orient_describe"]
DISTINCT_FRAMES = 64


def make_frames(n, seed0, w=W, h=H, distinct=DISTINCT_FRAMES):
    """n synthetic frames: `distinct` different ones (distinct/2 scenes x 2 views; consecutive frames are shifted /
    re-noised copies, i.e. real matches), tiled to n. The data-dependent stages (FAST survivors, empty-cell retries,
    quadtree depth) therefore see 64 different inputs per batch."""
    from multiagent_orb_slam2_b200 import synth
    base = []
    for s in range(max(1, min(distinct, n) // 2)):
        a, b, _ = synth.shifted_pair("blocks", w, h, seed0 + s)
        base += [a, b]
    reps = (n + len(base) - 1) // len(base)
    return np.ascontiguousarray(np.stack((base * reps)[:n]))


class ClockSampler:
    Q = ("utilization.gpu,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.path = gpu, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                rows.append((float(f[0]), float(f[1]), float(f[2]), f[5:9]))
            except ValueError:
                continue
        os.unlink(self.path)
        loaded = [r for r in rows if r[0] >= 20.0] or rows  # samples taken while the GPU was busy
        reasons = set()
        for r in loaded:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if loaded:
            out = {"sm_mhz": float(np.median([r[1] for r in loaded])), "sm_max_mhz": float(max(r[2] for r in loaded)),
                   "reasons": sorted(reasons), "samples": len(rows), "samples_under_load": len([r for r in rows if r[0] >= 20.0])}
        return out


def bind_to_gpu_numa_node(gpu):
    """Pins this rank's host threads (and therefore its first-touch pinned buffers) to the CPUs that are
    local to its GPU, so that N ranks do not pull their frames across the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def measured_int8_peak(device=0):
    """Dense int8 tensor-core peak of this GPU: a bare tcgen05.mma kind::i8 issue loop with both operands in shared memory
    (tools/microbench/utcimma_peak.cu) run live when its binary is there, else the value recorded under profiles/, else
    2 x the measured bf16 cuBLAS figure."""
    exe = os.path.join(ROOT, "tools", "microbench", "utcimma_peak")
    try:
        if os.path.exists(exe):
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            ids = vis.split(",") if vis else None
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=(ids[device] if ids and device < len(ids) else str(device)))
            for line in subprocess.check_output([exe], text=True, timeout=120, env=env).splitlines():
                if "int8_tops_measured" in line:
                    return json.loads(line)["int8_tops_measured"], "measured live: bare tcgen05.mma kind::i8 issue loop (tools/microbench/utcimma_peak.cu)"
    except Exception:
        pass
    try:
        for line in open(os.path.join(ROOT, "profiles", "r02_utcimma_peak.log")):
            if "int8_tops_measured" in line:
                return json.loads(line)["int8_tops_measured"], "profiles/r02_utcimma_peak.log (bare tcgen05.mma kind::i8 issue loop on a B200 of this pool)"
    except Exception:
        pass
    return 2 * peaks()[0].get("bf16_tflops", 2250.0), "2 x the measured dense bf16 cuBLAS figure of MEASURED_PEAKS.json"


# ------------------------------------------------------------------------------------------------------
def cpu_reference_run(frames, threads):
    """Reference CPU path on `frames`: extraction with the reference's own ORBextractor.cc
    (oracle/_ref) when present, else the oracle port; matching with the oracle's Hamming loop.
    Returns (seconds, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_lib as O
    from concurrent.futures import ThreadPoolExecutor
    kind = "reference" if O.ref_available("canonical") else "port"
    n = len(frames)
    local = threading.local()

    def extract(i):
        ex = getattr(local, "ex", None)
        if ex is None:
            ex = local.ex = (O.RefExtractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH) if kind == "reference"
                             else O.OracleExtractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH))
        return ex(frames[i])[1]

    def match(i):
        return O.knn2(descs[i], descs[(i + 1) % n])

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as pool:
        descs = list(pool.map(extract, range(n)))
        list(pool.map(match, range(n)))
    return time.perf_counter() - t0, kind


def cv2_primitive_bound(frame, reps=5):
    """Single-thread lower bound of the extraction's OpenCV primitives through cv2: the 7 resizes,
    the per-cell FAST calls (incl. ~815 Python call overheads) and the 8 Gaussian blurs of one 640x480 frame; no quadtree,
    orientation or descriptors. None when cv2 is not importable."""
    try:
        import cv2
    except Exception:
        return None
    cv2.setNumThreads(1)
    scale = [np.float32(1.0)]
    for _ in range(1, NLEVELS):
        scale.append(np.float32(scale[-1] * np.float32(SCALE)))

    def once():
        lv = [frame]
        for l in range(1, NLEVELS):
            inv = np.float32(1.0) / scale[l]
            sz = (int(np.rint(np.float32(W) * inv)), int(np.rint(np.float32(H) * inv)))
            lv.append(cv2.resize(lv[-1], sz, interpolation=cv2.INTER_LINEAR))
        fast_ini = cv2.FastFeatureDetector_create(INI_TH, True)
        fast_min = cv2.FastFeatureDetector_create(MIN_TH, True)
        ncand = 0
        for im in lv:
            hh, ww = im.shape
            minb, maxx, maxy = 16, ww - 16, hh - 16
            ncols, nrows = max(1, (maxx - minb) // 30), max(1, (maxy - minb) // 30)
            wc, hc = -(-(maxx - minb) // ncols), -(-(maxy - minb) // nrows)
            for i in range(nrows):
                y0 = minb + i * hc
                if y0 >= maxy - 3:
                    continue
                y1 = min(y0 + hc + 6, maxy)
                for j in range(ncols):
                    x0 = minb + j * wc
                    if x0 >= maxx - 6:
                        continue
                    x1 = min(x0 + wc + 6, maxx)
                    k = fast_ini.detect(im[y0:y1, x0:x1])
                    if not k:
                        k = fast_min.detect(im[y0:y1, x0:x1])
                    ncand += len(k)
        for im in lv:
            cv2.GaussianBlur(im, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        return ncand

    once()
    t0 = time.perf_counter()
    for _ in range(reps):
        ncand = once()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    return {"ms_per_frame_1_thread": ms, "frames_per_s_1_thread": 1e3 / ms, "fast_candidates": int(ncand),
            "what": "cv2 %s primitives only (7 resizes, per-cell FAST with Python call overhead, 8 blurs), 1 thread" % cv2.__version__}

# ------------------------------------------------------------------------------------------------------
def extra_config_legs(args, dev, local_rank):
    """The other named configurations of BASELINE.json (rank 0, N=1): C2 KITTI-shape stereo (2 x 1241x376, 2000 kp, extract
    L+R + Frame::ComputeStereoMatches on the device), C3 EuRoC-shape 752x480 batched 64 frames/launch with full
    frame-to-frame matching, and the drop-in's real operating point: one frame per call through the C++ facade."""
    import ctypes as C
    import torch
    from multiagent_orb_slam2_b200 import _lib, synth
    from multiagent_orb_slam2_b200.extractor import ORBextractor
    from multiagent_orb_slam2_b200.frontend import AgentFrontend
    L = _lib.lib()
    pk, _ = peaks()
    out = {}
    st = torch.cuda.current_stream(dev).cuda_stream

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    def stage_times(ex, d_frames, n):
        _lib.check(L.orbx_set_stage_timing(ex._h, 1))
        ms = (C.c_float * len(STAGES))()
        acc = np.zeros(len(STAGES))
        for _ in range(3):
            ex.extract_device(d_frames.data_ptr(), d_frames.stride(1), d_frames.stride(0), n, st)
            _lib.check(L.orbx_stage_times(ex._h, ms))
            acc += np.array(list(ms))
        _lib.check(L.orbx_set_stage_timing(ex._h, 0))
        alg = (C.c_double * len(STAGES))()
        _lib.check(L.orbx_algorithmic_bytes(ex._h, alg))
        return {name: {"ms": acc[i] / 3, "alg_bytes_per_frame": alg[i],
                       "frac_hbm": (alg[i] * n / (acc[i] / 3 * 1e-3) / 1e9 / pk["hbm_gbs"]) if acc[i] > 0 else 0.0} for i, name in enumerate(STAGES)}

    # ---- C2: KITTI-shape stereo ------------------------------------------------------------------------------------
    KW, KH, KN, KB = 1241, 376, 2000, 64
    pairs = [synth.stereo_pair("blocks", KW, KH, 700 + s) for s in range(32)]
    left = np.ascontiguousarray(np.stack([p[0] for p in pairs] * (KB // 32)))
    right = np.ascontiguousarray(np.stack([p[1] for p in pairs] * (KB // 32)))
    exl = ORBextractor(KN, SCALE, NLEVELS, INI_TH, MIN_TH, KW, KH, device=local_rank, max_batch=KB)
    exr = ORBextractor(KN, SCALE, NLEVELS, INI_TH, MIN_TH, KW, KH, device=local_rank, max_batch=KB)
    dl, dr = torch.from_numpy(left).to(dev), torch.from_numpy(right).to(dev)
    cap = exl.cap
    ur = torch.empty((KB, cap), dtype=torch.float32, device=dev); dz = torch.empty_like(ur)
    sad = torch.empty((KB, cap), dtype=torch.int32, device=dev); kept = torch.zeros(KB, dtype=torch.int32, device=dev)
    mbf, mb = 386.1448, 386.1448 / 718.856

    def stereo_step():
        exl.extract_device(dl.data_ptr(), dl.stride(1), dl.stride(0), KB, st)
        exr.extract_device(dr.data_ptr(), dr.stride(1), dr.stride(0), KB, st)
        _lib.check(L.orbm_stereo_match_batch_device(exl._h, exr._h, KB, mbf, mb, C.c_void_p(ur.data_ptr()), C.c_void_p(dz.data_ptr()),
                                                    C.c_void_p(sad.data_ptr()), C.c_void_p(kept.data_ptr()), cap, C.c_void_p(st)))

    ms = timed(stereo_step, 5)
    ms_extract = timed(lambda: (exl.extract_device(dl.data_ptr(), dl.stride(1), dl.stride(0), KB, st),
                                exr.extract_device(dr.data_ptr(), dr.stride(1), dr.stride(0), KB, st)), 5)
    # end to end through the host-buffer calls: both images up, keypoints / descriptors / stereo coordinates down
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        exl.extract_batch(left); exr.extract_batch(right)
        u = np.empty(cap, np.float32); d = np.empty(cap, np.float32); k = C.c_int()
        for f in range(KB):
            _lib.check(L.orbm_stereo_match(exl._h, exr._h, f, mbf, mb, u.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), cap, C.byref(k)))
    e2e_s = (time.perf_counter() - t0) / reps
    out["kitti_stereo"] = {"workload": "KITTI-shape stereo 1241x376 left/right, 2000 kp per image, 8 levels, stereo Hamming matching + SAD refinement",
                           "stereo_pairs_per_step": KB, "distinct_pairs": 32, "ms_per_step": ms, "value": KB / (ms * 1e-3), "unit": "stereo frames/s",
                           "ms_extract_LR": ms_extract, "ms_stereo_match": ms - ms_extract, "matched_per_frame": float(kept.float().mean().item()),
                           "e2e": {"value": KB / e2e_s, "unit": "stereo frames/s", "how": "orbx_extract_batch (L, R) + orbm_stereo_match per frame, host buffers, synchronous",
                                   "h2d_bytes_per_step": 2 * KB * KW * KH, "d2h_bytes_per_step": KB * (2 * cap * 60 + cap * 8)},
                           "stages_left": stage_times(exl, dl, KB)}
    del exl, exr, dl, dr, ur, dz, sad, kept

    # ---- C3: EuRoC-shape, 64 frames per launch, full frame-to-frame matching -----------------------------------------
    EW, EH, EN, EB = 752, 480, 1200, 64
    frames = make_frames(EB, 900, EW, EH)
    fe = AgentFrontend(EW, EH, EN, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank, max_batch=EB, nnratio=NNRATIO, th=TH)
    d_frames = torch.from_numpy(frames).to(dev)
    ms = timed(lambda: fe.process_device(d_frames), 10)
    ms_match = timed(lambda: fe.match_consecutive(EB), 10)
    h_frames = torch.from_numpy(frames).pin_memory()
    o = fe.pinned_outputs()
    fe.process_async(h_frames, o); torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(10):
        fe.process_async(h_frames, o)
        torch.cuda.current_stream(dev).synchronize()
        _ = int(o["counts"].numpy()[:EB].sum())
    e2e_s = (time.perf_counter() - t0) / 10
    out["euroc_b64"] = {"workload": "EuRoC-shape 752x480, 64 frames per launch, 1200 kp, full frame-to-frame matching (ratio 0.9, TH_LOW 50)",
                        "frames_per_step": EB, "ms_per_step": ms, "value": EB / (ms * 1e-3), "unit": "frames/s", "ms_matching": ms_match,
                        "keypoints_per_frame": float(o["counts"].numpy()[:EB].mean()),
                        "e2e": {"value": EB / e2e_s, "unit": "frames/s", "how": "one buffer, upload -> kernels -> download, synchronised every step",
                                "h2d_bytes_per_step": EB * EW * EH, "d2h_bytes_per_step": EB * fe.d2h_bytes_per_frame()},
                        "stages": stage_times(fe.ex, d_frames, EB)}
    del fe, d_frames

    # ---- the drop-in's operating point: batch 1 through the C++ facade --------------------------------------------------
    try:
        exe = os.path.join(ROOT, "multiagent_orb_slam2_b200", "lib", "bench_single_frame")
        src = os.path.join(ROOT, "tools", "bench", "single_frame.cc")
        lib = os.path.join(ROOT, "multiagent_orb_slam2_b200", "lib")
        if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
            subprocess.check_call(["g++", "-std=c++14", "-O2", "-I" + os.path.join(ROOT, "include"), src, "-o", exe, "-L" + lib, "-lorb_b200",
                                   "-Wl,-rpath," + lib])
        with tempfile.NamedTemporaryFile(suffix=".raw", delete=False) as f:
            make_frames(16, 300).tofile(f)
            raw = f.name
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local_rank)))
        js = subprocess.check_output([exe, str(W), str(H), raw, "16", str(NFEAT), "200"], text=True, env=env, timeout=300)
        os.unlink(raw)
        sf = json.loads(js)
        sf["what"] = ("C++ facade ORB_SLAM2::ORBextractor::operator() on ONE 640x480 host image per call (median of 200 calls, wall clock incl. "
                      "H2D, kernels, D2H and the vector<cv::KeyPoint> fill); SearchForInitialization on a 2 x nFeatures frame pair")
        sf["frames_per_s_single_stream"] = 1e6 / sf["extract_us"]
        out["single_frame"] = sf
    except Exception as e:  # the headline must not depend on a host compiler being present
        out["single_frame"] = {"unavailable": repr(e)[:300]}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = max(cores * 8, 64)
    frames = make_frames(sample, 0)
    for _ in range(args.warmup):
        cpu_reference_run(frames[:cores], cores)
    t = []
    kind = "port"
    for _ in range(args.steps):
        dt, kind = cpu_reference_run(frames, cores)
        t.append(dt)
    total = sum(t)
    v = sample * args.steps / total
    what = "%d frames per step, %d host threads, one extractor per thread" % (sample, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": what},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "sample": what,
                         "ms_per_frame_per_thread": 1e3 * cores / v, "cv2_bound": cv2_primitive_bound(frames[0])},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------
def run_b200(args):
    # NCCL reads its logging environment once, when the library initialises - which may be at `import torch`: set it first.
    # Its own log (communicator size, transport) is evidence of the N-rank run: INIT lines at INFO level, kept off stdout
    # (the contract is ONE JSON line there); every rank logs to a file, rank 0 echoes its file to stderr at the end.
    nccl_log = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        nccl_log = os.path.join(tempfile.gettempdir(), "orb_bench_nccl_%d_rank%s.log" % (os.getppid(), os.environ.get("RANK", "0")))
        os.environ["NCCL_DEBUG_FILE"] = nccl_log
    import torch
    import torch.distributed as dist
    from multiagent_orb_slam2_b200 import _lib
    from multiagent_orb_slam2_b200.frontend import AgentFrontend

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()   # creates the communicator now, so that its INIT lines are in the log
    comm = None
    if world > 1 and rank == 0:
        nr = None
        try:
            import re
            import glob
            for f in sorted(glob.glob(nccl_log + "*")):
                for line in open(f, errors="replace"):
                    m = re.search(r"nranks (\d+)", line)
                    if m:
                        nr = int(m.group(1))
                        break
                if nr is not None:
                    break
        except Exception:
            pass
        comm = {"backend": "nccl", "world_size": world, "nranks_in_nccl_log": nr,
                "data_path": "extraction / matching: none (replicas); cross-map: peer loads + flags over NVLink (orbm_knn2_allgather); NCCL "
                             "carries the timing all-reduce, the barriers and the 64-byte window handles only"}
    B = args.batch
    L = _lib.lib()

    frames = make_frames(B, 1000 * rank)
    fe = AgentFrontend(W, H, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank, max_batch=B, nnratio=NNRATIO, th=TH)
    d_frames = torch.from_numpy(frames).to(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ("value") -----------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # sampled through the value, e2e and stage-timing legs; idle samples are dropped
    for _ in range(args.warmup):
        fe.process_device(d_frames)
    barrier()
    launches0 = L.orb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fe.process_device(d_frames)
    e1.record()
    barrier()
    launches = L.orb_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = world * B * args.steps / (ms_total * 1e-3)
    import ctypes as C
    # counts live in handle-owned memory: read them through the ABI download
    pin = fe.pinned_outputs()
    _lib.check(L.orbx_download_results(fe.ex._h, B, C.c_void_p(pin["kps"].data_ptr()), C.c_void_p(pin["desc"].data_ptr()), fe.cap,
                                       C.c_void_p(pin["counts"].data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    torch.cuda.synchronize(dev)
    kp_per_frame = float(pin["counts"].numpy()[:B].mean())
    matches_per_frame = float((fe.match[:B].cpu().numpy() >= 0).sum() / B)

    # ---- end to end through the host-buffer API ("e2e"): four agents' worth of buffers in flight (three measured 156 - 167 k
    # frames/s from box to box, four 168 k: one more buffer absorbs the host's jitter between synchronise and enqueue), so that
    # the upload of step i+1, the kernels of step i and the download of step i-1 overlap --------------------
    h_frames = torch.from_numpy(frames).pin_memory()
    DEPTH = 4
    fes = [fe] + [AgentFrontend(W, H, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank, max_batch=B, nnratio=NNRATIO, th=TH)
                  for _ in range(DEPTH - 1)]
    streams = [torch.cuda.Stream(dev) for _ in range(DEPTH)]
    outs = [f.pinned_outputs() for f in fes]

    def consume(out):
        """What the facade's operator() does with a finished step on the host (src/ORBextractor.cc:1072-1101 fills
        vector<cv::KeyPoint>): every frame's count read, its keypoint records and matches walked once."""
        cnt = out["counts"].numpy()[:B]
        kp = out["kps"].numpy()[:B].view(np.float32).reshape(B, -1, 6)
        return int(cnt.sum()), float(kp[:, :, 4].sum()), int((out["match"].numpy()[:B] >= 0).sum())

    def e2e_step(i):
        k = i % DEPTH
        with torch.cuda.stream(streams[k]):
            fes[k].process_async(h_frames, outs[k])

    for i in range(max(args.warmup, 2 * DEPTH)):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    marks = []
    for i in range(args.steps):
        e2e_step(i)
        marks.append(time.perf_counter() - t0)
        if i >= DEPTH - 1:
            j = (i - (DEPTH - 1)) % DEPTH
            streams[j].synchronize()  # the host consumes an earlier step's results while the later ones run
            consume(outs[j])
    for s in streams:
        s.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    if rank == 0:
        sys.stderr.write("e2e enqueue times (ms): " + " ".join("%.2f" % (m * 1e3) for m in marks) + "\n")
    barrier()
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    # the floor of the end-to-end number on this box: the same pinned frames copied host -> device and nothing else, all
    # ranks at once (they share the host's PCIe switches / memory controllers), max over ranks
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d_tmp = torch.empty_like(d_frames)
    for _ in range(2):
        d_tmp.copy_(h_frames, non_blocking=True)
    barrier()
    f0.record()
    for _ in range(8):
        d_tmp.copy_(h_frames, non_blocking=True)
    f1.record()
    barrier()
    h2d_floor_ms = max_over_ranks(f0.elapsed_time(f1)) / 8
    del d_tmp

    # ---- per-stage device time of the dominant kernels (roofline), rank 0 -----------------------------
    roof = None
    stage_ms = None
    if rank == 0:
        _lib.check(L.orbx_set_stage_timing(fe.ex._h, 1))
        acc = np.zeros(len(STAGES))
        reps = max(3, min(args.steps, 10))
        ms = (C.c_float * len(STAGES))()
        for _ in range(reps):
            fe.ex.extract_device(d_frames.data_ptr(), d_frames.stride(1), d_frames.stride(0), B, torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(L.orbx_stage_times(fe.ex._h, ms))
            acc += np.array(list(ms))
        _lib.check(L.orbx_set_stage_timing(fe.ex._h, 0))
        stage_ms = (acc / reps).tolist()
        alg = (C.c_double * len(STAGES))()
        _lib.check(L.orbx_algorithmic_bytes(fe.ex._h, alg))
        # matching stage, timed alone
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for _ in range(reps):
            fe.match_consecutive(B)
        m1.record()
        torch.cuda.synchronize(dev)
        match_ms = m0.elapsed_time(m1) / reps
        pk, pk_src = peaks()
        dom = int(np.argmax(stage_ms))
        per_stage = {}
        for i, name in enumerate(STAGES):
            gbs = alg[i] * B / (stage_ms[i] * 1e-3) / 1e9 if stage_ms[i] > 0 else 0.0
            per_stage[name] = {"ms": stage_ms[i], "alg_bytes_per_frame": alg[i], "gbs": gbs, "frac_hbm": gbs / pk["hbm_gbs"]}
        cmp_per_step = float((pin["counts"].numpy()[:B].astype(np.float64) * np.roll(pin["counts"].numpy()[:B], -1)).sum())
        per_stage["hamming_knn2"] = {"ms": match_ms, "gcmp_per_s": cmp_per_step / (match_ms * 1e-3) / 1e9}
        a = per_stage[STAGES[dom]]
        # DRAM bytes per frame of each kernel from this round's `ncu --set full` captures (profiles/r02_ncu_traffic.json names
        # them), scaled to this launch's B frames; null when the file is missing
        try:
            ncu_traffic_per_frame = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))["bytes_per_frame"]
        except Exception:
            ncu_traffic_per_frame = {}
        roof = {"kernel": STAGES[dom], "bound": "hbm", "achieved": a["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": a["frac_hbm"],
                "traffic": ncu_traffic_per_frame[STAGES[dom]] * B if STAGES[dom] in ncu_traffic_per_frame else None,
                "algorithmic_bytes_per_launch": a["alg_bytes_per_frame"] * B,
                "peak_source": pk_src + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                "note": "bound is stated against HBM because the contract offers hbm | tensor; the dominant kernel (FAST) is in fact bound "
                        "by instruction issue and the shared-memory pipe (ncu at B=512, profiles/r02_ncu_fast_b512_summary.md: 78 % of "
                        "issue slots, 72 % of the shared-memory wavefront peak, 1.33 warp instructions per pixel, DRAM 9 %); its DRAM "
                        "traffic (1.16 x) matches its algorithmic bytes. The HBM-streaming stages are pyramid_resize and gaussian_blur "
                        "(see stages)",
                "stages": per_stage}

    # ---- MapFusion cross-map matching (BASELINE config 5): G Hamming cmp/s over all ranks ------------------
    # orbm_knn2_allgather (csrc/xmap.cu): the exchange is peer loads over NVLink inside the operand expansion, no collective
    # library call on the data path; one map per rank (2 maps at N=1), and a second leg with 2 maps whose query rows are
    # split over all N ranks.
    mapf = None
    if not args.no_mapfusion:
        from multiagent_orb_slam2_b200 import mapfusion, synth
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        rows = args.map_rows
        base = synth.descriptors(rows, 4242)
        popc_roof = 148 * 16 * 1.965e9 / 8 * world

        def xmap_leg(n_maps, reps):
            cm = mapfusion.CrossMapMatcher(rows, n_maps, 0.75)
            host = {m: synth.descriptors_fast(rows, 5000 + m, base, 60) for m in cm.owned}
            local = [torch.from_numpy(host[m]).to(dev) for m in cm.owned]
            rpm = [rows] * n_maps
            cm.match(local, rpm)  # warm-up (peer mappings, kernel load)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(reps):
                res, _r = cm.match(local, rpm)
            g1.record()
            barrier()
            ms = max_over_ranks(g0.elapsed_time(g1)) / reps
            # parity outside the timed region: sampled rows of every pair this rank owns against the oracle (needs every
            # database map on the host: regenerate it from its seed)
            import oracle_lib as O
            checked = 0
            rng = np.random.default_rng(rank)
            for (a, b), (idx, d1, d2, _m) in res.items():
                sample = np.sort(rng.choice(rows, 64, replace=False))
                dbm = host[b] if b in host else synth.descriptors_fast(rows, 5000 + b, base, 60)
                oi, o1, o2 = O.knn2(host[a][sample], dbm)
                ok = (np.array_equal(idx.cpu().numpy()[sample], oi) and np.array_equal(d1.cpu().numpy()[sample], o1)
                      and np.array_equal(d2.cpu().numpy()[sample], o2))
                if not ok:
                    raise SystemExit("bench.py: cross-map result of pair (%d, %d) differs from the oracle on rank %d" % (a, b, rank))
                checked += len(sample)
            accepted = int(sum(int((v[3] >= 0).sum()) for v in res.values()))
            tot = torch.tensor([checked, accepted], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(tot)
            launches_step = None
            cm.close()
            cmps = float(n_maps * (n_maps - 1)) * rows * rows
            return {"maps": n_maps, "rows_per_map": rows, "directed_pairs": n_maps * (n_maps - 1), "ms_per_step": ms,
                    "value": cmps / (ms * 1e-3) / 1e9, "unit": "Gcmp/s", "parity_rows_checked": int(tot[0].item()),
                    "accepted_matches": int(tot[1].item())}, cmps

        n_maps = max(2, world)
        leg, cmps = xmap_leg(n_maps, 5)
        mf_ms = leg["ms_per_step"]
        # the same step with the POPC kernel (north star: "popc-bound uint4-vectorised kernel") on ONE GPU-resident pair list
        popc = None
        if rank == 0:
            from multiagent_orb_slam2_b200 import _lib as _orb_lib
            LL = _orb_lib.lib()
            a_set = torch.from_numpy(synth.descriptors_fast(rows, 5000, base, 60)).to(dev)
            b_set = torch.from_numpy(synth.descriptors_fast(rows, 5001, base, 60)).to(dev)
            o3 = [torch.empty(rows, dtype=torch.int32, device=dev) for _ in range(3)]
            stp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _orb_lib.check(LL.orbm_set_knn2_backend(1))
            run = lambda: _orb_lib.check(LL.orbm_knn2_device(C.c_void_p(a_set.data_ptr()), rows, C.c_void_p(b_set.data_ptr()), rows,
                                                             *[C.c_void_p(o.data_ptr()) for o in o3], stp))
            run()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            run()
            h1.record()
            torch.cuda.synchronize(dev)
            _orb_lib.check(LL.orbm_set_knn2_backend(0))
            pms = h0.elapsed_time(h1)
            popc = {"value": float(rows) * rows / (pms * 1e-3) / 1e9, "unit": "Gcmp/s per GPU", "ms_per_pair": pms,
                    "frac_of_plain_popc_roofline": float(rows) * rows / (pms * 1e-3) / (popc_roof / world),
                    "what": "one 200k x 200k pair with the POPC kernel pinned (orb_b200_debug.h): knn2_kernel, carry-save popcount on the integer pipes"}
            del a_set, b_set, o3
        barrier()
        split = None
        if world > 2:
            split, _ = xmap_leg(2, 5)   # fewer maps than GPUs: query rows of the 2 maps split over all ranks
            split["what"] = "2 maps on %d GPUs: the 128-row query tiles of both pairs dealt evenly to all ranks" % world
        int8_peak, int8_src = measured_int8_peak(local_rank) if rank == 0 else (1.0, "")
        barrier()   # the other ranks keep their GPUs idle meanwhile (shared power / clocks are not an issue, but keep the step clean)
        mapf = dict(leg)
        mapf.update({"metric": "G Hamming cmp/s, cross-map brute force + ratio test",
                     "exchange": "fused: peer loads over NVLink inside the operand-expansion kernel (orbm_knn2_allgather), no collective call",
                     "frac_of_plain_popc_roofline": cmps / (mf_ms * 1e-3) / popc_roof,
                     "popc_roofline": "148 SM x 16 POPC/clk x 1.965 GHz / 8 POPC per cmp per GPU (measured 15.3/clk/SM)",
                     "popc_kernel": popc, "fewer_maps_than_gpus": split,
                     "kernel": "knn2_mma_kernel (tcgen05.mma kind::i8, 128x128x256 tiles, top-2 out of TMEM)",
                     "tensor": {"achieved_int8_tops": cmps * 512 / (mf_ms * 1e-3) / 1e12 / world, "peak_int8_tops_per_gpu": int8_peak,
                                "frac": cmps * 512 / (mf_ms * 1e-3) / 1e12 / world / int8_peak, "peak_source": int8_src}})

    # ---- DBoW2 vocabulary transform of the batch's descriptors (Frame::ComputeBoW), rank 0 -----
    bow = None
    if rank == 0 and not args.no_bow:
        from multiagent_orb_slam2_b200 import synth
        from multiagent_orb_slam2_b200.vocabulary import ORBVocabulary
        voc = ORBVocabulary(synth.vocabulary_fast(10, 6, 7), device=local_rank)  # ORBvoc shape: k=10, L=6, 1.1 M nodes
        nfeat = B * fe.cap
        ow = torch.empty(nfeat, dtype=torch.int32, device=dev); on = torch.empty_like(ow)
        owt = torch.empty(nfeat, dtype=torch.float64, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        call = lambda: _lib.check(L.orbv_descend_device(voc._h, C.c_void_p(fe.d_desc), nfeat, 4, C.c_void_p(ow.data_ptr()),
                                                        C.c_void_p(on.data_ptr()), C.c_void_p(owt.data_ptr()), st))
        for _ in range(3):
            call()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(10):
            call()
        b1.record()
        torch.cuda.synchronize(dev)
        bms = b0.elapsed_time(b1) / 10
        bow = {"metric": "vocabulary descents/s (k=10, L=6 synthetic tree, levelsup 4)", "value": nfeat / (bms * 1e-3), "unit": "features/s",
               "features_per_launch": nfeat, "ms_per_launch": bms, "hamming_per_feature": 60}
        del voc

    other = None
    if rank == 0 and world == 1 and not args.no_other_configs:
        other = extra_config_legs(args, dev, local_rank)

    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample = max(cores * 24, 192)  # ~15-20 CPU-seconds of reference work
        cf = make_frames(sample, 0)
        cpu_reference_run(cf[:cores], cores)
        dt, kind = cpu_reference_run(cf, cores)
        cpu = {"value": sample / dt, "unit": "frames/s", "cores": cores, "kind": kind,
               "sample": "%d frames extract+match on %d host threads (%.1f s)" % (sample, cores, dt),
               "ms_per_frame_per_thread": 1e3 * cores * dt / sample, "cv2_bound": cv2_primitive_bound(cf[0])}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "frames_per_step_per_gpu": B, "distinct_frames_per_batch": min(DISTINCT_FRAMES, B), "agents": world, "parallelism": "one agent stream per GPU, no collective",
                       "host_cpus_per_rank": numa,
                       "l2": "inputs larger than L2 (%d MB of frames per step)" % (B * W * H // 2**20),
                       "keypoints_per_frame": kp_per_frame, "matches_per_frame": matches_per_frame},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": B * fe.h2d_bytes_per_frame(),
                    "d2h_bytes_per_step": B * fe.d2h_bytes_per_frame(), "ms_per_step": e2e_ms / args.steps,
                    "h2d_floor_ms": h2d_floor_ms, "h2d_floor_gbs_per_rank": B * fe.h2d_bytes_per_frame() / (h2d_floor_ms * 1e-3) / 1e9,
                    "h2d_floor_note": "pure pinned host -> device copy of one step's frames on every rank at once, max over ranks: the "
                                      "platform bound of e2e (uploads of different steps cannot overlap each other on one PCIe link)",
                    "how": "pinned host frames -> orbx_upload_frames/extract_staged/orbm_knn2_batched/download, 3 buffers in flight"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "mapfusion": mapf, "bow_transform": bow, "other_configs": other,
            "comm": comm,
        }))
    if world > 1:
        dist.destroy_process_group()
        if rank == 0 and nccl_log:
            import glob
            found = sorted(glob.glob(nccl_log + "*"))
            sys.stderr.write("NCCL log files: %r\n" % [(f, os.path.getsize(f)) for f in found])
            for f in found[:1]:
                lines = [l.rstrip() for l in open(f, errors="replace")]
                keep = [l for l in lines if "nranks" in l or "NCCL version" in l or "comm 0x" in l][:16]
                sys.stderr.write("NCCL log (rank 0, %s):\n" % f + "\n".join(keep or lines[:16]) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=512, help="frames per step per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-mapfusion", action="store_true", help="skip the cross-map Hamming leg")
    ap.add_argument("--no-bow", action="store_true", help="skip the vocabulary transform leg")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the KITTI-stereo / EuRoC-B64 / single-frame legs")
    ap.add_argument("--map-rows", type=int, default=200000, help="descriptors per map in the cross-map leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
