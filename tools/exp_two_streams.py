"""Does running two half-batches on two streams overlap the latency-bound stages (quadtree, describe) of one with the
issue-bound FAST of the other? One frontend x 512 frames against two frontends x 256 frames on two streams, offset or not."""
import numpy as np
import torch

from multiagent_orb_slam2_b200 import synth
from multiagent_orb_slam2_b200.frontend import AgentFrontend

dev = torch.device("cuda", 0)
W, H = 640, 480
kinds = ["blocks", "blurnoise"]
frames = np.stack([synth.image(kinds[i % 2], W, H, i) for i in range(64)])
big = torch.from_numpy(np.concatenate([frames] * 8)).to(dev)  # 512 frames


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


one = AgentFrontend(W, H, max_batch=512)
print("1 x 512 on one stream: %.3f ms" % timed(lambda: one.process_device(big)))
del one
torch.cuda.empty_cache()
a, b = AgentFrontend(W, H, max_batch=256), AgentFrontend(W, H, max_batch=256)
sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
ha, hb = big[:256], big[256:]


def two():
    cur = torch.cuda.current_stream()
    sa.wait_stream(cur); sb.wait_stream(cur)
    with torch.cuda.stream(sa):
        a.process_device(ha)
    with torch.cuda.stream(sb):
        b.process_device(hb)
    cur.wait_stream(sa); cur.wait_stream(sb)


print("2 x 256 on two streams: %.3f ms" % timed(two))
sa2, sb2 = torch.cuda.Stream(dev, priority=-1), torch.cuda.Stream(dev, priority=0)


def two_prio():
    cur = torch.cuda.current_stream()
    sa2.wait_stream(cur); sb2.wait_stream(cur)
    with torch.cuda.stream(sa2):
        a.process_device(ha)
    with torch.cuda.stream(sb2):
        b.process_device(hb)
    cur.wait_stream(sa2); cur.wait_stream(sb2)


print("2 x 256, first stream high priority: %.3f ms" % timed(two_prio))
q = [AgentFrontend(W, H, max_batch=128) for _ in range(4)]
ss = [torch.cuda.Stream(dev) for _ in range(4)]


def four():
    cur = torch.cuda.current_stream()
    for k in range(4):
        ss[k].wait_stream(cur)
        with torch.cuda.stream(ss[k]):
            q[k].process_device(big[128 * k:128 * (k + 1)])
    for k in range(4):
        cur.wait_stream(ss[k])


print("4 x 128 on four streams: %.3f ms" % timed(four))
