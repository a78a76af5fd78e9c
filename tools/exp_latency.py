"""GPU experiment: single-frame latency of ORBextractor.__call__ (operator()) through the host API."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiagent_orb_slam2_b200 import ORBextractor, synth
img = synth.image("blocks", 640, 480, 0)
ex = ORBextractor(1000, 1.2, 8, 20, 7)
for _ in range(20): ex(img)
t = []
for _ in range(200):
    t0 = time.perf_counter(); k, d = ex(img); t.append(time.perf_counter() - t0)
t = np.array(t) * 1e6
print("operator() 640x480 single frame: median %.1f us, p10 %.1f, p90 %.1f  (%d keypoints) -> %.0f frames/s single stream" % (np.median(t), np.percentile(t, 10), np.percentile(t, 90), len(k), 1e6 / np.median(t)))
