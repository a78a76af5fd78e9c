"""Experiment (library built with ORB_EXTRA_NVCC_FLAGS=-DORB_QT_PROFILE): phase clocks of the level-0 quadtree block of frame 0
(batch of 4: one quadtree launch over all levels, block (0, 0) = level 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiagent_orb_slam2_b200 import ORBextractor, synth
import torch
imgs = np.stack([synth.image("blocks", 640, 480, i) for i in range(4)])
ex = ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=4)
for _ in range(3):
    ex.extract_batch(imgs)
torch.cuda.synchronize()
