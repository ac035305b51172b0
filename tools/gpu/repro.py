import sys, numpy as np
sys.path.insert(0, "optical-flow-fpga_b200")
import of_b200 as ofb
rng = np.random.default_rng(0)
p = rng.integers(0, 255, (240, 320)).astype(np.float32)
c = np.roll(p, 1, axis=1)
u, v = ofb.lk_pyramidal(p, c, 3, 5, 3, mode=ofb.MODE_FAST)
print("ok", float(np.abs(u).mean()))
