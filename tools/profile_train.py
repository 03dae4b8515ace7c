"""One sCM training step of Swift-B under a profiler (tools only):  ncu --metrics gpu__time_duration.sum
--clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/profile_train.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
t = bench._train_setup(dev, 0, 1, 1)
for _ in range(steps):
    t["one_step"](None)
torch.cuda.synchronize()
print("done")
