"""One profiled TRAINING step of the bench workload (for ncu): warm up, then exactly one HotPathTrainer.step inside
cudaProfilerStart/Stop.   ncu --profile-from-start off ... python tools/one_train_step.py [drop]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dpmn_b200.pipeline import DPMNHotPath  # noqa: E402
from dpmn_b200.train import HotPathTrainer  # noqa: E402

drop = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
dev = torch.device("cuda:0")
model = DPMNHotPath(precision="fp16", drop=drop)
pg, cm = bench.synth_weights(2)
bench.load_weights(model, pg, cm)
model = model.to(dev).train()
tr = HotPathTrainer(model)
psn, p1, p2 = bench.synth_inputs(1, bench.BATCH)
hr = torch.from_numpy(np.random.default_rng(3).uniform(0, 1, (bench.BATCH, 4, 32, 128)).astype(np.float32)).to(dev)
args = (torch.from_numpy(psn).to(dev), [torch.from_numpy(a).to(dev) for a in p1], [torch.from_numpy(a).to(dev) for a in p2], hr)
tr.step(*args)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step(*args)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("one training step done")
