"""One profiled step of the bench workload (for ncu): warm up, then exactly one hot-path forward inside
cudaProfilerStart/Stop.  Usage under ncu:  ncu --profile-from-start off ... python tools/one_step.py [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dpmn_b200.pipeline import DPMNHotPath  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
dev = torch.device("cuda:0")
model = DPMNHotPath(precision=prec)
pg, cm = bench.synth_weights(2)
bench.load_weights(model, pg, cm)
model = model.to(dev).eval()
psn, p1, p2 = bench.synth_inputs(1, bench.BATCH)
args = (torch.from_numpy(psn).to(dev), [torch.from_numpy(a).to(dev) for a in p1], [torch.from_numpy(a).to(dev) for a in p2])
with torch.no_grad():
    for _ in range(2):
        model(*args)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    model(*args)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("one step done")
