#!/bin/bash
# launch list of the weighted epilogue kernels on the configs[4] shard (tensor-core kernel pinned)
mkdir -p gpurun_out
cat > /tmp/ep.py <<'P'
import sys
sys.path.insert(0, "/root/repo")
import torch, bench
from viewport_entropy_toolkit_b200 import EntropyConfig, get_engine
dev = torch.device("cuda")
p = bench.synth_on_device(torch, 450, 1_000_000, 20265000, dev, chunk=32)
eng = get_engine(100, 200, [200], EntropyConfig(fov_angle=90.0, power_factor=2.0), dev)
eng.set_option("cuda_graph", "off")
for wk in ("i8", "fp64"):
    eng.set_option("weighted_kernel", wk)
    for _ in range(3):
        eng.spatial(p, want_per_k=False, want_hist0=False)
torch.cuda.synchronize()
P
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 60 --csv --log-file gpurun_out/r02_epilogue_launches.csv python /tmp/ep.py > gpurun_out/ncu_ep.log 2>&1
echo "ncu rc=$?"
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02_epilogue_launches.csv")) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
for r in rows[1:]:
    print(r[ki][:60], r[vi])
P
