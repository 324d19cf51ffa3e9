#!/bin/bash
# launch list of BASELINE configs[3] (100k users x 3600 frames, tile counts 200/500/1000, transition entropy)
mkdir -p gpurun_out
cat > /tmp/c4.py <<'P'
import sys
sys.path.insert(0, "/root/repo")
import torch, bench
from viewport_entropy_toolkit_b200 import EntropyConfig, get_engine
dev = torch.device("cuda")
p = bench.synth_on_device(torch, 3600, 100_000, 4, dev)
eng = get_engine(100, 200, [200, 500, 1000], EntropyConfig(use_weight_distribution=False), dev)
eng.set_option("cuda_graph", "off")
for _ in range(2):
    eng.transition(p, want_pairs0=False, want_per_k=False)
torch.cuda.synchronize()
P
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 60 --csv --log-file gpurun_out/r02_c4_launches.csv python /tmp/c4.py > gpurun_out/ncu_c4.log 2>&1
echo "ncu rc=$?"
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02_c4_launches.csv")) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
gi = hdr.index("Grid Size"); bi = hdr.index("Block Size")
for r in rows[1:]:
    print(r[ki][:70], r[gi], r[bi], r[vi])
P
