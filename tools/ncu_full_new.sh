#!/bin/bash
# --set full captures of the kernels added at the end of round 2: k_relabel_rows_group (configs[3]) and
# k_whist_i8 / k_whist_i8_finish / k_cnt_planes (configs[4] shard, tensor-core histogram split over the cells)
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
cat > /tmp/ep.py <<'P'
import sys
sys.path.insert(0, "/root/repo")
import torch, bench
from viewport_entropy_toolkit_b200 import EntropyConfig, get_engine
dev = torch.device("cuda")
p = bench.synth_on_device(torch, 450, 1_000_000, 20265000, dev, chunk=32)
eng = get_engine(100, 200, [200], EntropyConfig(fov_angle=90.0, power_factor=2.0), dev)
eng.set_option("cuda_graph", "off")
eng.set_option("weighted_kernel", "i8")
for _ in range(2):
    eng.spatial(p, want_per_k=False, want_hist0=False)
torch.cuda.synchronize()
P
ncu --set full --clock-control none --import-source on -k regex:k_relabel_rows_group -s 1 -c 1 -o gpurun_out/r02_relabel_full -f python /tmp/c4.py > gpurun_out/ncu_full_relabel.log 2>&1
echo "ncu relabel rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_whist_i8 -s 4 -c 2 -o gpurun_out/r02_whist_split_full -f python /tmp/ep.py > gpurun_out/ncu_full_whist.log 2>&1
echo "ncu whist rc=$?"
ls -la gpurun_out/*.ncu-rep
