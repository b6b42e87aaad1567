#!/bin/bash
# what does the GPU box look like?
nproc; lscpu | grep -E "Model name|Socket|Thread|NUMA node\(s\)"; free -g | head -2
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv
nvidia-smi topo -m 2>/dev/null | head -5
ls /root/reference 2>&1 | head -2
