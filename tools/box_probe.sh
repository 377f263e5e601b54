#!/bin/bash
# what the GPU box offers besides the GPU: cores, RAM, scratch space and how fast it is written / read back
echo "== nproc $(nproc)"; free -g | head -2; df -h / /tmp /dev/shm 2>/dev/null
lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" 
nvidia-smi --query-gpu=index,name,memory.total --format=csv
nvidia-smi topo -m 2>/dev/null | head -12
for d in /tmp /dev/shm; do
  t0=$(date +%s.%N); dd if=/dev/zero of=$d/jb_probe.bin bs=64M count=64 conv=fdatasync 2>&1 | tail -1
  t1=$(date +%s.%N); dd if=$d/jb_probe.bin of=/dev/null bs=64M 2>&1 | tail -1; rm -f $d/jb_probe.bin
done
