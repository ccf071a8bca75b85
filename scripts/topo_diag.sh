#!/bin/bash
# Host topology of the GPU box (NUMA nodes, GPU <-> CPU affinity): context for the multi-GPU end-to-end numbers.
nproc; lscpu | grep -iE "numa|socket|model name|^cpu\(s\)"; nvidia-smi topo -m 2>/dev/null | cut -c1-200
cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; free -g | head -2
