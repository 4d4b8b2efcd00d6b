#!/bin/bash
# what the GPU box looks like to a rank: GPUs, PCIe/NUMA placement, cpuset, memory nodes
nvidia-smi -L
nvidia-smi topo -m 2>/dev/null | sed 's/\x1b\[[0-9;]*m//g'
nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv
for d in /sys/bus/pci/devices/*; do
  if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ -e $d/numa_node ]; then echo "$(basename $d) class=$(cat $d/class) numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist 2>/dev/null)"; fi
done | grep "class=0x0302\|class=0x0300" 
echo "nodes online: $(cat /sys/devices/system/node/online 2>/dev/null)"
for n in /sys/devices/system/node/node*; do echo "$(basename $n): cpus $(cat $n/cpulist) mem $(grep MemTotal $n/meminfo | awk '{print $4,$5}')"; done
echo "cpuset.cpus.effective: $(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null)"
echo "cpuset.mems.effective: $(cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null)"
echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"
echo "affinity: $(python -c 'import os;print(sorted(os.sched_getaffinity(0)))' | cut -c1-200)"
lscpu | grep -i "model name\|socket\|numa\|^cpu(s)"
