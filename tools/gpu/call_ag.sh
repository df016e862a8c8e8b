#!/bin/bash
mkdir -p gpurun_out
{
nvidia-smi topo -m
echo "--- nodes"; ls /sys/devices/system/node/ | head; cat /sys/devices/system/node/node*/cpulist 2>/dev/null
echo "--- cpuset"; cat /proc/self/status | grep -i "cpus_allowed_list\|mems_allowed_list"
nproc
echo "--- gpu numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d $(cat $d/numa_node) $(cat $d/class)"; fi; done
free -g | head -2
cat /sys/devices/system/node/node*/meminfo 2>/dev/null | grep -i "MemTotal\|MemFree"
} > gpurun_out/r2ag_topo.log 2>&1
cat gpurun_out/r2ag_topo.log | head -60
