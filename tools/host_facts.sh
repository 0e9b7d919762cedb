#!/bin/bash
# What the box is: PCIe link per GPU, CPU / NUMA layout, GPU topology (VERDICT r1 task 4 artefacts).
out=${1:-gpurun_out/r2_host_facts.txt}
{
  echo "== nvidia-smi topo -m"; nvidia-smi topo -m
  echo "== PCIe link per GPU (nvidia-smi -q)"
  nvidia-smi -q | grep -E "^GPU 0000|Product Name|Bus Id|PCIe Generation|Link Width|Max  |Current  |Device Current|Device Max|Host Max" | sed 's/^ *//'
  echo "== lscpu"; lscpu | grep -E "Model name|Socket|Core|Thread|^CPU\(s\)|NUMA|Hypervisor|Virtualization|L3"
  echo "== numactl -H"; (numactl -H 2>/dev/null || for n in /sys/devices/system/node/node*; do echo "$n cpus $(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done)
  echo "== GPU local cpulist / numa_node"
  for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ -f $d/local_cpulist ]; then echo "$(basename $d) class $(cat $d/class) numa $(cat $d/numa_node) cpus $(cat $d/local_cpulist) link $(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done
  echo "== memory"; free -g | head -2
} > "$out" 2>&1
