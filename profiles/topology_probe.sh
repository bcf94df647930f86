#!/bin/bash
# Host / PCIe / NUMA topology of a GPU box, for the host-buffer (e2e) path. Writes to stdout.
echo "== nproc / affinity"; nproc; grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
echo "== lscpu"; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core|^CPU\(s\)"
echo "== nodes"; ls /sys/devices/system/node/ | grep node; for n in /sys/devices/system/node/node*; do echo "$n cpus=$(cat $n/cpulist) $(grep -E 'MemTotal|MemFree' $n/meminfo | tr -s ' ' | tr '\n' ' ')"; done
echo "== cgroup"; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective /sys/fs/cgroup/cpu.max /sys/fs/cgroup/memory.max 2>/dev/null
echo "== nvidia-smi topo"; nvidia-smi topo -m
echo "== gpu pcie"; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current --format=csv
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ -e $d/numa_node ]; then echo "$d numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist) class=$(cat $d/class)"; fi; done
echo "== virtualization"; (systemd-detect-virt 2>/dev/null; grep -m1 hypervisor /proc/cpuinfo | head -c 200; cat /sys/class/dmi/id/product_name 2>/dev/null)
echo "== numactl"; which numactl && numactl -H
