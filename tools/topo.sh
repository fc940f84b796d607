nvidia-smi topo -m 2>&1 | head -14
nproc; python -c "import os; print(len(os.sched_getaffinity(0)))"
cat /sys/devices/system/node/online 2>/dev/null
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class 2>/dev/null; then echo $(basename $d) numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist); fi; done
