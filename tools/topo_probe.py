"""Host topology as seen from a GPU box: NUMA nodes, the node / CPU affinity of every visible GPU (sysfs and NVML), nvidia-smi's
matrix.  Diagnostic for bench.py's bind_to_gpu_numa_node."""
import glob
import os
import subprocess


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e:
        return "ERR %s" % e


print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for n in sorted(glob.glob("/sys/devices/system/node/node*")):
    print(n, open(n + "/cpulist").read().strip(), sh("grep MemTotal %s/meminfo" % n))
print(sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"))
try:
    import pynvml
    pynvml.nvmlInit()
    for i in range(pynvml.nvmlDeviceGetCount()):
        h = pynvml.nvmlDeviceGetHandleByIndex(i)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        b = bus.lower()
        b = b[4:] if len(b.split(":")[0]) == 8 else b
        try:
            node = open("/sys/bus/pci/devices/%s/numa_node" % b).read().strip()
        except Exception as e:
            node = "ERR %s" % e
        words = (os.cpu_count() + 63) // 64
        try:
            aff = list(pynvml.nvmlDeviceGetCpuAffinity(h, words))
        except Exception as e:
            aff = "ERR %s" % e
        try:
            maff = list(pynvml.nvmlDeviceGetMemoryAffinity(h, 4, 0))
        except Exception as e:
            maff = "ERR %s" % e
        print("gpu", i, bus, "sysfs numa_node", node, "nvml cpu affinity", [hex(a) for a in aff] if isinstance(aff, list) else aff, "mem affinity", maff)
except Exception as e:
    print("nvml", e)
print(sh("nvidia-smi topo -m"))
print(sh("cat /proc/self/status | grep -i -E 'cpus_allowed_list|mems_allowed_list'"))
