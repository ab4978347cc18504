"""Host topology facts that decide where pinned staging memory can live (run on the GPU box)."""
import glob, os, subprocess
def rd(p):
    try: return open(p).read().strip()
    except Exception as e: return f"<{e.__class__.__name__}>"
print("nodes:", sorted(os.path.basename(p) for p in glob.glob("/sys/devices/system/node/node[0-9]*")))
for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    print(os.path.basename(p), "cpulist", rd(p + "/cpulist"), "|", rd(p + "/meminfo").splitlines()[0:2])
for k in ("Cpus_allowed_list", "Mems_allowed_list"):
    for line in rd("/proc/self/status").splitlines():
        if line.startswith(k): print(line)
print("cpuset.mems:", rd("/sys/fs/cgroup/cpuset.mems.effective"), "cpuset.cpus:", rd("/sys/fs/cgroup/cpuset.cpus.effective"))
print("nproc:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
print(subprocess.run("lscpu | egrep -i 'model name|socket|numa|thread|hypervisor|^CPU\\(s\\)'", shell=True, capture_output=True, text=True).stdout)
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
print(subprocess.run("nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv", shell=True, capture_output=True, text=True).stdout)
for d in glob.glob("/sys/bus/pci/devices/*"):
    if rd(d + "/vendor") == "0x10de" and rd(d + "/class").startswith("0x0302"):
        print(os.path.basename(d), "numa_node", rd(d + "/numa_node"), "local_cpulist", rd(d + "/local_cpulist"))
print(subprocess.run("grep -E 'MemTotal|MemFree|Hugepagesize' /proc/meminfo; dmidecode -t memory 2>/dev/null | egrep -i 'speed|locator' | head -20", shell=True, capture_output=True, text=True).stdout)
