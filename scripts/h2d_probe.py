"""Raw host->device ceiling of the box: N processes (one per GPU), pinned buffers, concurrent cudaMemcpyAsync.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/h2d_probe.py [numa]
With `numa` every process first binds itself to the cores of its GPU's NUMA node (from sysfs), so that the pinned
pages it then allocates and touches are local to the GPU's PCIe root.  Prints one JSON line (rank 0)."""
import glob, json, os, sys, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
mode = sys.argv[1] if len(sys.argv) > 1 else "default"
GB = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0


def gpu_numa_node(index):
    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id  # torch >= 2.x
    except Exception:
        bus = None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        busid = pynvml.nvmlDeviceGetPciInfo(h).busId
        busid = busid.decode() if isinstance(busid, bytes) else busid
        busid = busid.lower()
        if len(busid.split(":")[0]) == 8:
            busid = busid[4:]
        with open(f"/sys/bus/pci/devices/{busid}/numa_node") as f:
            return int(f.read().strip()), busid
    except Exception as e:
        return -1, str(e)


def node_cpus(node):
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            out = []
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                out += list(range(int(a), int(b or a) + 1))
            return out
    except Exception:
        return []


torch.cuda.set_device(local)
node, busid = gpu_numa_node(local)
allowed = sorted(os.sched_getaffinity(0))
bound = False
if mode == "numa" and node >= 0:
    cpus = [c for c in node_cpus(node) if c in allowed]
    if cpus:
        os.sched_setaffinity(0, cpus); bound = True
if world > 1:
    dist.init_process_group("gloo")
n = int(GB * (1 << 30)) // 8
host = torch.empty(n, dtype=torch.float64, pin_memory=True)
host.fill_(1.0)                                  # first touch under the (possibly NUMA-bound) affinity
dev = torch.empty(n, dtype=torch.float64, device="cuda")
dev.copy_(host, non_blocking=True); torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 5
t0 = time.perf_counter()
for _ in range(reps):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = reps * n * 8 / dt / 1e9
if world > 1:
    t = torch.tensor([gbs, dt], dtype=torch.float64)
    lst = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(lst, t)
    per = [float(v[0]) for v in lst]; tmax = max(float(v[1]) for v in lst)
else:
    per, tmax = [gbs], dt
nodes = len(glob.glob("/sys/devices/system/node/node[0-9]*"))
if rank == 0:
    print(json.dumps({"probe": "h2d", "mode": mode, "n_gpus": world, "buffer_GB": GB, "per_gpu_GBs": [round(p, 2) for p in per],
                      "aggregate_GBs": round(world * reps * n * 8 / tmax / 1e9, 2), "gpu0_numa_node": node, "gpu0_bus": busid,
                      "numa_nodes": nodes, "host_cpus_allowed": len(allowed), "bound": bound}))
if world > 1:
    dist.destroy_process_group()
