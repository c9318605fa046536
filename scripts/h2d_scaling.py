"""Why the host-logits e2e leg stops scaling at N = 8: per-rank pinned H2D bandwidth with all
N ranks copying at once, plus the box's topology. Run under torchrun; prints one JSON line on rank 0."""
import json, os, subprocess, sys, time
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 2 << 30
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
reps = 6
for _ in range(reps):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
bw = torch.tensor([reps * nbytes / dt / 1e9], dtype=torch.float64, device=dev)
allbw = [torch.zeros_like(bw) for _ in range(world)]
if world > 1:
    dist.all_gather(allbw, bw)
else:
    allbw = [bw]
aff = sorted(os.sched_getaffinity(0))
affs = [None] * world
if world > 1:
    dist.all_gather_object(affs, (len(aff), aff[0], aff[-1]))
else:
    affs = [(len(aff), aff[0], aff[-1])]
if rank == 0:
    def sh(cmd):
        try:
            return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout
        except Exception as e:
            return f"unavailable: {e}"
    per = [round(float(x.item()), 1) for x in allbw]
    print(json.dumps({"n_gpus": world, "pinned_h2d_gbs_per_rank": per, "aggregate_gbs": round(sum(per), 1),
                      "rank_cpu_affinity(count,first,last)": affs, "host_cpu_count": os.cpu_count(),
                      "nvidia_smi_topo": sh("nvidia-smi topo -m | head -20"), "numa": sh("numactl -H 2>/dev/null || lscpu | grep -i numa"),
                      "meminfo": sh("grep -E 'MemTotal|MemAvailable' /proc/meminfo")}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
