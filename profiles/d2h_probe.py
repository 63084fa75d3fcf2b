"""Per-rank host <-> device copy bandwidth on a multi-GPU box, alone and all together (why does the
end-to-end number scale 1.77x on 8 GPUs).  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 profiles/d2h_probe.py

Every rank copies a 512 MiB device buffer to pinned host memory (and back) with CUDA-event timing: (a) one rank at a
time while the others idle, (b) all ranks at once, (c) all at once after binding each process (and therefore the
first-touch of its pinned buffer) to the CPUs NVML reports as local to its GPU, (d) with cudaHostAllocWriteCombined.
Rank 0 prints one JSON object.
"""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

MB = 1 << 20
SIZE = 512 * MB


def bandwidth(dst, src, reps=6):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return reps * SIZE / (e0.elapsed_time(e1) * 1e-3) / 1e9


def host_alloc(flags):
    """cudaHostAlloc with explicit flags -> (uint8 tensor view, pointer)."""
    rt = ctypes.CDLL("libcudart.so.12")
    ptr = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(SIZE), ctypes.c_uint(flags))
    if rc != 0:
        return None, None
    buf = (ctypes.c_uint8 * SIZE).from_address(ptr.value)
    t = torch.frombuffer(buf, dtype=torch.uint8)
    return t, ptr


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.empty(SIZE, dtype=torch.uint8, device="cuda")
    res = {}

    def gather(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [round(float(o.item()), 1) for o in out]

    def run(tag, host):
        alone_d2h, alone_h2d = 0.0, 0.0
        for r in range(world):  # one rank at a time
            dist.barrier()
            if r == rank:
                alone_d2h, alone_h2d = bandwidth(host, dev), bandwidth(dev, host)
            dist.barrier()
        dist.barrier()
        all_d2h = bandwidth(host, dev, reps=12)
        dist.barrier()
        all_h2d = bandwidth(dev, host, reps=12)
        dist.barrier()
        res[tag] = {"d2h_alone_GBps": gather(alone_d2h), "h2d_alone_GBps": gather(alone_h2d),
                    "d2h_all_GBps": gather(all_d2h), "h2d_all_GBps": gather(all_h2d)}
        res[tag]["d2h_all_sum"] = round(sum(res[tag]["d2h_all_GBps"]), 1)
        res[tag]["h2d_all_sum"] = round(sum(res[tag]["h2d_all_GBps"]), 1)

    aff0 = sorted(os.sched_getaffinity(0))
    run("default_pinned", torch.empty(SIZE, dtype=torch.uint8).pin_memory())
    # bind to the GPU's CPUs, then allocate (first touch on that node)
    from topo_descriptors_b200 import bands

    bound = bands.bind_to_gpu_numa(local)
    aff1 = sorted(os.sched_getaffinity(0))
    run("numa_bound_pinned", torch.empty(SIZE, dtype=torch.uint8).pin_memory())
    wc, _keep = host_alloc(0x04)  # cudaHostAllocWriteCombined
    if wc is not None:
        run("write_combined", wc)
    if rank == 0:
        info = {"world": world, "size_MiB": SIZE // MB, "affinity_before": f"{aff0[0]}-{aff0[-1]} ({len(aff0)} cpus)",
                "affinity_after_bind": f"{aff1[0]}-{aff1[-1]} ({len(aff1)} cpus)", "bind_ok": bool(bound)}
        try:
            info["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
            info["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[-2500:]
            info["lscpu"] = [l for l in subprocess.run(["lscpu"], capture_output=True, text=True, timeout=20).stdout.splitlines()
                             if any(k in l for k in ("Model name", "Socket", "NUMA", "CPU(s):"))]
        except Exception as e:  # noqa: BLE001
            info["topo_error"] = str(e)
        print(json.dumps({"info": info, "results": res}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
