"""Multi-GPU band check (run under torchrun on a GPU box; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Every rank computes its row band of TPI / STD / gradient through bands.sweep (halo exchange over NCCL) and
compares it BIT FOR BIT with the same rows of the whole-image result computed locally on its own GPU.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from topo_descriptors_b200 import bands, device as dev  # noqa: E402
from topo_descriptors_b200.device import DeviceDEM  # noqa: E402
from topo_descriptors_b200.synth import fractal_dem  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    ny, nx = 1500, 1111
    bad = 0
    for integer in (True, False):
        z = fractal_dem(ny, nx, seed=5, integer=integer)
        whole = DeviceDEM(torch.from_numpy(z).to(device))
        ctx = bands.BandContext(ny, nx, rank, world)
        core = whole.tensor[ctx.r0 : ctx.r1].contiguous()
        sizes = [5, 21, 67, 201, 301]
        sigmas = [1.25, 5.25, 16.75, 50.25, 75.25]
        rx = (torch.full((nx,), 25.0, dtype=torch.float64, device=device), 0)
        ry = (torch.full((ny,), -25.0, dtype=torch.float64, device=device), 0)
        got = {}
        bands.sweep(core, ctx, sizes, sigmas, rx, ry, sink=lambda n, i, t: got.__setitem__((n, i), t.clone()))
        # like with like: the bands run the cached multi-size sweep (shared planes; for float DEMs that means the
        # fixed-point scale of the largest size), so the whole-image reference runs it too
        whole.share_disc_planes(max(sizes))
        for i, size in enumerate(sizes):
            ref = {"tpi": dev.tpi(whole, size, pair_std=True), "std": dev.std(whole, size)}
            g = DeviceDEM(dev.gauss(whole, sigmas[i], sigmas[i]))
            outs = dev.gradient_from_smooth(g, g, rx[0], 0, ry[0], 0)
            ref.update(dict(zip(("dx", "dy", "slope", "aspect"), outs)))
            if i == len(sizes) - 1:
                whole.release_disc_planes()
            for name, r in ref.items():
                same = torch.equal(got[(name, i)], r[ctx.r0 : ctx.r1])
                if not same:
                    bad += 1
                    d = (got[(name, i)] - r[ctx.r0 : ctx.r1]).abs().max().item()
                    print(f"rank {rank}: MISMATCH {name} size {size} integer={integer} max diff {d}", flush=True)
    # valley / ridge and Sx bands (z-score statistics are reduced per band: equal up to the float32 rounding of
    # mean / std, so compare with a tolerance and count direction flips)
    from topo_descriptors_b200 import topo  # noqa: E402
    from topo_descriptors_b200.synth import dem_dataset  # noqa: E402

    z = fractal_dem(ny, nx, seed=6)
    whole = DeviceDEM(torch.from_numpy(z).to(device))
    ctx = bands.BandContext(ny, nx, rank, world)
    core = whole.tensor[ctx.r0 : ctx.r1].contiguous()
    for sigma in (None, 2.25):
        want = topo.valley_ridge(whole, 21, "valley", [0, 0.15, 0.3], sigma)
        n, d = bands.valley_ridge_band(core, ctx, 21, "valley", [0, 0.15, 0.3], sigma)
        dn = (n - want[0][ctx.r0 : ctx.r1]).abs().max().item()
        flips = (d != want[1][ctx.r0 : ctx.r1]).float().mean().item()
        if dn > 1e-4 or flips > 1e-3:
            bad += 1
            print(f"rank {rank}: MISMATCH valley_ridge sigma={sigma} norm diff {dn} direction flips {flips}", flush=True)
    plan = topo._sx_plan(dem_dataset(z, res=25.0), [45.0, 200.0], 2000.0, 10.0, 15, 0.0)
    ref = topo._sx_device(whole, plan, 10.0)
    if not torch.equal(bands.sx_band(core, ctx, plan, 10.0), ref[:, ctx.r0 : ctx.r1]):
        bad += 1
        print(f"rank {rank}: MISMATCH sx band", flush=True)
    t = torch.tensor([bad], device=device)
    dist.all_reduce(t)
    if rank == 0:
        print("mgpu_check:", "OK (bands bit-identical to the whole image)" if t.item() == 0 else f"{int(t.item())} mismatches")
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
