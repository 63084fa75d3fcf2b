"""Device -> host copy rate of one GPU as a function of the pinned destination ring: does a ring that fits the host's
last-level cache (DMA writes can land there) beat a ring of whole 1 GiB outputs?  16 GiB per measurement."""
import json
import sys

import torch

dev = torch.device("cuda", 0)
src = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
src.fill_(7)
out = []
for chunk_mb, slots in ((1024, 3), (256, 4), (128, 4), (64, 4), (32, 4), (16, 8)):
    chunk = chunk_mb << 20
    ring = [torch.empty(chunk, dtype=torch.uint8).pin_memory() for _ in range(slots)]
    per = (1 << 30) // chunk
    def run(reps):
        k = 0
        for _ in range(reps):
            for c in range(per):
                ring[k % slots].copy_(src[c * chunk:(c + 1) * chunk], non_blocking=True)
                k += 1
    run(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(16)
    e1.record()
    torch.cuda.synchronize()
    gbps = 16 * (1 << 30) / (e0.elapsed_time(e1) * 1e-3) / 1e9
    out.append({"chunk_MiB": chunk_mb, "slots": slots, "GBps": round(gbps, 2)})
    print(out[-1], file=sys.stderr, flush=True)
    del ring
print(json.dumps({"d2h_ring_probe": out}))
