#!/usr/bin/env python
"""Device-to-host copy bandwidth into pinned memory, by chunk size and number of streams (the e2e path's ceiling)."""
import time, torch
dev = torch.device("cuda", 0)
tot = 4 << 30
src = torch.empty(tot, dtype=torch.uint8, device=dev)
dst = torch.empty(tot, dtype=torch.uint8, pin_memory=True)
for chunk_mb in (64, 256, 1024, 4096):
    for nstreams in (1, 2):
        chunk = chunk_mb << 20
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
        torch.cuda.synchronize()
        best = 0.0
        for rep in range(3):
            t0 = time.perf_counter()
            for i, off in enumerate(range(0, tot, chunk)):
                with torch.cuda.stream(streams[i % nstreams]):
                    dst[off:off + chunk].copy_(src[off:off + chunk], non_blocking=True)
            torch.cuda.synchronize()
            best = max(best, tot / (time.perf_counter() - t0) / 1e9)
        print(f"chunk {chunk_mb:5d} MiB x {nstreams} stream(s): {best:.1f} GB/s", flush=True)
# host -> device for comparison
t0 = time.perf_counter(); src.copy_(dst, non_blocking=True); torch.cuda.synchronize()
print(f"H2D 4 GiB: {tot / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
