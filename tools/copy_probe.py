"""How long do the per-tick copies take on this box?  Back-to-back pinned copies of the tick's sizes, CUDA events around them."""
import torch
dev = torch.device("cuda", 0)
for nbytes, name in [(139264, "h2d"), (131072, "d2h"), (16384, "h2d"), (16384, "d2h"), (1 << 20, "h2d"), (1 << 20, "d2h"), (64 << 20, "h2d"), (64 << 20, "d2h")]:
    hbuf = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
    dbuf = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(4)]
    for ns in (1, 4):
        streams = [torch.cuda.Stream() for _ in range(ns)]
        reps = 400 if nbytes < (8 << 20) else 20
        def run():
            for k in range(reps):
                with torch.cuda.stream(streams[k % ns]):
                    if name == "h2d":
                        dbuf[k % 4].copy_(hbuf[k % 4], non_blocking=True)
                    else:
                        hbuf[k % 4].copy_(dbuf[k % 4], non_blocking=True)
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time
        t0 = time.perf_counter()
        e0.record(); run()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("%s %8d B x %d on %d stream(s): %.2f us per copy wall (%.2f us host enqueue), %.1f GB/s"
              % (name, nbytes, reps, ns, (t2 - t0) / reps * 1e6, (t1 - t0) / reps * 1e6, nbytes * reps / (t2 - t0) / 1e9))
# both directions at once
nb = 139264
h1 = torch.empty(nb, dtype=torch.uint8).pin_memory(); d1 = torch.empty(nb, dtype=torch.uint8, device=dev)
h2 = torch.empty(131072, dtype=torch.uint8).pin_memory(); d2 = torch.empty(131072, dtype=torch.uint8, device=dev)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
import time
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(400):
        with torch.cuda.stream(sa):
            d1.copy_(h1, non_blocking=True)
        with torch.cuda.stream(sb):
            h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
print("h2d 139264 + d2h 131072 on two streams: %.2f us per pair" % ((t2 - t0) / 400 * 1e6))
import subprocess
print(subprocess.run(["nvidia-smi", "-q", "-d", "PERFORMANCE"], capture_output=True, text=True).stdout[:0])
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv; nproc; lscpu | grep -i 'numa\\|model name\\|socket'", shell=True, capture_output=True, text=True).stdout)
