"""Host <-> device copy rates of N ranks on one box, alone and all at once (the end-to-end step of bench.py moves
174 MB each way per rank and step):  torchrun --nproc-per-node N profiles/pcie_scaling.py
Prints the topology the container can see, then per test the per-rank GB/s (min / median / max over ranks) and the sum.
Plain cudaMemcpyAsync between a cudaHostAlloc'ed buffer and device memory through torch; no engine code involved."""
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

MB = 174.2  # one step's unary / Q in MB


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        for cmd in ("nvidia-smi topo -m", "lscpu | grep -E 'Model name|Socket|NUMA|^CPU\\(s\\)'", "ls /sys/devices/system/node/",
                    "cat /sys/devices/system/node/node*/meminfo | grep MemTotal", "nproc", "cat /proc/self/status | grep -E 'Cpus_allowed_list|Mems_allowed_list'"):
            r = subprocess.run(cmd, shell=True, capture_output=True, text=True)
            print("$ " + cmd + "\n" + (r.stdout + r.stderr).rstrip(), flush=True)
    n = int(MB * 1e6 / 4)
    host_in = torch.empty(n, dtype=torch.float32).pin_memory()
    host_out = torch.empty(n, dtype=torch.float32).pin_memory()
    host_in.fill_(1.0)
    host_out.fill_(0.0)   # first touch by this rank's CPU
    dev_a = torch.empty(n, dtype=torch.float32, device="cuda")
    dev_b = torch.ones(n, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=6):
        fn()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        dev_a.copy_(host_in, non_blocking=True)

    def d2h():
        host_out.copy_(dev_b, non_blocking=True)

    def both():
        with torch.cuda.stream(s1):
            dev_a.copy_(host_in, non_blocking=True)
        with torch.cuda.stream(s2):
            host_out.copy_(dev_b, non_blocking=True)

    def d2h_chunks(k):
        def f():
            step = n // k
            for c in range(k):
                with torch.cuda.stream(s1 if c % 2 == 0 else s2):
                    host_out[c * step:(c + 1) * step].copy_(dev_b[c * step:(c + 1) * step], non_blocking=True)
        return f

    def report(name, secs, bytes_per_rank, active):
        gbs = torch.tensor([bytes_per_rank / secs / 1e9 if active else 0.0], dtype=torch.float64, device="cuda")
        if world > 1:
            allv = [torch.zeros_like(gbs) for _ in range(world)]
            dist.all_gather(allv, gbs)
            v = sorted(float(x.item()) for x in allv if float(x.item()) > 0)
        else:
            v = [float(gbs.item())]
        if rank == 0:
            print("%-52s per rank GB/s min %.1f med %.1f max %.1f   sum %.1f" % (name, v[0], v[len(v) // 2], v[-1], sum(v)), flush=True)

    B = n * 4
    # each rank alone (the others wait at the barrier inside timed())
    for r in range(world):
        if rank == r:
            ta, tb = timed(h2d), timed(d2h)
        else:
            barrier(); barrier()
            ta = tb = 1.0
        # keep ranks aligned
        if world > 1:
            dist.barrier()
        if rank == r:
            print("rank %d alone: H2D %.1f GB/s, D2H %.1f GB/s" % (r, B / ta / 1e9, B / tb / 1e9), flush=True)
    report("all ranks H2D at once", timed(h2d), B, True)
    report("all ranks D2H at once", timed(d2h), B, True)
    report("all ranks H2D + D2H at once (two streams)", timed(both), 2 * B, True)
    report("all ranks D2H at once, 4 chunks on two streams", timed(d2h_chunks(4)), B, True)
    report("all ranks D2H at once, 16 chunks on two streams", timed(d2h_chunks(16)), B, True)
    # half the ranks upload while the other half download
    if world > 1:
        t = timed(h2d if rank % 2 == 0 else d2h)
        report("even ranks H2D while odd ranks D2H (even ranks)", t, B, rank % 2 == 0)
        report("even ranks H2D while odd ranks D2H (odd ranks)", t, B, rank % 2 == 1)
        for k in (2, 4):
            if k < world:
                t = timed(d2h) if rank < k else (barrier(), 1.0)[1]
                report("only ranks 0..%d D2H at once" % (k - 1), t, B, rank < k)
    # small result: the int16 label map (4 MB)
    small_h = torch.empty(1920 * 1080, dtype=torch.int16).pin_memory()
    small_d = torch.zeros(1920 * 1080, dtype=torch.int16, device="cuda")
    report("all ranks D2H of the 4 MB label map at once", timed(lambda: small_h.copy_(small_d, non_blocking=True), reps=20), small_h.numel() * 2, True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
