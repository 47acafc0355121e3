"""Diagnostic: how the reference CPU arm of bench.py scales with worker processes / threads on this host."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "effective", bench.effective_cores(), "loadavg", os.getloadavg())
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us", "/sys/fs/cgroup/cpuset.cpus.effective"):
    try:
        print(f, open(f).read().strip())
    except Exception as e:
        print(f, "n/a")
os.system("lscpu | egrep 'Model name|Socket|Core|Thread|NUMA node\\(s\\)' ; free -g | head -2")
n = 256
streams = bench.make_batch(n, 0, 16)
mirror = np.random.default_rng(0).integers(0, 2, n)
for procs, threads in ((1, 1), (1, 8), (8, 1), (16, 1), (32, 1), (64, 1), (128, 1), (16, 8), (32, 4), (64, 2)):
    r = bench.CpuReference(streams, procs * threads)
    r.threads, r.procs = threads, procs
    if r.pool is not None:
        r.pool.close(); r.pool.join()
    import multiprocessing as mp
    r.pool = mp.get_context("fork").Pool(procs) if procs > 1 else None
    m = n if procs * threads >= 8 else 16
    r.run(min(m, 2 * procs * threads), mirror)
    dt, _ = r.run(m, mirror)
    print(f"procs {procs:4d} x threads {threads}: {m / dt:8.1f} img/s", flush=True)
    r.close()
