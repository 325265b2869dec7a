"""How the CPU baseline of bench.py --impl reference scales on this box: affinity / cgroup limits and instances/s of the
oracle port with 16, 32, 64, 128 single-threaded worker processes (2 instances each).  CPU only."""
import json
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (pins BLAS threads before NumPy loads)

if __name__ == "__main__":
    info = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/proc/loadavg"):
        try:
            info[path] = open(path).read().strip()
        except OSError:
            pass
    print(json.dumps(info))
    for workers in (16, 32, 64, 128):
        if workers > (os.cpu_count() or 1):
            break
        with mp.get_context("spawn").Pool(workers, initializer=bench._cpu_init) as pool:
            pool.map(bench._cpu_worker, [(0, 1)] * workers)
            t0 = time.perf_counter()
            times = pool.map(bench._cpu_worker, [(2 * w, 2 * w + 2) for w in range(workers)], chunksize=1)
            dt = time.perf_counter() - t0
        print(json.dumps({"workers": workers, "instances": 2 * workers, "wall_s": dt, "inst_per_s": 2 * workers / dt,
                          "mean_worker_s": sum(times) / len(times), "max_worker_s": max(times)}))
