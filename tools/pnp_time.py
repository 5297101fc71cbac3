"""EPnP-RANSAC batch timing (bench.bench_pnp) - run under `ncu --metrics gpu__time_duration.sum -k regex:k_pnp` for the
per-kernel split of one corb_pnp_iterate_batch call."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    print(json.dumps(bench.bench_pnp(0, with_cpu=len(sys.argv) > 2, reps=reps)))
