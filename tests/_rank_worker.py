"""Worker for tests/test_multirank_cpu.py: run under torch.distributed.run with the gloo backend.
Exercises bench.py's multi-rank plumbing on CPU: whole-video sharding, barrier, max/sum over
ranks — and tracks each rank's shard of videos with the CPU oracle so the union can be compared
with a single-process run.  (The GPU arm uses the same Ranks class with backend="nccl".)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import Oracle  # noqa: E402


def main():
    out_dir = sys.argv[1]
    ranks = bench.Ranks(backend="gloo")
    total = 6
    mine = bench.shard_videos(total, ranks.world, ranks.rank)
    orc = Oracle()
    res = {}
    for v in mine:
        rng = np.random.default_rng(v)
        H, W = 96, 128
        f = np.full((H, W), 128, np.uint8)
        cy, cx = int(rng.integers(20, H - 20)), int(rng.integers(20, W - 20))
        yy, xx = np.ogrid[0:H, 0:W]
        f[(yy - cy) ** 2 + (xx - cx) ** 2 <= 25] = 0
        r = orc.step(f, 128, 10, True, (21, 21), (cy + 3, cx - 2), dense=True)
        res[v] = [r.i, r.j, cy + 1, cx + 1]
    ranks.barrier()
    t_local = 1.0 + ranks.rank                      # pretend device time of this rank
    t_max = ranks.max_over_ranks(t_local)
    n_sum = ranks.sum_over_ranks(float(len(mine)))
    per_rank = ranks.gather(t_local)                # per-rank times in rank order (reported beside the max)
    ranks.barrier()
    with open(os.path.join(out_dir, f"rank{ranks.rank}.json"), "w") as fh:
        json.dump({"rank": ranks.rank, "world": ranks.world, "mine": mine, "res": res, "t_max": t_max, "n_sum": n_sum,
                   "per_rank": per_rank}, fh)
    ranks.close()


if __name__ == "__main__":
    main()
