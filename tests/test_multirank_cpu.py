"""world_size-2 gloo test (CPU) of the multi-rank path: videos are the shard unit, no
data-path collective, timing is the max over ranks (bench.py contract)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gloo_sharding_and_reductions(tmp_path, oracle):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "_rank_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    outs = [json.load(open(tmp_path / f"rank{k}.json")) for k in range(2)]
    assert [o["world"] for o in outs] == [2, 2]
    # whole videos, disjoint, complete: video_id mod world
    assert outs[0]["mine"] == [0, 2, 4] and outs[1]["mine"] == [1, 3, 5]
    # max over ranks of the per-rank time; sum of per-rank unit counts
    assert outs[0]["t_max"] == outs[1]["t_max"] == 2.0
    assert outs[0]["n_sum"] == outs[1]["n_sum"] == 6.0
    assert outs[0]["per_rank"] == outs[1]["per_rank"] == [1.0, 2.0]
    # every shard tracked its own videos correctly (found the disk centre)
    merged = {}
    for o in outs:
        merged.update(o["res"])
    assert sorted(int(k) for k in merged) == list(range(6))
    for v, (i, j, ci, cj) in merged.items():
        assert (i, j) == (ci, cj), v


def test_shard_plan_properties():
    sys.path.insert(0, ROOT)
    import bench
    for world in (1, 2, 4, 8):
        seen = []
        for rank in range(world):
            seen += bench.shard_videos(256, world, rank)
        assert sorted(seen) == list(range(256))
    # strong scaling = BASELINE configs[2] as written: 256 videos in total → 256 / world per rank
    assert [len(bench.shard_videos(256, 8, r)) for r in range(8)] == [32] * 8
    assert bench.workload_config("strong", 8)["videos_per_gpu"] == 32 and bench.workload_config("weak", 8)["videos_per_gpu"] == 256
    a = bench.algorithmic_per_window()
    # tw=25 / 45x45 → 0.9009 M MAC, 1.804 MFLOP, 11,897 B (u8 frames)
    assert a["mac"] == 900900 and a["flops"] == 1803825 and a["bytes"] == 11897
    ff = bench.algorithmic_per_window(65, 1080, 1920)
    assert abs(ff["mac"] - 555.1e6) < 0.1e6


def test_reference_arm_runs_on_rank0_only(tmp_path):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""           # non-zero ranks exit 0 without work
    env["RANK"] = "0"; env["LOCAL_RANK"] = "0"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["positions_correct"] is True
