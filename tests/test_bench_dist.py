"""CPU tests of bench.py's multi-rank plumbing (world_size 2, gloo) and of its reference arm."""
import json
import os
import subprocess
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local_ms = [10.0 + 5.0 * rank, 100.0 - 7.0 * rank]      # rank 1 is slower on the first, faster on the second
    got = bench.global_max(local_ms, torch.device("cpu"))
    q.put((rank, got))
    dist.barrier()
    dist.destroy_process_group()


def test_global_max_over_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1] == [15.0, 100.0]


def test_job_throughput_is_weak_scaling_aggregate():
    one = bench.job_throughput(1, 10, 100.0)
    assert one == 8 * 10 / 0.1
    assert bench.job_throughput(8, 10, 100.0) == 8 * one      # same time, 8x the pairs


def test_live_pairs_matches_survey():
    assert bench.live_pairs(38, 63, 8) == 513536


def test_reference_arm_prints_one_json_line_and_only_rank0():
    env = dict(os.environ, OMP_NUM_THREADS="8")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frame-pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["metric"] == bench.METRIC
    env["RANK"], env["WORLD_SIZE"] = "1", "2"
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=60)
    assert out.returncode == 0 and out.stdout.strip() == ""
