"""CPU tests of the N>1 plumbing: the shard planner of the in-process multi-GPU scheduler, and the one-rank-per-GPU
bench harness under torch.distributed with the gloo backend (world_size 2)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,g,max_b", [(256, 8, 256), (256, 4, 256), (1, 8, 256), (7, 8, 256), (100, 3, 16), (1000, 8, 64),
                                        (64, 8, 256), (17, 2, 4)])
def test_shard_plan_covers_the_batch_exactly_once(pkg, n, g, max_b):
    shards = pkg.plan_shards(n, g, max_b, 8, 0)
    covered = sorted((off, off + cnt) for _, off, cnt in shards)
    assert covered[0][0] == 0 and covered[-1][1] == n
    for (a0, a1), (b0, b1) in zip(covered, covered[1:]):
        assert a1 == b0                                     # contiguous, no overlap, no gap
    assert all(0 < cnt <= max_b for _, _, cnt in shards)    # never more than the arena holds
    assert all(0 <= r < g for r, _, _ in shards)
    per = {}
    for r, _, cnt in shards:
        per[r] = per.get(r, 0) + cnt
    if n >= 8 * 2 and g > 1:
        assert len(per) == min(g, n // 8)
        assert max(per.values()) - min(per.values()) <= 1   # balanced contiguous split
    else:
        assert len(per) == 1                                 # too small to split: one replica


def test_small_batches_round_robin_over_replicas(pkg):
    assert [pkg.plan_shards(1, 8, 256, 8, rr)[0][0] for rr in range(10)] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]
    assert pkg.plan_shards(0, 8) == []
    # bs256 on 8 GPUs: 32 images each at offsets g*32
    assert pkg.plan_shards(256, 8) == [(g, 32 * g, 32) for g in range(8)]


def test_default_min_shard_keeps_mid_size_requests_on_few_gpus(pkg):
    """Engine default (B200_ENGINE_MIN_SHARD = 32): a request is only cut into shards of >= 32 samples, so under mixed concurrent
    traffic a 64-image request occupies two GPUs, not eight, and anything below 64 goes whole to one GPU (round-robin)."""
    assert pkg.plan_shards(64, 8, 256, 32, 0) == [(0, 0, 32), (1, 32, 32)]
    assert pkg.plan_shards(128, 8, 256, 32, 0) == [(g, 32 * g, 32) for g in range(4)]
    assert pkg.plan_shards(256, 8, 256, 32, 0) == [(g, 32 * g, 32) for g in range(8)]
    assert pkg.plan_shards(63, 8, 256, 32, 5) == [(5, 0, 63)]
    assert pkg.plan_shards(100, 2, 256, 32, 0) == [(0, 0, 50), (1, 50, 50)]


_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
os.environ["B200_BENCH_BACKEND"] = "gloo"
import bench
rank, local, world, reduce_max, barrier, host_barrier = bench._dist()
assert world == 2 and rank in (0, 1) and local == rank
barrier()
m = reduce_max(10.0 + rank)          # MAX over ranks of the per-rank time
assert m == 11.0, m
barrier()
host_barrier()
import os, sys
os.write(1, ("RANK_OK %d\\n" % rank).encode())   # ONE write per rank: two ranks share the pipe and print() may interleave
'''


def _free_port():
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:  # a fixed port collided with a run moments earlier (TIME_WAIT)
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(args, env_extra=None, timeout=300):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = None
    for _ in range(2):  # one retry: the rendezvous port can still be taken between the probe and the launch
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), *args]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
        if r.returncode == 0:
            break
    return r


def test_bench_dist_helpers_with_gloo_world_size_2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT))
    r = _torchrun([str(script)])
    assert r.returncode == 0, r.stderr[-2000:]
    assert "RANK_OK 0" in r.stdout and "RANK_OK 1" in r.stdout


def test_reference_arm_under_torchrun_prints_one_line_from_rank0(repo_dir):
    r = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["unit"] == "img/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
