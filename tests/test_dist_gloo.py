"""world_size-2 gloo test of the N>1 plumbing bench.py uses (runs on CPU): barrier, MAX-over-ranks
timing, deterministic cost-balanced sharding that covers every image exactly once."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, %r)
    from rocjpeg_b200 import dist
    rank, world = dist.init("gloo")
    assert world == 2
    costs = [((i * 7919) %% 97) + 1 for i in range(37)]
    shards = dist.shard_by_cost(costs, world)
    mine = shards[rank]
    dist.barrier()
    mx = dist.max_over_ranks([1.0 + rank, 5.0 - rank])
    total = dist.sum_over_ranks([len(mine), sum(costs[i] for i in mine)])
    with open(os.path.join(sys.argv[1], "rank%%d.json" %% rank), "w") as f:
        json.dump({"rank": rank, "mine": mine, "max": mx, "total": total}, f)
    dist.finalize()
""") % ROOT


def test_two_rank_gloo_plumbing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script), str(tmp_path)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    import json

    lines = [json.load(open(tmp_path / ("rank%d.json" % k))) for k in (0, 1)]
    by_rank = {l["rank"]: l for l in lines}
    assert sorted(by_rank[0]["mine"] + by_rank[1]["mine"]) == list(range(37))
    assert not set(by_rank[0]["mine"]) & set(by_rank[1]["mine"])
    for l in lines:
        assert l["max"] == [2.0, 5.0]
        assert l["total"][0] == 37.0
    costs = [((i * 7919) % 97) + 1 for i in range(37)]
    loads = [sum(costs[i] for i in by_rank[r]["mine"]) for r in (0, 1)]
    assert abs(loads[0] - loads[1]) <= max(costs)


def test_shard_by_cost_single_rank():
    sys.path.insert(0, ROOT)
    from rocjpeg_b200 import dist

    assert dist.shard_by_cost([3, 1, 2], 1) == [[0, 1, 2]]
    assert dist.max_over_ranks([1.5]) == [1.5]
