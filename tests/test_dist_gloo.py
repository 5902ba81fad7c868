"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: sequence sharding + the throughput reduction."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, %r)
    from orb_slam3_rgbl_b200 import dist as D
    r = D.Reporter("gloo")
    seqs = D.sequences_of_rank(8, r.rank, r.world)
    frames = 10 * len(seqs) + r.rank            # rank-dependent work
    elapsed = 1.0 + 0.5 * r.rank                # rank 1 is the slow one
    r.barrier()
    fps = r.throughput(frames, elapsed)
    mx = r.max_over_ranks(float(r.rank))
    print(json.dumps({"rank": r.rank, "world": r.world, "seqs": seqs, "fps": fps, "mx": mx}))
    r.close()
""") % str(ROOT)


def test_two_rank_gloo_sharding_and_reduction(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=120)
        assert p.returncode == 0, e[-2000:]
        outs.append(__import__("json").loads(o.strip().splitlines()[-1]))
    outs.sort(key=lambda d: d["rank"])
    assert outs[0]["seqs"] == [0, 2, 4, 6] and outs[1]["seqs"] == [1, 3, 5, 7]          # disjoint cover
    total = (10 * 4 + 0) + (10 * 4 + 1)
    for o in outs:
        assert o["world"] == 2 and abs(o["fps"] - total / 1.5) < 1e-9 and o["mx"] == 1.0   # slowest rank's clock


def test_single_process_is_identity():
    sys.path.insert(0, str(ROOT))
    from orb_slam3_rgbl_b200 import dist as D
    os.environ.pop("WORLD_SIZE", None); os.environ.pop("RANK", None)
    r = D.Reporter()
    assert r.world == 1 and r.throughput(32, 2.0) == 16.0 and D.sequences_of_rank(3, 0, 1) == [0, 1, 2]
