"""World-size-2 gloo test of the N>1 host logic (CPU): contiguous frame sharding, the counter all-reduce and
the max-over-ranks timing reduce must reproduce the single-process result.  The oracle stands in for the GPU
decoder here (tests may use it as the checker); the GPU kernels themselves are covered by -m gpu."""
import os
import subprocess
import sys

import numpy as np

import common

WORKER = r'''
import os, sys, json
import numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import common
from oracle import polar_oracle as po
from quantized_decoder_polar_codes_b200 import distributed as D
rank, local_rank, world = D.init("gloo")
kw, x, truth = common.make_case("SCLLUTDecoder", N=128, K=32, L=4, B=101, seed=7, tables="minsum", ebn0_db=0.5)
lo, hi = D.shard_range(x.shape[0], rank, world)
dec = po.OracleDecoder("SCLLUTDecoder", **kw)
y = dec.decode(x[lo:hi])
err = (y != truth[lo:hi])
cnt = torch.tensor([int(err.sum()), int(err.any(axis=1).sum()), hi - lo], dtype=torch.int64)
D.allreduce_counters(cnt)
tmax = D.max_over_ranks(1.0 + rank)
D.barrier()
if rank == 0:
    print("RESULT " + json.dumps({"cnt": cnt.tolist(), "tmax": tmax, "world": world}))
torch.distributed.destroy_process_group()
'''


def test_shard_ranges_cover_batch():
    from quantized_decoder_polar_codes_b200 import distributed as D
    for B in [0, 1, 7, 100, 101, 131072]:
        for world in [1, 2, 3, 4, 8]:
            parts = [D.shard_range(B, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == B
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_counters_match_single_process(tmp_path):
    from oracle import polar_oracle as po
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", str(script), common.ROOT],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][0]
    import json
    res = json.loads(line[7:])
    kw, x, truth = common.make_case("SCLLUTDecoder", N=128, K=32, L=4, B=101, seed=7, tables="minsum", ebn0_db=0.5)
    y = po.OracleDecoder("SCLLUTDecoder", **kw).decode(x)
    err = y != truth
    assert res["world"] == 2
    assert res["cnt"] == [int(err.sum()), int(err.any(axis=1).sum()), 101]
    assert res["tmax"] == 2.0
