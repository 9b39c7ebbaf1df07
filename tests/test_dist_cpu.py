"""World-size-2 gloo test of the multi-GPU plumbing (runs on CPU): sharding covers every stream
exactly once and the weight/embedding broadcast makes all ranks identical."""
import os
import subprocess
import sys
import textwrap

from lookoncetohear_b200 import dist as l2h_dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 2048, 2049):
        for world in (1, 2, 3, 8):
            spans = [l2h_dist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


WORKER = textwrap.dedent("""
    import os, sys, torch
    sys.path.insert(0, %r)
    from lookoncetohear_b200 import Net, dist as d, synth
    from lookoncetohear_b200.configs import TSH_PARAMS
    rank, world = d.init_from_env("gloo")
    torch.manual_seed(100 + rank)                       # ranks start with DIFFERENT weights
    net = Net(**TSH_PARAMS)
    emb = synth.embedding(1, seed0=10 + rank)
    d.broadcast_module(net, src=0)
    d.broadcast_tensor(emb, src=0)
    s = sum(float(p.double().sum()) for p in net.parameters()) + float(emb.double().sum())
    sums = d.gather_counts(s)
    lo, hi = d.shard_range(2048, rank, world)
    sizes = d.gather_counts(hi - lo)
    assert net._dirty
    if rank == 0:
        assert abs(sums[0] - sums[1]) < 1e-9, sums
        assert sum(sizes) == 2048, sizes
        print("OK")
    torch.distributed.destroy_process_group()
""") % ROOT


def test_broadcast_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0][0]
