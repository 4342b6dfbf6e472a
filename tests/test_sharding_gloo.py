"""World-size-2 gloo test of the multi-GPU host logic of bench.py: frames shard over ranks with no data-path
collective; the only collectives are the timing barrier / MAX reduction and (for C4) the one-off vocabulary broadcast."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'rgbd-pl-slam_b200'))
    import numpy as np, torch, torch.distributed as dist
    import argparse, bench
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    a = argparse.Namespace(batch=4, width=160, height=120)
    frames = bench.make_frames(a, rank)
    # every rank owns a disjoint shard of the global frame set
    digest = torch.tensor([int(frames.astype(np.int64).sum())])
    allsum = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(allsum, digest)
    assert len({int(t) for t in allsum}) == world, 'ranks must generate different shards'
    # timing reduction: MAX over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == 10.0 + world - 1
    # one-off broadcast of a flattened vocabulary blob from rank 0 (init only; nothing on the per-frame path)
    voc = torch.arange(1000, dtype=torch.int32) if rank == 0 else torch.zeros(1000, dtype=torch.int32)
    dist.broadcast(voc, 0)
    assert int(voc.sum()) == 499500
    dist.barrier()
    if rank == 0:
        print('gloo sharding ok', world)
    dist.destroy_process_group()
""") % (ROOT, ROOT)


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "gloo sharding ok 2" in r.stdout
