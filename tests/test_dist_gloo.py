"""world_size-2 gloo test of the multi-GPU host logic: plan broadcast + batch sharding (no GPU needed)."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import acvm_b200
    from acvm_b200 import acir_builder as ab
    from acvm_b200.dist import broadcast_bytes, gather_status_counts, shard_range
    import plan_interp
    blob = None
    data, inputs, _ = ab.synthetic_arith_circuit(200)
    if rank == 0:
        _, blob = acvm_b200.compile_plan_host(data, inputs, 16)
    got = broadcast_bytes(blob, src=0)
    plan = plan_interp.PlanBlob(got)
    lo, hi = shard_range(11, rank, world)
    inp = ab.synthetic_inputs(hi - lo, first_instance=lo)
    n_ok = 0
    digest = 0
    for i in range(hi - lo):
        iw = {w: int.from_bytes(inp[(i * 8 + k) * 32:(i * 8 + k + 1) * 32], "big") for k, w in enumerate(inputs)}
        st, wm = plan_interp.run_plan(plan, iw)
        n_ok += st[0] == "Solved"
        digest ^= hash(tuple(sorted(wm.items())))
    tot_ok, tot_fail = gather_status_counts(n_ok, (hi - lo) - n_ok)
    q.put((rank, lo, hi, len(got), tot_ok, tot_fail))
    dist.destroy_process_group()


def test_broadcast_and_shard_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, n0, ok0, f0), (r1, lo1, hi1, n1, ok1, f1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 6, 6, 11)
    assert n0 == n1 > 0
    assert ok0 == ok1 == 11 and f0 == f1 == 0


def test_shard_range_covers_batch():
    from acvm_b200.dist import shard_range
    for batch in (0, 1, 7, 8192, 65536):
        for world in (1, 2, 3, 8):
            r = [shard_range(batch, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == batch
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
