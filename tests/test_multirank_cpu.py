"""World-size-2 test of the multi-GPU host logic on CPU (gloo): contiguous cell-range partition, per-rank slabs whose
concatenation equals the single-process result, and the max-over-ranks timing reduction bench.py uses.  The per-rank
compute here is the ORACLE (no GPU in this container); on GPUs each rank calls integ.* on its own range instead and the
data path has no collective."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gemlab_b200 as g
    from gemlab_b200 import Block, GeoKind
    from oracle import pyoracle as po

    mesh = Block.new_cube(1.0).set_ndiv([4, 3, 5]).subdivide(GeoKind.Hex8)
    mesh = g.jitter_interior_points(mesh, 0.02)
    b, e = g.partition(mesh.ncell, world, rank)
    dd = g.mandel_matrix(g.lin_elasticity(480.0, 1 / 3, False))
    part = po.mesh_integ("mat_10_bdb", 3, mesh.coords, mesh.kinds, mesh.conn_off, mesh.conn, b, e, coef=dd)
    assert part.size == (e - b) * 576
    np.save(os.path.join(tmpdir, f"part{rank}.npy"), part)
    # bench.py's timing reduction: MAX over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    dist.barrier()
    if rank == 0:
        full = po.mesh_integ("mat_10_bdb", 3, mesh.coords, mesh.kinds, mesh.conn_off, mesh.conn, coef=dd)
        parts = [np.load(os.path.join(tmpdir, f"part{r}.npy")) for r in range(world)]
        np.testing.assert_array_equal(np.concatenate(parts), full)
    dist.destroy_process_group()


def test_two_rank_partition_concatenates_to_the_single_rank_result(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
