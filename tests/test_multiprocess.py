"""CPU test of the N>1 launch path (gloo, world_size 2): parameter PODs are
broadcast from rank 0, frames are sharded with no overlap, timing is the max
over ranks.  The GPU run uses the same helpers with the nccl backend."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200"))
    import bench
    import pcc_attr_b200 as pb

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if rank == 0:
        params, qpset = pb.RahtParams(), pb.QpSet()
        params.prediction_enabled, params.prediction_search_range, params.raht_extension = 1, 2500, 1
        for i in range(19):
            params.pred_weight_parent[i] = i + 1
        qpset.num_layers, qpset.max_qp = 1, 51
        qpset.layers[0][0], qpset.layers[0][1] = 34, -2
    else:
        params, qpset = pb.RahtParams(), pb.QpSet()
    raw = bench.broadcast_pods(bytes(params) + bytes(qpset), dist, torch.device("cpu"))
    p = pb.RahtParams.from_buffer_copy(raw[:C.sizeof(pb.RahtParams)])
    q = pb.QpSet.from_buffer_copy(raw[C.sizeof(pb.RahtParams):])
    t = bench.reduce_timing([10.0 + rank, 5.0 - rank], dist, torch.device("cpu"))
    out[rank] = (p.prediction_search_range, list(p.pred_weight_parent), q.layers[0][0], q.layers[0][1],
                 bench.frame_seeds(rank, 4), t)
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0[:4] == r1[:4] == (2500, list(range(1, 20)), 34, -2)   # PODs arrived intact
    assert not set(r0[4]) & set(r1[4])                              # disjoint frames
    assert r0[5] == r1[5] == [11.0, 5.0]                            # max over ranks
