"""Two slab engines on two GPUs exchanging over NCCL, against the single-slab engine (rank 0)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("nl-partsol_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))

from nlps_b200 import engine  # noqa: E402
from slabcases import COMPARE, merge, moving_block  # noqa: E402
from util import assert_close, field_scales  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    nsteps = 60
    P = moving_block(nsteps=nsteps)
    axis, cuts = engine.slab_cuts(P, world)
    comm = engine.NcclComm(rank, world, local)
    eng = engine.Engine(P, device=local, slab=dict(rank=rank, world=world, axis=axis, cuts=cuts, comm=comm,
                                                   migrate_every=4))
    assert eng.initialize_lme() == 0
    assert eng.run(0, nsteps) == 0, eng.error()
    f, ids = eng.download_local()
    counts, lists = eng.lists()
    mine = (f, ids, counts, lists, eng.migrated_count())
    eng.close()
    comm.close()
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    if rank == 0:
        assert sum(r[4] for r in allr) > 100
        m = merge([r[:4] for r in allr], P.np_)
        e1 = engine.Engine(P, device=local)
        assert e1.initialize_lme() == 0 and e1.run(0, nsteps) == 0
        f1 = e1.download()
        c1, l1 = e1.lists()
        e1.close()
        assert np.array_equal(m["I0"], f1["I0"]) and np.array_equal(m["_lists"], l1) and np.array_equal(m["_counts"], c1)
        sc = field_scales(P)
        for k in COMPARE:
            assert_close(m[k], f1[k], "nccl slabs vs single: " + k, scale=sc.get(k))
        print("NCCL slabs OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
