"""world_size-2 gloo test of the multi-GPU host logic (mytrim_b200/dist.py): contiguous sharding by
global primary index + all-reduce of the tally blocks reproduces the single-rank result exactly.
The per-rank engine here is the CPU oracle; the GPU path uses the same helpers with NCCL."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL = 301
BINS = 2048
SEED = 99


def _rank_tallies(lo, hi):
    from mytrim_b200 import capi, dist as mdist
    from tests import util
    with util.OracleEngine(util.ORC_RNG_PHILOX, tally_mask=capi.TALLY_VAC_DEPTH) as orc:
        c = util.setup_engine(orc, "cu_on_cu_1keV")
        orc.run(util.primaries_for(c, hi - lo), seed=SEED, first_index=lo)
        vac, repl = orc.vac_depth()
        return mdist.pack_host_tallies(orc.counters(), vac, repl, BINS)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from mytrim_b200 import dist as mdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = mdist.shard_range(N_TOTAL, rank, world)
    u, f = _rank_tallies(lo, hi)
    tu, tf = torch.from_numpy(u), torch.from_numpy(f)
    tu[12] = 1000 + rank            # rank-local bookkeeping slot (list length): must survive the join untouched
    red = mdist.TallyReducer(tu, tf)
    before = tu.clone()
    total_u, total_f = red.reduce(write_back=False)          # out of place: the blocks keep this rank's share
    assert torch.equal(tu, before)
    pu, pf = red.per_rank()
    assert torch.equal(pu[rank], before) and torch.equal(pu[:, :mdist.N_ADDITIVE_COUNTERS].sum(dim=0),
                                                         total_u[:mdist.N_ADDITIVE_COUNTERS])
    mdist.reduce_tallies(tu, tf)                             # in place (one collective)
    assert torch.equal(tu, total_u) and torch.equal(tf, total_f) and int(tu[12]) == 1000 + rank
    tu[12] = 0
    np.save(os.path.join(out_dir, "u%d.npy" % rank), tu.numpy())
    np.save(os.path.join(out_dir, "f%d.npy" % rank), tf.numpy())
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    from mytrim_b200 import dist as mdist
    for n in (0, 1, 7, 1000, 12345):
        for w in (1, 2, 3, 8):
            r = [mdist.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def test_two_rank_gloo_reduction_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    u_all, f_all = _rank_tallies(0, N_TOTAL)
    for rank in range(2):
        u = np.load(tmp_path / ("u%d.npy" % rank))
        f = np.load(tmp_path / ("f%d.npy" % rank))
        assert np.array_equal(u, u_all)            # integer tallies: exact
        assert np.allclose(f, f_all, rtol=1e-12)   # f64 sums: order of addition differs
