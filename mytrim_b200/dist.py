"""Multi-GPU plumbing: how primaries shard over ranks and how tallies are joined.

The hot path shards trivially: every primary and its whole cascade is independent,
Philox streams are keyed by the GLOBAL primary index, so rank r of W simply takes a contiguous index
range and no data-path collective exists.  The only exchange is the final tally reduction —
the analogue of runmytrim's threadJoin (runmytrim.C:316-323) — one all-reduce(sum) over the
additive u64 block [counters | histograms] and one over the f64 block [EelTotal, EnucTotal].
Works on any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np

# layout of the engine's u64 tally block (mtb_types.h: CNT_*)
N_ADDITIVE_COUNTERS = 9   # vacancies .. hist_clamped
IDX_STACK_MAX = 9         # reduced with max, not sum
N_COUNTER_SLOTS = 16      # histograms start here


def shard_range(n_total, rank, world):
    """Contiguous [lo, hi) range of global primary indices owned by `rank`."""
    lo = n_total * rank // world
    hi = n_total * (rank + 1) // world
    return lo, hi


def reduce_tallies(u64_block, f64_block, group=None):
    """In-place all-reduce of the additive tallies.  `u64_block` is an int64 view of the engine's
    u64 block (sums of non-negative counts are bit-identical in either signedness)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(u64_block[:N_ADDITIVE_COUNTERS], group=group)
    dist.all_reduce(u64_block[IDX_STACK_MAX:IDX_STACK_MAX + 1], op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(u64_block[N_COUNTER_SLOTS:], group=group)
    dist.all_reduce(f64_block, group=group)


class DeviceView:
    """__cuda_array_interface__ wrapper so torch can alias the engine's device tallies."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def tally_tensors(engine):
    """(int64 tensor, float64 tensor) aliasing the engine's device tally blocks."""
    import torch
    pu, nu, pf, nf = engine.tally_device_views()
    return (torch.as_tensor(DeviceView(pu, nu, "<i8"), device="cuda"),
            torch.as_tensor(DeviceView(pf, nf, "<f8"), device="cuda"))


def pack_host_tallies(counters, vac, repl, bins):
    """Builds the u64/f64 blocks from host-side results (used by the CPU tests)."""
    u = np.zeros(N_COUNTER_SLOTS + 2 * bins, dtype=np.int64)
    names = ["vacancies_created", "replacements", "steps", "ions", "primaries", "recoils_queued", "lost",
             "left_sample", "hist_clamped", "stack_max"]
    for i, k in enumerate(names):
        u[i] = counters[k]
    u[N_COUNTER_SLOTS:N_COUNTER_SLOTS + len(vac)] = vac
    u[N_COUNTER_SLOTS + bins:N_COUNTER_SLOTS + bins + len(repl)] = repl
    f = np.array([counters["EelTotal"], counters["EnucTotal"]], dtype=np.float64)
    return u, f
