"""Multi-GPU plumbing: how primaries shard over ranks and how tallies are joined.

The hot path shards trivially: every primary and its whole cascade is independent,
Philox streams are keyed by the GLOBAL primary index, so rank r of W simply takes a contiguous index
range and no data-path collective exists.  The only exchange is the final tally reduction —
the analogue of runmytrim's threadJoin (runmytrim.C:316-323) — ONE all-gather of the tally block
[u64 counters | u64 histograms | f64 EelTotal, EnucTotal] followed by a local reduction (TallyReducer).
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


class TallyReducer:
    """The tally join as ONE collective: every rank contributes its whole tally block
    [u64 counters | u64 histograms | f64 Eel, Enuc] (~260 KB) to a single all-gather into a staging tensor torch
    allocated, and reduces the W copies locally in rank order — sum for the additive slots, max for the stack
    high-water mark, FP64 sum for the energies.  Compared with four all-reduce calls on the engine's own memory this
    is one NCCL launch on buffers NCCL has seen in the warm-up (the first all-reduce on memory torch did not allocate
    paid ~50 ms of set-up inside the timed job in round 1), and the FP64 totals are bit-identical on every rank and
    independent of NCCL's algorithm choice.

    `reduce()` returns the job totals and, with write_back=True, also stores them into the engine's blocks (so that
    mtb_get_counters / mtb_get_vac_depth on any rank read the totals).  Writing back is NOT idempotent: the blocks
    then hold totals, and reducing them again would count every rank's share W times — reduce once per job, or
    use write_back=False and keep accumulating."""

    def __init__(self, u64_block, f64_block, group=None):
        import torch
        import torch.distributed as dist
        self.u64, self.f64, self.group = u64_block, f64_block, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.nu, self.nf = u64_block.numel(), f64_block.numel()
        self.mine = torch.empty(self.nu + self.nf, dtype=torch.int64, device=u64_block.device)
        self.all = torch.empty(self.world * (self.nu + self.nf), dtype=torch.int64, device=u64_block.device)

    def reduce(self, write_back=True):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return self.u64, self.f64
        self.mine[:self.nu].copy_(self.u64)
        self.mine[self.nu:].copy_(self.f64.view(torch.int64))
        dist.all_gather_into_tensor(self.all, self.mine, group=self.group)
        blocks = self.all.view(self.world, self.nu + self.nf)
        u = blocks[:, :self.nu].sum(dim=0)
        u[IDX_STACK_MAX] = blocks[:, IDX_STACK_MAX].max()
        # slots 10..15 are rank-local bookkeeping (work counter, list lengths, error flag): not joined
        u[IDX_STACK_MAX + 1:N_COUNTER_SLOTS] = self.u64[IDX_STACK_MAX + 1:N_COUNTER_SLOTS]
        f = blocks[:, self.nu:].contiguous().view(torch.float64).sum(dim=0)
        if write_back:
            self.u64.copy_(u)
            self.f64.copy_(f)
        return u, f

    def per_rank(self):
        """The W contributions of the last reduce() as (W, n_u64) / (W, n_f64) tensors (self-checks)."""
        import torch
        blocks = self.all.view(self.world, self.nu + self.nf)
        return blocks[:, :self.nu], blocks[:, self.nu:].contiguous().view(torch.float64)


def reduce_tallies(u64_block, f64_block, group=None):
    """In-place join of the tally blocks over the ranks (one collective, see TallyReducer).  `u64_block` is an int64
    view of the engine's u64 block (sums of non-negative counts are bit-identical in either signedness).  Call it
    once per job: the blocks hold job totals afterwards."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    TallyReducer(u64_block, f64_block, group).reduce(write_back=True)


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
