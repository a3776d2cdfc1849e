"""CPU test of the N>1 host logic (kit4b_b200/dist.py): world_size-2 gloo run of the
broadcast -> per-rank query shard -> gather/concatenate flow.  The CUDA engine is replaced by a
stand-in that answers each shard with the oracle, so what is verified here is the sharding,
the collective plumbing and the concatenation - the kernels have their own -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, random_genome


class OracleEngine:
    """Stand-in for kit4b_b200.dist.CudaEngine on CPU (test infrastructure)."""

    def __init__(self):
        self.device = torch.device("cpu")
        self.K = None

    @classmethod
    def for_k(cls, K):
        e = cls()
        e.K = K
        return e

    def empty_image(self, length):
        return torch.empty(length, dtype=torch.uint8)

    def pack(self, concat, K):
        self.K = K
        img = torch.from_numpy(np.ascontiguousarray(concat).copy())
        return img, img.numpy(), bool((concat[(concat != 7)] > 3).any())

    def adopt(self, image, length, K, has_non_acgt):
        self.K = K
        return image.numpy()

    def compute(self, packed, both, b, e, out):
        from oracle import hamm_oracle as ho
        hd = ho.exhaustive_brute(packed, self.K, both, b, e)
        out[: e - b] = torch.from_numpy(hd[b:e].view(np.int16).copy())
        return 1

    # band-engine stand-ins: bootstrap = nothing learnt yet, bands = this rank's share of the
    # reference's sweep offsets (a partition of the pair matrix), answered by the oracle
    def new_best(self, length, K):
        return torch.full((length,), K + 1, dtype=torch.int32)

    def bootstrap(self, packed, both, b, e, best):
        return 0

    def bands(self, packed, both, part, nparts, best):
        from oracle import hamm_oracle as ho
        n = len(packed)
        for s in range(1 + part, n + 2, nparts):
            hd = ho.exhaustive_sliding_sweep(packed, self.K, both, s, s)
            torch.minimum(best, torch.from_numpy(hd.astype(np.int32)), out=best)
        return 1

    # slab-wise form: 3 slabs, slab j = every third sweep offset of this rank's share
    def slab_count(self, packed, both, nparts):
        return 3

    def slabs(self, packed, both, part, nparts, slab_begin, slab_end, best):
        from oracle import hamm_oracle as ho
        n = len(packed)
        mine = list(range(1 + part, n + 2, nparts))
        for slab in range(slab_begin, slab_end):
            for s in mine[slab::3]:
                hd = ho.exhaustive_sliding_sweep(packed, self.K, both, s, s)
                torch.minimum(best, torch.from_numpy(hd.astype(np.int32)), out=best)
        return slab_end - slab_begin

    # targeted stand-ins: the oracle answers this rank's slice of the probes
    def seed(self, probes, targets, both, clamp, core_len, b, e, best):
        from oracle import hamm_oracle as ho
        R = self.K // core_len - 1
        assert self.K // (R + 1) == core_len
        h = ho.targeted_brute(targets, probes, self.K, R, both).astype(np.int32)
        h[h == 0xFF] = self.K + 1
        part = torch.from_numpy(h[b:e])
        torch.minimum(best[b:e], part, out=best[b:e])
        return 1

    def seed_part(self, probes, targets, both, clamp, core_len, part, nparts, best):
        # stand-in for a share of the index buckets: this rank answers every probe against ITS share of
        # the target K-mers (interleaved starts); the union over the ranks is the whole target set
        from oracle import hamm_oracle as ho
        R = self.K // core_len - 1
        assert self.K // (R + 1) == core_len
        tvalid = ho.valid_starts(targets, self.K)
        pvalid = ho.valid_starts(probes, self.K)
        tpos = np.flatnonzero(tvalid)[part::nparts]
        h = np.full(len(probes), self.K + 1, dtype=np.int32)
        cpl = ho.CPL
        tk = np.stack([targets[p:p + self.K] for p in tpos]) if len(tpos) else np.zeros((0, self.K), np.uint8)
        for q in np.flatnonzero(pvalid):
            km = probes[q:q + self.K]
            d = (tk != km).sum(axis=1).min(initial=self.K + 1)
            if both:
                d = min(d, (tk != cpl[km[::-1]]).sum(axis=1).min(initial=self.K + 1))
            h[q] = d
        torch.minimum(best, torch.from_numpy(h), out=best)
        return 1

    def targeted_finalize(self, probes, best, clamp):
        out = torch.minimum(best, torch.tensor(clamp, dtype=torch.int32)).to(torch.int16)
        from oracle import hamm_oracle as ho
        out[torch.from_numpy(~ho.valid_starts(probes, self.K))] = self.K + 1
        return out

    def finalize(self, packed, best):
        return best.to(torch.int16)


def _worker(rank, world, port, K, both, qb, qe, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kit4b_b200.dist import exhaustive_distributed
    concat = random_genome(77, [900, 41, 700]) if rank == 0 else None
    res = exhaustive_distributed(concat, K, both, qb, qe, engine=OracleEngine())
    if rank == 0:
        ret["res"] = res
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def _worker_bands(rank, world, port, K, both, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kit4b_b200.dist import exhaustive_distributed_bands
    concat = random_genome(78, [300, 30, 260]) if rank == 0 else None
    res = exhaustive_distributed_bands(concat, K, both, engine=OracleEngine.for_k(K))
    if rank == 0:
        ret["res"] = res
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def _worker_targeted(rank, world, port, K, R, both, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kit4b_b200.dist import targeted_distributed
    target, probes = _targeted_case() if rank == 0 else (None, None)
    res = targeted_distributed(target, probes, K, R, both, engine=OracleEngine.for_k(K))
    if rank == 0:
        ret["res"] = res
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def _targeted_case():
    target = random_genome(79, [2500, 1200])
    probes = np.ascontiguousarray(np.concatenate([target[300:700], [7], random_genome(80, [300])]), dtype=np.uint8)
    probes[40] = (probes[40] + 1) % 4
    return target, probes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("K,both,qb,qe", [(25, True, 0, None), (40, False, 100, 1500)])
def test_two_rank_shards_equal_single_process(oracle, K, both, qb, qe):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), K, both, qb, qe, ret), nprocs=2, join=True)
    concat = random_genome(77, [900, 41, 700])
    want = oracle.exhaustive_brute(concat, K, both, qb, len(concat) if qe is None else qe)
    assert np.array_equal(ret["res"], want)


def test_shard_bounds_partition():
    from kit4b_b200.dist import shard_bounds
    for qb, qe, w in [(0, 10, 3), (5, 5, 2), (0, 1_000_003, 8), (7, 9, 4)]:
        b = shard_bounds(qb, qe, w)
        assert b[0][0] == qb and b[-1][1] == qe
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("K,both", [(20, True), (25, False)])
def test_two_rank_pair_matrix_partition_with_min_allreduce(oracle, K, both):
    """exhaustive_distributed_bands: partitioned pair matrix + all_reduce(MIN) == full result."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_bands, args=(2, _free_port(), K, both, ret), nprocs=2, join=True)
    concat = random_genome(78, [300, 30, 260])
    assert np.array_equal(ret["res"], oracle.exhaustive_brute(concat, K, both))


@pytest.mark.parametrize("K,R,both", [(32, 3, True), (25, 2, False)])
def test_two_rank_targeted_probe_shards(oracle, K, R, both):
    """targeted_distributed: both packed sets broadcast, probes sharded, all_reduce(MIN) == full result."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_targeted, args=(2, _free_port(), K, R, both, ret), nprocs=2, join=True)
    target, probes = _targeted_case()
    assert np.array_equal(ret["res"], oracle.targeted_brute(target, probes, K, R, both))


def _worker_error(rank, world, port, which, ret):
    """Rank 0 fails before the first broadcast; every rank must raise instead of blocking in it."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kit4b_b200 import dist as kd

    class FailingEngine(OracleEngine):
        def pack(self, concat, K):
            raise MemoryError("pack failed on rank 0")

    try:
        if which == "wildcard":
            target, probes = _targeted_case() if rank == 0 else (None, None)
            if rank == 0:
                probes[10] = 4  # N in a probe: rejected by rank 0 only (it alone holds the probes)
            kd.targeted_distributed(target, probes, 32, 3, True, engine=OracleEngine.for_k(32))
        elif which == "pack_bands":
            concat = random_genome(78, [300, 30, 260]) if rank == 0 else None
            kd.exhaustive_distributed_bands(concat, 20, True, engine=FailingEngine.for_k(20))
        else:
            concat = random_genome(77, [900, 41, 700]) if rank == 0 else None
            kd.exhaustive_distributed(concat, 25, True, engine=FailingEngine.for_k(25))
        ret[rank] = "no error"
    except Exception as exc:  # noqa: BLE001 - the test inspects the type
        ret[rank] = type(exc).__name__
    dist.barrier()  # reachable on every rank: nobody is stuck in the broadcast
    dist.destroy_process_group()


@pytest.mark.parametrize("which,rank0_error", [("wildcard", "ValueError"), ("pack_bands", "MemoryError"),
                                                ("pack_shards", "MemoryError")])
def test_rank0_failure_is_raised_on_every_rank(which, rank0_error):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_error, args=(2, _free_port(), which, ret), nprocs=2, join=True)
    assert ret[0] == rank0_error and ret[1] == "RuntimeError"


def _worker_bands_lag(rank, world, port, lag, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kit4b_b200 import dist as kd
    K, both = 20, True
    concat = random_genome(78, [300, 30, 260])
    eng = OracleEngine.for_k(K)
    _, packed, _ = eng.pack(concat, K)
    best = eng.new_best(len(concat), K)
    kd.bands_slabwise(eng, packed, both, rank, world, best, lag=lag)
    if rank == 0:
        ret["res"] = best.numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lag", [0, 1, 2])
def test_slab_exchange_schedule_does_not_change_the_result(oracle, lag):
    """bands_slabwise: synchronous (lag 0) and overlapped (lag >= 1) exchanges give the same minima."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_bands_lag, args=(2, _free_port(), lag, ret), nprocs=2, join=True)
    concat = random_genome(78, [300, 30, 260])
    want = oracle.exhaustive_brute(concat, 20, True).astype(np.int32)
    got = ret["res"]
    valid = want <= 20
    assert np.array_equal(got[valid], want[valid])
