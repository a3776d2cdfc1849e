"""One-process-per-GPU drivers of the engines (torchrun / torch.distributed).

Mirrors the reference's only parallel axis - independent workers with private result arrays
that are merged at the end (ngskit4b/hammings.cpp:2752-2766 thread blocks, :2855-2867 merge;
-m2/-m3 node slices :2660-2689, :1126-1343).  Every driver starts with ONE broadcast of the
bit-plane packed sequence set(s) from rank 0 (NCCL over NVLink on GPUs); what is partitioned
depends on the engine:
  * exhaustive_distributed        POPC all-pairs engine: QUERY K-mers are sharded, no exchange after
                                  the broadcast, per-rank minima are gathered and concatenated on
                                  the host of rank 0;
  * exhaustive_distributed_bands  diagonal-band engine: the PAIR MATRIX is partitioned (every cell
                                  lowers both K-mers of its pair), every rank keeps a complete
                                  minima array and the arrays meet in all_reduce(MIN) after the
                                  query-sharded bootstrap and after every slab (asynchronous, one slab
                                  behind: bands_slabwise), so each rank thresholds against what all
                                  ranks have found;
  * targeted_distributed          seed-and-verify engine: the index BUCKETS are sharded (each rank
                                  builds 1/world of the index and answers all probes for it), one
                                  all_reduce(MIN) at the end.
A failure on rank 0 before the first broadcast (bad parameters, out of memory while packing) is
carried to every rank in the metadata broadcast and raised everywhere - no rank is left waiting in
a collective.

torch is plumbing here (device buffers, streams, process group); the arithmetic is the CUDA
engine behind the C ABI (kit4b_b200.hamm).  `engine` is pluggable so that the host-side
shard/broadcast/gather logic can be exercised with the gloo backend on CPU in the tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import hamm


def shard_bounds(q_begin: int, q_end: int, world: int) -> List[Tuple[int, int]]:
    """Even split of [q_begin, q_end) by position (same formula as run_sharded in k4b_capi.cu)."""
    span = max(0, q_end - q_begin)
    return [(q_begin + span * r // world, q_begin + span * (r + 1) // world) for r in range(world)]


class CudaEngine:
    """Device side of one rank: pack (rank 0), adopt a broadcast image (others), run a shard."""

    def __init__(self, device: torch.device):
        self.device = device

    def empty_image(self, length: int) -> torch.Tensor:
        return torch.empty(hamm.packed_image_bytes(length), dtype=torch.uint8, device=self.device)

    def pack(self, concat: np.ndarray, K: int):
        """rank 0: H2D of the 1-byte/base concat + pack kernels into a torch-owned image."""
        host = torch.from_numpy(np.ascontiguousarray(concat, dtype=np.uint8))
        d_concat = host.to(self.device, non_blocking=False)
        image = self.empty_image(len(concat))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        packed = hamm.Packed.from_device_into(d_concat.data_ptr(), len(concat), K, image.data_ptr(),
                                              image.numel(), stream)
        packed._image = image  # the image tensor lives exactly as long as the handle built on it
        return image, packed, bool(packed.has_non_acgt)

    def adopt(self, image: torch.Tensor, length: int, K: int, has_non_acgt: bool):
        packed = hamm.Packed.from_image(image.data_ptr(), image.numel(), length, K, has_non_acgt)
        packed._image = image
        return packed

    def compute(self, packed, both: bool, b: int, e: int, out: torch.Tensor) -> int:
        """Enqueues the shard [b,e) on the current stream; out: int16[e-b] on the device."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        return hamm.allpairs_min_device(packed, packed, both, True, b, e, out.data_ptr(), 0, stream)

    # ---- diagonal-band engine: the pair matrix is partitioned, minima meet in all_reduce(MIN) ----
    def new_best(self, length: int, K: int) -> torch.Tensor:
        best = torch.empty(length, dtype=torch.int32, device=self.device)
        hamm.best_init_device(best.data_ptr(), length, K, torch.cuda.current_stream(self.device).cuda_stream)
        return best

    def bootstrap(self, packed, both: bool, b: int, e: int, best: torch.Tensor) -> int:
        hamm.diag_bootstrap_device(packed, both, b, e, best.data_ptr(),
                                   torch.cuda.current_stream(self.device).cuda_stream)
        return 1

    def bands(self, packed, both: bool, part: int, nparts: int, best: torch.Tensor) -> int:
        return hamm.diag_bands_device(packed, both, part, nparts, best.data_ptr(),
                                      torch.cuda.current_stream(self.device).cuda_stream)

    def slab_count(self, packed, both: bool, nparts: int) -> int:
        return hamm.diag_slab_count(packed, both, nparts)

    def slabs(self, packed, both: bool, part: int, nparts: int, slab_begin: int, slab_end: int,
              best: torch.Tensor) -> int:
        return hamm.diag_slabs_device(packed, both, part, nparts, slab_begin, slab_end, best.data_ptr(),
                                      torch.cuda.current_stream(self.device).cuda_stream)

    # ---- targeted mode: seed-and-verify engine on a probe range / on a share of the index buckets ----
    def seed(self, probes, targets, both: bool, clamp: int, core_len: int, b: int, e: int, best: torch.Tensor) -> int:
        return hamm.targeted_seed_device(probes, targets, both, clamp, core_len, b, e, best.data_ptr(),
                                         torch.cuda.current_stream(self.device).cuda_stream)

    def seed_part(self, probes, targets, both: bool, clamp: int, core_len: int, part: int, nparts: int,
                  best: torch.Tensor) -> int:
        return hamm.targeted_seed_part_device(probes, targets, both, clamp, core_len, part, nparts, best.data_ptr(),
                                              torch.cuda.current_stream(self.device).cuda_stream)

    def targeted_finalize(self, probes, best: torch.Tensor, clamp: int) -> torch.Tensor:
        out = torch.empty(best.numel(), dtype=torch.int16, device=self.device)
        hamm.targeted_finalize_device(probes, best.data_ptr(), clamp, out.data_ptr(),
                                      torch.cuda.current_stream(self.device).cuda_stream)
        return out

    def finalize(self, packed, best: torch.Tensor) -> torch.Tensor:
        out = torch.empty(best.numel(), dtype=torch.int16, device=self.device)
        hamm.best_finalize_device(packed, best.data_ptr(), out.data_ptr(),
                                  torch.cuda.current_stream(self.device).cuda_stream)
        return out


def _broadcast_packed(engine, concat, K: int, rank: int, world: int, group, extra=(), check=None):
    """rank 0 packs `concat` (after running `check()`, if given); a status word, the length, the
    non-ACGT flag and the integers of `extra` travel in ONE metadata broadcast, the packed image in
    a second one.  A failure on rank 0 is raised on EVERY rank (ranks != 0 get a RuntimeError naming
    it), so nobody is left waiting in a collective.  Returns (packed handle, length, extra values)."""
    dev = engine.device
    meta = torch.zeros(3 + len(extra), dtype=torch.int64, device=dev)
    packed = image = err = None
    if rank == 0:
        try:
            if check is not None:
                check()
            image, packed, non_acgt = engine.pack(concat, K)
            meta = torch.tensor([0, len(concat), int(non_acgt), *[int(v) for v in extra]], dtype=torch.int64, device=dev)
        except Exception as exc:  # carried to the other ranks below, then re-raised here
            err = exc
            meta[0] = 1
    if world > 1:
        dist.broadcast(meta, src=0, group=group)
    vals = [int(v) for v in meta.tolist()]
    if vals[0]:
        if err is not None:
            raise err
        raise RuntimeError("rank 0 failed before the broadcast of the packed sequence set (its exception has the cause)")
    length, non_acgt = vals[1], vals[2]
    if rank != 0:
        image = engine.empty_image(length)
    if world > 1:
        dist.broadcast(image, src=0, group=group)  # the one broadcast of this packed sequence set
    if rank != 0:
        packed = engine.adopt(image, length, K, bool(non_acgt))
    return packed, length, vals[3:]


def exhaustive_distributed_bands(concat: Optional[np.ndarray], K: int, both: bool, engine=None,
                                 group=None) -> Optional[np.ndarray]:
    """Full all-vs-all minima on the diagonal-band engine over the ranks of `group`.

    The symmetric formulation visits every unordered pair once and lowers both K-mers, so the
    PAIR MATRIX (interleaved groups of diagonals) is partitioned instead of the queries; every
    rank keeps a complete minima array and the arrays meet in all_reduce(MIN): after the sharded
    bootstrap and (overlapped with the next slab, see bands_slabwise) after every slab.  rank 0
    passes the concat and gets uint16[len(concat)] (K+1 where no K-mer starts); others get None."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    if engine is None:
        engine = CudaEngine(torch.device("cuda", torch.cuda.current_device()))
    packed, length, _ = _broadcast_packed(engine, concat, K, rank, world, group)
    best = engine.new_best(length, K)
    b, e = shard_bounds(0, length, world)[rank]
    engine.bootstrap(packed, both, b, e, best)
    if world > 1:
        dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
    bands_slabwise(engine, packed, both, rank, world, best, group)
    if rank != 0:
        return None
    return engine.finalize(packed, best).cpu().numpy().view(np.uint16)


def bands_slabwise(engine, packed, both: bool, rank: int, world: int, best: torch.Tensor, group=None,
                   lag: int = 1) -> int:
    """This rank's part of the pair matrix, slab by slab (every rank runs the same number of slabs).

    After every slab the minima of all ranks meet in all_reduce(MIN), so that each rank thresholds
    against what all of them have found.  The exchange works on a SNAPSHOT of `best` and is
    asynchronous: slab j starts as soon as the exchange issued after slab j-1-lag has landed (its
    result is folded in with an element-wise minimum), so with lag = 1 a collective overlaps the
    following slab and no rank idles at a slab boundary waiting for the slowest one; lag = 0 is the
    synchronous schedule.  Thresholds only need to be upper bounds of the final minima, which any
    state of `best` is - the results do not depend on the schedule.  All exchanges are drained
    before returning, the last one holding every rank's final minima.  Returns the kernel launches."""
    if world == 1:
        return engine.bands(packed, both, 0, 1, best)
    launches = 0
    pending = []  # (work, snapshot) of exchanges in flight, oldest first

    def land(n_keep: int):
        while len(pending) > n_keep:
            work, snap = pending.pop(0)
            work.wait()  # NCCL: the current stream waits for the collective; the host does not block
            torch.minimum(best, snap, out=best)

    for slab in range(engine.slab_count(packed, both, world)):
        land(lag)
        launches += engine.slabs(packed, both, rank, world, slab, slab + 1, best)
        snap = best.clone()
        pending.append((dist.all_reduce(snap, op=dist.ReduceOp.MIN, group=group, async_op=True), snap))
    land(0)
    return launches


def targeted_distributed(target: Optional[np.ndarray], probes: Optional[np.ndarray], K: int, R: int, both: bool,
                         engine=None, group=None) -> Optional[np.ndarray]:
    """Targeted mode (-m0 -I) over the ranks of `group` on the seed-and-verify engine: rank 0 passes
    the assembly's sequence area and the probe concat (pure ACGT; probe sets with N / InDel go
    through hamm.targeted, which adds the wildcard pass), both packed sets are broadcast once,
    every rank builds ITS SHARE OF THE INDEX BUCKETS (1/world of the index) and answers every probe
    K-mer for those buckets, the minima meet in one all_reduce(MIN).  rank 0 gets
    uint8[len(probes)] (0xFF where no K-mer starts), others None."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    if engine is None:
        engine = CudaEngine(torch.device("cuda", torch.cuda.current_device()))
    core = K // (R + 1)
    if R < 1 or R > 10 or core < 4:  # same arguments on every rank: raised everywhere
        raise ValueError("R outside 1..10 or K/(R+1) < 4 (hammings.cpp:399-404)")
    clamp = K // core  # the reference's "not found" value (SfxArray.cpp:4462-4463)

    def check_probes():  # rank 0 only (it alone holds the probes): travels in the status word
        if ((probes >= 4) & (probes < 7)).any():
            raise ValueError("probe K-mers with N / InDel need the wildcard pass: use kit4b_b200.targeted")

    q_img, q_len, _ = _broadcast_packed(engine, probes, K, rank, world, group, check=check_probes)
    t_img, _, _ = _broadcast_packed(engine, target, K, rank, world, group)
    best = engine.new_best(q_len, K)
    engine.seed_part(q_img, t_img, both, clamp, core, rank, world, best)
    if world > 1:
        dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
    if rank != 0:
        return None
    out = engine.targeted_finalize(q_img, best, clamp).cpu().numpy().view(np.uint16)
    res = np.full(q_len, 0xFF, dtype=np.uint8)
    ok = out <= K
    res[ok] = out[ok].astype(np.uint8)
    return res


def exhaustive_distributed(concat: Optional[np.ndarray], K: int, both: bool, q_begin: int = 0,
                           q_end: Optional[int] = None, engine=None, group=None) -> Optional[np.ndarray]:
    """All-vs-all minima with queries sharded over the ranks of `group` (POPC engine).

    rank 0 passes the concat (others pass None) and gets uint16[len(concat)] laid out like the
    reference's m_pHamDist (K+1 where nothing was computed); other ranks get None."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    if engine is None:
        engine = CudaEngine(torch.device("cuda", torch.cuda.current_device()))
    dev = engine.device

    # --- metadata, then ONE broadcast of the packed target set ---
    extra = ()
    if rank == 0:
        extra = (q_begin, len(concat) if q_end is None else min(q_end, len(concat)))
    packed, length, (qb, qe) = _broadcast_packed(engine, concat, K, rank, world, group,
                                                 extra=extra if rank == 0 else (0, 0))

    # --- every rank: its own query slice against all targets, no further exchange ---
    bounds = shard_bounds(qb, qe, world)
    b, e = bounds[rank]
    width = max(2, max(hi - lo for lo, hi in bounds))
    width += width & 1  # even, so the int16 minima travel as int32 words (gloo has no int16)
    out = torch.full((width,), K + 1, dtype=torch.int16, device=dev)
    if e > b:
        engine.compute(packed, both, b, e, out)

    # --- gather per-rank minima on rank 0 and concatenate on the host ---
    wire = out.view(torch.int32)
    parts = [torch.empty_like(wire) for _ in range(world)] if rank == 0 else None
    dist.gather(wire, parts, dst=0, group=group)
    if rank != 0:
        return None
    result = np.full(length, K + 1, dtype=np.uint16)
    for (lo, hi), t in zip(bounds, parts):
        if hi > lo:
            result[lo:hi] = t.cpu().numpy().view(np.uint16)[: hi - lo]
    return result
