"""One-process-per-GPU sharding of the exhaustive engine (torchrun / torch.distributed).

Mirrors the reference's only parallel axis - independent workers with private result arrays
that are merged at the end (ngskit4b/hammings.cpp:2752-2766 thread blocks, :2855-2867 merge;
-m2/-m3 node slices :2660-2689, :1126-1343) - but shards QUERY K-mers instead of sweep offsets,
so no merge is needed: rank r owns the queries whose flat start position lies in its slice and
compares them with every target.  The only collective on the data path is ONE broadcast of
the bit-plane packed target set from rank 0 (NCCL over NVLink on GPUs); per-rank minima are
gathered to rank 0 and concatenated on the host.

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
        self.keep = []  # tensors that must outlive the handles built on them

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
        self.keep.append(image)
        return image, packed, bool(packed.has_non_acgt)

    def adopt(self, image: torch.Tensor, length: int, K: int, has_non_acgt: bool):
        self.keep.append(image)
        return hamm.Packed.from_image(image.data_ptr(), image.numel(), length, K, has_non_acgt)

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

    # ---- targeted mode: seed-and-verify engine on a probe range ----
    def seed(self, probes, targets, both: bool, clamp: int, core_len: int, b: int, e: int, best: torch.Tensor) -> int:
        return hamm.targeted_seed_device(probes, targets, both, clamp, core_len, b, e, best.data_ptr(),
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


def exhaustive_distributed_bands(concat: Optional[np.ndarray], K: int, both: bool, engine=None,
                                 group=None) -> Optional[np.ndarray]:
    """Full all-vs-all minima on the diagonal-band engine over the ranks of `group`.

    The symmetric formulation visits every unordered pair once and lowers both K-mers, so the
    PAIR MATRIX (interleaved groups of diagonals) is partitioned instead of the queries; every
    rank keeps a complete minima array and the arrays meet in all_reduce(MIN) - once after the
    sharded bootstrap, once after the bands.  rank 0 passes the concat and gets
    uint16[len(concat)] (K+1 where no K-mer starts); other ranks get None."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    if engine is None:
        engine = CudaEngine(torch.device("cuda", torch.cuda.current_device()))
    dev = engine.device
    meta = torch.zeros(2, dtype=torch.int64, device=dev)
    packed = None
    image = None
    if rank == 0:
        image, packed, non_acgt = engine.pack(concat, K)
        meta = torch.tensor([len(concat), int(non_acgt)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(meta, src=0, group=group)
    length, non_acgt = (int(v) for v in meta.tolist())
    if rank != 0:
        image = engine.empty_image(length)
    if world > 1:
        dist.broadcast(image, src=0, group=group)  # the one broadcast of the packed sequence set
    if rank != 0:
        packed = engine.adopt(image, length, K, bool(non_acgt))
    best = engine.new_best(length, K)
    b, e = shard_bounds(0, length, world)[rank]
    engine.bootstrap(packed, both, b, e, best)
    if world > 1:
        dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
    bands_slabwise(engine, packed, both, rank, world, best, group)
    if rank != 0:
        return None
    return engine.finalize(packed, best).cpu().numpy().view(np.uint16)


def bands_slabwise(engine, packed, both: bool, rank: int, world: int, best: torch.Tensor, group=None) -> int:
    """This rank's part of the pair matrix, slab by slab, with all_reduce(MIN) after every slab
    (every rank runs the same number of slabs).  Returns the number of kernel launches."""
    if world == 1:
        return engine.bands(packed, both, 0, 1, best)
    launches = 0
    for slab in range(engine.slab_count(packed, both, world)):
        launches += engine.slabs(packed, both, rank, world, slab, slab + 1, best)
        dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
    return launches


def targeted_distributed(target: Optional[np.ndarray], probes: Optional[np.ndarray], K: int, R: int, both: bool,
                         engine=None, group=None) -> Optional[np.ndarray]:
    """Targeted mode (-m0 -I) over the ranks of `group` on the seed-and-verify engine: rank 0 passes
    the assembly's sequence area and the probe concat (pure ACGT; probe sets with N / InDel go
    through hamm.targeted, which adds the wildcard pass), both packed sets are broadcast once,
    every rank indexes the assembly and answers its slice of the probes, the minima meet in one
    all_reduce(MIN).  rank 0 gets uint8[len(probes)] (0xFF where no K-mer starts), others None."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    if engine is None:
        engine = CudaEngine(torch.device("cuda", torch.cuda.current_device()))
    dev = engine.device
    core = K // (R + 1)
    if R < 1 or R > 10 or core < 4:
        raise ValueError("R outside 1..10 or K/(R+1) < 4 (hammings.cpp:399-404)")
    clamp = K // core  # the reference's "not found" value (SfxArray.cpp:4462-4463)
    if rank == 0 and ((probes >= 4) & (probes < 7)).any():
        raise ValueError("probe K-mers with N / InDel need the wildcard pass: use kit4b_b200.targeted")
    handles = []
    for concat in (target, probes):
        meta = torch.zeros(2, dtype=torch.int64, device=dev)
        packed = image = None
        if rank == 0:
            image, packed, non_acgt = engine.pack(concat, K)
            meta = torch.tensor([len(concat), int(non_acgt)], dtype=torch.int64, device=dev)
        if world > 1:
            dist.broadcast(meta, src=0, group=group)
        length, non_acgt = (int(v) for v in meta.tolist())
        if rank != 0:
            image = engine.empty_image(length)
        if world > 1:
            dist.broadcast(image, src=0, group=group)
        if rank != 0:
            packed = engine.adopt(image, length, K, bool(non_acgt))
        handles.append((packed, length))
    (t_img, _), (q_img, q_len) = handles
    best = engine.new_best(q_len, K)
    b, e = shard_bounds(0, q_len, world)[rank]
    engine.seed(q_img, t_img, both, clamp, core, b, e, best)
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
    """All-vs-all minima with queries sharded over the ranks of `group`.

    rank 0 passes the concat (others pass None) and gets uint16[len(concat)] laid out like the
    reference's m_pHamDist (K+1 where nothing was computed); other ranks get None."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    if engine is None:
        engine = CudaEngine(torch.device("cuda", torch.cuda.current_device()))
    dev = engine.device

    # --- metadata, then ONE broadcast of the packed target set ---
    meta = torch.zeros(4, dtype=torch.int64, device=dev)
    packed = None
    image = None
    if rank == 0:
        image, packed, non_acgt = engine.pack(concat, K)
        qe = len(concat) if q_end is None else min(q_end, len(concat))
        meta = torch.tensor([len(concat), int(non_acgt), q_begin, qe], dtype=torch.int64, device=dev)
    dist.broadcast(meta, src=0, group=group)
    length, non_acgt, qb, qe = (int(v) for v in meta.tolist())
    if rank != 0:
        image = engine.empty_image(length)
    dist.broadcast(image, src=0, group=group)
    if rank != 0:
        packed = engine.adopt(image, length, K, bool(non_acgt))

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
