"""Tile-sharded multi-GPU rendering (BASELINE config 4): one process per GPU, Gaussians replicated,
each rank renders a contiguous band of 16-pixel tile rows, and the per-Gaussian gradients -- the only exchange step
of the path -- are reduced to their owner rank (Gaussian index blocks) over NCCL.

The reference has no multi-GPU path at all (`cfg.gpus` only sets CUDA_VISIBLE_DEVICES, apps/train.py:136-137).
"""
from typing import List, Tuple

import torch
import torch.distributed as dist

GRAD_FLOATS_PRECOMP = 17   # means3D 3 + means2D 3 + opacity 1 + scales 3 + rotations 4 + colors 3


def tile_row_partition(image_height: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, near-equal bands of tile rows; band r belongs to rank r.  Bands can be empty when there are more
    ranks than tile rows."""
    gy = (image_height + 15) // 16
    base, extra = divmod(gy, world_size)
    out, start = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


def owner_chunk(num_gaussians: int, world_size: int) -> int:
    """Gaussians per owner: ceil(N/R) rounded up to a multiple of 256 (LGR_OWNER_CHUNK in the C header)."""
    return ((num_gaussians + world_size - 1) // world_size + 255) // 256 * 256


def owner_partition(num_gaussians: int, world_size: int) -> List[Tuple[int, int]]:
    """Gaussian index blocks [lo,hi) owning the reduced gradient rows (equal chunk, last ranks may be short)."""
    chunk = owner_chunk(num_gaussians, world_size)
    return [(min(num_gaussians, r * chunk), min(num_gaussians, (r + 1) * chunk)) for r in range(world_size)]


def pack_grads(grads) -> torch.Tensor:
    """(dmeans3D, dmeans2D, dopacities, dscales, drotations, dcolors) -> one (N, 17) row-major buffer."""
    dm3, dm2, dop, dsc, drot, dcol = grads
    return torch.cat([dm3, dm2, dop.reshape(-1, 1), dsc, drot, dcol], dim=1)


def unpack_grads(buf: torch.Tensor):
    return buf[:, 0:3], buf[:, 3:6], buf[:, 6], buf[:, 7:10], buf[:, 10:14], buf[:, 14:17]


def rows_to_shard(rows: torch.Tensor, lo: int, hi: int, shard: torch.Tensor = None) -> torch.Tensor:
    """Add packed gradient rows (M, LGR_ROW_FLOATS) whose id lies in [lo,hi) into the dense owner shard
    (hi-lo, LGR_ROW_FLOATS); columns 0..16 are the 17 gradient floats in pack_grads order."""
    import ctypes
    from . import _capi
    lib = _capi.load()
    if shard is None:
        shard = torch.zeros((max(hi - lo, 0), _capi.LGR_ROW_FLOATS), dtype=torch.float32, device=rows.device)
    if rows.shape[0]:
        rows = rows.contiguous()
        _capi.check(lib.lgr_grad_scatter_add(int(rows.shape[0]), ctypes.c_void_p(rows.data_ptr()), int(lo), int(hi),
                                             ctypes.c_void_p(shard.data_ptr()),
                                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'lgr_grad_scatter_add')
    return shard


def exchange_rows_to_owners(rows: torch.Tensor, send_counts, num_gaussians: int, group=None) -> torch.Tensor:
    """The path's only collective: every rank holds packed gradient rows grouped by owner (send_counts[o] rows for owner
    o); one NCCL all-to-all moves them to their owners, which add them into their dense shard.  Returns this rank's
    shard (chunk, LGR_ROW_FLOATS); [:, :17] are the summed gradients of Gaussians owner_partition(N)[rank]."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = rows.device
    sc = torch.tensor(list(send_counts), dtype=torch.int64, device=dev)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = rc.tolist()
    recv = torch.empty((sum(recv_counts), rows.shape[1]), dtype=rows.dtype, device=dev)
    dist.all_to_all_single(recv, rows.contiguous(), output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
    lo, hi = owner_partition(num_gaussians, world)[rank]
    shard = torch.zeros((owner_chunk(num_gaussians, world), rows.shape[1]), dtype=torch.float32, device=dev)
    return rows_to_shard(recv, lo, hi, shard)


def reduce_to_owners(packed: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the per-rank partial gradients; rank r receives rows owner_partition(N)[r] (zero padded to the chunk).
    NCCL: one reduce_scatter over NVLink.  gloo (CPU tests): all_reduce + slice, same result."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n, f = packed.shape
    chunk = owner_chunk(n, world)
    if n != chunk * world:
        pad = torch.zeros((chunk * world - n, f), dtype=packed.dtype, device=packed.device)
        packed = torch.cat([packed, pad], dim=0)
    if dist.get_backend(group) == 'nccl':
        out = torch.empty((chunk, f), dtype=packed.dtype, device=packed.device)
        dist.reduce_scatter_tensor(out, packed.contiguous(), op=dist.ReduceOp.SUM, group=group)
        return out
    full = packed.clone()
    dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full[rank * chunk:(rank + 1) * chunk].clone()


class PeerExchange:
    """Fused gradient exchange over NVLink peer memory (torch symmetric memory): the per-Gaussian backward kernel stores
    each packed gradient row directly into its owner rank's staging buffer, a device-side barrier follows, and the
    owner adds what it received into its dense shard.  No NCCL call and no host synchronisation on the data path."""

    def __init__(self, num_gaussians: int, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _capi
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n = int(num_gaussians)
        self.chunk = owner_chunk(self.n, self.world)
        dev = torch.device('cuda', torch.cuda.current_device())
        floats = _capi.LGR_STAGE_HEADER_FLOATS + self.world * self.chunk * _capi.LGR_ROW_FLOATS
        self.stage = symm.empty(floats, dtype=torch.float32, device=dev)
        self.stage[:_capi.LGR_STAGE_HEADER_FLOATS].zero_()
        self.handle = symm.rendezvous(self.stage, self.group)
        self.peer_ptrs = torch.tensor(list(self.handle.buffer_ptrs), dtype=torch.int64, device=dev)
        self.lo, self.hi = owner_partition(self.n, self.world)[self.rank]
        self.handle.barrier()

    def backward(self, state, grad_image, means3D, opacities, scales, rotations, colors_precomp) -> torch.Tensor:
        """Blend backward + per-Gaussian backward with rows pushed to the owners; returns this rank's reduced shard
        (owner_chunk, LGR_ROW_FLOATS): [:, :17] gradients in pack_grads order, [:, 18] the radius."""
        import ctypes
        from . import _capi
        from .rasterizer import rasterize_backward
        lib = _capi.load()
        self.handle.barrier()           # every owner has consumed the previous step's rows
        rasterize_backward(state, grad_image, means3D, opacities, scales, rotations, colors_precomp, None,
                           peer_stage=self.peer_ptrs, my_rank=self.rank)
        self.handle.barrier()           # all rows have landed
        shard = torch.zeros((self.chunk, _capi.LGR_ROW_FLOATS), dtype=torch.float32, device=self.stage.device)
        _capi.check(lib.lgr_grad_scatter_add_staged(ctypes.c_void_p(self.stage.data_ptr()), self.world, self.chunk, self.lo,
                                                    self.hi, ctypes.c_void_p(shard.data_ptr()),
                                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                    'lgr_grad_scatter_add_staged')
        return shard
