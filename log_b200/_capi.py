"""ctypes binding of include/log_b200_raster.h (the C ABI).  Fails loudly if the library is missing:
there is no CPU path in this package."""
import ctypes
import os

from .build import LIB_PATH

_i32, _i64, _f32, _vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

LGR_FILTER_ADD, LGR_FILTER_MAX, LGR_FILTER_NONE = 0, 1, 2
LGR_SPLAT_FLOATS = 12
LGR_GRAD_FLOATS = 12
LGR_META_INTS = 8
LGR_TILE_SCRATCH_INTS = 33
LGR_ABI_VERSION = 16
LGR_STAGE_HEADER_FLOATS = 64
LGR_ROW_FLOATS = 20

EXPORTS = ('lgr_abi_version', 'lgr_sort_smem_capacity', 'lgr_compute_radius', 'lgr_forward_project',
           'lgr_forward_render', 'lgr_forward_render_device_sized', 'lgr_backward', 'lgr_grad_scatter_add', 'lgr_grad_scatter_add_staged', 'lgr_point_compact', 'lgr_sparse_adam', 'lgr_profile_enable', 'lgr_profile_collect',
           'lgr_profile_kernel_name', 'lgr_shard_send', 'lgr_shard_recv_bin', 'lgr_blend_backward', 'lgr_shard_return_rows',
           'lgr_shard_gather', 'lgr_shard_recv_bin_aux', 'lgr_shard_return_packed', 'lgr_shard_gather_packed', 'lgr_tree_traverse', 'lgr_mark_visible')
LGR_SHARD_MAX_RANKS = 32
LGR_PROFILE_KERNELS = 12


class LgrView(ctypes.Structure):
    """struct lgr_view (include/log_b200_raster.h)."""
    _fields_ = [('image_height', _i32), ('image_width', _i32), ('tanfovx', _f32), ('tanfovy', _f32),
                ('scale_modifier', _f32), ('sh_degree', _i32), ('sh_coeffs', _i32), ('filter_mode', _i32),
                ('want_aux', _i32), ('tile_row_begin', _i32), ('tile_row_end', _i32),
                ('num_owners', _i32), ('raw_params', _i32), ('band_ids_d', _vp), ('band_blk_d', _vp), ('band_count_d', _vp), ('band_rows_d', _vp), ('band_dsplat_d', _vp), ('tile_rank_d', _vp), ('gather_index_d', _vp), ('pid_map_d', _vp), ('contrib_d', _vp), ('last_contrib_d', _vp),
                ('region_count_d', _vp), ('region_cap', _i64), ('num_regions', _i32), ('reserved0', _i32), ('cov3D_precomp_d', _vp), ('dcov3D_d', _vp),
                ('viewmatrix_d', _vp), ('projmatrix_d', _vp), ('campos_d', _vp), ('bg_d', _vp)]


class LgrShardLayout(ctypes.Structure):
    """struct lgr_shard_layout (multi-GPU shard mode): float offsets into every rank's exchange buffer."""
    _fields_ = [('num_ranks', _i32), ('my_rank', _i32), ('cap', _i64), ('off_count', _i64), ('off_splat', _i64),
                ('off_radii', _i64), ('off_gid', _i64), ('off_dsplat', _i64), ('off_weight', _i64), ('off_pcount', _i64)]


class LgrTree(ctypes.Structure):
    """struct lgr_tree: the level-of-Gaussian tree tables (LoG/model/tensor_tree.py)."""
    _fields_ = [('num_points', _i64), ('num_nodes', _i64), ('max_child', _i32), ('max_level', _i32),
                ('node_index_d', _vp), ('tree_d', _vp)]


def tree_scratch_ints(num_points, slots):
    """LGR_TREE_SCRATCH_INTS."""
    return 8 + 2 * num_points + slots + 2 * ((slots + 255) // 256) + 2 + (slots + 3) // 4


def shard_send_ints(n_local, r):
    """LGR_SHARD_SEND_INTS: int32 scratch of lgr_shard_send (kept until lgr_shard_gather)."""
    return 2 * r * ((max(n_local, 1) + 255) // 256) + r


class LgrError(RuntimeError):
    pass


_lib = None


def bind(lib):
    """Declare restype / argtypes of every entry point of include/log_b200_raster.h on a loaded library handle."""
    lib.lgr_abi_version.restype = ctypes.c_int
    lib.lgr_sort_smem_capacity.restype = _i32
    lib.lgr_mark_visible.restype = ctypes.c_int
    lib.lgr_mark_visible.argtypes = [_i64, _vp, _vp, _vp, _vp]
    lib.lgr_compute_radius.restype = ctypes.c_int
    lib.lgr_compute_radius.argtypes = [_i64, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _vp, _vp]
    lib.lgr_forward_project.restype = ctypes.c_int
    lib.lgr_forward_project.argtypes = [ctypes.POINTER(LgrView), _i64] + [_vp] * 13
    lib.lgr_forward_render.restype = ctypes.c_int
    lib.lgr_forward_render.argtypes = [ctypes.POINTER(LgrView), _i64, _i64, _i32, _i32] + [_vp] * 16
    lib.lgr_forward_render_device_sized.restype = ctypes.c_int
    lib.lgr_forward_render_device_sized.argtypes = [ctypes.POINTER(LgrView), _i64, _i64] + [_vp] * 16
    lib.lgr_sparse_adam.restype = ctypes.c_int
    lib.lgr_sparse_adam.argtypes = [_i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, ctypes.c_double, ctypes.c_double,
                                    ctypes.c_double, ctypes.c_double, _vp]
    lib.lgr_point_compact.restype = ctypes.c_int
    lib.lgr_point_compact.argtypes = [_i64, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.lgr_backward.restype = ctypes.c_int
    lib.lgr_backward.argtypes = [ctypes.POINTER(LgrView), _i64, _i64] + [_vp] * 23 + [_i32, _i64, _vp]
    lib.lgr_grad_scatter_add_staged.restype = ctypes.c_int
    lib.lgr_grad_scatter_add_staged.argtypes = [_vp, _i32, _i64, _i64, _i64, _vp, _vp]
    lib.lgr_grad_scatter_add.restype = ctypes.c_int
    lib.lgr_grad_scatter_add.argtypes = [_i64, _vp, _i64, _i64, _vp, _vp]
    lay = ctypes.POINTER(LgrShardLayout)
    lib.lgr_shard_send.restype = ctypes.c_int
    lib.lgr_shard_send.argtypes = [ctypes.POINTER(LgrView), lay, _i64, _i64, _vp, _vp, _vp, _vp, _vp]
    lib.lgr_shard_recv_bin.restype = ctypes.c_int
    lib.lgr_shard_recv_bin.argtypes = [ctypes.POINTER(LgrView), lay, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.lgr_blend_backward.restype = ctypes.c_int
    lib.lgr_blend_backward.argtypes = [ctypes.POINTER(LgrView), _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.lgr_shard_return_rows.restype = ctypes.c_int
    lib.lgr_shard_return_rows.argtypes = [lay, _vp, _i64, _vp, _i32, _i64, _vp, _vp]
    lib.lgr_shard_gather.restype = ctypes.c_int
    lib.lgr_shard_gather.argtypes = [ctypes.POINTER(LgrView), lay, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.lgr_shard_recv_bin_aux.restype = ctypes.c_int
    lib.lgr_shard_recv_bin_aux.argtypes = [ctypes.POINTER(LgrView), lay, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.lgr_shard_return_packed.restype = ctypes.c_int
    lib.lgr_shard_return_packed.argtypes = [lay, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.lgr_shard_gather_packed.restype = ctypes.c_int
    lib.lgr_shard_gather_packed.argtypes = [ctypes.POINTER(LgrView), lay, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.lgr_tree_traverse.restype = ctypes.c_int
    lib.lgr_tree_traverse.argtypes = [ctypes.POINTER(LgrTree), _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _vp, _i64, _f32,
                                      _i32, _vp, _vp, _vp, _vp]
    lib.lgr_profile_enable.restype = ctypes.c_int
    lib.lgr_profile_enable.argtypes = [ctypes.c_int]
    lib.lgr_profile_collect.restype = ctypes.c_int
    lib.lgr_profile_collect.argtypes = [_vp, _vp, _i32]
    lib.lgr_profile_kernel_name.restype = ctypes.c_char_p
    lib.lgr_profile_kernel_name.argtypes = [ctypes.c_int]
    return lib


def load():
    """Load liblog_b200_raster.so; raise (never fall back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LgrError(f'{LIB_PATH} not found: run `python -m log_b200.build` (or __graft_entry__.build()). '
                       'log_b200 has no CPU fallback.')
    lib = bind(ctypes.CDLL(LIB_PATH))
    if lib.lgr_abi_version() != LGR_ABI_VERSION:
        raise LgrError('liblog_b200_raster.so ABI version mismatch: rebuild with `python -m log_b200.build --force`')
    _lib = lib
    return lib


def current_stream(device=None):
    """The caller's CUDA stream ON `device` (torch's current stream of that device, not of whatever device happens to
    be current) as the void* the C ABI takes."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name):
    """Every tensor handed to the C ABI must live on a CUDA device: there is no CPU path in this package."""
    if not t.is_cuda:
        raise LgrError(f'{name} is on {t.device}: log_b200 rasterises on CUDA only (no CPU fallback)')


def check(rc, what):
    if rc == 0:
        return
    if rc > 0:
        raise LgrError(f'{what}: CUDA error {rc}')
    names = {-1: 'bad argument', -2: 'instance buffer capacity', -3: 'unsupported'}
    raise LgrError(f'{what}: {names.get(rc, rc)}')


def profile_enable(on=True):
    check(load().lgr_profile_enable(1 if on else 0), 'lgr_profile_enable')


def profile_collect():
    """-> {kernel_name: (total_ms, launches)} since the last enable/collect."""
    lib = load()
    ms = (ctypes.c_double * LGR_PROFILE_KERNELS)()
    cnt = (_i32 * LGR_PROFILE_KERNELS)()
    check(lib.lgr_profile_collect(ms, cnt, LGR_PROFILE_KERNELS), 'lgr_profile_collect')
    return {lib.lgr_profile_kernel_name(k).decode(): (ms[k], cnt[k]) for k in range(LGR_PROFILE_KERNELS)}


def owner_chunk(n, r):
    """LGR_OWNER_CHUNK: Gaussians per owner rank, a multiple of 256 so that no projection CTA straddles two owners."""
    return ((n + r - 1) // r + 255) // 256 * 256
