"""ORACLE (test infrastructure, NOT product code): ctypes binding of oracle/lgr_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.
`build()` compiles the C restatement with gcc (make -C oracle); building the checker is not using it.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'liblgr_oracle.so')
_lib = None

FILTER_ADD, FILTER_MAX, FILTER_NONE = 0, 1, 2


def build(force=False):
    src = os.path.join(HERE, 'lgr_oracle.c')
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', HERE, '-s'])
    return LIB


def _camera_struct(real):
    class Cam(ctypes.Structure):
        _fields_ = [('image_height', ctypes.c_int32), ('image_width', ctypes.c_int32),
                    ('tanfovx', real), ('tanfovy', real),
                    ('viewmatrix', real * 16), ('projmatrix', real * 16), ('campos', real * 3), ('bg', real * 3),
                    ('scale_modifier', real), ('sh_degree', ctypes.c_int32), ('sh_K', ctypes.c_int32),
                    ('filter_mode', ctypes.c_int32)]
    return Cam


_Cam64 = _camera_struct(ctypes.c_double)
_Cam32 = _camera_struct(ctypes.c_float)


def lib():
    global _lib
    if _lib is None:
        build()
        # two OpenMP runtimes live in one process (torch's bundled libgomp and the system one this .so links):
        # spinning waiters of both oversubscribe the cores, so make ours sleep between parallel regions.
        os.environ.setdefault('OMP_WAIT_POLICY', 'passive')
        _lib = ctypes.CDLL(LIB)
        for pfx in ('lgo64_', 'lgo32_'):
            getattr(_lib, pfx + 'render').restype = ctypes.c_int64
            getattr(_lib, pfx + 'num_threads').restype = ctypes.c_int
    return _lib


def num_threads():
    return lib().lgo64_num_threads()


def set_num_threads(n):
    lib().lgo64_set_num_threads(int(n))
    lib().lgo32_set_num_threads(int(n))


def _cam(cam, np_dtype, filter_mode, sh_K):
    """cam: any object with the fields of oracle.torch_dense.Camera (tensors or arrays)."""
    C = _Cam64 if np_dtype == np.float64 else _Cam32
    c = C()
    c.image_height, c.image_width = int(cam.image_height), int(cam.image_width)
    c.tanfovx, c.tanfovy = float(cam.tanfovx), float(cam.tanfovy)
    to = lambda a: np.asarray(a.detach().cpu().numpy() if hasattr(a, 'detach') else a, dtype=np_dtype).reshape(-1)
    c.viewmatrix[:] = to(cam.viewmatrix).tolist()
    c.projmatrix[:] = to(cam.projmatrix).tolist()
    c.campos[:] = to(cam.campos).tolist()
    c.bg[:] = to(cam.bg).tolist()
    c.scale_modifier = float(cam.scale_modifier)
    c.sh_degree = int(cam.sh_degree)
    c.sh_K = int(sh_K)
    c.filter_mode = int(filter_mode)
    return c


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _np(a, dt):
    if a is None:
        return None
    if hasattr(a, 'detach'):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=dt)


def compute_radius(cam, means3D, scales, rotations, dtype=np.float64):
    dt = np.dtype(dtype).type
    m, s, r = _np(means3D, dt), _np(scales, dt), _np(rotations, dt)
    out = np.zeros(m.shape[0], dtype=dt)
    c = _cam(cam, dt, FILTER_MAX, 0)
    fn = lib().lgo64_compute_radius if dt == np.float64 else lib().lgo32_compute_radius
    fn(ctypes.byref(c), ctypes.c_int64(m.shape[0]), _p(m), _p(s), _p(r), _p(out))
    return out


def render(cam, means3D, opacities, scales, rotations, colors_precomp=None, shs=None, filter_mode=FILTER_ADD,
           dL_dimage=None, dtype=np.float64, want_aux=True):
    """Forward (+ backward if dL_dimage given).  Returns dict of numpy arrays."""
    dt = np.dtype(dtype).type
    m, o, s, r = _np(means3D, dt), _np(opacities, dt).reshape(-1), _np(scales, dt), _np(rotations, dt)
    col, sh = _np(colors_precomp, dt), _np(shs, dt)
    N = m.shape[0]
    K = 0 if sh is None else sh.shape[1]
    H, W = int(cam.image_height), int(cam.image_width)
    c = _cam(cam, dt, filter_mode, K)
    out = dict(image=np.zeros((3, H, W), dt), radii=np.zeros(N, np.int32),
               point_id_pixel=np.zeros((H, W), np.int32) if want_aux else None,
               point_weight_pixel=np.zeros((H, W), dt) if want_aux else None,
               point_weight=np.zeros(N, dt) if want_aux else None, final_T=np.zeros((H, W), dt))
    g = dict(dmeans3D=None, dmeans2D=None, dopacities=None, dscales=None, drotations=None, dcolors=None, dshs=None)
    G = _np(dL_dimage, dt)
    if G is not None:
        g = dict(dmeans3D=np.zeros((N, 3), dt), dmeans2D=np.zeros((N, 3), dt), dopacities=np.zeros(N, dt),
                 dscales=np.zeros((N, 3), dt), drotations=np.zeros((N, 4), dt),
                 dcolors=np.zeros((N, 3), dt) if col is not None else None,
                 dshs=np.zeros((N, K, 3), dt) if sh is not None else None)
    fn = lib().lgo64_render if dt == np.float64 else lib().lgo32_render
    D = fn(ctypes.byref(c), ctypes.c_int64(N), _p(m), _p(o), _p(s), _p(r), _p(col), _p(sh),
           _p(out['image']), _p(out['radii']), _p(out['point_id_pixel']), _p(out['point_weight_pixel']),
           _p(out['point_weight']), _p(out['final_T']), _p(G),
           _p(g['dmeans3D']), _p(g['dmeans2D']), _p(g['dopacities']), _p(g['dscales']), _p(g['drotations']),
           _p(g['dcolors']), _p(g['dshs']))
    if D < 0:
        raise MemoryError('lgr_oracle: allocation failed')
    out['n_instances'] = int(D)
    out.update(g)
    return out
