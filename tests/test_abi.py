"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/*.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'log_b200_raster.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(lgr_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_path():
    names = declared_functions()
    for want in ('lgr_compute_radius', 'lgr_forward_project', 'lgr_forward_render', 'lgr_backward', 'lgr_abi_version'):
        assert want in names


def test_library_loads_and_exports_every_declared_symbol(built):
    from log_b200 import _capi
    from log_b200.build import LIB_PATH
    lib = ctypes.CDLL(LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert set(declared_functions()) == set(_capi.EXPORTS)
    lib.lgr_abi_version.restype = ctypes.c_int
    assert lib.lgr_abi_version() == _capi.LGR_ABI_VERSION          # pure host call, no GPU needed
    assert _capi.load().lgr_sort_smem_capacity() > 1024


def test_library_is_sm_100a_only(built):
    from log_b200.build import LIB_PATH
    out = subprocess.run(['cuobjdump', '-lelf', LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


def test_view_struct_layout_matches_header():
    from log_b200._capi import LgrView
    src = open(HEADER).read()
    body = re.search(r'typedef struct lgr_view \{(.*?)\} lgr_view;', src, flags=re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for decl in body.split(';'):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if not decl.startswith('const') else decl.split('*', 1)[1]
        fields += [n.strip(' *') for n in names.split(',')]
    assert fields == [f[0] for f in LgrView._fields_]


def test_ctypes_arity_matches_header(built):
    """Every binding in log_b200/_capi.py declares exactly as many arguments as the C prototype in the header."""
    from log_b200 import _capi
    lib = _capi.load()
    src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    protos = dict(re.findall(r'\b(lgr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S))
    checked = 0
    for name, params in protos.items():
        fn = getattr(lib, name)
        if fn.argtypes is None:
            continue
        n_params = 0 if params.strip() in ('', 'void') else params.count(',') + 1
        assert len(fn.argtypes) == n_params, (name, len(fn.argtypes), n_params)
        checked += 1
    assert checked >= 6


def test_ctypes_argument_kinds_match_header(built):
    """Beyond the count: every bound argument has the C parameter's kind (pointer / int32 / int64 / float / double)."""
    from log_b200 import _capi
    lib = _capi.load()
    src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    protos = dict(re.findall(r'\b(lgr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S))
    scalar = {'int64_t': ctypes.c_int64, 'int32_t': ctypes.c_int32, 'int': ctypes.c_int, 'float': ctypes.c_float,
              'double': ctypes.c_double}
    for name, params in protos.items():
        fn = getattr(lib, name)
        if fn.argtypes is None or params.strip() in ('', 'void'):
            continue
        for k, (decl, bound) in enumerate(zip(params.split(','), fn.argtypes)):
            decl = decl.strip()
            if '*' in decl:
                assert bound is ctypes.c_void_p or issubclass(bound, ctypes._Pointer), (name, k, decl, bound)
            else:
                ctype = decl.replace('const', '').split()[0]
                want = scalar[ctype]
                assert ctypes.sizeof(bound) == ctypes.sizeof(want) and bound(1).value == want(1).value, (name, k, decl, bound)
                assert (bound in (ctypes.c_float, ctypes.c_double)) == (want in (ctypes.c_float, ctypes.c_double)), (name, k, decl)


def test_shard_layout_struct_matches_header():
    from log_b200._capi import LgrShardLayout
    src = open(HEADER).read()
    body = re.search(r'typedef struct lgr_shard_layout \{(.*?)\} lgr_shard_layout;', src, flags=re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields, kinds = [], []
    for decl in body.split(';'):
        decl = decl.strip()
        if not decl:
            continue
        ctype, names = decl.split(None, 1)
        for nm in names.split(','):
            fields.append(nm.strip())
            kinds.append({'int32_t': ctypes.c_int32, 'int64_t': ctypes.c_int64}[ctype])
    assert fields == [f[0] for f in LgrShardLayout._fields_]
    assert kinds == [f[1] for f in LgrShardLayout._fields_]
    assert ctypes.sizeof(LgrShardLayout) == 8 + 8 * 8


def test_integration_md_binding_matches_the_struct():
    """INTEGRATION.md section 3 shows the ctypes binding a maintainer would add: its field list must be the real one."""
    import re
    from log_b200 import _capi
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    block = text[text.index('class LgrView(ctypes.Structure)'):]
    block = block[:block.index(']\n') + 1]
    block = re.sub(r'#[^\n]*', '', block)
    shown = re.findall(r"\('(\w+)',\s*(\w+)\)", block)
    kinds = {'i32': ctypes.c_int32, 'i64': ctypes.c_int64, 'f32': ctypes.c_float, 'vp': ctypes.c_void_p}
    assert [(n, kinds[k]) for n, k in shown] == list(_capi.LgrView._fields_)
