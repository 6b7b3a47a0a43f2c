"""ctypes binding of the C-ABI CUDA library `librih_b200.so`; signatures are parsed from include/rih_b200.h.

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import re

from . import _build

HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'include', 'rih_b200.h')
_PROTO_RE = re.compile(r'^\s*((?:const\s+)?[\w ]+?\*?)\s*\b(rih_\w+)\s*\(([^)]*)\)\s*;', re.M)


def _ctype(param):
    p = param.strip()
    if p in ('void', ''):
        return None
    if '*' in p:
        return ctypes.c_void_p
    t = ' '.join(p.split()[:-1])  # drop the parameter name
    t = t.replace('const ', '').strip()
    table = {'int': ctypes.c_int, 'float': ctypes.c_float, 'double': ctypes.c_double, 'long long': ctypes.c_longlong,
             'unsigned long long': ctypes.c_ulonglong, 'size_t': ctypes.c_size_t, 'cudaStream_t': ctypes.c_void_p,
             'unsigned int': ctypes.c_uint, 'uint32_t': ctypes.c_uint32}
    if t not in table:
        raise RuntimeError('rih_b200.h: unknown parameter type %r' % p)
    return table[t]


def parse_header(path=HEADER):
    """-> {name: (restype_str, [ctypes types])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    text = re.sub(r'//[^\n]*', '', text)
    protos = {}
    for m in _PROTO_RE.finditer(text):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3)
        args = [_ctype(p) for p in params.split(',')]
        protos[name] = (ret, [a for a in args if a is not None])
    return protos


_lib = None
CALLS = [0]   # number of C-ABI kernel-launching calls issued by this process (bench.py reports it as gpu_launches)


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building first if the sources changed and nvcc is present). Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # no nvcc on this box: accept a prebuilt library if one exists
            if not os.path.exists(path):
                raise RuntimeError('renderih_b200: CUDA library is not built and cannot be built here: %s' % e)
    if not os.path.exists(path):
        raise RuntimeError('renderih_b200: %s missing -- run `python -m renderih_b200._build`' % path)
    lib = ctypes.CDLL(path)
    for name, (ret, args) in parse_header().items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
        fn.restype = ctypes.c_char_p if 'char' in ret else ctypes.c_int
        fn.argtypes = args
    _lib = lib
    return lib


def call(name, *args):
    """Invoke a C-ABI entry point; raise RuntimeError(rih_last_error()) on a non-zero status."""
    lib = _lib if _lib is not None else load()
    CALLS[0] += 1
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError('%s failed (status %d): %s' % (name, rc, lib.rih_last_error().decode()))
