"""Stand-ins for the reference's third-party imports (pyopencl, arraycontext, pytools, cgen, mako)
that execute its kernels serially on the CPU.  None of these packages is installed in this image.

Test infrastructure only (see ``tests/refexec/__init__.py``).  Arrays are numpy arrays behind a thin
``Array`` wrapper with pyopencl's array interface; the kernel classes restate pyopencl's published
semantics (``ListOfListsBuilder``, ``ElementwiseKernel`` / ``ElementwiseTemplate``,
``GenericScanKernel`` / ``ScanTemplate``, ``ReductionKernel`` / ``ReductionTemplate``) as one work
item at a time, compiled with g++ from the kernel text the reference renders itself.
"""
from __future__ import annotations

import ctypes
import functools
import hashlib
import os
import re
import subprocess
import sys
import threading
import types
from dataclasses import dataclass
from typing import Any

import numpy as np

from .minimako import Template as MiniMakoTemplate

TRACE = bool(os.environ.get("REFEXEC_TRACE"))
HERE = os.path.dirname(os.path.abspath(__file__))
CACHE_DIR = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "refexec_cache")


# {{{ arrays

class Buffer:
    """What ``Array.data`` returns: a handle on the same memory."""

    def __init__(self, ary: np.ndarray):
        self.ary = ary

    @property
    def ptr(self) -> int:
        return self.ary.ctypes.data


class Event:
    def wait(self):
        pass


def wait_for_events(events):
    pass


def enqueue_marker(queue, wait_for=None):
    return Event()


def _unwrap(x):
    return x._a if isinstance(x, Array) else x


class Array:
    """pyopencl.array.Array over a numpy array (no queue, no offsets: views are numpy views)."""

    __array_priority__ = 100

    def __init__(self, a, queue=None):
        self._a = np.asarray(a)
        self.queue = queue
        self.events: list[Event] = [Event()]
        self.allocator = None

    # pyopencl-isms
    @property
    def data(self):
        return Buffer(self._a)

    base_data = data
    offset = 0

    def get(self, queue=None, **kw):
        return self._a.copy()

    def get_async(self, queue=None, **kw):
        return self._a.copy(), Event()

    def with_queue(self, queue):
        return self

    def add_event(self, evt):
        pass

    def finish(self):
        pass

    def fill(self, value, queue=None, wait_for=None):
        self._a[...] = value
        return self

    def copy(self, queue=None):
        return Array(self._a.copy(), self.queue)

    def astype(self, dtype, queue=None):
        return Array(self._a.astype(dtype), self.queue)

    def view(self, dtype=None):
        return Array(self._a.view(dtype), self.queue)

    def reshape(self, *shape, **kw):
        return Array(self._a.reshape(*shape, **kw), self.queue)

    def transpose(self, *axes):
        return Array(self._a.transpose(*axes), self.queue)

    def ravel(self, order="C"):
        return Array(self._a.ravel(order), self.queue)

    def setitem(self, subscript, value, queue=None, wait_for=None):
        self[subscript] = value

    def set(self, ary, queue=None, **kw):
        self._a[...] = ary

    def any(self, queue=None, wait_for=None):
        return Array(np.asarray(self._a.any()), self.queue)

    def all(self, queue=None, wait_for=None):
        return Array(np.asarray(self._a.all()), self.queue)

    @property
    def T(self):
        return Array(self._a.T, self.queue)

    @property
    def dtype(self):
        return self._a.dtype

    @property
    def shape(self):
        return self._a.shape

    @property
    def size(self):
        return self._a.size

    @property
    def ndim(self):
        return self._a.ndim

    @property
    def strides(self):
        return self._a.strides

    @property
    def nbytes(self):
        return self._a.nbytes

    @property
    def flags(self):
        return self._a.flags

    def __len__(self):
        return len(self._a)

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            idx = tuple(_unwrap(i) for i in idx)
        else:
            idx = _unwrap(idx)
        if not isinstance(idx, tuple):
            idx = (idx,)
        if not any(i is Ellipsis for i in idx):
            idx = idx + (Ellipsis,)          # an integer subscript gives a 0-d *view*, as in pyopencl
        r = self._a[idx]
        return Array(r, self.queue)

    def __setitem__(self, idx, value):
        if isinstance(idx, tuple):
            idx = tuple(_unwrap(i) for i in idx)
        else:
            idx = _unwrap(idx)
        self._a[idx] = _unwrap(value)

    def __iter__(self):
        for i in range(len(self._a)):
            yield self[i]

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def __int__(self):
        return int(self._a)

    def __index__(self):
        return int(self._a)

    def __float__(self):
        return float(self._a)

    def __bool__(self):
        return bool(self._a)

    def __repr__(self):
        return f"fakecl.Array({self._a!r})"


def _binop(name):
    def op(self, other):
        return Array(getattr(self._a, name)(_unwrap(other)), self.queue)
    return op


for _n in ("add", "sub", "mul", "truediv", "floordiv", "mod", "and", "or", "xor", "lshift",
           "rshift", "eq", "ne", "lt", "le", "gt", "ge", "radd", "rsub", "rmul", "rand", "ror",
           "pow"):
    setattr(Array, f"__{_n}__", _binop(f"__{_n}__"))
for _n in ("neg", "invert", "abs"):
    setattr(Array, f"__{_n}__", (lambda n: lambda self: Array(getattr(self._a, n)(), self.queue))(
        f"__{_n}__"))
Array.__hash__ = object.__hash__


def _inplace(name):
    def op(self, other):
        getattr(self._a, name)(_unwrap(other))
        return self
    return op


for _n in ("iadd", "isub", "imul", "iand", "ior"):
    setattr(Array, f"__{_n}__", _inplace(f"__{_n}__"))


def to_device(queue, ary, allocator=None, **kw):
    return Array(np.array(ary, copy=True), queue)


def empty(queue, shape, dtype, order="C", allocator=None):
    return Array(np.zeros(shape, dtype, order=order), queue)   # zeros: deterministic "garbage"


def zeros(queue, shape, dtype, order="C", allocator=None):
    return Array(np.zeros(shape, dtype, order=order), queue)


def empty_like(ary, queue=None, allocator=None):
    return Array(np.zeros_like(_unwrap(ary)), getattr(ary, "queue", None))


zeros_like = empty_like


def arange(queue, *args, dtype=None, allocator=None, **kw):
    return Array(np.arange(*args, dtype=dtype), queue)


def cumsum(ary, output_dtype=None, queue=None, wait_for=None, return_event=False):
    a = _unwrap(ary)
    r = Array(np.cumsum(a, dtype=output_dtype or a.dtype), getattr(ary, "queue", None))
    return (r, Event()) if return_event else r


def array_sum(ary, dtype=None, queue=None, slice=None):
    a = _unwrap(ary)
    return Array(np.asarray(a.sum(dtype=dtype or a.dtype)), getattr(ary, "queue", None))


def array_max(ary, queue=None):
    return Array(np.asarray(_unwrap(ary).max()), getattr(ary, "queue", None))


def array_min(ary, queue=None):
    return Array(np.asarray(_unwrap(ary).min()), getattr(ary, "queue", None))


def array_take(ary, indices, out=None, queue=None, wait_for=None):
    r = _unwrap(ary).reshape(-1)[_unwrap(indices)]      # pyopencl indexes the flat buffer
    if out is not None:
        out._a[...] = r
        return out
    return Array(r, getattr(ary, "queue", None))


def multi_put(arrays, dest_indices, dest_shape=None, out=None, queue=None, wait_for=None):
    idx = _unwrap(dest_indices)
    if out is None:
        out = [Array(np.zeros(dest_shape, _unwrap(a).dtype), queue) for a in arrays]
    for o, a in zip(out, arrays):
        o._a[idx] = _unwrap(a)
    return out


def multi_take(arrays, indices, out=None, queue=None):
    idx = _unwrap(indices)
    res = [Array(_unwrap(a)[idx], queue) for a in arrays]
    if out is not None:
        for o, r in zip(out, res):
            o._a[...] = r._a
        return out
    return res


def concatenate(arrays, axis=0, queue=None, allocator=None):
    return Array(np.concatenate([_unwrap(a) for a in arrays], axis=axis), queue)


def enqueue_copy(queue, dest, src, **kw):
    d = dest.ary if isinstance(dest, Buffer) else _unwrap(dest)
    s = src.ary if isinstance(src, Buffer) else _unwrap(src)
    nbytes = kw.get("byte_count")
    if nbytes is None:
        d.reshape(-1).view(np.uint8)[:s.nbytes] = s.reshape(-1).view(np.uint8)
    else:
        so, do = kw.get("src_offset", 0), kw.get("dst_offset", kw.get("dest_offset", 0))
        d.reshape(-1).view(np.uint8)[do:do + nbytes] = s.reshape(-1).view(np.uint8)[so:so + nbytes]
    return Event()


class CommandQueue:
    def __init__(self, context=None):
        self.context = context or Context()
        self.device = self.context.devices[0]

    def finish(self):
        pass


class Platform:
    name = "refexec"
    vendor = "refexec"


class Device:
    platform = Platform()
    name = "refexec serial CPU"
    vendor = "refexec"
    version = "serial"
    max_work_group_size = 1
    type = 2


class Context:
    def __init__(self):
        self.devices = [Device()]

# }}}


# {{{ dtype <-> C

_CTYPES = {
    np.dtype(np.int8): "char", np.dtype(np.uint8): "unsigned char",
    np.dtype(np.int16): "short", np.dtype(np.uint16): "unsigned short",
    np.dtype(np.int32): "int", np.dtype(np.uint32): "unsigned int",
    np.dtype(np.int64): "long", np.dtype(np.uint64): "unsigned long",
    np.dtype(np.float32): "float", np.dtype(np.float64): "double",
    np.dtype(np.bool_): "bool",
}
_REGISTERED: dict[Any, str] = {}          # struct dtypes -> C name
_REGISTERED_BY_NAME: dict[str, np.dtype] = {}


class _VecTypes(dict):
    pass


_VEC_DTYPES: set = set()


def _make_vec_types():
    vt = _VecTypes()
    for base, cname in ((np.float32, "float"), (np.float64, "double"), (np.int32, "int")):
        for n in (2, 3, 4):
            names = ["x", "y", "z", "w"][:n] + (["padding0"] if n == 3 else [])
            dt = np.dtype({"names": names, "formats": [base] * len(names)})
            vt[np.dtype(base), n] = dt
            _REGISTERED[dt] = f"{cname}{n}"
            _VEC_DTYPES.add(dt)
    return vt


def dtype_to_ctype(dtype):
    if dtype is None:
        raise ValueError("dtype may not be None")
    dtype = np.dtype(dtype)
    if dtype in _CTYPES:
        return _CTYPES[dtype]
    if dtype in _REGISTERED:
        return _REGISTERED[dtype]
    raise ValueError(f"unable to map dtype {dtype}")


def get_or_register_dtype(c_names, dtype=None):
    if isinstance(c_names, str):
        c_names = [c_names]
    if dtype is None:
        return _REGISTERED_BY_NAME[c_names[0]]
    dtype = np.dtype(dtype)
    if dtype in _REGISTERED:
        return dtype
    _REGISTERED[dtype] = c_names[0]
    for c in c_names:
        _REGISTERED_BY_NAME[c] = dtype
    return dtype


def match_dtype_to_c_struct(device, name, dtype, context=None):
    """numpy's aligned struct layout is the C layout for scalar members (all the reference uses)."""
    dtype = np.dtype(dtype)
    fields = sorted(dtype.fields.items(), key=lambda kv: kv[1][1])
    aligned = np.dtype([(n, f[0]) for n, f in fields], align=True)
    lines = [f"typedef struct {{"]
    for n, f in fields:
        lines.append(f"  {dtype_to_ctype(f[0])} {n};")
    lines.append(f"}} {name};")
    return aligned, "\n".join(lines) + "\n"


def dtype_to_c_struct(device, dtype):
    dtype = np.dtype(dtype)
    if dtype.fields is None or dtype in _VEC_DTYPES:
        return ""
    name = _REGISTERED[dtype]
    fields = sorted(dtype.fields.items(), key=lambda kv: kv[1][1])
    body = "".join(f"  {dtype_to_ctype(f[0])} {n};\n" for n, f in fields)
    return f"typedef struct {{\n{body}}} {name};\n"


class _Arg:
    def __init__(self, dtype, name):
        self.dtype = None if dtype is None else (dtype if isinstance(dtype, str) else np.dtype(dtype))
        self.name = name

    def ctype(self):
        return self.dtype if isinstance(self.dtype, str) else dtype_to_ctype(self.dtype)


class VectorArg(_Arg):
    def __init__(self, dtype, name, with_offset=False):
        super().__init__(dtype, name)
        self.with_offset = with_offset

    def declarator(self):
        return f"{self.ctype()} *{self.name}"


class ScalarArg(_Arg):
    def declarator(self):
        return f"{self.ctype()} {self.name}"


class OtherArg(_Arg):
    def __init__(self, declarator, name):
        super().__init__(None, name)
        self._decl = declarator

    def declarator(self):
        return self._decl


def parse_arg_list(arguments, with_offset=False):
    """``"T *a, U b"`` -> Arg objects whose dtype is the C type *string* (aliases included)."""
    if not isinstance(arguments, str):
        return list(arguments)
    arguments = re.sub(r"/\*.*?\*/", "", arguments, flags=re.S)
    arguments = re.sub(r"//[^\n]*", "", arguments)
    result = []
    for piece in arguments.split(","):
        piece = piece.strip()
        if not piece:
            continue
        piece = re.sub(r"\b(__global|global|const|restrict|__restrict__)\b", " ", piece)
        m = re.match(r"^(.*?)(\*?)\s*([A-Za-z_][A-Za-z_0-9]*)$", piece.strip(), re.S)
        ctype = " ".join(m.group(1).split())
        if ctype.endswith("*"):
            ctype = ctype[:-1].strip()
            result.append(VectorArg(ctype, m.group(3), with_offset))
        elif m.group(2):
            result.append(VectorArg(ctype, m.group(3), with_offset))
        else:
            result.append(ScalarArg(ctype, m.group(3)))
    return result

# }}}


# {{{ compile and call

_CXXFLAGS = ["-O1", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-fno-fast-math",
             "-w", "-fpermissive", f"-I{HERE}"]


def _translate_opencl(src: str) -> str:
    """The two OpenCL C constructs C++ parses differently: vector literals ``(T)(a, b, c)`` and
    ``inline`` free functions (fine as they are).  ``(coord_vec_t)(x, y)`` becomes
    ``coord_vec_t(x, y)``; ``(coord_vec_t) 0.5`` already broadcasts through the constructor."""
    return re.sub(r"\(\s*(\w*coord_vec_t|double[234]|float[234]|int[234])\s*\)\s*\(", r"\1(", src)


def compile_module(source: str):
    os.makedirs(CACHE_DIR, exist_ok=True)
    source = _translate_opencl(source)
    with open(os.path.join(HERE, "cl_shim.h"), "rb") as f:
        shim = f.read()
    key = hashlib.sha256(source.encode() + shim + " ".join(_CXXFLAGS).encode()).hexdigest()[:24]
    so = os.path.join(CACHE_DIR, f"k_{key}.so")
    if not os.path.exists(so):
        cpp = os.path.join(CACHE_DIR, f"k_{key}.{os.getpid()}.{threading.get_ident()}.cpp")
        with open(cpp, "w") as f:
            f.write(source)
        tmp = so + f".tmp{os.getpid()}.{threading.get_ident()}"
        proc = subprocess.run(["g++", *_CXXFLAGS, cpp, "-o", tmp], capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"g++ failed on {cpp}:\n{proc.stderr[:6000]}")
        os.replace(tmp, so)
        if not TRACE:
            os.unlink(cpp)
    return ctypes.CDLL(so)


_SCALAR_INFO_SRC = """
#include <type_traits>
template <class T> static int scalar_code() {
    return (std::is_floating_point<T>::value ? 1 : (std::is_signed<T>::value ? 2 : 3)) * 100
           + (int) sizeof(T);
}
"""


def _np_scalar_dtype(code: int) -> np.dtype:
    kind, size = divmod(code, 100)
    return np.dtype({1: "f", 2: "i", 3: "u"}[kind] + str(size))


class _CompiledKernel:
    """``extern "C" void run(void **args, long *sizes)`` plus ``scalar_codes``: argument marshalling
    shared by all the kernel classes.  *args*: list of Arg; values: Array / Buffer / scalars."""

    def __init__(self, lib, args):
        self.lib = lib
        self.args = args
        n = len(args)
        codes = (ctypes.c_int * max(n, 1))()
        lib.scalar_codes(codes)
        self.scalar_dtypes = [
            None if isinstance(a, VectorArg)
            else (a.dtype if isinstance(a.dtype, np.dtype) else _np_scalar_dtype(codes[i]))
            for i, a in enumerate(args)]

    def pack(self, values):
        if len(values) != len(self.args):
            raise TypeError(f"expected {len(self.args)} kernel arguments "
                            f"({[a.name for a in self.args]}), got {len(values)}")
        keep = []
        ptrs = (ctypes.c_void_p * max(len(values), 1))()
        for i, (a, v) in enumerate(zip(self.args, values)):
            if isinstance(a, VectorArg):
                if v is None:
                    ptrs[i] = None
                    continue
                arr = v.ary if isinstance(v, Buffer) else _unwrap(v)
                if not isinstance(arr, np.ndarray):
                    raise TypeError(f"argument {a.name}: expected an array, got {type(v)}")
                if not (arr.flags.c_contiguous or arr.flags.f_contiguous):
                    raise ValueError(f"argument {a.name}: non-contiguous array")
                keep.append(arr)
                ptrs[i] = arr.ctypes.data
            else:
                v = _unwrap(v)
                if self.scalar_dtypes[i].fields is not None:
                    box = np.zeros(1, self.scalar_dtypes[i])
                    box[0] = v
                else:
                    box = np.array([v]).astype(self.scalar_dtypes[i])
                keep.append(box)
                ptrs[i] = box.ctypes.data
        return ptrs, keep


def _unpack_source(args, type_prefix=""):
    """C++ that turns ``void **a`` into named locals, and the scalar_codes table."""
    unpack, codes = [], []
    for i, a in enumerate(args):
        if isinstance(a, VectorArg):
            unpack.append(f"    {a.ctype()} *{a.name} = ({a.ctype()} *) refexec_args[{i}];")
            codes.append(f"    out[{i}] = 0;")
        else:
            unpack.append(f"    {a.ctype()} {a.name} = *({a.ctype()} *) refexec_args[{i}];")
            codes.append(f"    out[{i}] = scalar_code<{a.ctype()}>();")
    return "\n".join(unpack), "\n".join(codes)

# }}}


# {{{ ListOfListsBuilder (pyopencl.algorithm)

_LIST_BUILDER_LOCK = threading.RLock()


@dataclass
class BuiltList:
    count: int | None = None
    starts: Array | None = None
    lists: Array | None = None
    num_nonempty_lists: int | None = None
    nonempty_indices: Array | None = None
    compressed_indices: Array | None = None


class ListOfListsBuilder:
    """``generate(LIST_ARG_DECL USER_ARG_DECL index_type i)`` is run for i = 0..n_objects-1; what it
    ``APPEND_<name>``s becomes row i of list <name> (pyopencl's count / scan / write passes give
    exactly this)."""

    def __init__(self, context, list_names_and_dtypes, generate_template, arg_decls,
                 count_sharing=None, devices=None, name_prefix="plb_build_list", options=None,
                 preamble="", debug=False, complex_kernel=False,
                 eliminate_empty_output_lists=()):
        self.list_names_and_dtypes = [(n, np.dtype(d)) for n, d in list_names_and_dtypes]
        self.args = parse_arg_list(arg_decls)
        self.eliminate = list(eliminate_empty_output_lists or [])
        self.count_sharing = count_sharing or {}
        lists = self.list_names_and_dtypes
        list_decl = "".join(f"plb_list<{dtype_to_ctype(d)}> *plb_{n}, " for n, d in lists)
        list_args = "".join(f"plb_{n}, " for n, d in lists)
        user_decl = "".join(a.declarator() + ", " for a in self.args)
        user_args = "".join(a.name + ", " for a in self.args)
        appends = "\n".join(f"#define APPEND_{n}(value) plb_{n}->append(value)" for n, d in lists)
        unpack, codes = _unpack_source(self.args)
        nl = len(lists)
        src = f"""#include "cl_shim.h"
{_SCALAR_INFO_SRC}
typedef int index_type;
#define LIST_ARG_DECL {list_decl}
#define LIST_ARGS {list_args}
#define USER_ARG_DECL {user_decl}
#define USER_ARGS {user_args}
{appends}
{preamble}
{generate_template}

{"".join(f"static plb_list<{dtype_to_ctype(d)}> g_{n};" + chr(10) for n, d in lists)}
extern "C" void scalar_codes(int *out) {{
{codes}
}}
extern "C" void run(void **refexec_args, long n_objects, const int *omit, long **starts) {{
{unpack}
{"".join(f"    g_{n}.items.clear(); g_{n}.omitted = omit[{k}] != 0; plb_list<{dtype_to_ctype(d)}> *plb_{n} = &g_{n};" + chr(10) for k, (n, d) in enumerate(lists))}
    for (long i = 0; i < n_objects; ++i) {{
{"".join(f"        starts[{k}][i] = (long) g_{n}.items.size();" + chr(10) for k, (n, d) in enumerate(lists))}
        generate(LIST_ARGS USER_ARGS i);
    }}
{"".join(f"    starts[{k}][n_objects] = (long) g_{n}.items.size();" + chr(10) for k, (n, d) in enumerate(lists))}
}}
extern "C" void fetch(int k, void *dst) {{
{"".join(f"    if (k == {k}) memcpy(dst, g_{n}.items.data(), g_{n}.items.size() * sizeof(g_{n}.items[0]));" + chr(10) for k, (n, d) in enumerate(lists))}
}}
"""
        self.kernel = _CompiledKernel(compile_module(src), self.args)

    def __call__(self, queue, n_objects, *args, allocator=None, omit_lists=(), wait_for=None):
        n_objects = int(_unwrap(n_objects))
        lists = self.list_names_and_dtypes
        ptrs, keep = self.kernel.pack(args)
        omit = (ctypes.c_int * len(lists))(*[int(n in omit_lists) for n, _ in lists])
        starts_np = [np.zeros(n_objects + 1, np.int64) for _ in lists]
        starts_ptrs = (ctypes.c_void_p * len(lists))(*[s.ctypes.data for s in starts_np])
        with _LIST_BUILDER_LOCK:      # the generated module keeps its lists in statics
            return self._run(queue, n_objects, ptrs, omit, starts_np, starts_ptrs, omit_lists)

    def _run(self, queue, n_objects, ptrs, omit, starts_np, starts_ptrs, omit_lists):
        lists = self.list_names_and_dtypes
        self.kernel.lib.run(ptrs, ctypes.c_long(n_objects), omit, starts_ptrs)
        index_dtype = np.int32 if n_objects < np.iinfo(np.int32).max else np.int64
        result = {}
        for k, (name, dtype) in enumerate(lists):
            if name in omit_lists:
                result[name] = BuiltList()
                continue
            starts = starts_np[k]
            count = int(starts[-1])
            data = np.zeros(count, dtype)
            if count:
                self.kernel.lib.fetch(ctypes.c_int(k), ctypes.c_void_p(data.ctypes.data))
            if name in self.count_sharing:
                # a list that shares its counts has no starts of its own
                result[name] = BuiltList(count=count, lists=Array(data, queue))
                continue
            if name in self.eliminate:
                counts = np.diff(starts)
                nonempty = np.nonzero(counts)[0]
                compressed_indices = np.concatenate([[0], np.cumsum(counts != 0)])
                result[name] = BuiltList(
                    count=count,
                    starts=Array(np.concatenate([starts[nonempty], [count]]).astype(index_dtype),
                                 queue),
                    lists=Array(data, queue),
                    num_nonempty_lists=len(nonempty),
                    nonempty_indices=Array(nonempty.astype(index_dtype), queue),
                    compressed_indices=Array(compressed_indices.astype(index_dtype), queue))
            else:
                result[name] = BuiltList(count=count, starts=Array(starts.astype(index_dtype), queue),
                                         lists=Array(data, queue))
        return result, Event()

class KeyValueSorter:
    """pyopencl.algorithm.KeyValueSorter: values grouped by key with a stable sort; returns
    ``(starts, lists, event)`` with ``nkeys + 1`` starts."""

    def __init__(self, context):
        self.context = context

    def __call__(self, queue, keys, values, nkeys, starts_dtype, allocator=None, wait_for=None):
        k, v = _unwrap(keys), _unwrap(values)
        order = np.argsort(k, kind="stable")
        starts = np.zeros(int(nkeys) + 1, starts_dtype)
        np.cumsum(np.bincount(k, minlength=int(nkeys)), out=starts[1:])
        return Array(starts, queue), Array(v[order], queue), Event()

# }}}


# {{{ elementwise (pyopencl.elementwise)

class ElementwiseKernel:
    def __init__(self, context, arguments, operation, name="elwise_kernel", options=None,
                 preamble="", **kwargs):
        self.args = parse_arg_list(arguments)
        self.name = name
        unpack, codes = _unpack_source(self.args)
        src = f"""#include "cl_shim.h"
{_SCALAR_INFO_SRC}
#define PYOPENCL_ELWISE_CONTINUE continue
{preamble}
extern "C" void scalar_codes(int *out) {{
{codes}
}}
extern "C" void run(void **refexec_args, long start, long stop, long step, long n) {{
{unpack}
    for (long i = start; i < stop; i += step) {{
        {operation};
    }}
}}
"""
        self.kernel = _CompiledKernel(compile_module(src), self.args)

    def __call__(self, *args, range=None, slice=None, queue=None, wait_for=None, **kw):
        if kw:
            raise TypeError(f"unexpected keyword arguments {list(kw)}")
        ptrs, keep = self.kernel.pack(args)
        n = None
        for a, v in zip(self.args, args):
            if isinstance(a, VectorArg) and v is not None:
                arr = v.ary if isinstance(v, Buffer) else _unwrap(v)
                n = arr.size
                break
        rng = range if range is not None else slice
        if rng is not None:
            start = int(_unwrap(rng.start)) if rng.start is not None else 0
            stop = int(_unwrap(rng.stop)) if rng.stop is not None else n
            step = int(_unwrap(rng.step)) if rng.step is not None else 1
        else:
            start, stop, step = 0, n, 1
        if TRACE:
            print(f"[refexec] elementwise {self.name} range=({start},{stop},{step}) n={n}",
                  flush=True)
        self.kernel.lib.run(ptrs, ctypes.c_long(int(start)), ctypes.c_long(int(stop)),
                            ctypes.c_long(int(step)), ctypes.c_long(int(n or 0)))
        return Event()


def _render_typedefs(type_aliases, declare_types=()):
    """pyopencl's KernelTemplateBase: struct dtypes (explicitly listed or met in an alias) are
    declared once, then one typedef per alias."""
    out = []
    declared = set()

    def declare(dt):
        dt = np.dtype(dt)
        if dt.fields is None or dt in declared or dt in _VEC_DTYPES:
            return
        for _n, f in dt.fields.items():
            declare(f[0])
        declared.add(dt)
        out.append(dtype_to_c_struct(None, dt))

    for dt in declare_types:
        declare(dt)
    for _name, dt in type_aliases:
        declare(dt)
    for name, dt in type_aliases:
        out.append(f"typedef {dtype_to_ctype(dt)} {name};")
    return "\n".join(out) + "\n"


class _KernelTemplateBase:
    def render(self, text, type_aliases, var_values, context=None):
        if text is None:
            return None
        var_dict = dict(var_values)
        var_dict.setdefault("np", np)
        var_dict.update({name: np.dtype(dt) for name, dt in type_aliases})
        return str(MiniMakoTemplate(text, strict_undefined=True).render(**var_dict))

    def render_args(self, arguments, more_arguments, type_aliases, var_values):
        if isinstance(arguments, str):
            arguments = self.render(arguments, type_aliases, var_values)
        args = parse_arg_list(arguments)
        if isinstance(more_arguments, str):
            more_arguments = self.render(more_arguments, type_aliases, var_values)
        return args + parse_arg_list(more_arguments or [])


class ElementwiseTemplate(_KernelTemplateBase):
    def __init__(self, arguments, operation, name="elwise", preamble="", template_processor=None):
        self.arguments, self.operation, self.name, self.preamble = \
            arguments, operation, name, preamble

    def build(self, context, type_aliases=(), var_values=(), more_preamble="",
              more_arguments=(), declare_types=(), options=None):
        type_aliases, var_values = tuple(type_aliases), tuple(var_values)
        preamble = (_render_typedefs(type_aliases, declare_types)
                    + self.render(self.preamble + more_preamble, type_aliases, var_values))
        return ElementwiseKernel(
            context, self.render_args(self.arguments, more_arguments, type_aliases, var_values),
            self.render(self.operation, type_aliases, var_values), name=self.name,
            preamble=preamble)

# }}}


# {{{ scan (pyopencl.scan)

class GenericScanKernel:
    """Serial restatement: ``item_i = scan_expr(a=item_{i-1}, b=input_expr(i))`` with
    ``across_seg_boundary = is_segment_start_expr(i)``; ``prev_item`` is the neutral element at
    i = 0 and at segment starts; ``last_item`` is ``item_{N-1}``."""

    def __init__(self, ctx, dtype, arguments, input_expr, scan_expr, neutral, output_statement,
                 is_segment_start_expr=None, input_fetch_exprs=None, index_dtype=np.int32,
                 name_prefix="scan", options=None, preamble="", devices=None):
        self.args = parse_arg_list(arguments)
        self.name = name_prefix
        scan_t = dtype_to_ctype(dtype)
        index_t = dtype_to_ctype(index_dtype)
        unpack, codes = _unpack_source(self.args)
        fetch = ""
        for name, arg_name, offset in (input_fetch_exprs or []):
            arg = next(a for a in self.args if a.name == arg_name)
            # pyopencl: the value at i+offset (offset 0 or -1); out-of-range reads are not used
            fetch += (f"        {arg.ctype()} {name} = {arg_name}[(i + ({offset})) < 0 ? 0 : "
                      f"i + ({offset})];\n")
        seg = is_segment_start_expr
        src = f"""#include "cl_shim.h"
{_SCALAR_INFO_SRC}
typedef {index_t} index_type;
{preamble}
typedef {scan_t} scan_type;
extern "C" void scalar_codes(int *out) {{
{codes}
}}
static inline scan_type scan_op(scan_type a, scan_type b, bool across_seg_boundary) {{
    return {scan_expr};
}}
extern "C" void run(void **refexec_args, long N_) {{
{unpack}
    const index_type N = (index_type) N_;
    std::vector<scan_type> items(N_ > 0 ? N_ : 1);
    std::vector<char> seg_start(N_ > 0 ? N_ : 1);
    for (index_type i = 0; i < N; ++i) {{
{fetch}
        scan_type my_input = {input_expr};
        bool is_seg_start = {("(" + seg + ")") if seg else "false"};
        seg_start[i] = is_seg_start;
        if (i == 0) items[i] = my_input;
        else items[i] = scan_op(items[i - 1], my_input, is_seg_start);
    }}
    if (N == 0) return;
    const scan_type last_item = items[N - 1];
    for (index_type i = 0; i < N; ++i) {{
        scan_type item = items[i];
        scan_type prev_item = {neutral};
        if (i > 0 && !seg_start[i]) prev_item = items[i - 1];
        {{ {output_statement}; }}
    }}
}}
"""
        self.kernel = _CompiledKernel(compile_module(src), self.args)

    def __call__(self, *args, allocator=None, size=None, queue=None, wait_for=None, **kw):
        if kw:
            raise TypeError(f"unexpected keyword arguments {list(kw)}")
        ptrs, keep = self.kernel.pack(args)
        if size is None:
            for a, v in zip(self.args, args):
                if isinstance(a, VectorArg) and v is not None:
                    size = (v.ary if isinstance(v, Buffer) else _unwrap(v)).size
                    break
        if TRACE:
            print(f"[refexec] scan {self.name} size={int(_unwrap(size))}", flush=True)
        self.kernel.lib.run(ptrs, ctypes.c_long(int(_unwrap(size))))
        return Event()


class ScanTemplate(_KernelTemplateBase):
    def __init__(self, arguments, input_expr, scan_expr, neutral, output_statement,
                 is_segment_start_expr=None, input_fetch_exprs=None, name_prefix="scan",
                 preamble="", template_processor=None):
        self.arguments = arguments
        self.input_expr = input_expr
        self.scan_expr = scan_expr
        self.neutral = neutral
        self.output_statement = output_statement
        self.is_segment_start_expr = is_segment_start_expr
        self.input_fetch_exprs = input_fetch_exprs or []
        self.name_prefix = name_prefix
        self.preamble = preamble

    def build(self, context, type_aliases=(), var_values=(), more_preamble="",
              more_arguments=(), declare_types=(), options=None, devices=None,
              scan_cls=None):
        type_aliases, var_values = tuple(type_aliases), tuple(var_values)
        r = functools.partial(self.render, type_aliases=type_aliases, var_values=var_values)
        scan_dtype = dict(type_aliases)["scan_t"]
        index_dtype = dict(type_aliases).get("index_t", np.int32)
        preamble = (_render_typedefs(type_aliases, declare_types)
                    + r(self.preamble + more_preamble))
        return GenericScanKernel(
            context, scan_dtype,
            self.render_args(self.arguments, more_arguments, type_aliases, var_values),
            r(self.input_expr), r(self.scan_expr), r(self.neutral), r(self.output_statement),
            is_segment_start_expr=r(self.is_segment_start_expr),
            input_fetch_exprs=self.input_fetch_exprs, index_dtype=index_dtype,
            name_prefix=self.name_prefix, preamble=preamble)

# }}}


# {{{ reduction (pyopencl.reduction)

class ReductionKernel:
    """Serial left fold in index order.  (pyopencl reduces in a tree; the reference only reduces
    with min / max / integer sums, for which the order does not change the result.)"""

    def __init__(self, ctx, dtype_out, neutral, reduce_expr, map_expr=None, arguments=None,
                 name="reduce_kernel", options=None, preamble=""):
        self.args = parse_arg_list(arguments)
        self.dtype_out = np.dtype(dtype_out)
        out_t = dtype_to_ctype(dtype_out)
        unpack, codes = _unpack_source(self.args)
        src = f"""#include "cl_shim.h"
{_SCALAR_INFO_SRC}
{preamble}
typedef {out_t} out_type;
extern "C" void scalar_codes(int *out) {{
{codes}
}}
static inline out_type reduce_op(out_type a, out_type b) {{ return {reduce_expr}; }}
extern "C" void run(void **refexec_args, long n, void *result) {{
{unpack}
    out_type acc = {neutral};
    for (long i = 0; i < n; ++i) {{
        out_type mapped = {map_expr or "in[i]"};
        acc = reduce_op(acc, mapped);
    }}
    *(out_type *) result = acc;
}}
"""
        self.kernel = _CompiledKernel(compile_module(src), self.args)

    def __call__(self, *args, queue=None, wait_for=None, return_event=False, out=None,
                 allocator=None, range=None, slice=None):
        ptrs, keep = self.kernel.pack(args)
        n = None
        for a, v in zip(self.args, args):
            if isinstance(a, VectorArg) and v is not None:
                n = (v.ary if isinstance(v, Buffer) else _unwrap(v)).size
                break
        res = np.zeros((), self.dtype_out)
        self.kernel.lib.run(ptrs, ctypes.c_long(int(n)), ctypes.c_void_p(res.ctypes.data))
        result = Array(res, queue)
        return (result, Event()) if return_event else result


class ReductionTemplate(_KernelTemplateBase):
    def __init__(self, arguments, neutral, reduce_expr, map_expr=None, is_segment_start_expr=None,
                 input_fetch_exprs=None, name_prefix="reduce", preamble="",
                 template_processor=None):
        self.arguments, self.neutral, self.reduce_expr, self.map_expr = \
            arguments, neutral, reduce_expr, map_expr
        self.name_prefix, self.preamble = name_prefix, preamble

    def build(self, context, type_aliases=(), var_values=(), more_preamble="",
              more_arguments=(), declare_types=(), options=None, devices=None):
        type_aliases, var_values = tuple(type_aliases), tuple(var_values)
        r = functools.partial(self.render, type_aliases=type_aliases, var_values=var_values)
        preamble = (_render_typedefs(type_aliases, declare_types)
                    + r(self.preamble + more_preamble))
        return ReductionKernel(
            context, dict(type_aliases)["reduction_t"], r(self.neutral), r(self.reduce_expr),
            r(self.map_expr),
            self.render_args(self.arguments, more_arguments, type_aliases, var_values),
            name=self.name_prefix, preamble=preamble)

# }}}


# {{{ arraycontext

class _ActxNumpy:
    def __init__(self, actx):
        self.actx = actx

    def zeros(self, shape, dtype):
        return Array(np.zeros(shape, dtype), self.actx.queue)

    def empty(self, shape, dtype):
        return Array(np.zeros(shape, dtype), self.actx.queue)

    def zeros_like(self, a):
        return _map_leaves(a, lambda x: Array(np.zeros_like(_unwrap(x)), self.actx.queue))

    def __getattr__(self, name):
        fn = getattr(np, name)

        def wrapped(*args, **kwargs):
            args = [_unwrap(a) if not isinstance(a, (list, tuple)) else [_unwrap(x) for x in a]
                    for a in args]
            r = fn(*args, **{k: _unwrap(v) for k, v in kwargs.items()})
            return Array(r, self.actx.queue) if isinstance(r, np.ndarray) or np.isscalar(r) \
                else r
        return wrapped


def _map_leaves(obj, f):
    import dataclasses
    if isinstance(obj, (Array, np.ndarray)):
        if isinstance(obj, np.ndarray) and obj.dtype == object:
            out = np.empty(obj.shape, object)
            for i in np.ndindex(obj.shape):
                out[i] = _map_leaves(obj[i], f)
            return out
        return f(obj)
    if dataclasses.is_dataclass(obj) and not isinstance(obj, type):
        changes = {fld.name: _map_leaves(getattr(obj, fld.name), f)
                   for fld in dataclasses.fields(obj) if fld.init}
        return dataclasses.replace(obj, **changes)
    return obj


class PyOpenCLArrayContext:
    def __init__(self, queue=None, allocator=None, **kw):
        self.queue = queue or CommandQueue()
        self.context = self.queue.context
        self.allocator = allocator
        self.np = _ActxNumpy(self)

    def from_numpy(self, a):
        if a is None:
            return None
        return _map_leaves(a, lambda x: Array(np.array(x, copy=True), self.queue)
                           if isinstance(x, np.ndarray) else x) \
            if not np.isscalar(a) else a

    def to_numpy(self, a):
        if isinstance(a, (int, float, np.generic)) or a is None:
            return a
        return _map_leaves(a, lambda x: _unwrap(x).copy() if isinstance(x, Array) else x)

    def freeze(self, a):
        return a

    def thaw(self, a):
        return a

    def zeros(self, shape, dtype):
        return Array(np.zeros(shape, dtype), self.queue)

    def empty(self, shape, dtype):
        return Array(np.zeros(shape, dtype), self.queue)

    def call_loopy(self, *a, **k):
        raise NotImplementedError


ArrayContext = PyOpenCLArrayContext

# }}}


# {{{ pytools / cgen

def memoize_method(f):
    cache_name = f"_memoize_dic_{f.__name__}"

    @functools.wraps(f)
    def wrapper(self, *args, **kwargs):
        key = (args, tuple(sorted(kwargs.items())))
        dic = self.__dict__.setdefault(cache_name, {}) if hasattr(self, "__dict__") else {}
        try:
            return dic[key]
        except KeyError:
            dic[key] = r = f(self, *args, **kwargs)
            return r
        except TypeError:
            return f(self, *args, **kwargs)
    return wrapper


def memoize(*args, **kw):
    def deco(f):
        cache: dict = {}

        @functools.wraps(f)
        def wrapper(*a, **k):
            use_kw = kw.get("use_kwargs", False)
            key = (a, tuple(sorted(k.items()))) if (use_kw or k) else a
            keyf = kw.get("key")
            if keyf is not None:
                key = keyf(*a, **k)
            try:
                return cache[key]
            except KeyError:
                cache[key] = r = f(*a, **k)
                return r
            except TypeError:
                return f(*a, **k)
        return wrapper
    if len(args) == 1 and callable(args[0]) and not kw:
        return deco(args[0])
    return deco


def keyed_memoize_method(key, cache_dict_name=None):
    """pytools.keyed_memoize_method: memoize on ``key(*args, **kwargs)``."""
    def deco(f):
        name = cache_dict_name or f"_memoize_dic_{f.__name__}"

        @functools.wraps(f)
        def wrapper(self, *args, **kwargs):
            k = key(*args, **kwargs)
            dic = self.__dict__.setdefault(name, {})
            try:
                return dic[k]
            except KeyError:
                dic[k] = r = f(self, *args, **kwargs)
                return r
        return wrapper
    return deco


class ProcessLogger:
    def __init__(self, *a, **k):
        pass

    def done(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass


DebugProcessLogger = ProcessLogger


def log_process(logger=None, description=None, long_threshold_seconds=None):
    def deco(f):
        return f
    if callable(logger) and not hasattr(logger, "debug"):
        return logger
    return deco


def div_ceil(a, b):
    return -(-a // b)


def single_valued(iterable, equality_pred=None):
    it = iter(iterable)
    first = next(it)
    for other in it:
        if not (other == first):
            raise ValueError("not single-valued")
    return first


def product(iterable):
    r = 1
    for x in iterable:
        r *= x
    return r


def partition(pred, iterable):
    yes, no = [], []
    for x in iterable:
        (yes if pred(x) else no).append(x)
    return yes, no


class Record:
    def __init__(self, valuedict=None, exclude=None, **kwargs):
        self.__dict__["_fields"] = set()
        for k, v in {**(valuedict or {}), **kwargs}.items():
            if exclude and k in exclude:
                continue
            setattr(self, k, v)
            self._fields.add(k)

    def copy(self, **kwargs):
        d = {k: getattr(self, k) for k in self._fields}
        d.update(kwargs)
        return type(self)(**d)

    def get_copy_kwargs(self, **kwargs):
        d = {k: getattr(self, k) for k in self._fields}
        d.update(kwargs)
        return d

    def register_fields(self, new_fields):
        self._fields.update(new_fields)

    def __getattr__(self, name):
        raise AttributeError(name)


def _obj_array_module():
    m = types.ModuleType("pytools.obj_array")

    def new_1d(objs):
        objs = list(objs)
        out = np.empty(len(objs), object)
        for i, o in enumerate(objs):
            out[i] = o
        return out

    def vectorize(f, ary):
        if isinstance(ary, np.ndarray) and ary.dtype == object:
            out = np.empty(ary.shape, object)
            for i in np.ndindex(ary.shape):
                out[i] = f(ary[i])
            return out
        return f(ary)

    class _Sub:
        def __class_getitem__(cls, item):
            return cls

    m.new_1d = new_1d
    m.vectorize = vectorize
    m.ObjectArray1D = _Sub
    m.ObjectArray = _Sub
    m.ObjectArray2D = _Sub
    m.make_obj_array = new_1d
    return m


class CgenEnum:
    """cgen.Enum: upper-case class attributes are the members."""

    @classmethod
    def get_flag_names_and_values(cls):
        return [(n, getattr(cls, n)) for n in sorted(dir(cls)) if n[0].isupper() and n == n.upper()]

    @classmethod
    def get_c_defines_lines(cls):
        return [f"#define {cls.c_value_prefix}{n} {v}" for n, v in cls.get_flag_names_and_values()]

    @classmethod
    def get_c_defines(cls):
        return "\n".join(cls.get_c_defines_lines())

    @classmethod
    def get_c_typedef_line(cls):
        return f"typedef {dtype_to_ctype(cls.dtype)} {cls.c_name};"

    @classmethod
    def get_c_typedef(cls):
        return f"\n\n{cls.get_c_typedef_line()}\n\n"

    @classmethod
    def stringify_value(cls, val):
        return "|".join(n for n, v in cls.get_flag_names_and_values() if val & v)

# }}}


# {{{ pymbolic (only `var`, arithmetic and `evaluate`, as boxtree/cost.py uses them)

class _Expr:
    def __init__(self, fn, text):
        self.fn, self.text = fn, text

    def __repr__(self):
        return self.text

    @staticmethod
    def lift(x):
        return x if isinstance(x, _Expr) else _Expr(lambda ctx: x, repr(x))


def _expr_op(symbol, op):
    def forward(self, other):
        o = _Expr.lift(other)
        return _Expr(lambda ctx: op(self.fn(ctx), o.fn(ctx)), f"({self.text} {symbol} {o.text})")

    def backward(self, other):
        o = _Expr.lift(other)
        return _Expr(lambda ctx: op(o.fn(ctx), self.fn(ctx)), f"({o.text} {symbol} {self.text})")
    return forward, backward


import operator as _operator  # noqa: E402

for _sym, _name, _op in (("+", "add", _operator.add), ("-", "sub", _operator.sub),
                         ("*", "mul", _operator.mul), ("/", "truediv", _operator.truediv),
                         ("**", "pow", _operator.pow)):
    _f, _b = _expr_op(_sym, _op)
    setattr(_Expr, f"__{_name}__", _f)
    setattr(_Expr, f"__r{_name}__", _b)


def pymbolic_var(name):
    return _Expr(lambda ctx: ctx[name], name)


def pymbolic_evaluate(expr, context=None, **kw):
    return _Expr.lift(expr).fn(context or {})

# }}}


# {{{ module table

def build_modules() -> dict[str, types.ModuleType]:
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    vec_types = _make_vec_types()
    cl_array = mod("pyopencl.array", Array=Array, to_device=to_device, empty=empty, zeros=zeros,
                   empty_like=empty_like, zeros_like=zeros_like, arange=arange, cumsum=cumsum,
                   sum=array_sum, max=array_max, min=array_min, take=array_take,
                   multi_put=multi_put, multi_take=multi_take,
                   concatenate=concatenate)
    cltypes = mod("pyopencl.cltypes", vec_types=vec_types)
    tools = mod("pyopencl.tools", dtype_to_ctype=dtype_to_ctype, ScalarArg=ScalarArg,
                VectorArg=VectorArg, OtherArg=OtherArg, dtype_to_c_struct=dtype_to_c_struct,
                get_or_register_dtype=get_or_register_dtype,
                match_dtype_to_c_struct=match_dtype_to_c_struct, parse_arg_list=parse_arg_list)
    algorithm = mod("pyopencl.algorithm", ListOfListsBuilder=ListOfListsBuilder,
                    BuiltList=BuiltList, KeyValueSorter=KeyValueSorter)
    elementwise = mod("pyopencl.elementwise", ElementwiseKernel=ElementwiseKernel,
                      ElementwiseTemplate=ElementwiseTemplate)
    scan = mod("pyopencl.scan", GenericScanKernel=GenericScanKernel, ScanTemplate=ScanTemplate)
    reduction = mod("pyopencl.reduction", ReductionKernel=ReductionKernel,
                    ReductionTemplate=ReductionTemplate)
    typing_mod = mod("pyopencl.typing", WaitList=object, Allocator=object)
    cl = mod("pyopencl", array=cl_array, cltypes=cltypes, tools=tools, algorithm=algorithm,
             elementwise=elementwise, scan=scan, reduction=reduction, typing=typing_mod,
             Context=Context, CommandQueue=CommandQueue, Event=Event, WaitList=object,
             wait_for_events=wait_for_events, enqueue_copy=enqueue_copy,
             enqueue_marker=enqueue_marker, Buffer=Buffer)
    cl.__path__ = []
    obj_array = _obj_array_module()
    pytools = mod("pytools", memoize_method=memoize_method, memoize=memoize,
                  keyed_memoize_method=keyed_memoize_method,
                  ProcessLogger=ProcessLogger, DebugProcessLogger=DebugProcessLogger,
                  log_process=log_process, div_ceil=div_ceil, single_valued=single_valued,
                  partition=partition, product=product, Record=Record, obj_array=obj_array)
    pytools.__path__ = []
    arraycontext = mod("arraycontext", Array=Array, ArrayContext=ArrayContext,
                       ArrayOrContainer=object, ArrayOrContainerT=object,
                       PyOpenCLArrayContext=PyOpenCLArrayContext)
    mako = mod("mako")
    mako.__path__ = []
    mako_template = mod("mako.template", Template=MiniMakoTemplate)
    cgen = mod("cgen", Enum=CgenEnum)
    characterize = mod("pyopencl.characterize", has_struct_arg_count_bug=lambda dev, ctx=None: False)
    cl.characterize = characterize
    pymbolic = mod("pymbolic", var=pymbolic_var, evaluate=pymbolic_evaluate)
    arraycontext.ArrayContextFactory = object
    return {m.__name__: m for m in (cl, cl_array, cltypes, tools, algorithm, elementwise, scan,
                                    reduction, typing_mod, pytools, obj_array, arraycontext,
                                    mako, mako_template, cgen, characterize, pymbolic)}

# }}}
