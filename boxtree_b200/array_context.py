"""Torch-tensor stand-in for the reference's ``PyOpenCLArrayContext``.

The hot path only uses a thin slice of the array-context surface
(``boxtree/array_context.py:104-238``): ``from_numpy``,
``to_numpy`` (both working on whole containers), ``freeze``/``thaw`` and a
queue.  Here the "queue" is a CUDA stream and arrays are ``torch.Tensor`` s in
HBM.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import dataclasses
from typing import Any

import numpy as np
import torch


class TorchArrayContext:
    """``actx`` object accepted by :class:`TreeBuilder` / :class:`FMMTraversalBuilder`."""

    array_types = (torch.Tensor,)

    def __init__(self, device: str | torch.device | None = None,
                 stream: torch.cuda.Stream | None = None) -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("boxtree_b200 needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device if device is not None else "cuda:0")
        self._stream = stream

    # -- the reference's `actx.queue`: a CUDA stream
    @property
    def stream(self) -> torch.cuda.Stream:
        return self._stream if self._stream is not None else torch.cuda.current_stream(self.device)

    queue = stream

    @property
    def stream_handle(self) -> int:
        return self.stream.cuda_stream

    # -- allocation helpers (torch owns every buffer)
    def zeros(self, shape, dtype) -> torch.Tensor:
        return torch.zeros(shape, dtype=_torch_dtype(dtype), device=self.device)

    def empty(self, shape, dtype) -> torch.Tensor:
        return torch.empty(shape, dtype=_torch_dtype(dtype), device=self.device)

    # -- containers
    def from_numpy(self, obj: Any) -> Any:
        return _map_container(obj, self._from_numpy_leaf)

    def to_numpy(self, obj: Any) -> Any:
        return _map_container(obj, self._to_numpy_leaf)

    def read_back(self, t: torch.Tensor) -> np.ndarray:
        """Small device tensor -> numpy through PINNED memory and a wait on the stream.  (``.cpu()``
        goes through pageable memory: the copy is staged and the call returns ~0.1 ms later --
        noticeable for the handful of control words the builders read back per step.)"""
        if not t.is_cuda:
            return t.numpy().copy()
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        with torch.cuda.stream(self.stream):
            host.copy_(t, non_blocking=True)
        self.stream.synchronize()
        return host.numpy().copy()

    def freeze(self, obj: Any) -> Any:
        return obj

    def thaw(self, obj: Any) -> Any:
        return obj

    def _from_numpy_leaf(self, a):
        if isinstance(a, np.ndarray) and a.dtype != object:
            return torch.from_numpy(np.ascontiguousarray(a)).to(self.device, non_blocking=False)
        return a

    def _to_numpy_leaf(self, a):
        if isinstance(a, torch.Tensor):
            return a.detach().cpu().numpy()
        return a


_NP2TORCH = {
    np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
    np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
    np.dtype(np.uint8): torch.uint8, np.dtype(np.int8): torch.int8,
}


def _torch_dtype(dt):
    if isinstance(dt, torch.dtype):
        return dt
    return _NP2TORCH[np.dtype(dt)]


def numpy_dtype_of(t: torch.Tensor) -> np.dtype:
    for k, v in _NP2TORCH.items():
        if v == t.dtype:
            return k
    raise TypeError(f"unsupported tensor dtype {t.dtype}")


def _map_container(obj, leaf):
    """Apply *leaf* to every array of a (possibly nested) container."""
    if obj is None or isinstance(obj, (bool, int, float, str, np.generic, np.dtype)):
        return obj
    if isinstance(obj, torch.Tensor):
        return leaf(obj)
    if isinstance(obj, np.ndarray):
        if obj.dtype == object:
            out = np.empty(obj.shape, dtype=object)
            for idx in np.ndindex(obj.shape):
                out[idx] = _map_container(obj[idx], leaf)
            return out
        return leaf(obj)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map_container(o, leaf) for o in obj)
    if dataclasses.is_dataclass(obj) and not isinstance(obj, type):
        changes = {}
        memo: dict[int, Any] = {}
        for f in dataclasses.fields(obj):
            v = getattr(obj, f.name)
            # keep aliasing (e.g. `targets is sources`) intact across conversion
            if id(v) in memo:
                changes[f.name] = memo[id(v)]
            else:
                changes[f.name] = memo[id(v)] = _map_container(v, leaf)
        return dataclasses.replace(obj, **changes)
    return obj


def make_obj_array(items) -> np.ndarray:
    """1-D numpy object array, like ``pytools.obj_array.new_1d``."""
    out = np.empty(len(items), dtype=object)
    for i, it in enumerate(items):
        out[i] = it
    return out
