"""Order-independent 64-bit hash of a batch of per-line results (SURVEY.md §8d "parity check at scale").

    H = sum over lines i of mix(global line index, ext_id[i], spans[i, :])   (mod 2^64)

The same function over numpy arrays (results computed on the CPU) and over torch CUDA tensors (device results of
gorp_extract_*_device); tests/ and bench.py compare the two on the same lines. Because H is a plain sum it can be
accumulated per shard / per rank and added up (all-reduce) — the whole-corpus value does not depend on how the lines
were sharded over GPUs.
"""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1
_A, _B, _C1, _C2 = 0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB


def _weights(stride: int):
    """odd 64-bit weights, one per span entry (splitmix64 of the column index)"""
    w = []
    for k in range(stride):
        z = (k + 1) * _A & _M64
        z = ((z ^ (z >> 30)) * _C1) & _M64
        z = ((z ^ (z >> 27)) * _C2) & _M64
        w.append((z ^ (z >> 31)) | 1)
    return w


def hash_numpy(first_line: int, ext_id: np.ndarray, spans: np.ndarray) -> int:
    n = len(ext_id)
    with np.errstate(over="ignore"):
        idx = np.arange(first_line + 1, first_line + 1 + n, dtype=np.uint64)
        x = idx * np.uint64(_A) + (ext_id.astype(np.int64) + 0x9E37).astype(np.uint64) * np.uint64(_B)
        if spans.size:
            sp = (spans.astype(np.int64) + 2).astype(np.uint64)
            for k, w in enumerate(_weights(spans.shape[1])):
                x += sp[:, k] * np.uint64(w)
        x ^= x >> np.uint64(31)
        x *= np.uint64(_C1)
        x ^= x >> np.uint64(29)
        return int(x.sum(dtype=np.uint64))


def _i64(v: int) -> int:  # two's complement view of an unsigned 64-bit constant
    return v - (1 << 64) if v >= (1 << 63) else v


def hash_torch(first_line: int, ext_id, spans, chunk: int = 1 << 22) -> int:
    """ext_id: int32 CUDA tensor [n]; spans: int32 CUDA tensor [n, stride] (stride may be 0). int64 arithmetic wraps like
    uint64; logical right shifts are emulated with a mask."""
    import torch
    n = ext_id.numel()
    stride = spans.shape[1] if spans.dim() == 2 else 0
    w = torch.tensor([_i64(v) for v in _weights(stride)], dtype=torch.int64, device=ext_id.device) if stride else None
    total = 0
    for lo in range(0, n, chunk):
        hi = min(lo + chunk, n)
        idx = torch.arange(first_line + 1 + lo, first_line + 1 + hi, dtype=torch.int64, device=ext_id.device)
        x = idx * _i64(_A) + (ext_id[lo:hi].to(torch.int64) + 0x9E37) * _i64(_B)
        if stride:
            x = x + ((spans[lo:hi].to(torch.int64) + 2) * w[None, :]).sum(dim=1)
        x = x ^ ((x >> 31) & ((1 << 33) - 1))
        x = x * _i64(_C1)
        x = x ^ ((x >> 29) & ((1 << 35) - 1))
        total = (total + (int(x.sum().item()) & _M64)) & _M64
    return total


class DeviceArray:
    """Wraps a raw device pointer for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 2}


def device_results(dres, dev):
    """(ext_id, spans) of a gorp_device_result as torch tensors that alias the engine's buffers (valid until the next call)."""
    import torch
    n, stride = int(dres.n_lines), int(dres.span_stride)
    if n == 0:
        return torch.empty(0, dtype=torch.int32, device=dev), torch.empty((0, stride), dtype=torch.int32, device=dev)
    ext = torch.as_tensor(DeviceArray(dres.d_ext_id, (n,), "<i4"), device=dev)
    if stride:
        sp = torch.as_tensor(DeviceArray(dres.d_spans, (n, stride), "<i4"), device=dev)
    else:
        sp = torch.empty((n, 0), dtype=torch.int32, device=dev)
    return ext, sp
