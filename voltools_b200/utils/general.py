"""Device-string handling and output-shape helpers (mirror of voltools/utils/general.py).

The reference's launch-dimension heuristics (general.py:9-58) have no counterpart: launch geometry is
chosen inside libvoltools_b200.so.
"""
from typing import List, Tuple

import numpy as np


def get_available_devices() -> List[str]:
    """general.py:61-80.  This package is the GPU path only: there is no 'cpu' device (no CPU fallback).

    'gpu' = the current CUDA device, 'gpu:X' = ordinal X.
    """
    from .. import _native
    n = _native.device_count()
    if n == 0:
        return []
    return ['gpu'] + [f'gpu:{i}' for i in range(n)]


def device_index(device: str) -> int:
    """'gpu' -> -1 (current device), 'gpu:3' -> 3   (general.py:84-88 parses device[4:] the same way)."""
    return int(device[4:]) if device[4:] else -1


def switch_to_device(device: str):
    """general.py:84-88.  Kept for API compatibility; the library calls take the ordinal explicitly and
    restore the caller's device, so nothing relies on this."""
    idx = device_index(device)
    if idx >= 0:
        import torch
        torch.cuda.set_device(idx)


def compute_post_transform_dimensions(shape: Tuple[int, int, int], transform_m: np.ndarray) \
        -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """general.py:91-123: padding before/after and the new shape such that the transformed volume fits.

    The 8 corners (0 or dim on each axis) are pushed through the inverse matrix (input -> output index),
    rounded to integers; negative extents pad before, extents beyond `shape` pad after.
    """
    dims = np.asarray(tuple(shape) + (1,))
    corners = np.array([[(k >> axis) & 1 for k in range(8)] for axis in range(3)] + [[1] * 8]) * dims[:, None]
    corners[3, :] = 1
    try:
        inverse = np.linalg.inv(transform_m)
    except np.linalg.LinAlgError as e:
        print('Something went wrong. Transform matrix should have been affine but still couldnt inverse...')
        raise e
    moved = np.round(inverse @ corners).astype(int)
    pad_before = -np.minimum(moved, 0).min(axis=1)
    over = moved - dims[:, None]
    pad_after = np.maximum(over, 0).max(axis=1)
    new_dims = pad_before + dims + pad_after
    return pad_before[:3], pad_after[:3], new_dims[:3]
