"""Host-side 4x4 float32 matrix builders.

Mirror of voltools/utils/matrices.py:22-154 (same names, arguments, defaults, errors and - to float32
rounding - the same values).  A matrix maps OUTPUT index (a0, a1, a2, 1) to INPUT index, numpy axis order,
the convention of scipy.ndimage.affine_transform.  Euler angles follow Gohlke's `transformations.py`
naming ('sxyz' ... 'rzyz': static/rotating frame + three axes), composed here from elementary rotations
rather than from the closed-form table the reference uses.
"""
from itertools import product
from typing import Tuple, Union

import numpy as np

AVAILABLE_UNITS = ['rad', 'deg']


def _valid_orders():
    orders = []
    for frame in 'sr':
        for a, b, c in product('xyz', repeat=3):
            if a != b and b != c:  # consecutive axes differ (proper Euler a==c, Tait-Bryan all distinct)
                orders.append(frame + a + b + c)
    return orders


AVAILABLE_ROTATIONS = _valid_orders()  # the same 24 conventions as the reference (matrices.py:8-18)


def translation_matrix(translation: Union[Tuple[float, float, float], np.ndarray],
                       dtype: np.dtype = np.float32) -> np.ndarray:
    """matrices.py:22-27: moving the content by +t means sampling at -t."""
    m = np.identity(4, dtype=dtype)
    m[:3, 3] = np.negative(np.asarray(translation[:3], dtype=dtype))
    return m


def _elementary(axis: str, angle: float) -> np.ndarray:
    c, s = np.cos(angle), np.sin(angle)
    if axis == 'x':
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)
    if axis == 'y':
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def rotation_matrix(rotation: Union[Tuple[float, float, float], np.ndarray],
                    rotation_units: str = 'deg', rotation_order: str = 'rzxz',
                    dtype: np.dtype = np.float32) -> np.ndarray:
    """matrices.py:30-90."""
    if rotation_units not in AVAILABLE_UNITS:
        raise ValueError(f'Rotation units must be one of {AVAILABLE_UNITS}')
    if rotation_order not in AVAILABLE_ROTATIONS:
        raise ValueError(f'Rotation order must be one of {AVAILABLE_ROTATIONS}')

    angles = np.asarray(rotation, dtype=np.float64)[:3]
    if rotation_units == 'deg':
        angles = np.deg2rad(angles)
    angles = -angles  # the reference's "CCW notation" (matrices.py:47)

    axes = rotation_order[1:]
    if rotation_order[0] == 'r':  # rotating frame == static frame with the sequence reversed
        axes, angles = axes[::-1], angles[::-1]
    r = np.identity(3)
    for axis, angle in zip(axes, angles):  # static frame: later rotations multiply from the left
        r = _elementary(axis, angle) @ r

    m = np.identity(4, dtype=dtype)
    m[:3, :3] = r
    return m


def shear_matrix(coefficients: Union[Tuple[float, float, float], np.ndarray],
                 dtype: np.dtype = np.float32) -> np.ndarray:
    """matrices.py:93-99: upper-triangular shear, coefficients -> m[0,1], m[0,2], m[1,2]."""
    m = np.identity(4, dtype)
    m[0, 1], m[0, 2], m[1, 2] = coefficients[0], coefficients[1], coefficients[2]
    return m


def scale_matrix(coefficients: Union[Tuple[float, float, float], np.ndarray],
                 dtype: np.dtype = np.float32) -> np.ndarray:
    """matrices.py:102-108."""
    m = np.identity(4, dtype)
    m[0, 0], m[1, 1], m[2, 2] = coefficients[0], coefficients[1], coefficients[2]
    return m


def transform_matrix(scale: Union[Tuple[float, float, float], np.ndarray] = None,
                     shear: Union[Tuple[float, float, float], np.ndarray] = None,
                     rotation: Union[Tuple[float, float, float], np.ndarray] = None,
                     rotation_units: str = 'deg', rotation_order: str = 'rzxz',
                     translation: Union[Tuple[float, float, float], np.ndarray] = None,
                     center: Union[Tuple[float, float, float], np.ndarray] = None,
                     dtype: np.dtype = np.float32) -> np.ndarray:
    """matrices.py:111-154: M = T(translation) . T(-center) . R . Sh . Sc . T(center), products in `dtype`."""
    factors = []
    if translation is not None:
        factors.append(translation_matrix(translation, dtype))
    if center is not None:
        factors.append(translation_matrix(tuple(-1 * c for c in center), dtype))
    if rotation is not None:
        factors.append(rotation_matrix(rotation, rotation_units, rotation_order, dtype))
    if shear is not None:
        factors.append(shear_matrix(shear, dtype))
    if scale is not None:
        factors.append(scale_matrix(scale, dtype))
    if center is not None:
        factors.append(translation_matrix(center, dtype))

    m = np.identity(4, dtype=dtype)
    for f in factors:
        m = np.dot(m, f)
    m /= m[3, 3]
    return m
