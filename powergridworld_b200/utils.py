"""Host-side mirrors of gridworld/utils.py:9-53 for user code (policies, plotting).
The simulator itself applies these transforms inside the CUDA kernels
(csrc/component_math.cuh)."""
import numpy as np

from powergridworld_b200 import spaces


def to_scaled(x, low, high):
    x = np.clip(x, low, high)
    return (2 * x - (low + high)) / (high - low)


def to_raw(y, low, high, eps=1e-4):
    y = np.clip(y, -np.ones_like(y), np.ones_like(y))
    return (y * (high - low) + (high + low)) / 2.


def maybe_rescale_box_space(box, rescale=True):
    if rescale:
        return spaces.Box(low=-1., high=1., shape=box.shape, dtype=box.dtype)
    return box
