"""Drop-in for the reference's ``python/lucas_kanade_core.py`` backed by the B200 kernels.

Same module name, function names, argument meaning and return types as the reference, so
``optical_flow_verifier.py`` and friends import it unchanged when this directory precedes
the reference's ``python/`` on ``PYTHONPATH``.  All arithmetic runs in libof_b200.so on
the GPU; there is no CPU path.

Arithmetic mode: ``OF_B200_MODE=exact`` (default) reproduces the reference bit for bit on
any float32 input; ``OF_B200_MODE=fast`` uses the register-marching throughput kernel,
which is bit-identical on uint8-valued frames.
"""

from typing import Tuple

import numpy as np
import numpy.typing as npt

import of_b200

FloatImage = npt.NDArray[np.float32]


def compute_gradients(frame_prev: FloatImage, frame_curr: FloatImage) -> Tuple[FloatImage, FloatImage, FloatImage]:
    """(Ix, Iy, It): Sobel (true convolution, /8, symmetric border) on the frame average and
    It = prev - curr.  Replaces reference lucas_kanade_core.py:15-45."""
    return of_b200.gradients(frame_prev, frame_curr)


def lucas_kanade_single_scale(
    frame_prev: FloatImage, frame_curr: FloatImage, window_size: int = 5
) -> Tuple[FloatImage, FloatImage]:
    """(u, v) flow fields from one fused GPU pass.  Replaces reference
    lucas_kanade_core.py:48-70; the window_size // 2 border is zero."""
    return of_b200.lk_single_scale(frame_prev, frame_curr, window_size)


def lucas_kanade_from_gradients(
    Ix: FloatImage, Iy: FloatImage, It: FloatImage, window_size: int = 5
) -> Tuple[FloatImage, FloatImage]:
    """(u, v) from caller-supplied gradients: window sums + 2x2 Cramer solve, |det| > 1e-4.
    Replaces reference lucas_kanade_core.py:73-135."""
    return of_b200.lk_from_gradients(Ix, Iy, It, window_size)
