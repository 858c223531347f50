"""Ground-truth flow loader -- ``get_flow_image_from_path`` of the reference's datamanager
(``freegaussian/datamanager/freegaussian_datamanager.py:211-236``; file names ``interflow_n{k}/*.npy``,
``dataparser/freegaussian_dataparser.py:1164-1166``): ``np.load(path) * scale_factor`` resized to (height, width) with
nearest-neighbour sampling, returned as a ``[height, width, 2]`` tensor.

Host-side file I/O, not a kernel: it is here so that the flow supervision the renderer's ``meta["flow"]`` is compared
with can be fed without OpenCV.  ``cv2.resize(..., INTER_NEAREST)`` samples source pixel ``floor(dst * src / dst_size)``;
the same index arithmetic is done here with numpy.  ``pinned=True`` returns the tensor in pinned memory (as float16 when
``half=True``) for the asynchronous host->device staging ``bench.py`` uses for its per-step targets.
"""

from __future__ import annotations

from pathlib import Path
from typing import Union

import numpy as np
import torch


def _nearest_indices(dst: int, src: int) -> np.ndarray:
    # OpenCV INTER_NEAREST: sx = floor(dx * (src / dst)), clamped
    idx = np.floor(np.arange(dst, dtype=np.float64) * (src / dst)).astype(np.int64)
    return np.minimum(idx, src - 1)


def load_flow_image(filepath: Union[str, Path], height: int, width: int, scale_factor: float = 1.0,
                    pinned: bool = False, half: bool = False) -> torch.Tensor:
    """``[height, width, 2]`` flow image from a ``*.npy`` file (any other suffix raises, as the reference does)."""
    filepath = Path(filepath)
    if filepath.suffix != ".npy":
        raise ValueError(f"Unsupported flow image format: {filepath.suffix}")
    image = np.load(filepath) * scale_factor
    assert image.ndim == 3 and image.shape[2] == 2, image.shape
    if image.shape[:2] != (height, width):
        image = image[_nearest_indices(height, image.shape[0])][:, _nearest_indices(width, image.shape[1])]
    t = torch.from_numpy(np.ascontiguousarray(image))
    if half:
        t = t.to(torch.float16)
    return t.pin_memory() if pinned else t
