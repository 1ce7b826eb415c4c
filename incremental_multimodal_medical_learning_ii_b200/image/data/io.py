"""Image file loading (reference ``health_multimodal/image/data/io.py``).  Host-side decode is outside the GPU hot
path; JPEG/PNG go through PIL, DICOM / NIfTI need the optional ``pydicom`` / ``SimpleITK`` packages."""
from __future__ import annotations

from pathlib import Path

import numpy as np
from PIL import Image


def remap_to_uint8(array: np.ndarray, percentiles=None) -> np.ndarray:
    """Linearly remap intensities to [0, 255] (reference io.py:16-35)."""
    array = array.astype(float)
    if percentiles is not None:
        lo, hi = np.percentile(array, percentiles)
        array = array.clip(lo, hi)
    array -= array.min()
    array /= max(array.max(), 1e-12)
    array *= 255
    return array.astype(np.uint8)


def load_image(path: Path) -> Image.Image:
    """Load a chest X-ray as an 8-bit grey PIL image (reference io.py:38-71)."""
    path = Path(path)
    suffixes = path.suffixes
    if ".nii" in suffixes:
        import SimpleITK as sitk
        image = Image.fromarray(remap_to_uint8(sitk.GetArrayFromImage(sitk.ReadImage(str(path)))[0]))
    elif path.suffix == ".dcm":
        import pydicom
        image = Image.fromarray(remap_to_uint8(pydicom.dcmread(path).pixel_array))
    elif path.suffix.lower() in (".jpg", ".jpeg", ".png"):
        image = Image.open(path)
    else:
        raise ValueError(f"Image type not supported, filename was: {path}")
    return image.convert("L")
