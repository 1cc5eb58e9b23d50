"""Image IO helpers with the reference's contract (src/depth_pro/utils.py:42-112).

RESTATED FROM THE REFERENCE: `load_rgb` and `fpx_from_f35` are contract-bound host IO that SURVEY.md marks out of scope
for the GPU path ("kept as-is; only its output contract matters") -- the EXIF key chain, the orientation table and the
35 mm formula ARE the contract (the returned `f_px` feeds `DepthPro.infer`), so this file follows the reference's control
flow closely on purpose.  Nothing here touches the GPU.

Host-side only: PIL decode, EXIF orientation, EXIF 35 mm focal length -> f_px.  HEIC needs
`pillow_heif`, which is optional here.
"""

from __future__ import annotations

import logging
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np

LOGGER = logging.getLogger(__name__)


def fpx_from_f35(width: float, height: float, f_mm: float = 50) -> float:
    """35 mm-equivalent focal length (mm) -> pixels (utils.py:42-44)."""
    return f_mm * np.sqrt(width**2.0 + height**2.0) / np.sqrt(36**2 + 24**2)


def extract_exif(img_pil) -> Dict[str, Any]:
    from PIL import ExifTags, TiffTags

    exif = img_pil.getexif()
    out = {ExifTags.TAGS[k]: v for k, v in exif.get_ifd(0x8769).items() if k in ExifTags.TAGS}
    out.update({TiffTags.TAGS_V2[k].name: v for k, v in exif.items() if k in TiffTags.TAGS_V2})
    return out


def load_rgb(path: Union[Path, str], auto_rotate: bool = True, remove_alpha: bool = True
             ) -> Tuple[np.ndarray, Optional[List[bytes]], Optional[float]]:
    """Return (uint8 HWC image, icc profile, f_px or None) exactly like the reference's load_rgb."""
    from PIL import Image

    path = Path(path)
    if path.suffix.lower() == ".heic":
        try:
            import pillow_heif
        except ImportError as err:  # pragma: no cover
            raise RuntimeError("HEIC input needs the optional pillow_heif package") from err
        img_pil = pillow_heif.open_heif(path, convert_hdr_to_8bit=True).to_pillow()
    else:
        img_pil = Image.open(path)

    exif = extract_exif(img_pil)
    icc_profile = img_pil.info.get("icc_profile", None)
    if auto_rotate:
        orientation = exif.get("Orientation", 1)
        if orientation == 3:
            img_pil = img_pil.transpose(Image.ROTATE_180)
        elif orientation == 6:
            img_pil = img_pil.transpose(Image.ROTATE_270)
        elif orientation == 8:
            img_pil = img_pil.transpose(Image.ROTATE_90)
        elif orientation != 1:
            LOGGER.warning(f"Ignoring image orientation {orientation}.")

    img = np.array(img_pil)
    if img.ndim < 3 or img.shape[2] == 1:
        img = np.dstack((img, img, img))
    if remove_alpha:
        img = img[:, :, :3]

    f_35mm = exif.get("FocalLengthIn35mmFilm",
                      exif.get("FocalLenIn35mmFilm", exif.get("FocalLengthIn35mmFormat", None)))
    f_px = fpx_from_f35(img.shape[1], img.shape[0], f_35mm) if f_35mm is not None and f_35mm > 0 else None
    return img, icc_profile, f_px
