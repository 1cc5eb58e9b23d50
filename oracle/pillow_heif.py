"""TEST INFRASTRUCTURE ONLY — stub of `pillow_heif` (absent from this image).

The reference imports it at `src/depth_pro/utils.py:8-12,69`; only HEIC files need it.
"""


def register_heif_opener():
    return None


def open_heif(*args, **kwargs):
    raise RuntimeError("pillow_heif stub: HEIC decoding is not available in this image")
