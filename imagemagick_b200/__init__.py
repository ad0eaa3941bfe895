"""magick-b200: ImageMagick's per-pixel hot path (separable / 2-D convolution, erode /
dilate, filtered resize, sRGB<->Lab) as hand-written sm_100a CUDA kernels behind the
reference's own MagickCore operator names.  See DESIGN.md / INTEGRATION.md."""
from .api import *  # noqa: F401,F403
from .api import Image, KernelInfo  # noqa: F401
from ._lib import MagickB200Error, LIB_PATH  # noqa: F401
