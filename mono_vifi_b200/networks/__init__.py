"""Drop-in for the reference's `networks` package (networks/__init__.py)."""
from .monodepth2 import DepthEncoder, DepthDecoder  # noqa: F401
from .posenet import ResnetEncoder, PoseDecoder, resnet_multiimage_input  # noqa: F401
from . import monodepth2, posenet  # noqa: F401
