"""Drop-in for the reference's `networks` package (networks/__init__.py:1-6: star-imports of every network file)."""
from .monodepth2 import DepthEncoder, DepthDecoder  # noqa: F401
from .posenet import ResnetEncoder, PoseDecoder, resnet_multiimage_input  # noqa: F401
from .fusion_module import FusionModule  # noqa: F401
from .IFRNet import IFRNet  # noqa: F401
from . import monodepth2, posenet, DHRNet, LiteMono, fusion_module, hrnet_encoder  # noqa: F401
from . import IFRNet as _ifrnet_module  # noqa: F401
