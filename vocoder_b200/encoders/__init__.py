from .convnext import ConvNeXtEncoder  # noqa: F401
from .vocos_backbone import VocosBackbone  # noqa: F401
