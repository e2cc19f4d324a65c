"""Same dotted path as the reference's fish_vocoder/modules/encoders/convnext.py, backed by vocoder_b200."""
from vocoder_b200.encoders.convnext import ConvNeXtBlock, ConvNeXtEncoder, LayerNorm  # noqa: F401
