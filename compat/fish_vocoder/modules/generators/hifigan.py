"""Same dotted path as the reference's fish_vocoder/modules/generators/hifigan.py, backed by vocoder_b200."""
from vocoder_b200.generators.hifigan import HiFiGANGenerator, ParralelBlock, ResBlock1  # noqa: F401
from vocoder_b200.generators._mrf import same_padding as get_padding  # noqa: F401
