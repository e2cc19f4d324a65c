"""Same dotted path as the reference's fish_vocoder/modules/generators/bigvgan.py, backed by vocoder_b200."""
from vocoder_b200.generators.bigvgan import AMPBlock, BigVGANGenerator, Snake, SnakeBeta  # noqa: F401
