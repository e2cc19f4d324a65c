"""Same dotted path as the reference's fish_vocoder/modules/generators/refinegan.py, backed by vocoder_b200."""
from vocoder_b200.generators.refinegan import RefineGANGenerator  # noqa: F401
