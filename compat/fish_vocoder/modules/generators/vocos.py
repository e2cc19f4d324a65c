"""Same dotted path as the reference's fish_vocoder/modules/generators/vocos.py, backed by vocoder_b200."""
from vocoder_b200.generators.vocos import ISTFTHead  # noqa: F401
