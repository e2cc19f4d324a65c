"""Same dotted path as the reference's fish_vocoder/modules/generators/unify.py, backed by vocoder_b200."""
from vocoder_b200.generators.unify import UnifyGenerator  # noqa: F401
