"""Same dotted path as the reference's fish_vocoder/data/transforms/spectrogram.py, backed by vocoder_b200."""
from vocoder_b200.transforms.spectrogram import LinearSpectrogram, LogMelSpectrogram  # noqa: F401
