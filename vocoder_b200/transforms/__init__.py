from .spectrogram import LinearSpectrogram, LogMelSpectrogram, slaney_mel_filterbank  # noqa: F401
