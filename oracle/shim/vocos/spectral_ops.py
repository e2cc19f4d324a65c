"""TEST INFRASTRUCTURE ONLY - stand-in for ``vocos.spectral_ops.ISTFT`` (vocos==0.0.2).

Imported by the reference at fish_vocoder/modules/generators/vocos.py:3 (pinned pdm.lock:1348-1349);
the package is neither vendored nor installed here, so its published algorithm is restated
(SURVEY.md section 8c).  Parity status: UNPINNED against the real wheel; pinned instead by the
known-answer test "ISTFT_same inverts the reference's own STFT framing"
(fish_vocoder/data/transforms/spectrogram.py:29-49) in tests/test_oracle_cpu.py.
"""
import torch
from torch import nn


class ISTFT(nn.Module):
    def __init__(self, n_fft: int, hop_length: int, win_length: int, padding: str = "same"):
        super().__init__()
        if padding not in ("center", "same"):
            raise ValueError("Padding must be 'center' or 'same'.")
        self.padding = padding
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.win_length = win_length
        self.register_buffer("window", torch.hann_window(win_length))

    def forward(self, spec: torch.Tensor) -> torch.Tensor:
        if self.padding == "center":
            return torch.istft(spec, self.n_fft, self.hop_length, self.win_length,
                               self.window, center=True)
        pad = (self.win_length - self.hop_length) // 2
        B, N, T = spec.shape
        frames = torch.fft.irfft(spec, self.n_fft, dim=1, norm="backward")
        frames = frames * self.window[None, :, None]
        out_len = (T - 1) * self.hop_length + self.win_length
        y = torch.nn.functional.fold(
            frames, output_size=(1, out_len), kernel_size=(1, self.win_length),
            stride=(1, self.hop_length))[:, 0, 0, pad:-pad]
        wsq = self.window.square().expand(1, T, -1).transpose(1, 2)
        env = torch.nn.functional.fold(
            wsq, output_size=(1, out_len), kernel_size=(1, self.win_length),
            stride=(1, self.hop_length)).squeeze()[pad:-pad]
        assert (env > 1e-11).all()
        return y / env
