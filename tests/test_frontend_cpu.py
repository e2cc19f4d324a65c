"""Mel front-end: oracle pinned on reference-generated goldens, filterbank restatement, state-dict
contract and the C-ABI exports.  CPU only."""
import numpy as np
import pytest
import torch

from tests.util import load_golden

FIXTURES = ["frontend_24k", "frontend_44k"]


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_reference_golden(name):
    from oracle import frontend as O
    kw, sd, ins, out, extra = load_golden(name)
    y = ins["audio"]
    lin = O.linear_spectrogram(y, sd["spectrogram.window"], kw["n_fft"], kw["hop_length"], kw["win_length"])
    mel = O.log_mel_spectrogram(y, sd["spectrogram.window"], sd["mel_scale.fb"], kw["n_fft"], kw["hop_length"],
                                kw["win_length"])
    assert lin.shape == tuple(extra["out_linear"].shape) and mel.shape == out.shape
    assert float((lin - torch.from_numpy(extra["out_linear"])).abs().max()) <= 1e-5
    assert float((mel - out).abs().max()) <= 1e-5
    # frames: T = L // hop for the reference's padding (spectrogram.py:29-37)
    assert mel.shape[-1] == y.shape[-1] // kw["hop_length"]


@pytest.mark.parametrize("name", FIXTURES)
def test_slaney_filterbank_restatement_matches_torchaudio_buffer(name):
    from vocoder_b200.transforms import slaney_mel_filterbank
    kw, sd, *_ = load_golden(name)
    fb = slaney_mel_filterbank(kw["n_fft"] // 2 + 1, 0.0, kw["sample_rate"] // 2, kw["n_mels"], kw["sample_rate"])
    assert fb.shape == sd["mel_scale.fb"].shape
    assert float((fb - sd["mel_scale.fb"]).abs().max()) <= 2e-7


@pytest.mark.parametrize("name", FIXTURES)
def test_state_dict_contract_and_strict_load(name):
    from vocoder_b200.transforms import LogMelSpectrogram
    kw, sd, *_ = load_golden(name)
    m = LogMelSpectrogram(**kw)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    m.load_state_dict(sd, strict=True)
    assert float((m.spectrogram.window - torch.hann_window(kw["win_length"])).abs().max()) == 0.0
    x = torch.rand(3, 5) + 1e-7
    assert torch.equal(m.compress(x), torch.log(torch.clamp(x, min=1e-5)))


def test_front_end_refuses_cpu_tensors():
    from vocoder_b200 import cabi
    from vocoder_b200.transforms import LogMelSpectrogram
    m = LogMelSpectrogram(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=256, n_mels=100)
    with pytest.raises(cabi.FvError):
        m(torch.zeros(1, 4096))


def test_framed_dft_as_row_conv_identity():
    """The identity the kernel path rests on: with the padded signal as rows of `hop` samples, torch.stft(center=False)
    equals a conv over rows with n_fft/hop taps of window * {cos, -sin} (numpy, fp64)."""
    rng = np.random.RandomState(0)
    n_fft, hop, T = 64, 16, 9
    y = rng.standard_normal(hop * (T + n_fft // hop - 1))
    win = np.hanning(n_fft + 1)[:-1]
    F = n_fft // 2 + 1
    n = np.arange(n_fft)
    ang = 2 * np.pi * np.outer(np.arange(F), n) / n_fft
    basis = np.concatenate([np.cos(ang) * win, -np.sin(ang) * win], 0)            # [2F, n_fft]
    rows = y.reshape(-1, hop)
    got = np.zeros((T, 2 * F))
    for j in range(n_fft // hop):
        got += rows[j:j + T] @ basis[:, j * hop:(j + 1) * hop].T
    ref = torch.stft(torch.from_numpy(y), n_fft, hop_length=hop, window=torch.from_numpy(win), center=False,
                     return_complex=True).numpy()                                  # [F, T]
    assert np.abs(got[:, :F] - ref.real.T).max() < 1e-10 and np.abs(got[:, F:] - ref.imag.T).max() < 1e-10
