"""CPU tests of the inference helpers: checkpoint ingest, yaml instantiation, receptive-field
model, wav writer.  The forward itself needs CUDA and is covered by tests/test_parity_gpu.py."""
import os
import wave

import numpy as np
import pytest
import torch

from tests.util import load_golden
from vocoder_b200 import inference as inf
from vocoder_b200.generators import BigVGANGenerator, HiFiGANGenerator, UnifyGenerator


def test_lightning_checkpoint_ingest(tmp_path):
    kwargs, sd, _, _, _ = load_golden("hifigan_small_ref")
    ckpt = {"state_dict": {**{"generator." + k: v for k, v in sd.items()},
                           "discriminators.mpd.discriminators.0.convs.0.bias": torch.zeros(3),
                           "mel_transforms.input.spectrogram.window": torch.zeros(8)},
            "epoch": 3}
    path = tmp_path / "step_000001.ckpt"
    torch.save(ckpt, path)
    cfg = {"_target_": "fish_vocoder.modules.generators.hifigan.HiFiGANGenerator", **kwargs}
    gen = inf.load_generator(cfg, str(path), device="cpu")
    assert isinstance(gen, HiFiGANGenerator)
    for k, v in gen.state_dict().items():
        assert torch.equal(v, sd[k])
    assert inf.generator_state_dict(sd).keys() == sd.keys()  # bare state dict passes through


def test_instantiate_nested_yaml_like_vocos():
    cfg = {"_target_": "fish_vocoder.modules.generators.unify.UnifyGenerator",
           "backbone": {"_target_": "fish_vocoder.modules.encoders.convnext.ConvNeXtEncoder", "input_channels": 20,
                        "depths": [1, 2], "dims": [32, 48], "drop_path_rate": 0.2, "kernel_size": 7},
           "head": {"_target_": "fish_vocoder.modules.generators.vocos.ISTFTHead", "dim": 48, "n_fft": 64,
                    "hop_length": 16, "win_length": 64, "padding": "same"}}
    gen = inf.instantiate(cfg)
    assert isinstance(gen, UnifyGenerator)
    _, sd, _, _, _ = load_golden("vocos_small_ref")
    gen.load_state_dict(sd, strict=True)
    with pytest.raises(KeyError):
        inf.instantiate({"_target_": "fish_vocoder.modules.discriminators.mpd.MPD"})


def test_context_frames_covers_measured_receptive_field():
    # SURVEY 8c KAT 9: cfg A reaches ~3.23 k samples = 12.6 frames either side; the model must not be smaller
    g = HiFiGANGenerator(hop_length=256, upsample_rates=(8, 8, 2, 2), upsample_kernel_sizes=(16, 16, 4, 4), num_mels=80,
                         use_template=False)
    assert 13 <= inf.context_frames(g) <= 40
    b = BigVGANGenerator(hop_length=512, num_mels=100, use_template=False)
    assert 18 <= inf.context_frames(b) <= 60
    assert inf.hop_of(g) == 256


def test_write_wav_and_load_mel(tmp_path):
    wav = torch.stack([torch.linspace(-1.2, 1.2, 100), torch.zeros(100)])
    p = tmp_path / "a.wav"
    inf.write_wav(str(p), wav, 24000)
    with wave.open(str(p)) as f:
        assert f.getnchannels() == 2 and f.getframerate() == 24000 and f.getnframes() == 100
        pcm = np.frombuffer(f.readframes(100), dtype=np.int16).reshape(100, 2)
    assert pcm[0, 0] == -32767 and pcm[-1, 0] == 32767 and (pcm[:, 1] == 0).all()
    mel = torch.randn(37, 20)  # [T, n_mels] -> transposed like test.py:81-82
    torch.save(mel, tmp_path / "m.pt")
    got = inf.load_mel(str(tmp_path / "m.pt"), 20)
    assert got.shape == (1, 20, 37)
    np.save(tmp_path / "m.npy", torch.randn(2, 20, 9).numpy())
    assert inf.load_mel(str(tmp_path / "m.npy"), 20).shape == (2, 20, 9)
    assert list(inf.iter_inputs(str(tmp_path))) == [str(tmp_path / "m.npy"), str(tmp_path / "m.pt")]
