"""Round-2 additions to tests/golden (run only in the build container: needs /root/reference):

    python oracle/make_golden_extra.py

* vocos_center_stress       reference ConvNeXtEncoder + ISTFTHead(padding="center") (generators/vocos.py:33-38,69 ->
                            vocos.spectral_ops.ISTFT "center" = torch.istft(center=True); with the head's 2*n_fft outputs
                            the spectrum is two-sided and ATen cuts it to the first n_fft/2+1 rows before the c2r transform)
* bigvgan_snake_mix_stress  reference BigVGANGenerator(activation=Snake) (activation_post = plain Snake, bigvgan.py:335-337)
                            whose first-stage AMPBlocks are rebuilt as AMPBlock(activation=Snake, snake_logscale=False) and
                            whose second-stage blocks as AMPBlock(activation=SnakeBeta, snake_logscale=False)
                            (bigvgan.py:139-146,60-71,122-135): plain Snake and the non-logscale parameterisation.
* bigvgan_template_stress   reference BigVGANGenerator(use_template=True): the noise_convs / template path of BigVGAN
                            (bigvgan.py:300-317,357-359), three kernel sizes, stages of 32 and 16 channels.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import (OUT, BigVGANGenerator, ConvNeXtEncoder, ISTFTHead, UnifyGenerator, mel_input, save,  # noqa: E402
                         stress_init)
from fish_vocoder.modules.generators.bigvgan import AMPBlock, Snake, SnakeBeta  # noqa: E402


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    vk = dict(backbone=dict(input_channels=20, depths=[1, 2], dims=[32, 48], drop_path_rate=0.2, kernel_size=7),
              head=dict(dim=48, n_fft=64, hop_length=16, win_length=64, padding="center"))
    torch.manual_seed(0)
    m = UnifyGenerator(backbone=ConvNeXtEncoder(**vk["backbone"]), head=ISTFTHead(**vk["head"])).eval()
    stress_init(m)
    sd = m.state_dict()
    g = torch.Generator().manual_seed(3)
    for k in sd:
        if k.endswith("weight") and sd[k].ndim >= 2:
            sd[k] = sd[k] * 3.0
        if k.endswith("norm.weight") or (k.endswith(".weight") and sd[k].ndim == 1):
            sd[k] = 1.0 + 0.2 * torch.randn(sd[k].shape, generator=g)
    m.load_state_dict(sd)
    x = mel_input(2, 20, 10)
    y = m.head(m.backbone(x))[:, None, :]
    save("vocos_center_stress", vk, m, {"mel": x}, y)

    bk = dict(hop_length=16, upsample_rates=[4, 2, 2], upsample_kernel_sizes=[8, 4, 4],
              resblock_kernel_sizes=[3, 7], resblock_dilation_sizes=[[1, 3, 5]] * 2,
              num_mels=20, upsample_initial_channel=64, use_template=False,
              pre_conv_kernel_size=7, post_conv_kernel_size=7)
    torch.manual_seed(0)
    m = BigVGANGenerator(activation=Snake, **bk).eval()
    nk = len(bk["resblock_kernel_sizes"])
    for j, (k, d) in enumerate(zip(bk["resblock_kernel_sizes"], bk["resblock_dilation_sizes"])):
        m.resblocks[j] = AMPBlock(32, k, tuple(d), activation=Snake, snake_logscale=False)
        m.resblocks[nk + j] = AMPBlock(16, k, tuple(d), activation=SnakeBeta, snake_logscale=False)
    m = m.eval()
    stress_init(m)
    sd = m.state_dict()
    g = torch.Generator().manual_seed(11)
    for k in list(sd):  # non-logscale alpha / beta must stay positive: 1 + small noise
        if (k.endswith(".act.alpha") or k.endswith(".act.beta")) and (k.startswith("resblocks.0.") or k.startswith(
                "resblocks.1.") or k.startswith("resblocks.2.") or k.startswith("resblocks.3.")):
            sd[k] = 1.0 + 0.3 * torch.rand(sd[k].shape, generator=g)
    m.load_state_dict(sd)
    x = mel_input(2, 20, 12)
    save("bigvgan_snake_mix_stress", bk, m, {"mel": x}, m(x))

    # BigVGAN with the template path (use_template=True is the constructor DEFAULT, bigvgan.py:262,300-317,357-359:
    # x = ups[i](x) + noise_convs[i](template) before every AMP stage); the last stage has 16 channels (row-pair convs)
    tk = dict(hop_length=8, upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4],
              resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3,
              num_mels=20, upsample_initial_channel=64, use_template=True,
              pre_conv_kernel_size=7, post_conv_kernel_size=7)
    torch.manual_seed(0)
    m = BigVGANGenerator(**tk).eval()
    stress_init(m)
    x = mel_input(2, 20, 13)
    torch.manual_seed(9)
    tpl = torch.randn(2, 1, 13 * 8) * 0.3
    save("bigvgan_template_stress", tk, m, {"mel": x, "template": tpl}, m(x, tpl))


if __name__ == "__main__":
    main()
