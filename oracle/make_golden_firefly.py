"""Generate tests/golden/firefly_small_stress.npz: the UNMODIFIED reference UnifyGenerator(ConvNeXtEncoder, HiFiGANGenerator)
composition of configs/model/generator/firefly-gan-base.yaml (5 upsample stages, pre/post conv kernel 13) at reduced width,
SURVEY-8d stress weights.  Run only in the build container:  python oracle/make_golden_firefly.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (path setup + helpers; its main() is not run)

from fish_vocoder.modules.encoders.convnext import ConvNeXtEncoder  # noqa: E402
from fish_vocoder.modules.generators.hifigan import HiFiGANGenerator  # noqa: E402
from fish_vocoder.modules.generators.unify import UnifyGenerator  # noqa: E402


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    kw = dict(
        backbone=dict(input_channels=24, depths=[1, 1, 2, 1], dims=[32, 48, 64, 96], drop_path_rate=0.2, kernel_size=7),
        head=dict(hop_length=64, upsample_rates=[4, 2, 2, 2, 2], upsample_kernel_sizes=[8, 4, 4, 4, 4],
                  resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3, num_mels=96,
                  upsample_initial_channel=128, use_template=False, pre_conv_kernel_size=13, post_conv_kernel_size=13))
    torch.manual_seed(0)
    m = UnifyGenerator(backbone=ConvNeXtEncoder(**kw["backbone"]), head=HiFiGANGenerator(**kw["head"])).eval()
    G.stress_init(m)
    x = G.mel_input(2, 24, 11)
    y = m(x)
    G.save("firefly_small_stress", kw, m, {"mel": x}, y)


if __name__ == "__main__":
    main()
