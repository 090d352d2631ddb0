"""`loss_config.target: vidtok.modules.losses.GeneralLPIPSWithDiscriminator` appears in every config but is a
training-only component (LPIPS download + discriminator); the engine accepts and skips it."""
import torch.nn as nn


class GeneralLPIPSWithDiscriminator(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
