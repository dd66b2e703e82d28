"""lvmhelpers/utils.py counterpart: the decoder side of the drop-in boundary."""
import torch.nn as nn


class DeterministicWrapper(nn.Module):
    """Pass-through wrapper for the decoder (reference lvmhelpers/utils.py:5-17).

    The marginalisation methods call ``decoder(z, decoder_input)``; no sampling
    happens on top of the wrapped agent.
    """

    def __init__(self, agent):
        super().__init__()
        self.agent = agent

    def forward(self, *args, **kwargs):
        return self.agent(*args, **kwargs)
