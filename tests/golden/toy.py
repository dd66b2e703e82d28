"""Toy encoder/decoder problems shared by tests/golden/make_golden.py (reference side) and
tests/test_golden_gpu.py (our side).  TEST INFRASTRUCTURE."""
import torch


def toy_problem(seed, B, d_in, latent, d_out):
    g = torch.Generator().manual_seed(seed)
    return dict(
        x=torch.randn(B, d_in, generator=g),
        labels=torch.randn(B, d_out, generator=g),
        enc_w=torch.randn(latent, d_in, generator=g) * 0.5,
        enc_b=torch.randn(latent, generator=g) * 0.1,
        dec_w=torch.randn(d_out, latent, generator=g) * 0.3,   # structured: Linear(latent, d_out)
        emb=torch.randn(latent, d_out, generator=g),           # categorical: Embedding(latent, d_out)
    )


def make_modules(prob, categorical):
    enc = torch.nn.Linear(prob["x"].shape[1], prob["enc_w"].shape[0])
    enc.weight.data.copy_(prob["enc_w"])
    enc.bias.data.copy_(prob["enc_b"])
    if categorical:
        dec = torch.nn.Embedding(*prob["emb"].shape)
        dec.weight.data.copy_(prob["emb"])
    else:
        dec = torch.nn.Linear(prob["dec_w"].shape[1], prob["dec_w"].shape[0], bias=False)
        dec.weight.data.copy_(prob["dec_w"])
    return enc, dec


def loss_fun(encoder_input, z, decoder_input, decoder_output, labels):
    per_row = ((decoder_output - labels) ** 2).sum(-1)
    return per_row, {"acc": (decoder_output.sum(-1) > 0).float()}


class CatDecoder(torch.nn.Module):
    def __init__(self, emb):
        super().__init__()
        self.emb = emb

    def forward(self, z, decoder_input):
        return self.emb(z)


class BitDecoder(torch.nn.Module):
    def __init__(self, lin):
        super().__init__()
        self.lin = lin

    def forward(self, z, *args):
        return self.lin(z)


