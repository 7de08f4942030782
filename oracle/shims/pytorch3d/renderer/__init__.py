import torch

from .cameras import CamerasBase, PerspectiveCameras


class HarmonicEmbedding(torch.nn.Module):
    """[sin(x_d * 2^k) | cos(x_d * 2^k) | x], d-major / k-minor (reference use: embedding.py:44).

    Version caveat: recent pytorch3d evaluates the cos block as
    sin(x + pi/2); we use cos(x) and say so in DESIGN.md.
    """

    def __init__(self, n_harmonic_functions=6, omega_0=1.0, logspace=True, append_input=True):
        super().__init__()
        if logspace:
            freqs = 2.0 ** torch.arange(n_harmonic_functions, dtype=torch.float32)
        else:
            freqs = torch.linspace(1.0, 2.0 ** (n_harmonic_functions - 1), n_harmonic_functions, dtype=torch.float32)
        self.register_buffer("_frequencies", freqs * omega_0, persistent=False)
        self.append_input = append_input

    def forward(self, x, **kwargs):
        scaled = (x[..., None] * self._frequencies).reshape(*x.shape[:-1], -1)
        parts = [scaled.sin(), scaled.cos()]
        if self.append_input:
            parts.append(x)
        return torch.cat(parts, dim=-1)

    def get_output_dim(self, input_dims=3):
        return input_dims * (2 * len(self._frequencies) + int(self.append_input))
