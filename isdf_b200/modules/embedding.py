"""Icosahedral positional encoding -- host-side mirror of reference isdf/modules/embedding.py.

`PostionalEncoding` (sic, the reference's spelling, embedding.py:25) only carries the
hyper-parameters; the encoding itself is fused into the MLP kernels and never materialised on the
training path.  Calling the module directly encodes through `isdfb_pe_encode` (API parity)."""
import torch

from ..engine import Engine

_A, _B, _C, _D = 0.8506508, 0.5257311, 0.809017, 0.309017
ICO_DIRS = (
    (_A, 0, _B), (_C, 0.5, _D), (_B, _A, 0), (1, 0, 0), (_C, 0.5, -_D), (_A, 0, -_B), (_D, _C, -0.5),
    (0, _B, -_A), (0.5, _D, -_C), (0, 1, 0), (-_B, _A, 0), (-_D, _C, -0.5), (0, _B, _A), (-_D, _C, 0.5),
    (_D, _C, 0.5), (0.5, _D, _C), (0.5, -_D, _C), (0, 0, 1), (-0.5, _D, _C), (-_C, 0.5, _D), (-_C, 0.5, -_D),
)


def scale_input(tensor, transform=None, scale=None):
    """x -> scale * (R x + t)   (embedding.py:12-22); small torch helper for callers outside the kernels."""
    if transform is not None:
        tr = torch.as_tensor(transform, dtype=tensor.dtype, device=tensor.device)
        tensor = tensor @ tr[:3, :3].T + tr[:3, 3]
    if scale is not None:
        tensor = tensor * scale
    return tensor


class PostionalEncoding(torch.nn.Module):
    def __init__(self, min_deg=0, max_deg=6, scale=0.1, transform=None):
        super().__init__()
        if min_deg != 0:
            raise ValueError("isdf_b200 supports min_deg == 0 only (the reference never uses another value, "
                             "trainer.py:422)")
        self.min_deg, self.max_deg = min_deg, max_deg
        self.n_freqs = max_deg - min_deg + 1
        self.scale = scale
        self.transform = transform
        self.dirs = torch.tensor(ICO_DIRS, dtype=torch.float32).T.contiguous()      # [3, 21]
        self.embedding_size = 2 * self.dirs.shape[1] * self.n_freqs + 3
        self._engine = None

    def _get_engine(self, device):
        if self._engine is None or self._engine.device != torch.device(device):
            self._engine = Engine(device, self.n_freqs, 128, 1, self.scale, 1.0, transform=self.transform,
                                  precision="fp32", max_points=128)
        return self._engine

    def forward(self, tensor):
        return self._get_engine(tensor.device).pe_encode(tensor.float())
