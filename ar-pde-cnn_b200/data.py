"""On-device mirror of the reference's on-the-fly initial-condition data set.

`InitPeriodicCond2d` has the reference's constructor and indexing contract (2D-Burgers-SWAG/utils/burgerLoader2D.py:80-116:
`np.random.seed(index + 100000)`, `lam = randn(2, 2, (2*order+1)**2)`, `c = -1 + 2*rand(2)`, truncated Fourier series normalised to
max |f| = 2 and shifted by c) but synthesises the fields with `arpde_fourier_ic2d` on the GPU: the host only draws the 326 random
numbers of a sample.  `batch(indices)` builds a whole mini-batch in one launch -- the form the training loop wants."""
import numpy as np
import torch

from . import _lib as L


class InitPeriodicCond2d(torch.utils.data.Dataset):
    def __init__(self, ncells, nsamples, order=4, device='cuda'):
        super().__init__()
        self.order = order
        self.nsamples = nsamples
        self.ncells = ncells
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('arpde_b200.data.InitPeriodicCond2d synthesises on the GPU; there is no CPU path')

    def coefficients(self, indices):
        """The reference's random draws for the given sample indices: (lam [B,2,2,M], c [B,2]) as float64 numpy arrays."""
        m = (2 * self.order + 1) ** 2
        lam = np.empty((len(indices), 2, 2, m)); c = np.empty((len(indices), 2))
        for b, index in enumerate(indices):
            np.random.seed(int(index) + 100000)
            lam[b] = np.random.randn(2, 2, m)
            c[b] = -1 + np.random.rand(2) * 2
        return lam, c

    def batch(self, indices):
        """[len(indices), 2, ncells, ncells] float32 CUDA tensor: the stacked `__getitem__` results."""
        indices = list(indices)
        lam, c = self.coefficients(indices)
        lam_d = torch.from_numpy(lam).to(self.device); c_d = torch.from_numpy(c).to(self.device)
        out = torch.empty((len(indices), 2, self.ncells, self.ncells), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.call('arpde_fourier_ic2d', L.ptr(lam_d), L.ptr(c_d), L.ptr(out), len(indices), self.ncells, self.order, L.stream())
        return out

    def __getitem__(self, index):
        return self.batch([index])[0]

    def __len__(self):
        return self.nsamples
