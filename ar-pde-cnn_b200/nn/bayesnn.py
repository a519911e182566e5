"""BayesNN: the reference's Bayesian wrapper around a DenseED (*/nn/bayesNN.py:10-155, three functionally identical copies) as a
thin host-side shell around libarpde.

Contract kept for the unmodified drivers and the shipped checkpoints: constructor `(args, model)`, the noise-precision parameter is
attached to the WRAPPED network as `model.log_beta` (state-dict key `model.log_beta`, bayesNN.py:46-49) and initialised from numpy's
global RNG, `calc_neg_log_joint` returns a shape-(1,) tensor, `beta()` a python float, the two console lines of the constructor.
The one hot function, calc_neg_log_joint (bayesNN.py:98-129), is 2-3 launches of csrc/loss.cu behind `functional.neg_log_joint`."""
import os
from collections import namedtuple

import numpy as np
import torch
from torch import nn

from .. import functional as F_

# Gamma(shape, rate) prior on the noise precision beta and Student-t prior on every weight (bayesNN.py:29-38)
Priors = namedtuple('Priors', 'beta_shape beta_rate w_shape w_rate expected_noise')


def make_priors(dt, c0=0.2, order=3.0, beta_shape=10.0, w_shape=0.5, w_rate=10):
    """The expected noise precision is the inverse truncation error of the time integrator, 1 / (c0 dt^order)."""
    expected = 1.0 / (c0 * dt ** order)
    return Priors(beta_shape, beta_shape / expected, w_shape, w_rate, expected)


class BayesNN(nn.Module):
    def __init__(self, args, model):
        super().__init__()
        self.dt, self.device, self.model = args.dt, args.device, model
        pr = make_priors(args.dt)
        self.beta_prior_shape, self.beta_prior_rate = pr.beta_shape, pr.beta_rate
        self.w_prior_shape, self.w_prior_rate = pr.w_shape, pr.w_rate
        # one draw from numpy's global generator (not torch's): seeded drivers reproduce the reference's initial log_beta
        draw = np.random.gamma(pr.beta_shape, 1.0 / pr.beta_rate, size=(1))
        model.log_beta = nn.Parameter(torch.as_tensor(np.log(draw), dtype=torch.float32).to(self.device))
        print('[BayesNN]: Expected Noise: {:.03f}'.format(pr.expected_noise))
        print('[BayesNN]: Total number of parameters: {}'.format(self._num_parameters()))

    def _num_parameters(self):
        return sum(p.numel() for p in self.parameters())

    def beta(self):
        return float(self.model.log_beta.detach().exp())

    def forward(self, input):
        return self.model(input)

    def predict(self, x_test):
        # dead code upstream: it reads self.c0, which is never assigned (bayesNN.py:93); fail the same way
        raise AttributeError("'BayesNN' object has no attribute 'c0'")

    def calc_neg_log_joint(self, output, target, ntrain):
        """Negative log joint, shape (1,); differentiable w.r.t. output, target, log_beta and model.features.parameters()."""
        weights = list(self.model.features.parameters())
        return F_.neg_log_joint(output, target, ntrain, self.model.log_beta, weights, self.w_prior_shape, self.w_prior_rate,
                                self.beta_prior_shape, self.beta_prior_rate)

    @staticmethod
    def _ckpt_path(file_dir, epoch):
        return os.path.join(file_dir, 'torchNet_epoch%d.pth' % epoch)

    def saveNetwork(self, epoch, file_dir='./networks'):
        os.makedirs(file_dir, exist_ok=True)
        print('Epoch {}, Saving network!'.format(epoch))
        torch.save(self.state_dict(), self._ckpt_path(file_dir, epoch))

    def loadNetwork(self, epoch, file_dir='./networks'):
        print('[BayesNN]: Attempting to load network from epoch {:d}'.format(epoch))
        path = self._ckpt_path(file_dir, epoch)
        if not os.path.isfile(path):
            print('Error: Could not find PyTorch network')
            return
        self.load_state_dict(torch.load(path, map_location='cpu'))
        print('[BayesNN]: Pre-trained network loaded!')
