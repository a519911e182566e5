"""BayesNN -- mirrors reference */nn/bayesNN.py:10-155 (three functionally identical copies).

Same constructor, attributes, parameter names (`model.log_beta` is attached to the wrapped DenseED,
:46-49) and return shapes; calc_neg_log_joint runs as 2-3 launches of libarpde (csrc/loss.cu)."""
import os

import numpy as np
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from .. import functional as F_


class BayesNN(nn.Module):
    def __init__(self, args, model):
        super(BayesNN, self).__init__()
        self.dt = args.dt
        self.device = args.device
        # expected noise precision from the truncation error of the integrator (bayesNN.py:29-31)
        c0 = 0.2
        k0 = 3.0
        beta_mean = 1.0 / (c0 * args.dt ** (k0))
        self.beta_prior_shape = 10.
        self.beta_prior_rate = self.beta_prior_shape / beta_mean
        # Student-t prior on the weights (:37-38)
        self.w_prior_shape = 0.5
        self.w_prior_rate = 10
        self.model = model
        # numpy RNG, exactly like the reference (:46-47), so seeded runs draw the same log_beta
        beta0 = np.log(np.random.gamma(self.beta_prior_shape, 1. / self.beta_prior_rate, size=(1)))
        self.model.log_beta = Parameter(torch.Tensor(beta0).to(self.device))
        print('[BayesNN]: Expected Noise: {:.03f}'.format(beta_mean))
        print('[BayesNN]: Total number of parameters: {}'.format(self._num_parameters()))

    def _num_parameters(self):
        count = 0
        for name, param in self.named_parameters():
            count += param.numel()
        return count

    def beta(self):
        return self.model.log_beta.exp().item()

    def forward(self, input):
        return self.model(input)

    def predict(self, x_test):
        # the reference method is dead code that raises AttributeError (self.c0 is never set, bayesNN.py:93)
        raise AttributeError("'BayesNN' object has no attribute 'c0'")

    def calc_neg_log_joint(self, output, target, ntrain):
        """Negative log joint (shape-(1,) tensor), differentiable w.r.t. output, target, log_beta and
        every tensor of model.features.parameters()."""
        return F_.neg_log_joint(output, target, ntrain, self.model.log_beta, list(self.model.features.parameters()),
                                self.w_prior_shape, self.w_prior_rate, self.beta_prior_shape, self.beta_prior_rate)

    def saveNetwork(self, epoch, file_dir='./networks'):
        if not os.path.exists(file_dir):
            os.makedirs(file_dir)
        print('Epoch {}, Saving network!'.format(epoch))
        torch.save(self.state_dict(), file_dir + '/torchNet_epoch{:d}.pth'.format(epoch))

    def loadNetwork(self, epoch, file_dir='./networks'):
        try:
            print('[BayesNN]: Attempting to load network from epoch {:d}'.format(epoch))
            file_name = file_dir + '/torchNet_epoch{:d}.pth'.format(epoch)
            param_dict = torch.load(file_name, map_location=lambda storage, loc: storage)
        except FileNotFoundError:
            print('Error: Could not find PyTorch network')
            return
        self.load_state_dict(param_dict)
        print('[BayesNN]: Pre-trained network loaded!')
