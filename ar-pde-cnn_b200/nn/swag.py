"""SwagNN -- mirrors reference */nn/swag.py:10-231.

Buffer names (`<param name with dots replaced>_mean/_sq_mean/_cov_mat_sqrt`, `n_models`), checkpoint
format and method signatures are the reference's, so shipped .pth files load with strict=True.
`sample` draws its noise with the same torch calls in the same order as the reference
(randn_like per tensor, then normal_ on a [K,1] tensor when diagCov is False), so a seeded run
reproduces the reference's samples; the arithmetic is one libarpde launch over the flat parameter
vector (csrc/swag.cu).  `sample_flat` is the batched form used by the rollout engine."""
import copy
import os

import torch

from .. import _lib as L


class SwagNN(torch.nn.Module):
    def __init__(self, args, base, full_cov=True, max_models=1, eps=1e-30):
        super(SwagNN, self).__init__()
        self.device = args.device
        self.base = base
        self.full_cov = full_cov
        self.max_models = max_models
        self.eps = eps
        self.buildSwagParameters(base)

    def buildSwagParameters(self, baseModule):
        for (name, data) in list(baseModule.named_parameters()):
            if name is None:
                continue
            name = name.replace('.', '_')
            self.register_buffer('%s_mean' % name, data.new(data.size()).zero_())
            self.register_buffer('%s_sq_mean' % name, data.new(data.size()).zero_())
            if self.full_cov is True:
                self.register_buffer('%s_cov_mat_sqrt' % name, data.new_empty((self.max_models, data.numel())).zero_())
        self.register_buffer('n_models', torch.zeros([1], dtype=torch.long))

    def forward(self, *args, **kwargs):
        return self.base(*args, **kwargs)

    def calc_neg_log_joint(self, *args, **kwargs):
        return self.base.calc_neg_log_joint(*args, **kwargs)

    def swag_state_dict(self):
        return dict(filter(lambda v: (v[1].requires_grad == False and "base" not in v[0]), self.state_dict().items()))

    # ---------------------------------------------------------------------------------
    def _tables(self, module, diagCov):
        names = [n for n, _ in module.named_parameters()]
        means = [self.__getattr__('%s_mean' % n.replace('.', '_')) for n in names]
        sqs = [self.__getattr__('%s_sq_mean' % n.replace('.', '_')) for n in names]
        covs = None if diagCov else [self.__getattr__('%s_cov_mat_sqrt' % n.replace('.', '_')) for n in names]
        return names, means, sqs, covs

    def _launch(self, means, sqs, covs, z1, z2, S, scale, diagCov):
        L.require_cuda(*means)
        dev = means[0].device
        numels = [m.numel() for m in means]
        P = sum(numels)
        means = [m.contiguous() for m in means]
        sqs = [s.contiguous() for s in sqs]
        K = 0
        cov_arr = 0
        if not diagCov:
            covs = [c.contiguous() for c in covs]
            K = covs[0].shape[0]
            if any(c.shape[0] != K for c in covs):
                raise ValueError('cov_mat_sqrt buffers disagree on the number of stored models')
            cov_arr = L.ptr_array([c.data_ptr() for c in covs])
        out = torch.empty((S, P), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.call('arpde_swag_sample', L.ptr_array([m.data_ptr() for m in means]), L.ptr_array([s.data_ptr() for s in sqs]),
                   cov_arr, L.i64_array(numels), len(means), L.ptr(z1), L.ptr(z2), L.ptr(out), S, K, int(self.max_models),
                   float(scale), float(self.eps), 1 if diagCov else 0, L.stream())
        return out, numels

    def sample(self, module=None, scale=1.0, diagCov=False, copy_model=True, seed=None):
        if (module is None):
            module = self.base
        assert type(module) is type(self.base), 'Provided module is not the same as base module!'
        if seed is not None:
            torch.manual_seed(seed)
        module0 = copy.deepcopy(module) if copy_model else module
        names, means, sqs, covs = self._tables(module0, diagCov)
        # noise, drawn exactly as the reference does (swag.py:125 and :139), one tensor at a time
        z1s, z2s = [], []
        for i, m in enumerate(means):
            z1s.append(torch.randn_like(m).reshape(-1))
            if diagCov is False:
                z2s.append(covs[i].new_empty((covs[i].size(0), 1)).normal_().reshape(-1))
        z1 = torch.cat(z1s)
        z2 = torch.stack(z2s, 0) if z2s else None
        flat, numels = self._launch(means, sqs, covs, z1, z2, 1, scale, diagCov)
        off = 0
        for (name0, param), n in zip(list(module0.named_parameters()), numels):
            param.data = flat[0, off:off + n].view(param.shape)
            off += n
        return module0

    @torch.no_grad()
    def sample_flat(self, n_samples, scale=1.0, diagCov=False, generator=None, noise=None):
        """S posterior draws as one flat [S, P] tensor (named_parameters() order of self.base).
        noise: optional (z1 [S,P], z2 [S,n_tensors,K]) to reproduce given draws."""
        names, means, sqs, covs = self._tables(self.base, diagCov)
        dev = means[0].device
        P = sum(m.numel() for m in means)
        if noise is not None:
            z1, z2 = noise
        else:
            z1 = torch.randn((n_samples, P), device=dev, generator=generator)
            z2 = None if diagCov else torch.randn((n_samples, len(means), covs[0].shape[0]), device=dev, generator=generator)
        flat, numels = self._launch(means, sqs, covs, z1.contiguous(), None if z2 is None else z2.contiguous(),
                                    n_samples, scale, diagCov)
        return flat, names, numels

    def collect(self, module=None):
        """Running first/second moments + ring of the last max_models raw parameter vectors (swag.py:152-194).
        Called once per epoch: plain tensor arithmetic, format contract only."""
        if (module is None):
            module = self.base
        assert type(module) is type(self.base), 'Provided module is not the same as base module!'
        n = self.n_models.item()
        for name, param in list(module.named_parameters()):
            name = name.replace('.', '_')
            mean = self.__getattr__('%s_mean' % name)
            sq_mean = self.__getattr__('%s_sq_mean' % name)
            mean = mean * n / (n + 1.0) + param.data / (n + 1.0)
            sq_mean = sq_mean * n / (n + 1.0) + param.data ** 2 / (n + 1.0)
            if self.full_cov:
                cov_mat_sqrt = self.__getattr__('%s_cov_mat_sqrt' % name)
                # raw parameters are stored (not deviations) and [-(K-1):] keeps all rows when K == 1
                cov_mat_sqrt = torch.cat((cov_mat_sqrt[-(self.max_models - 1):], param.data.view(1, -1)), dim=0)
                self.__setattr__('%s_cov_mat_sqrt' % name, cov_mat_sqrt)
            self.__setattr__('%s_mean' % name, mean)
            self.__setattr__('%s_sq_mean' % name, sq_mean)
        self.n_models.add_(1)

    def saveModel(self, epoch, optimizer, scheduler, file_dir):
        if not os.path.exists(file_dir):
            os.makedirs(file_dir)
        print('[SWAG] Epoch {}, Saving SWAG model!'.format(epoch))
        state = {'epoch': epoch, 'state_dict': self.state_dict(),
                 'optimizer': optimizer.state_dict(), 'scheduler': scheduler.state_dict()}
        torch.save(state, file_dir + '/torchSwagModel_epoch{:d}.pth'.format(epoch))

    def loadModel(self, epoch, optimizer=None, scheduler=None, file_dir="."):
        try:
            file_name = file_dir + '/torchSwagModel_epoch{:d}.pth'.format(epoch)
            param_dict = torch.load(file_name, map_location=lambda storage, loc: storage)
            print('[SWAG] Found model at epoch: {:d}'.format(param_dict['epoch']))
        except FileNotFoundError:
            print('[SWAG] Error: Could not find PyTorch network')
            return
        self.load_state_dict(param_dict['state_dict'])
        if (not optimizer is None):
            optimizer.load_state_dict(param_dict['optimizer'])
            scheduler.load_state_dict(param_dict['scheduler'])
        print('[SWAG] Pre-trained swag model loaded!')
        print('[SWAG] Collected {:d} models'.format(self.n_models.item()))
        return optimizer, scheduler
