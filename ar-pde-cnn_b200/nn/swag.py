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


def _bufname(param_name, kind):
    """Buffer of a parameter: dots of the parameter path replaced (swag.py:50-58), kind in mean | sq_mean | cov_mat_sqrt."""
    return param_name.replace('.', '_') + '_' + kind


class SwagNN(torch.nn.Module):
    def __init__(self, args, base, full_cov=True, max_models=1, eps=1e-30):
        super().__init__()
        self.device, self.base = args.device, base
        self.full_cov, self.max_models, self.eps = full_cov, max_models, eps
        self.buildSwagParameters(base)

    def buildSwagParameters(self, baseModule):
        """First / second moment of every parameter and, for the low-rank covariance, max_models rows of flattened parameters."""
        for pname, p in baseModule.named_parameters():
            self.register_buffer(_bufname(pname, 'mean'), torch.zeros_like(p))
            self.register_buffer(_bufname(pname, 'sq_mean'), torch.zeros_like(p))
            if self.full_cov is True:
                self.register_buffer(_bufname(pname, 'cov_mat_sqrt'), p.new_zeros((self.max_models, p.numel())))
        self.register_buffer('n_models', torch.zeros(1, dtype=torch.long))

    def forward(self, *args, **kwargs):
        return self.base(*args, **kwargs)

    def calc_neg_log_joint(self, *args, **kwargs):
        return self.base.calc_neg_log_joint(*args, **kwargs)

    def swag_state_dict(self):
        """The SWAG buffers alone (everything in the state dict that is neither trainable nor part of the wrapped model)."""
        return {k: v for k, v in self.state_dict().items() if not v.requires_grad and 'base' not in k}

    # ---------------------------------------------------------------------------------
    def _tables(self, module, diagCov):
        names = [n for n, _ in module.named_parameters()]
        means = [getattr(self, _bufname(n, 'mean')) for n in names]
        sqs = [getattr(self, _bufname(n, 'sq_mean')) for n in names]
        covs = None if diagCov else [getattr(self, _bufname(n, 'cov_mat_sqrt')) for n in names]
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

    @torch.no_grad()
    def collect(self, module=None):
        """One more model into the running moments (weights n/(n+1), 1/(n+1)) and into the window of the last max_models parameter
        vectors (swag.py:152-194).  Quirks kept: the window stores RAW parameters, not deviations (sample() centres the non-zero
        rows), and with max_models == 1 nothing is ever dropped from it.  Once per epoch: plain tensor arithmetic."""
        module = self.base if module is None else module
        assert type(module) is type(self.base), 'Provided module is not the same as base module!'
        n = int(self.n_models)
        keep = self.max_models - 1
        for pname, p in module.named_parameters():
            w = p.detach()
            for kind, value in (('mean', w), ('sq_mean', w * w)):
                old = getattr(self, _bufname(pname, kind))
                setattr(self, _bufname(pname, kind), old * n / (n + 1.0) + value / (n + 1.0))
            if self.full_cov:
                window = getattr(self, _bufname(pname, 'cov_mat_sqrt'))
                window = window[-keep:] if keep > 0 else window
                setattr(self, _bufname(pname, 'cov_mat_sqrt'), torch.cat((window, w.reshape(1, -1)), dim=0))
        self.n_models += 1

    @staticmethod
    def _ckpt_path(file_dir, epoch):
        return os.path.join(file_dir, 'torchSwagModel_epoch%d.pth' % epoch)

    def saveModel(self, epoch, optimizer, scheduler, file_dir):
        """Checkpoint dictionary of the reference (swag.py:196-210): epoch | state_dict | optimizer | scheduler."""
        os.makedirs(file_dir, exist_ok=True)
        print('[SWAG] Epoch {}, Saving SWAG model!'.format(epoch))
        torch.save(dict(epoch=epoch, state_dict=self.state_dict(), optimizer=optimizer.state_dict(), scheduler=scheduler.state_dict()),
                   self._ckpt_path(file_dir, epoch))

    def loadModel(self, epoch, optimizer=None, scheduler=None, file_dir="."):
        path = self._ckpt_path(file_dir, epoch)
        if not os.path.isfile(path):
            print('[SWAG] Error: Could not find PyTorch network')
            return
        ckpt = torch.load(path, map_location='cpu')
        print('[SWAG] Found model at epoch: {:d}'.format(ckpt['epoch']))
        self.load_state_dict(ckpt['state_dict'])
        if optimizer is not None:
            optimizer.load_state_dict(ckpt['optimizer'])
            scheduler.load_state_dict(ckpt['scheduler'])
        print('[SWAG] Pre-trained swag model loaded!')
        print('[SWAG] Collected {:d} models'.format(int(self.n_models)))
        return optimizer, scheduler
