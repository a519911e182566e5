"""DenseED (dense encoder-decoder with circular padding), 1D and 2D, on libarpde.

Mirrors the reference module trees so that state_dict keys, parameter order, default initialisation
(and its RNG consumption) and the public methods are identical:
  2D        : 2D-Burgers-SWAG/nn/denseEDcirc2d.py:34-353
  1D Burgers: 1D-Burger-SWAG/nn/denseEDcirc.py:32-344 (3-conv stem :243-257)
  1D KS     : 1D-KS-SWAG/nn/denseEDcirc.py (single k7/s2 stem :243-245)
The modules are parameter containers; DenseED.forward compiles `features` into a short program of
fused circular-conv ops (BN+ReLU[+nearest x2] folded into the conv that follows, dense-block concat
written in place) and executes it through the C ABI (program.py).  There is no CPU path.

Circular padding: the reference passes padding=2*p because torch<=1.2 padded (padding+1)//2 | padding//2
per side, i.e. p per side (reference README.md:40).  Here `padding` holds the effective p directly.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib as L
from .. import program as P

ACT_IDS = {'tanh': 1, 'relu': 2, 'lrelu': 3, 'sigmoid': 4, 'softplus': 5}


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t.requires_grad for t in tensors if t is not None)


def _standalone_guard(*tensors):
    L.require_cuda(*tensors)


def _as4(x):
    """[N,C,L] -> [N,C,1,L]; returns (tensor4d, was_1d)."""
    if x.dim() == 3:
        return x.unsqueeze(2), True
    return x, False


class _CircConvMixin:
    """forward of a lone circular convolution (wrap pad (k-1)//2, k=4 -> 1)."""

    def forward(self, x):
        _standalone_guard(x, self.weight)
        if _needs_grad(x, self.weight):
            # differentiable: the convolution is its own one-op program (forward, data and weight gradient through libarpde)
            return P.run_module(self, x, self.training)
        x4, one_d = _as4(x.detach().contiguous())
        N, Cin, H, W = x4.shape
        k, s = self.kernel_size[-1], self.stride[-1]
        kh = 1 if one_d else k
        Ho, Wo = (H // s if not one_d else 1), W // s
        y = torch.empty((N, self.out_channels, Ho, Wo), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            L.call('arpde_conv_fwd', L.ptr(x4), Cin * H * W, L.ptr(self.weight.detach()), 0, 0, 0, 0, L.ptr(y), 0,
                   N, Cin, H, W, self.out_channels, kh, k, s, 0, 0, 0, self.out_channels, 0, N, L.stream())
        return y.squeeze(2) if one_d else y


class CircConv1d(_CircConvMixin, nn.Conv1d):
    def __init__(self, cin, cout, kernel_size, stride=1):
        p = 1 if kernel_size == 4 else (kernel_size - 1) // 2
        nn.Conv1d.__init__(self, cin, cout, kernel_size, stride=stride, padding=p, bias=False,
                           padding_mode='circular' if p else 'zeros')


class CircConv2d(_CircConvMixin, nn.Conv2d):
    def __init__(self, cin, cout, kernel_size, stride=1):
        p = 1 if kernel_size == 4 else (kernel_size - 1) // 2
        nn.Conv2d.__init__(self, cin, cout, kernel_size, stride=stride, padding=p, bias=False,
                           padding_mode='circular' if p else 'zeros')


class _BNMixin:
    def forward(self, x):
        _standalone_guard(x, self.weight)
        if _needs_grad(x, self.weight, self.bias):
            # generic differentiable path of a lone BatchNorm (only reached module by module: dropout / bilinear topologies)
            return nn.modules.batchnorm._BatchNorm.forward(self, x)
        x4, one_d = _as4(x.detach().contiguous())
        N, C, H, W = x4.shape
        dev = x.device
        scale = torch.empty((C,), device=dev)
        shift = torch.empty((C,), device=dev)
        y = torch.empty_like(x4)
        with torch.cuda.device(dev):
            if self.training:
                stats = torch.empty((2 * C,), device=dev, dtype=torch.float64)
                L.call('arpde_bn_stats', L.ptr(x4), C * H * W, N, C, H * W, L.ptr(stats), L.stream())
                L.call('arpde_bn_finalize_train', L.ptr(stats), float(N * H * W), L.ptr(self.weight.detach()), L.ptr(self.bias.detach()),
                       L.ptr(self.running_mean), L.ptr(self.running_var), L.ptr(self.num_batches_tracked), float(self.momentum),
                       float(self.eps), L.ptr(scale), L.ptr(shift), 0, 0, C, L.stream())
            else:
                L.call('arpde_bn_fold_eval', L.ptr(self.weight.detach()), L.ptr(self.bias.detach()), L.ptr(self.running_mean),
                       L.ptr(self.running_var), float(self.eps), L.ptr(scale), L.ptr(shift), C, L.stream())
            L.call('arpde_affine_act', L.ptr(x4), C * H * W, L.ptr(scale), L.ptr(shift), L.ptr(y), N, C, H, W, 0, 0, 0, L.stream())
        return y.squeeze(2) if one_d else y


class BatchNorm1d(_BNMixin, nn.BatchNorm1d):
    pass


class BatchNorm2d(_BNMixin, nn.BatchNorm2d):
    pass


class _Elementwise(nn.Module):
    act_id, up = 0, 0

    def forward(self, x):
        _standalone_guard(x)
        if _needs_grad(x):
            return self._torch_forward(x)
        x4, one_d = _as4(x.detach().contiguous())
        N, C, H, W = x4.shape
        up_w = self.up
        up_h = 0 if one_d else self.up
        y = torch.empty((N, C, H << up_h, W << up_w), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            L.call('arpde_affine_act', L.ptr(x4), C * H * W, 0, 0, L.ptr(y), N, C, H, W, up_h, up_w, self.act_id, L.stream())
        return y.squeeze(2) if one_d else y


class ReLU(_Elementwise):
    act_id = 2

    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace          # kept for repr parity; outputs are always fresh tensors

    def _torch_forward(self, x):
        return torch.relu(x)


class Activation(_Elementwise):
    def __init__(self, name):
        super().__init__()
        self.act_id = ACT_IDS[name]
        self.name = name

    def _torch_forward(self, x):
        return {'tanh': torch.tanh, 'relu': torch.relu, 'lrelu': lambda t: F.leaky_relu(t, 0.01), 'sigmoid': torch.sigmoid,
                'softplus': lambda t: F.softplus(t, beta=4)}[self.name](x)


class UpsamplingNearest1d(_Elementwise):
    up = 1

    def __init__(self, scale_factor=2.):
        super().__init__()
        self.scale_factor = scale_factor

    def _torch_forward(self, x):
        return F.interpolate(x, scale_factor=self.scale_factor, mode='nearest')


class UpsamplingNearest2d(UpsamplingNearest1d):
    pass


class UpsamplingLinear1d(nn.Module):
    """F.interpolate(mode='linear', align_corners=True) (1D-*/nn/denseEDcirc.py UpsamplingLinear1d): no fused form -- a model that
    contains it runs module by module (convolutions on libarpde, this op on torch)."""

    def __init__(self, scale_factor=2.):
        super().__init__()
        self.scale_factor = scale_factor

    def forward(self, x):
        L.require_cuda(x)
        return F.interpolate(x, scale_factor=self.scale_factor, mode='linear', align_corners=True)


class UpsamplingBilinear2d(nn.Module):
    """F.interpolate(mode='bilinear', align_corners=True) (denseEDcirc2d.py:24-31); see UpsamplingLinear1d."""

    def __init__(self, scale_factor=2.):
        super().__init__()
        self.scale_factor = scale_factor

    def forward(self, x):
        L.require_cuda(x)
        return F.interpolate(x, scale_factor=self.scale_factor, mode='bilinear', align_corners=True)


def activation(name, *args):
    key = {'tanh': 'tanh', 'Tanh': 'tanh', 'relu': 'relu', 'ReLU': 'relu', 'lrelu': 'lrelu', 'LReLU': 'lrelu',
           'sigmoid': 'sigmoid', 'Sigmoid': 'sigmoid', 'softplus': 'softplus', 'Softplus': 'softplus'}.get(name)
    if key is None:
        raise ValueError('Unknown activation function')
    return Activation(key)


def needs_generic(module, training):
    """True when a sub-tree cannot run as the fused program: active dropout (train mode, p > 0) or a (bi)linear upsampling."""
    for m in module.modules():
        if isinstance(m, (UpsamplingLinear1d, UpsamplingBilinear2d)):
            return True
        if isinstance(m, _DenseLayer) and hasattr(m, 'norm2'):
            return True        # the reference's (malformed) bottleneck dense layer ends in BatchNorm + ReLU, not in a convolution
        if training and isinstance(m, nn.modules.dropout._DropoutNd) and m.p > 0:
            return True
    return False


def generic_forward(module, x):
    """Module-by-module forward (the reference's own nn.Sequential walk): every convolution is a libarpde call with its own
    autograd node, BatchNorm / ReLU / dropout / upsampling are the torch ops the reference uses, the dense-layer concat is a
    torch.cat.  Only topologies the fused program cannot express come here."""
    if isinstance(module, _DenseLayer):
        y = x
        for ch in module.children():
            y = generic_forward(ch, y)
        return torch.cat([x, y], 1)
    if isinstance(module, nn.Sequential):
        for ch in module.children():
            x = generic_forward(ch, x)
        return x
    return module(x)


class _Fused(nn.Sequential):
    """A sub-tree that can run on its own (DenseED.forward_test walks top-level children)."""

    def forward(self, x):
        if needs_generic(self, self.training):
            return generic_forward(self, x)
        return P.run_module(self, x, self.training)


def _mods(nd):
    return (CircConv1d, BatchNorm1d, UpsamplingNearest1d) if nd == 1 else (CircConv2d, BatchNorm2d, UpsamplingNearest2d)


def _dropout(nd, p):
    return nn.Dropout1d(p=p) if nd == 1 else nn.Dropout2d(p=p)


def _upsampler(nd, upsample):
    if upsample == 'nearest':
        return (UpsamplingNearest1d if nd == 1 else UpsamplingNearest2d)(scale_factor=2)
    if upsample == 'linear':
        return (UpsamplingLinear1d if nd == 1 else UpsamplingBilinear2d)(scale_factor=2)
    raise NotImplementedError("upsample=None (ConvTranspose) is not reachable from DenseED, which hard-codes 'nearest'")


class _DenseLayer(_Fused):
    """x -> cat([x, conv3(relu(bn(x)))], 1)   (denseEDcirc2d.py:34-68)"""

    def __init__(self, nd, in_features, growth_rate, drop_rate=0., bn_size=8, bottleneck=False, padding=1):
        super().__init__()
        Conv, BN, _ = _mods(nd)
        self.add_module('norm1', BN(in_features))
        self.add_module('relu1', ReLU(inplace=True))
        self.add_module('conv1', Conv(in_features, growth_rate, 2 * padding + 1, 1))
        if bottleneck and in_features > bn_size * growth_rate:
            # The reference registers 'conv1' twice (denseEDcirc2d.py:50-56): the 3x3 conv overwrites the 1x1 bottleneck IN PLACE,
            # so the module order becomes norm1, relu1, conv1 (in -> growth, 3x3), norm2 (bn_size*growth channels), relu2 and the
            # forward fails in BatchNorm unless bn_size == 1.  Same modules / state_dict keys here, same failure at call time
            # (main.py forces bottleneck=False).
            self.add_module('norm2', BN(bn_size * growth_rate))
            self.add_module('relu2', ReLU(inplace=True))
        if drop_rate > 0:
            self.add_module('dropout', _dropout(nd, drop_rate))
        self.growth_rate = growth_rate


class _DenseBlock(_Fused):
    def __init__(self, nd, num_layers, in_features, growth_rate, drop_rate, bn_size=4, bottleneck=False):
        super().__init__()
        for i in range(num_layers):
            self.add_module('denselayer%d' % (i + 1),
                            _DenseLayer(nd, in_features + i * growth_rate, growth_rate, drop_rate=drop_rate,
                                        bn_size=bn_size, bottleneck=bottleneck))


class _Transition(_Fused):
    """denseEDcirc2d.py:82-162 with the defaults DenseED uses (bottleneck=True, upsample='nearest')."""

    def __init__(self, nd, in_features, out_features, down, bottleneck=True, drop_rate=0, padding=1, upsample='nearest'):
        super().__init__()
        Conv, BN, _ = _mods(nd)
        self.add_module('norm1', BN(in_features))
        self.add_module('relu1', ReLU(inplace=True))
        if down:
            if bottleneck:
                self.add_module('conv1', Conv(in_features, out_features, 1, 1))
                if drop_rate > 0:
                    self.add_module('dropout1', _dropout(nd, drop_rate))
                self.add_module('norm2', BN(out_features))
                self.add_module('relu2', ReLU(inplace=True))
                self.add_module('conv2', Conv(out_features, out_features, padding * 2 + 2, 2))
                if drop_rate > 0:
                    self.add_module('dropout2', _dropout(nd, drop_rate))
            else:
                self.add_module('conv1', Conv(in_features, out_features, padding * 2 + 2, 2))
                if drop_rate > 0:
                    self.add_module('dropout1', _dropout(nd, drop_rate))
        else:
            if bottleneck:
                self.add_module('conv1', Conv(in_features, out_features, 1, 1))
                if drop_rate > 0:
                    self.add_module('dropout1', _dropout(nd, drop_rate))
                self.add_module('norm2', BN(out_features))
                self.add_module('relu2', ReLU(inplace=True))
                self.add_module('upsample', _upsampler(nd, upsample))
                self.add_module('conv2', Conv(out_features, out_features, 3, 1))
            else:
                self.add_module('upsample', _upsampler(nd, upsample))
                self.add_module('conv2', Conv(in_features, out_features, 3, 1))
            if drop_rate > 0:
                self.add_module('dropout1', _dropout(nd, drop_rate))


def last_decoding(nd, in_features, out_channels, drop_rate=0., upsample='nearest'):
    """denseEDcirc2d.py:165-190"""
    Conv, BN, _ = _mods(nd)
    last_up = _Fused()
    last_up.add_module('norm1', BN(in_features))
    last_up.add_module('relu1', ReLU(True))
    last_up.add_module('conv1', Conv(in_features, in_features // 2, 1, 1))
    if drop_rate > 0.:
        last_up.add_module('dropout1', _dropout(nd, drop_rate))
    last_up.add_module('norm2', BN(in_features // 2))
    last_up.add_module('relu2', ReLU(True))
    last_up.add_module('upsample', _upsampler(nd, upsample))
    last_up.add_module('conv2', Conv(in_features // 2, in_features // 4, 3, 1))
    last_up.add_module('norm3', BN(in_features // 4))
    last_up.add_module('relu3', ReLU(True))
    last_up.add_module('conv3', Conv(in_features // 4, out_channels, 5, 1))
    return last_up


class _DenseEDBase(nn.Module):
    nd = 2
    stem = 'burgers2d'

    def __init__(self, in_channels, out_channels, blocks, growth_rate=16, init_features=48, bn_size=8, drop_rate=0,
                 bottleneck=False, out_activation=None):
        super().__init__()
        if len(blocks) > 1 and len(blocks) % 2 == 0:
            raise ValueError('length of blocks must be an odd number, but got {}'.format(len(blocks)))
        nd = self.nd
        Conv, BN, _ = _mods(nd)
        enc_block_layers = blocks[: len(blocks) // 2]
        dec_block_layers = blocks[len(blocks) // 2:]
        self.features = nn.Sequential()
        if self.stem == 'burgers2d':        # denseEDcirc2d.py:247-255
            self.features.add_module('In_conv1', Conv(in_channels, init_features, 5, 1))
            self.features.add_module('In_conv_norm1', BN(init_features))
            self.features.add_module('In_conv_relu1', ReLU(True))
            self.features.add_module('In_conv3', Conv(init_features, init_features // 2, 3, 2))
            num_features = init_features // 2
        elif self.stem == 'burgers1d':      # 1D-Burger-SWAG/nn/denseEDcirc.py:243-257
            self.features.add_module('In_conv', Conv(in_channels, init_features // 4, 5, 1))
            self.features.add_module('In_conv_norm', BN(init_features // 4))
            self.features.add_module('In_conv_relu', ReLU(True))
            self.features.add_module('In_conv2', Conv(init_features // 4, init_features // 2, 3, 1))
            self.features.add_module('In_conv_norm2', BN(init_features // 2))
            self.features.add_module('In_conv_relu2', ReLU(True))
            self.features.add_module('In_conv3', Conv(init_features // 2, init_features, 3, 2))
            num_features = init_features
        else:                               # 1D-KS-SWAG/nn/denseEDcirc.py:243-245
            self.features.add_module('In_conv', Conv(in_channels, init_features, 7, 2))
            num_features = init_features
        for i, num_layers in enumerate(enc_block_layers):
            self.features.add_module('EncBlock%d' % (i + 1),
                                     _DenseBlock(nd, num_layers, num_features, growth_rate, drop_rate, bn_size, bottleneck))
            num_features = num_features + num_layers * growth_rate
            self.features.add_module('TransDown%d' % (i + 1),
                                     _Transition(nd, num_features, num_features // 2, down=True, drop_rate=drop_rate))
            num_features = num_features // 2
        for i, num_layers in enumerate(dec_block_layers):
            self.features.add_module('DecBlock%d' % (i + 1),
                                     _DenseBlock(nd, num_layers, num_features, growth_rate, drop_rate, bn_size, bottleneck))
            num_features += num_layers * growth_rate
            if i < len(dec_block_layers) - 1:
                self.features.add_module('TransUp%d' % (i + 1),
                                         _Transition(nd, num_features, num_features // 2, down=False, drop_rate=drop_rate))
                num_features = num_features // 2
        self.features.add_module('LastTransUp', last_decoding(nd, num_features, out_channels, drop_rate=0., upsample='nearest'))
        if out_activation is not None:
            self.features.add_module(out_activation, activation(out_activation))
        if self.stem == 'burgers2d':        # only the 2D reference prints this (:304-305)
            print('# params {}, # conv layers {}'.format(*self._num_parameters_convlayers()))
            self.in_channels = in_channels

    def forward(self, x):
        if needs_generic(self.features, self.training):
            return generic_forward(self.features, x)
        return P.run_module(self.features, x, self.training)

    # ---- diagnostics of the reference kept for API parity (denseEDcirc2d.py:316-352): same console output, same return values
    @torch.no_grad()
    def forward_test(self, x):
        """Walk the top-level children one by one and print the shape after each."""
        print('input: {}'.format(x.shape))
        for child_name, child in self.features.named_children():
            x = child(x)
            print('{}: {}'.format(child_name, x.shape))
        return x

    def _num_parameters_convlayers(self):
        """(number of scalars, number of parameter tensors whose path contains 'conv')."""
        named = list(self.named_parameters())
        return sum(p.numel() for _, p in named), sum('conv' in n for n, _ in named)

    def _count_parameters(self):
        running = 0
        for pname, p in self.named_parameters():
            running += p.numel()
            print('{}\n{}\n{}'.format(pname, p.size(), p.numel()))
            print('num of parameters so far: {}'.format(running))

    def reset_parameters(self, verbose=False):
        """Re-initialise every sub-module that knows how to (convolutions, BatchNorm), not this container itself."""
        for sub in self.modules():
            reset = None if isinstance(sub, type(self)) else getattr(sub, 'reset_parameters', None)
            if callable(reset):
                reset()
                if verbose:
                    print('Reset parameters in {}'.format(sub))


class DenseED2d(_DenseEDBase):
    nd, stem = 2, 'burgers2d'


class DenseED1dBurgers(_DenseEDBase):
    nd, stem = 1, 'burgers1d'


class DenseED1dKS(_DenseEDBase):
    nd, stem = 1, 'ks1d'
