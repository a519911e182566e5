"""
oracle/arpde_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (torch functional ops, fp32 or fp64) of the reference's hot path:
DenseED forward (1D Burgers, 1D KS, 2D Burgers), the periodic finite-difference
Crank-Nicolson / backward-Euler integrators, the negative log joint, SWAG sampling
and the auto-regressive rollout.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module.  The product
package (ar-pde-cnn_b200/) never does.

Parity status: PINNED.  tests/golden/make_golden.py imports the real reference from
/root/reference (with the torch<=1.2 circular-padding shim of reference README.md:40) and
dumps inputs/outputs/gradients to tests/golden/*.npz; tests/test_oracle_golden.py checks
this restatement against every one of them.  The reference itself ships no tests, so
those fixtures (plus the shipped 1D checkpoints they embed) are the only pins that exist.

Every function cites the reference file:line it restates (paths relative to
/root/reference).  The reference's arithmetic lives in PyTorch (README.md:32-36 names
"PyTorch 1.1.0"; no lock file); gradients in this oracle come from torch autograd over the
restated forward, exactly as the reference obtains them.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# circular convolution (torch<=1.2 semantics == symmetric wrap pad (k-1)//2, k=4 -> 1)
#   reference: every nn.Conv{1,2}d(padding=2*p, padding_mode='circular'), README.md:40
# --------------------------------------------------------------------------------------


def eff_pad(k):
    """Effective wrap padding per side of a reference circular conv with kernel k."""
    return 1 if k == 4 else (k - 1) // 2


def conv_circ(x, w, stride=1):
    """Circular cross-correlation; x [N,C,L] or [N,C,H,W]; w [Co,Ci,k] / [Co,Ci,k,k]."""
    k = w.shape[-1]
    p = eff_pad(k)
    if x.dim() == 3:
        xp = F.pad(x, (p, p), mode='circular') if p else x
        return F.conv1d(xp, w, stride=stride)
    xp = F.pad(x, (p, p, p, p), mode='circular') if p else x
    return F.conv2d(xp, w, stride=stride)


def _bn(x, sd, name, training):
    """BatchNorm{1,2}d eps 1e-5 momentum 0.1 (torch defaults used by the reference)."""
    rm, rv = sd[name + '.running_mean'], sd[name + '.running_var']
    y = F.batch_norm(x, rm, rv, sd[name + '.weight'], sd[name + '.bias'],
                     training=training, momentum=0.1, eps=1e-5)
    if training and (name + '.num_batches_tracked') in sd:
        sd[name + '.num_batches_tracked'] += 1
    return y


def _up2(x):
    """nearest x2 (denseEDcirc2d.py:15-21, denseEDcirc.py:13-19): out[i] = in[i >> 1]."""
    if x.dim() == 3:
        return x.repeat_interleave(2, dim=2)
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


_ACT = {
    'tanh': torch.tanh, 'relu': torch.relu, 'lrelu': lambda t: F.leaky_relu(t, 0.01),
    'sigmoid': torch.sigmoid, 'softplus': lambda t: F.softplus(t, beta=4),
}


def densed_forward(system, sd, x, blocks, training=False, out_activation=None):
    """DenseED.forward restated as straight-line code over a state dict.

    system: 'burgers2d' (2D-Burgers-SWAG/nn/denseEDcirc2d.py:242-302,313-314),
            'burgers1d' (1D-Burger-SWAG/nn/denseEDcirc.py:243-257 stem, rest :258-300),
            'ks1d'      (1D-KS-SWAG/nn/denseEDcirc.py:243-245 stem).
    sd: dict 'features.<...>' -> tensor (running stats are updated in place when training).
    blocks: list of dense-block sizes (odd length); len 1 => decoder block only.
    """
    f = 'features.'
    if system == 'burgers2d':
        x = conv_circ(x, sd[f + 'In_conv1.weight'])
        x = torch.relu(_bn(x, sd, f + 'In_conv_norm1', training))
        x = conv_circ(x, sd[f + 'In_conv3.weight'], stride=2)
    elif system == 'burgers1d':
        x = conv_circ(x, sd[f + 'In_conv.weight'])
        x = torch.relu(_bn(x, sd, f + 'In_conv_norm', training))
        x = conv_circ(x, sd[f + 'In_conv2.weight'])
        x = torch.relu(_bn(x, sd, f + 'In_conv_norm2', training))
        x = conv_circ(x, sd[f + 'In_conv3.weight'], stride=2)
    elif system == 'ks1d':
        x = conv_circ(x, sd[f + 'In_conv.weight'], stride=2)
    else:
        raise ValueError(system)

    def dense_block(x, prefix, nlayers):
        # _DenseLayer.forward: cat([x, conv(relu(bn(x)))]) (denseEDcirc2d.py:58-68)
        for j in range(1, nlayers + 1):
            p = '%s.denselayer%d.' % (prefix, j)
            y = torch.relu(_bn(x, sd, p + 'norm1', training))
            y = conv_circ(y, sd[p + 'conv1.weight'])
            x = torch.cat([x, y], 1)
        return x

    n_enc = len(blocks) // 2 if len(blocks) > 1 else 0
    enc, dec = blocks[:n_enc], blocks[n_enc:]
    for i, nl in enumerate(enc):
        x = dense_block(x, f + 'EncBlock%d' % (i + 1), nl)
        p = f + 'TransDown%d.' % (i + 1)      # _Transition(down=True, bottleneck=True), :96-114
        x = torch.relu(_bn(x, sd, p + 'norm1', training))
        x = conv_circ(x, sd[p + 'conv1.weight'])
        x = torch.relu(_bn(x, sd, p + 'norm2', training))
        x = conv_circ(x, sd[p + 'conv2.weight'], stride=2)          # k4 s2, wrap pad 1
    for i, nl in enumerate(dec):
        x = dense_block(x, f + 'DecBlock%d' % (i + 1), nl)
        if i < len(dec) - 1:
            p = f + 'TransUp%d.' % (i + 1)    # _Transition(down=False, bottleneck=True), :123-145
            x = torch.relu(_bn(x, sd, p + 'norm1', training))
            x = conv_circ(x, sd[p + 'conv1.weight'])
            x = torch.relu(_bn(x, sd, p + 'norm2', training))
            x = conv_circ(_up2(x), sd[p + 'conv2.weight'])
    p = f + 'LastTransUp.'                    # last_decoding, denseEDcirc2d.py:165-190
    x = torch.relu(_bn(x, sd, p + 'norm1', training))
    x = conv_circ(x, sd[p + 'conv1.weight'])
    x = torch.relu(_bn(x, sd, p + 'norm2', training))
    x = conv_circ(_up2(x), sd[p + 'conv2.weight'])
    x = torch.relu(_bn(x, sd, p + 'norm3', training))
    x = conv_circ(x, sd[p + 'conv3.weight'])
    if out_activation is not None:
        x = _ACT[out_activation.lower()](x)
    return x


# --------------------------------------------------------------------------------------
# periodic finite-difference filters (cross-correlation, taps pre-divided in fp32,
# 1/dx^n applied afterwards) -- restated with torch.roll
# --------------------------------------------------------------------------------------
GRAD_TAPS = {   # burgerFiniteDifference.py:186-195, ksFiniteDifference.py:96-99
    3: ([-1., 0., 1.], 2.), 5: ([1., -8., 0., 8., -1.], 12.),
    7: ([-1., 9., -45., 0., 45., -9., 1.], 60.),
}
LAP_TAPS = {    # burgerFiniteDifference.py:62-65, ksFiniteDifference.py:58-61
    3: ([1., -2., 1.], 1.), 5: ([-1., 16., -30., 16., -1.], 12.),
}
D4_TAPS = {     # ksFiniteDifference.py:19-22
    5: ([1., -4., 6., -4., 1.], 1.), 7: ([-1., 12., -39., 56., -39., 12., -1.], 6.),
}


def _taps32(taps, den):
    """Tap values as the reference builds them: FloatTensor(taps) / den in fp32."""
    return (torch.tensor(taps, dtype=torch.float32) / den).tolist()


def filt1d(u, taps_den, scale):
    """out[i] = sum_k w[k] u[(i+k-p) mod L] / scale, along the last dim."""
    w = _taps32(*taps_den)
    p = (len(w) - 1) // 2
    out = None
    for k, wk in enumerate(w):
        if wk == 0.0:
            continue
        t = wk * torch.roll(u, shifts=p - k, dims=-1)
        out = t if out is None else out + t
    return out / scale


def filt2d(u, w3, scale):
    """3x3 periodic cross-correlation over the last two dims ([row=H, col=W])."""
    out = None
    for ky in range(3):
        for kx in range(3):
            wk = w3[ky][kx]
            if wk == 0.0:
                continue
            t = wk * torch.roll(u, shifts=(1 - ky, 1 - kx), dims=(-2, -1))
            out = t if out is None else out + t
    return out / scale


SOBEL_H = [[-1 / 8., 0., 1 / 8.], [-2 / 8., 0., 2 / 8.], [-1 / 8., 0., 1 / 8.]]   # burger2DFiniteDifference.py:69-71
SOBEL_V = [[-1 / 8., -2 / 8., -1 / 8.], [0., 0., 0.], [1 / 8., 2 / 8., 1 / 8.]]   # transpose, :80
LAPLACE_2D = [[.5, 0., .5], [0., -2., 0.], [.5, 0., .5]]                          # :23-27 (smoothed + transpose)


def burgers2d_rhs(uv, dx, nu):
    """R(u,v) of the 2D coupled Burgers system (burger2DFiniteDifference.py:185-203)."""
    u, v = uv[:, :1], uv[:, 1:]
    gh = lambda t: filt2d(t, SOBEL_H, dx)
    gv = lambda t: filt2d(t, SOBEL_V, dx)
    lap = lambda t: filt2d(t, LAPLACE_2D, dx ** 2)
    ru = -gh(0.5 * u ** 2) - v * gv(u) + nu * lap(u)
    rv = -u * gh(v) - gv(0.5 * v ** 2) + nu * lap(v)
    return torch.cat([ru, rv], 1)


def burgers2d_cn(uPred, uPred0, dt, dx, nu):
    """Burger2DIntegrate.crankNicolson (burger2DFiniteDifference.py:175-211)."""
    return uPred0 + 0.5 * dt * (burgers2d_rhs(uPred, dx, nu) + burgers2d_rhs(uPred0, dx, nu))


def burgers2d_be(uPred, uPred0, dt, dx, nu):
    """Burger2DIntegrate.backwardEuler (:148-173)."""
    return uPred0 + dt * burgers2d_rhs(uPred, dx, nu)


def burgers1d_cn(uPred, uPred0, dt, dx, nu, k1=3, k2=3):
    """BurgerIntegrate.crankNicolson (burgerFiniteDifference.py:284-302)."""
    g = lambda t: filt1d(t, GRAD_TAPS[k1], dx)
    lap = lambda t: filt1d(t, LAP_TAPS[k2], dx ** 2)
    return uPred0 - 0.5 * dt * ((uPred * g(uPred) + uPred0 * g(uPred0)) - nu * (lap(uPred) + lap(uPred0)))


def burgers1d_be(uPred, uPred0, dt, dx, nu, k1=3, k2=3):
    """BurgerIntegrate.backwardEuler (:264-282) -- note the literal double 0.5."""
    g = lambda t: filt1d(t, GRAD_TAPS[k1], dx)
    lap = lambda t: filt1d(t, LAP_TAPS[k2], dx ** 2)
    fluxdx = 0.5 * g(uPred ** 2)
    return uPred0 + dt * (-0.5 * fluxdx + nu * lap(uPred))


# ---- P4 variants no reference driver calls (restated for the parity tests of the product's roll path) ----------------------
SOBEL_H5 = [[v / 108. for v in row] for row in ([1, -8, 0, 8, -1], [2, -16, 0, 16, -2], [3, -24, 0, 24, -3], [2, -16, 0, 16, -2],
                                                [1, -8, 0, 8, -1])]                        # burger2DFiniteDifference.py:74-78
SOBEL_V5 = [list(r) for r in zip(*SOBEL_H5)]                                               # transpose, :84
LAPLACE_5 = [[v / 12. for v in row] for row in ([0, 0, -1, 0, 0], [0, 0, 16, 0, 0], [-1, 16, -60, 16, -1], [0, 0, 16, 0, 0],
                                                [0, 0, -1, 0, 0])]                         # :31-35


def filt2d_k(u, w, scale):
    """k x k periodic cross-correlation over the last two dims (gather form, independent of the product's roll form)."""
    k = len(w)
    p = (k - 1) // 2
    H, W = u.shape[-2:]
    iy = (torch.arange(H)[:, None] + torch.arange(k)[None, :] - p) % H          # [H, k]
    ix = (torch.arange(W)[:, None] + torch.arange(k)[None, :] - p) % W
    patches = u[..., iy[:, None, :, None], ix[None, :, None, :]]                # [..., H, W, k, k]
    return (patches * torch.tensor(w, dtype=u.dtype)).sum((-2, -1)) / scale


def burgers2d_rhs_k(uv, dx, nu, k1, k2):
    """R(u, v) with the 3x3 or 5x5 Sobel / Laplace stencils (burger2DFiniteDifference.py:185-203)."""
    u, v = uv[:, :1], uv[:, 1:]
    sh, sv = (SOBEL_H, SOBEL_V) if k1 == 3 else (SOBEL_H5, SOBEL_V5)
    lp = LAPLACE_2D if k2 == 3 else LAPLACE_5
    gh = lambda t: filt2d_k(t, sh, dx)
    gv = lambda t: filt2d_k(t, sv, dx)
    lap = lambda t: filt2d_k(t, lp, dx ** 2)
    return torch.cat([-gh(0.5 * u ** 2) - v * gv(u) + nu * lap(u), -u * gh(v) - gv(0.5 * v ** 2) + nu * lap(v)], 1)


def burgers2d_integrate_k(uPred, uPred0, dt, dx, nu, k1, k2, cn=True):
    if cn:
        return uPred0 + 0.5 * dt * (burgers2d_rhs_k(uPred, dx, nu, k1, k2) + burgers2d_rhs_k(uPred0, dx, nu, k1, k2))
    return uPred0 + dt * burgers2d_rhs_k(uPred, dx, nu, k1, k2)


def upwind_grad(u, dx):
    """GradFilter1D(kernel_size=2).__call__ = conditionalUpwind (burgerFiniteDifference.py:224-244): it RETURNS u2, the forward
    difference, regardless of the sign of u (reference quirk)."""
    idx = (torch.arange(u.shape[-1]) + 1) % u.shape[-1]
    return (u[..., idx] - u) / dx


def weno_flux_grad(u, dx):
    """FluxWENOFilter1D.__call__ (burgerFiniteDifference.py:89-166), gather form over the L+1 cell edges as the reference lays them out."""
    L_ = u.shape[-1]
    f = u ** 2 / 2.0

    def edges(fl):
        i = torch.arange(L_ + 1)
        at = lambda o: fl[..., (i + o) % L_]
        e1 = -0.5 * at(-2) + 1.5 * at(-1)
        e2 = 0.5 * at(0) + 0.5 * at(1)
        b1 = (at(-1) - at(-2)) ** 2
        b2 = (at(1) - at(0)) ** 2
        w1 = 1. / (3 * (1e-6 + b1) ** 2)
        w2 = 2. / (3 * (1e-6 + b2) ** 2)
        return (w1 * e1 + w2 * e2) / (w1 + w2)
    e = edges(f)
    er = torch.flip(edges(torch.flip(f, dims=[-1])), dims=[-1])
    g = (e[..., 1:] - e[..., :-1]) / dx
    gr = (er[..., 1:] - er[..., :-1]) / dx
    return torch.where(u < 0, g, gr)


def burgers1d_upwind(uPred, uPred0, dt, dx, nu, k2=3, cn=True):
    """BurgerIntegrate.crankNicolson / backwardEuler with grad_kernels[0] == 2 (:264-302)."""
    lap = lambda t: filt1d(t, LAP_TAPS[k2], dx ** 2)
    if cn:
        return uPred0 - 0.5 * dt * ((uPred * upwind_grad(uPred, dx) + uPred0 * upwind_grad(uPred0, dx)) - nu * (lap(uPred) + lap(uPred0)))
    return uPred0 + dt * (-0.5 * (0.5 * upwind_grad(uPred ** 2, dx)) + nu * lap(uPred))


def burgers1d_cn_limiter(uPred, uPred0, dt, dx, nu, k2=3):
    """BurgerIntegrate.crankNicolsonLimiter (:304-322)."""
    lap = lambda t: filt1d(t, LAP_TAPS[k2], dx ** 2)
    return uPred0 - 0.5 * dt * ((weno_flux_grad(uPred, dx) + weno_flux_grad(uPred0, dx)) - nu * (lap(uPred) + lap(uPred0)))


def _ks_rhs(u, dx, nu, k1, k2, k4):
    g = filt1d(u ** 2, GRAD_TAPS[k1], dx)
    lap = filt1d(u, LAP_TAPS[k2], dx ** 2)
    d4 = filt1d(u, D4_TAPS[k4], dx ** 4)
    return 0.5 * g + lap + nu * d4


def ks_cn(uPred, uPred0, dt, dx, nu=1.0, k1=5, k2=5, k4=7):
    """KSIntegrate.crankNicolson (ksFiniteDifference.py:155-173); main.py:219 uses [5,5,7]."""
    return uPred0 - 0.5 * dt * (_ks_rhs(uPred, dx, nu, k1, k2, k4) + _ks_rhs(uPred0, dx, nu, k1, k2, k4))


def ks_be(uPred, uPred0, dt, dx, nu=1.0, k1=5, k2=5, k4=7):
    """KSIntegrate.backwardEuler (:139-153)."""
    return uPred0 - dt * _ks_rhs(uPred, dx, nu, k1, k2, k4)


# --------------------------------------------------------------------------------------
# ground-truth FD solver (1D-Burger-SWAG/solver/fd_burger1D.py:32-98), numpy fp64
# --------------------------------------------------------------------------------------


def fd_burgers1d(u0, dt, nsteps, dx, nu, grad_order=(1, 2), rows=None):
    """burgers1d(): RK4 (:45-52) of u_t = -u u_x + nu u_xx with the upwind gradient (:36-42: backward difference where u > 0, forward
    otherwise; first order [-1, 1, 0] / [0, -1, 1] over dx or second order [1, -4, 3, 0, 0] / [0, 0, -3, 4, -1] over 2 dx, :58-68) and the
    central second derivative of order 2 or 4 (:70-79); periodic.  u0 [N, L] float64.  Returns float32 [N, nsteps (or len(rows)), L]:
    the state after every step, stored in float32 as the reference does (:82)."""
    u = np.asarray(u0, dtype=np.float64).copy()
    r = lambda a, k: np.roll(a, -k, axis=-1)                    # a[i + k]

    def rhs(v):
        if grad_order[0] == 1:
            back, fwd = (-r(v, -1) + v) / dx, (-v + r(v, 1)) / dx
        else:
            back, fwd = (r(v, -2) - 4 * r(v, -1) + 3 * v) / (2 * dx), (-3 * v + 4 * r(v, 1) - r(v, 2)) / (2 * dx)
        if grad_order[1] == 2:
            lap = (r(v, -1) - 2 * v + r(v, 1)) / dx ** 2
        else:
            lap = (-r(v, -2) / 12 + 4 * r(v, -1) / 3 - 5 * v / 2 + 4 * r(v, 1) / 3 - r(v, 2) / 12) / dx ** 2
        return -v * np.where(v > 0, back, fwd) + nu * lap
    out = []
    keep = None if rows is None else set(int(i) for i in rows)
    for i in range(nsteps):
        k1 = rhs(u); k2 = rhs(u + k1 * dt / 2); k3 = rhs(u + k2 * dt / 2); k4 = rhs(u + k3 * dt)
        u = u + (k1 + 2 * k2 + 2 * k3 + k4) * dt / 6
        if keep is None or i in keep:
            out.append(u.astype(np.float32))
    return np.stack(out, 1)


# --------------------------------------------------------------------------------------
# loss  (bayesNN.py:29-35, 98-129)
# --------------------------------------------------------------------------------------


def loss_consts(dt_args):
    """Hyper-parameters fixed in BayesNN.__init__ (bayesNN.py:25-38)."""
    beta_mean = 1.0 / (0.2 * dt_args ** 3.0)
    return dict(beta_prior_shape=10., beta_prior_rate=10. / beta_mean, w_prior_shape=0.5, w_prior_rate=10.)


def neg_log_joint(output, target, ntrain, log_beta, feature_params, dt_args):
    """BayesNN.calc_neg_log_joint; log_beta has shape (1,), result has shape (1,)."""
    c = loss_consts(dt_args)
    log_lik = ntrain / output.size(0) * (-0.5 * log_beta.exp() * (target - output).pow(2).sum()
                                         + 0.5 * target.numel() * log_beta)
    prior_w = torch.zeros((), dtype=output.dtype)
    for p in feature_params:
        prior_w = prior_w + torch.log1p(0.5 / c['w_prior_rate'] * p.pow(2)).sum()
    prior_w = prior_w * (-(c['w_prior_shape'] + 0.5))
    prior_lb = (c['beta_prior_shape'] - 1.0) * log_beta - log_beta.exp() * c['beta_prior_rate']
    return -log_lik - prior_w - prior_lb


# --------------------------------------------------------------------------------------
# SWAG sampling / collection (swag.py:80-150, 152-194), noise supplied by the caller
# --------------------------------------------------------------------------------------


def swag_sample(means, sq_means, covs, z1s, z2s, max_models, scale=1.0, diag=False, eps=1e-30):
    """Per-parameter-tensor posterior draw.

    means/sq_means: lists of tensors (parameter shapes); covs: list of [K, numel] (raw
    parameter rows, zeros = empty slots, swag.py:183-185); z1s: list like means
    (randn_like draw, :125); z2s: list of [K,1] (normal_ draw, :139, fresh per tensor).
    """
    out = []
    for i, (m, sq) in enumerate(zip(means, sq_means)):
        var = torch.clamp(sq - m ** 2, min=eps)
        sds = scale * torch.sqrt(var) * z1s[i]
        if diag:
            out.append(m + sds)
            continue
        c = covs[i]
        d = torch.where(c == 0, c, c - m.reshape(1, -1))
        cs = (scale / ((max_models - 1) ** 0.5)) * d.t().matmul(z2s[i]).view_as(m)
        out.append(m + np.sqrt(0.5) * (sds + cs))
    return out


def swag_collect(mean, sq_mean, cov, param, n, max_models):
    """One SwagNN.collect update for one tensor (swag.py:168-187)."""
    mean = mean * n / (n + 1.0) + param / (n + 1.0)
    sq_mean = sq_mean * n / (n + 1.0) + param ** 2 / (n + 1.0)
    # NOTE [-(K-1):] keeps every row when K == 1 ([-0:]), reproduced literally.
    cov = torch.cat((cov[-(max_models - 1):], param.reshape(1, -1)), dim=0)
    return mean, sq_mean, cov


# --------------------------------------------------------------------------------------
# auto-regressive rollout (main.py test()/testSample() loops)
# --------------------------------------------------------------------------------------


def rollout(system, step, ic, nic, tsteps):
    """AR rollout with the reference's history handling.

    burgers2d (2D-Burgers-SWAG/main.py:112-135): ic [B,2,H,W] repeated nic times; feed the
      last 2*nic channels; next history = cat(last 2*(nic-1), uPred).
    burgers1d (1D-Burger-SWAG/main.py:106-121): ic [B,C>=nic,L] (loader repeats the IC);
      feed last nic; append uPred[:,0:1].
    ks1d (1D-KS-SWAG/main.py:117-131): same as burgers1d with nic channels.
    Returns [B, tsteps+1, C_out, ...] with the initial state at index 0.
    """
    outs = []
    if system == 'burgers2d':
        inp = ic.repeat(1, nic, 1, 1)
        outs.append(ic)
        for _ in range(tsteps):
            up = step(inp[:, -2 * nic:])
            outs.append(up)
            inp = torch.cat([inp[:, -2 * (nic - 1):], up], 1)
    else:
        inp = ic
        outs.append(ic[:, :1])
        for _ in range(tsteps):
            up = step(inp[:, -nic:])
            outs.append(up[:, :1])
            inp = torch.cat([inp[:, -(nic - 1):], up[:, :1]], 1)
    return torch.stack(outs, 1)


def predictive_moments_2d(u_samples, betas):
    """2D-Burgers-SWAG/post/plotBARContour.py:85-104: u_samples [S, ...]; var uses mean beta."""
    mean = u_samples.mean(0)
    var = 1.0 / betas.mean() + (u_samples ** 2).mean(0) - mean ** 2
    return mean, var


def predictive_moments_1d(u_samples, betas):
    """1D-Burger-SWAG/post/plotBARContour.py:123,136-139: mean_s(1/beta_s + y_s^2) - mean^2."""
    mean = u_samples.mean(0)
    shp = (-1,) + (1,) * (u_samples.dim() - 1)
    var = (1.0 / betas.reshape(shp) + u_samples ** 2).mean(0) - mean ** 2
    return mean, var


# --------------------------------------------------------------------------------------
# synthetic initial conditions (the benchmark input spec)
# --------------------------------------------------------------------------------------


def ic_burgers2d(ncells, indices, order=4):
    """InitPeriodicCond2d.__getitem__ (2D-Burgers-SWAG/utils/burgerLoader2D.py:88-112)."""
    x = np.linspace(0, 1, ncells + 1)[:-1]
    xx, yy = np.meshgrid(x, x)
    aa, bb = np.meshgrid(np.arange(-order, order + 1), np.arange(-order, order + 1))
    k = np.stack((aa.flatten(), bb.flatten()), 1)
    kxly = (np.outer(k[:, 0], xx.flatten()) + np.outer(k[:, 1], yy.flatten())) * 2 * np.pi
    out = []
    for index in indices:
        np.random.seed(index + 100000)
        lam = np.random.randn(2, 2, (2 * order + 1) ** 2)
        c = -1 + np.random.rand(2) * 2
        f = np.dot(lam[0], np.cos(kxly)) + np.dot(lam[1], np.sin(kxly))
        f = 2 * f / np.amax(np.abs(f), axis=1, keepdims=True) + c[:, None]
        out.append(torch.FloatTensor(f).reshape(-1, ncells, ncells))
    return torch.stack(out, 0)


def error_metrics(mean_pred, target, nel, test_every=1):
    """testBayesianMSE's error curves (2D-Burgers-SWAG/post/plotMSE.py:131-137): mean_pred [N, T+1, 2, H, W], target [N, Tp, 2, H, W]."""
    n, tp = target.shape[0], target.shape[1]
    u_out = mean_pred.reshape(n, mean_pred.shape[1], -1)[:, ::test_every][:, :tp]
    u_target = target.reshape(n, tp, -1)
    mse_error = torch.mean(torch.pow(u_out.double() - u_target.double(), 2), dim=-1)
    ese_error = torch.pow(torch.sum(torch.pow(u_out, 2) / 2.0, dim=-1) / (nel ** 2) - torch.sum(torch.pow(u_target, 2) / 2.0, dim=-1) / (nel ** 2), 2)
    return mse_error, ese_error


def mse(a, b):
    return float(((a - b) ** 2).mean())


def rel_err(a, b):
    """max|a-b| / max|b| -- the relative error every parity test in tests/ uses."""
    a = torch.as_tensor(a, dtype=torch.float64).detach()
    b = torch.as_tensor(b, dtype=torch.float64).detach()
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)
