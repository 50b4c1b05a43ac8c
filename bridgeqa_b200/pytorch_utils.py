"""Layer containers of the hot path -- drop-in for the reference's
`lib/pointnet2/pytorch_utils.py`.

What matters for compatibility is (1) the class names / constructor keywords the SA and
FP modules and lib/solver.py use (SharedMLP, Conv1d/2d/3d, BatchNorm1d/2d/3d, FC,
BNMomentumScheduler, set_bn_momentum_default) and (2) the parameter names they produce,
because published VoteNet / BridgeQA checkpoints are keyed by them
(`...mlp_module.layer0.conv.weight`, `...layer0.bn.bn.running_mean`, ...;
pytorch_utils.py:24-36, 73-80, 104-157).  Both are preserved.  The dense math these
containers describe (1x1 conv without bias -> BatchNorm -> ReLU) is what the fused
tcgen05 kernels compute at inference; `fold_conv_bn()` turns one block into the
(W', b') the kernel consumes.
"""
import torch
import torch.nn as nn

_BN_FOR = {nn.Conv1d: nn.BatchNorm1d, nn.Conv2d: nn.BatchNorm2d, nn.Conv3d: nn.BatchNorm3d}


class _NormWrapper(nn.Sequential):
    """A one-element Sequential so the norm's parameters are named `<prefix>bn.*`
    under the owning block's own `bn` slot (=> `bn.bn.weight`)."""

    def __init__(self, channels, norm_cls, name=""):
        super().__init__()
        norm = norm_cls(channels)
        nn.init.constant_(norm.weight, 1.0)
        nn.init.constant_(norm.bias, 0.0)
        self.add_module(name + "bn", norm)


class BatchNorm1d(_NormWrapper):
    def __init__(self, in_size, *, name=""):
        super().__init__(in_size, nn.BatchNorm1d, name)


class BatchNorm2d(_NormWrapper):
    def __init__(self, in_size, name=""):
        super().__init__(in_size, nn.BatchNorm2d, name)


class BatchNorm3d(_NormWrapper):
    def __init__(self, in_size, name=""):
        super().__init__(in_size, nn.BatchNorm3d, name)


_WRAPPED_BN = {nn.Conv1d: BatchNorm1d, nn.Conv2d: BatchNorm2d, nn.Conv3d: BatchNorm3d}


class _ConvBlock(nn.Sequential):
    """[bn -> act ->] conv [-> bn -> act]; conv has a bias only when there is no BN."""

    conv_cls = None

    def __init__(self, in_size, out_size, *, kernel_size, stride, padding,
                 activation=None, bn=False, init=nn.init.kaiming_normal_, bias=True,
                 preact=False, name=""):
        super().__init__()
        use_bias = bias and not bn
        conv = self.conv_cls(in_size, out_size, kernel_size=kernel_size, stride=stride,
                             padding=padding, bias=use_bias)
        init(conv.weight)
        if use_bias:
            nn.init.constant_(conv.bias, 0)
        norm = _WRAPPED_BN[self.conv_cls](in_size if preact else out_size) if bn else None

        def add_norm_act():
            if norm is not None:
                self.add_module(name + "bn", norm)
            if activation is not None:
                self.add_module(name + "activation", activation)

        if preact:
            add_norm_act()
        self.add_module(name + "conv", conv)
        if not preact:
            add_norm_act()

    # -- training: conv -> [BatchNorm(batch stats) + ReLU] as one streaming autograd node ------
    def _conv_bn_relu(self):
        """(conv, norm) when this block is exactly conv -> BatchNorm -> ReLU, else None."""
        mods = list(self.children())
        if (len(mods) == 3 and isinstance(mods[0], self.conv_cls) and isinstance(mods[1], _NormWrapper)
                and isinstance(mods[2], nn.ReLU)):
            return mods[0], next(iter(mods[1].children()))
        return None

    def forward(self, x):
        from . import train_fused
        plan = self._conv_bn_relu() if (self.training and x.is_cuda) else None
        if plan is not None:
            conv, norm = plan
            if train_fused.conv_norm_supported(conv, norm, x):
                return train_fused.conv_bn_relu(x, conv, norm)       # tcgen05 conv + statistics in its epilogue
            y = conv(x)
            if train_fused.norm_supported(norm, y):
                return train_fused.bn_relu(y, norm)
            return mods_tail(self, y)
        return super().forward(x)

    def forward_pooled(self, x):
        """max over the last axis of forward(x) -- for the last block of an SA layer
        (pointnet2_modules.py:259-262); the (B,C,npoint,nsample) activation is not written when
        the fused training kernel applies."""
        from . import train_fused
        plan = self._conv_bn_relu() if (self.training and x.is_cuda) else None
        if plan is not None:
            conv, norm = plan
            if (train_fused.conv_norm_supported(conv, norm, x) and x.dim() == 4
                    and x.shape[0] * conv.out_channels <= 65535
                    and bool(train_fused.N.lib().bqa_bn_relu_max_supported(int(x.shape[3])))):
                return train_fused.conv_bn_relu_max(x, conv, norm)
            y = conv(x)
            if train_fused.norm_supported(norm, y) and train_fused.max_supported(y):
                return train_fused.bn_relu_max(y, norm)
            out = train_fused.bn_relu(y, norm) if train_fused.norm_supported(norm, y) else mods_tail(self, y)
        else:
            out = super().forward(x)
        return torch.nn.functional.max_pool2d(out, kernel_size=[1, out.size(3)]).squeeze(-1)


def mods_tail(block, y):
    """the modules after the conv of a [conv, bn, act] block, applied to the conv's output"""
    for m in list(block.children())[1:]:
        y = m(y)
    return y


class Conv1d(_ConvBlock):
    conv_cls = nn.Conv1d

    def __init__(self, in_size, out_size, *, kernel_size=1, stride=1, padding=0,
                 activation=nn.ReLU(inplace=True), bn=False, init=nn.init.kaiming_normal_,
                 bias=True, preact=False, name=""):
        super().__init__(in_size, out_size, kernel_size=kernel_size, stride=stride,
                         padding=padding, activation=activation, bn=bn, init=init, bias=bias,
                         preact=preact, name=name)


class Conv2d(_ConvBlock):
    conv_cls = nn.Conv2d

    def __init__(self, in_size, out_size, *, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                 activation=nn.ReLU(inplace=True), bn=False, init=nn.init.kaiming_normal_,
                 bias=True, preact=False, name=""):
        super().__init__(in_size, out_size, kernel_size=kernel_size, stride=stride,
                         padding=padding, activation=activation, bn=bn, init=init, bias=bias,
                         preact=preact, name=name)


class Conv3d(_ConvBlock):
    conv_cls = nn.Conv3d

    def __init__(self, in_size, out_size, *, kernel_size=(1, 1, 1), stride=(1, 1, 1),
                 padding=(0, 0, 0), activation=nn.ReLU(inplace=True), bn=False,
                 init=nn.init.kaiming_normal_, bias=True, preact=False, name=""):
        super().__init__(in_size, out_size, kernel_size=kernel_size, stride=stride,
                         padding=padding, activation=activation, bn=bn, init=init, bias=bias,
                         preact=preact, name=name)


class SharedMLP(nn.Sequential):
    """Stack of 1x1 Conv2d blocks named `layer{i}` (pytorch_utils.py:11-36).  The blocks
    share one activation instance, as in the reference (default argument evaluated once)."""

    def __init__(self, args, *, bn=False, activation=nn.ReLU(inplace=True), preact=False,
                 first=False, name=""):
        super().__init__()
        for i in range(len(args) - 1):
            plain_input = first and preact and i == 0   # raw input: no bn/act in front
            self.add_module(
                name + "layer{}".format(i),
                Conv2d(args[i], args[i + 1], bn=bn and not plain_input,
                       activation=None if plain_input else activation, preact=preact))

    def forward_pooled(self, x):
        """max_pool2d over nsample of forward(x), (B,C,npoint,nsample) -> (B,C,npoint), with the
        pooling folded into the last block's BatchNorm + ReLU node when training on the GPU."""
        blocks = list(self.children())
        for blk in blocks[:-1]:
            x = blk(x)
        return blocks[-1].forward_pooled(x)


class SharedMLPv2(nn.Sequential):
    """torch-native variant (conv with bias, separate bn/relu entries; no ReLU after the
    last layer) -- pytorch_utils.py:38-70."""

    def __init__(self, args, *, bn=False, activation=nn.ReLU(inplace=True), preact=False,
                 first=False, name=""):
        super().__init__()
        last = len(args) - 2
        for i in range(len(args) - 1):
            self.add_module(name + "layer{}".format(i),
                            nn.Conv2d(args[i], args[i + 1], kernel_size=1, stride=1, padding=0, bias=True))
            if bn:
                self.add_module(name + "bn{}".format(i), nn.BatchNorm2d(args[i + 1]))
            if i != last:
                self.add_module(name + "relu{}".format(i), nn.ReLU(inplace=True))


class FC(nn.Sequential):
    def __init__(self, in_size, out_size, *, activation=nn.ReLU(inplace=True), bn=False,
                 init=None, preact=False, name=""):
        super().__init__()
        fc = nn.Linear(in_size, out_size, bias=not bn)
        if init is not None:
            init(fc.weight)
        if not bn:
            nn.init.constant_(fc.bias, 0)

        def add_norm_act(width):
            if bn:
                self.add_module(name + "bn", BatchNorm1d(width))
            if activation is not None:
                self.add_module(name + "activation", activation)

        if preact:
            add_norm_act(in_size)
        self.add_module(name + "fc", fc)
        if not preact:
            add_norm_act(out_size)


def set_bn_momentum_default(bn_momentum):
    def fn(m):
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            m.momentum = bn_momentum

    return fn


class BNMomentumScheduler(object):
    """model.apply(setter(bn_lambda(epoch))) on every step (pytorch_utils.py:299-333;
    driven by lib/solver.py:271-279)."""

    def __init__(self, model, bn_lambda, last_epoch=-1, setter=set_bn_momentum_default):
        if not isinstance(model, nn.Module):
            raise RuntimeError("Class '{}' is not a PyTorch nn Module".format(type(model).__name__))
        self.model = model
        self.setter = setter
        self.lmbd = bn_lambda
        self.step(last_epoch + 1)
        self.last_epoch = last_epoch

    def step(self, epoch=None):
        if epoch is None:
            epoch = self.last_epoch + 1
        self.last_epoch = epoch
        self.model.apply(self.setter(self.lmbd(epoch)))


# ---- helpers for the fused kernels (not in the reference) --------------------------

def fold_conv_bn(block):
    """(W', b') of one eval-mode [1x1 conv -> BN] block: W' = W * g/sqrt(var+eps),
    b' = beta + (bias - mean) * g/sqrt(var+eps).  Returns fp32 tensors (C_out, C_in), (C_out,)."""
    conv = norm = None
    for m in block.modules():
        if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Conv3d)):
            conv = m
        elif isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            norm = m
    w = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(
        conv.out_channels, device=w.device)
    if norm is not None:
        scale = norm.weight.detach().float() * torch.rsqrt(norm.running_var.float() + norm.eps)
        w = w * scale[:, None]
        b = norm.bias.detach().float() + (b - norm.running_mean.float()) * scale
    return w.contiguous(), b.contiguous()
