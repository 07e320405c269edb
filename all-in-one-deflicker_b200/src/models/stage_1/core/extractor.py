"""RAFT `BasicEncoder` with the reference's module tree / state_dict keys
(src/models/stage_1/core/extractor.py:6-57,118-192).  Convolutions run in b200_conv2d; instance norm in
b200_instance_norm; batch norm (eval mode) is folded into the preceding convolution on the fly."""
import torch
import torch.nn as nn

from b200 import nn as K


_fold_cache = {}     # id(conv) -> (versions of every source tensor, folded weight, folded bias)


def _folded(conv, norm):
    """conv followed by eval-mode BatchNorm2d == conv with scaled weights / shifted bias.  The folded tensors are
    cached per convolution and per in-place version of their sources, so `b200.nn` keeps hitting its packed
    weight-image cache (which is keyed on the weight TENSOR) instead of re-packing every forward."""
    if not isinstance(norm, nn.BatchNorm2d):
        return conv.weight, conv.bias.detach()   # the parameter itself: b200.nn caches its packed images
    src = (conv.weight, conv.bias, norm.weight, norm.bias, norm.running_mean, norm.running_var)
    key = tuple((t.data_ptr(), t._version) for t in src)
    ent = _fold_cache.get(id(conv))
    if ent is None or ent[0] != key:
        s = norm.weight.detach() / torch.sqrt(norm.running_var + norm.eps)
        w = (conv.weight.detach() * s.view(-1, 1, 1, 1)).contiguous()
        b = ((conv.bias.detach() - norm.running_mean) * s + norm.bias.detach()).contiguous()
        ent = _fold_cache[id(conv)] = (key, w, b)
    return ent[1], ent[2]


def _conv_norm(conv, norm, x, relu=True):
    w, b = _folded(conv, norm)
    inst = isinstance(norm, nn.InstanceNorm2d)
    y = K.conv2d(x, w, b, stride=conv.stride[0], pad=conv.padding, act="relu" if (relu and not inst) else "none")
    return K.instance_norm(y, norm.eps, relu) if inst else y


def _make_norm(norm_fn, planes):
    if norm_fn == 'batch':
        return nn.BatchNorm2d(planes)
    if norm_fn == 'instance':
        return nn.InstanceNorm2d(planes)
    if norm_fn == 'none':
        return nn.Sequential()
    raise NotImplementedError("group norm is not used by the RAFT configuration of the reference")


class ResidualBlock(nn.Module):
    def __init__(self, in_planes, planes, norm_fn='group', stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1, self.norm2 = _make_norm(norm_fn, planes), _make_norm(norm_fn, planes)
        if stride != 1:
            self.norm3 = _make_norm(norm_fn, planes)
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, kernel_size=1, stride=stride), self.norm3)
        else:
            self.downsample = None

    def forward(self, x):
        y = _conv_norm(self.conv1, self.norm1, x)
        y = _conv_norm(self.conv2, self.norm2, y)
        if self.downsample is not None:
            x = _conv_norm(self.downsample[0], self.norm3, x, relu=False)
        return K.add_relu(x, y)


class BasicEncoder(nn.Module):
    def __init__(self, output_dim=128, norm_fn='batch', dropout=0.0):
        super().__init__()
        self.norm_fn = norm_fn
        self.norm1 = _make_norm(norm_fn, 64)
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        self.in_planes = 64
        self.layer1 = self._make_layer(64, stride=1)
        self.layer2 = self._make_layer(96, stride=2)
        self.layer3 = self._make_layer(128, stride=2)
        self.conv2 = nn.Conv2d(128, output_dim, kernel_size=1)
        self.dropout = None
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')

    def _make_layer(self, dim, stride=1):
        layers = (ResidualBlock(self.in_planes, dim, self.norm_fn, stride=stride),
                  ResidualBlock(dim, dim, self.norm_fn, stride=1))
        self.in_planes = dim
        return nn.Sequential(*layers)

    @torch.no_grad()
    def forward(self, x):
        is_list = isinstance(x, (tuple, list))
        if is_list:
            batch_dim = x[0].shape[0]
            x = torch.cat(x, dim=0)
        x = _conv_norm(self.conv1, self.norm1, x.float().contiguous())
        for layer in (self.layer1, self.layer2, self.layer3):
            for blk in layer:
                x = blk(x)
        x = K.conv2d(x, self.conv2.weight, self.conv2.bias.detach())
        if is_list:
            x = torch.split(x, [batch_dim, batch_dim], dim=0)
        return x
