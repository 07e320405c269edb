"""CPU restatement (oracle) of the stage-2 networks: the UNet neural filter and the TransformNet local
refinement (ConvLSTM with a zero state).  TEST INFRASTRUCTURE ONLY (see atlas_oracle.py).

Pinned bit-exactly against the reference modules (seeded random weights) by
tests/golden/make_golden_nets.py; state-dict keys are the reference's.
Reference: src/models/network_filter.py:8-107, src/models/network_local.py:7-188,
src/neural_filter_and_refinement.py:89-109.
"""
import torch
import torch.nn.functional as F


def _block(sd, name, prefix, x):
    x = F.relu(F.conv2d(x, sd[f"{name}.{prefix}conv1.weight"], None, padding=1))
    return F.relu(F.conv2d(x, sd[f"{name}.{prefix}conv2.weight"], None, padding=1))


def _up(sd, name, x):
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    return F.conv2d(x, sd[name + ".1.weight"], sd[name + ".1.bias"], padding=1)


def unet_forward(sd, x):
    """UNet.forward (network_filter.py:54-75)."""
    e1 = _block(sd, "encoder1", "enc1", x)
    e2 = _block(sd, "encoder2", "enc2", F.max_pool2d(e1, 2, 2))
    e3 = _block(sd, "encoder3", "enc3", F.max_pool2d(e2, 2, 2))
    e4 = _block(sd, "encoder4", "enc4", F.max_pool2d(e3, 2, 2))
    bt = _block(sd, "bottleneck", "bottleneck", F.max_pool2d(e4, 2, 2))
    d4 = _block(sd, "decoder4", "dec4", torch.cat((_up(sd, "upconv4", bt), e4), 1))
    d3 = _block(sd, "decoder3", "dec3", torch.cat((_up(sd, "upconv3", d4), e3), 1))
    d2 = _block(sd, "decoder2", "dec2", torch.cat((_up(sd, "upconv2", d3), e2), 1))
    d1 = _block(sd, "decoder1", "dec1", torch.cat((_up(sd, "upconv1", d2), e1), 1))
    return F.conv2d(d1, sd["conv.weight"], sd["conv.bias"])


def _rconv(sd, name, x, k, stride=1, upsample=None):
    """ConvLayer / UpsampleConvLayer (network_local.py:118-167): reflection pad k//2, conv; the norm
    layer is never applied (`self.norm in ["BN" or "IN"]` is `in ["BN"]`)."""
    if upsample:
        x = F.interpolate(x, scale_factor=upsample, mode="nearest")
    x = F.pad(x, (k // 2,) * 4, mode="reflect")
    return F.conv2d(x, sd[name + ".conv2d.weight"], sd.get(name + ".conv2d.bias"), stride=stride)


def transformnet_forward(sd, X, blocks=5):
    """TransformNet.forward (network_local.py:89-115) with prev_state=None; returns (Y, hidden, cell)."""
    lrelu = lambda t: F.leaky_relu(t, 0.2)
    e1a = lrelu(_rconv(sd, "conv1a", X[:, :6], 7))
    e1b = lrelu(_rconv(sd, "conv1b", X[:, 6:], 7))
    e2a = lrelu(_rconv(sd, "conv2a", e1a, 3, 2))
    e2b = lrelu(_rconv(sd, "conv2b", e1b, 3, 2))
    rb = lrelu(_rconv(sd, "conv3", torch.cat((e2a, e2b), 1), 3, 2))
    for b in range(blocks):
        t = lrelu(_rconv(sd, f"ResBlocks.{b}.conv1", rb, 3))
        rb = _rconv(sd, f"ResBlocks.{b}.conv2", t, 3) + rb
    hidden0 = torch.zeros_like(rb)
    gates = F.conv2d(torch.cat((rb, hidden0), 1), sd["convlstm.Gates.weight"], sd["convlstm.Gates.bias"], padding=1)
    i_g, r_g, o_g, c_g = gates.chunk(4, 1)
    cell = torch.sigmoid(r_g) * torch.zeros_like(rb) + torch.sigmoid(i_g) * torch.tanh(c_g)
    hidden = torch.sigmoid(o_g) * torch.tanh(cell)
    d2 = lrelu(_rconv(sd, "deconv1", hidden, 3, upsample=2))
    d1 = lrelu(_rconv(sd, "deconv2", torch.cat((d2, e2a), 1), 3, upsample=2))
    y = torch.tanh(_rconv(sd, "deconv3", torch.cat((d1, e1a), 1), 7))
    return y, hidden, cell
