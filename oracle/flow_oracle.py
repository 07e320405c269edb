"""CPU restatement (oracle) of the RAFT pieces on the hot path: all-pairs correlation pyramid,
windowed lookup, BasicUpdateBlock, convex upsampling.  TEST INFRASTRUCTURE ONLY (see atlas_oracle.py).

Pinned bit-exactly against the reference's own modules (seeded random weights; the reference ships no
checkpoints) by tests/golden/make_golden_nets.py.  State-dict keys are the reference's.
Reference: src/models/stage_1/core/{corr.py,update.py,raft.py,utils/utils.py}.
"""
import torch
import torch.nn.functional as F


def corr_pyramid(fmap1, fmap2, num_levels=4):
    """CorrBlock.__init__ (core/corr.py:16-31, 56-64): fmaps (B, C, H, W) -> list of (B*H*W, 1, h, w)."""
    b, c, h, w = fmap1.shape
    corr = torch.matmul(fmap1.view(b, c, h * w).transpose(1, 2), fmap2.view(b, c, h * w))
    corr = corr.view(b, h, w, 1, h, w) / torch.sqrt(torch.tensor(c).float())
    corr = corr.reshape(b * h * w, 1, h, w)
    out = [corr]
    for _ in range(num_levels - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        out.append(corr)
    return out


def bilinear_sampler(img, coords):
    """core/utils/utils.py:57-71."""
    H, W = img.shape[-2:]
    xg, yg = coords.split([1, 1], dim=-1)
    grid = torch.cat([2 * xg / (W - 1) - 1, 2 * yg / (H - 1) - 1], dim=-1)
    return F.grid_sample(img, grid, align_corners=True)


def corr_lookup(pyramid, coords, radius=4):
    """CorrBlock.__call__ (core/corr.py:33-54): coords (B, 2, H, W) -> (B, L*(2r+1)^2, H, W)."""
    r = radius
    coords = coords.permute(0, 2, 3, 1)
    b, h, w, _ = coords.shape
    outs = []
    for i, corr in enumerate(pyramid):
        d = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
        delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1)        # (dy, dx) added to (x, y): :41-47
        cl = coords.reshape(b * h * w, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
        outs.append(bilinear_sampler(corr, cl).view(b, h, w, -1))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def _conv(sd, name, x, padding=0, stride=1):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def update_block(sd, net, inp, corr, flow):
    """BasicUpdateBlock.forward (core/update.py:127-136) with BasicMotionEncoder (:79-99), SepConvGRU
    (:33-60), FlowHead (:6-14) and the mask head (:120-123); keys as in `update_block.state_dict()`."""
    cor = F.relu(_conv(sd, "encoder.convc1", corr))
    cor = F.relu(_conv(sd, "encoder.convc2", cor, 1))
    flo = F.relu(_conv(sd, "encoder.convf1", flow, 3))
    flo = F.relu(_conv(sd, "encoder.convf2", flo, 1))
    out = F.relu(_conv(sd, "encoder.conv", torch.cat([cor, flo], 1), 1))
    x = torch.cat([inp, torch.cat([out, flow], 1)], 1)
    h = net
    for tag, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([h, x], 1)
        z = torch.sigmoid(_conv(sd, "gru.convz" + tag, hx, pad))
        r = torch.sigmoid(_conv(sd, "gru.convr" + tag, hx, pad))
        q = torch.tanh(_conv(sd, "gru.convq" + tag, torch.cat([r * h, x], 1), pad))
        h = (1 - z) * h + z * q
    delta = _conv(sd, "flow_head.conv2", F.relu(_conv(sd, "flow_head.conv1", h, 1)), 1)
    mask = .25 * _conv(sd, "mask.2", F.relu(_conv(sd, "mask.0", h, 1)))
    return h, mask, delta


def convex_upsample(flow, mask):
    """RAFT.upsample_flow (core/raft.py:76-87)."""
    n, _, h, w = flow.shape
    mask = torch.softmax(mask.view(n, 1, 9, 8, 8, h, w), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(n, 2, 8 * h, 8 * w)
