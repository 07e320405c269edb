"""The loss functions the stage-1 scripts call, with the reference's signatures
(src/models/stage_1/loss_utils.py:134 `get_gradient_loss_single`, :227 `get_rigidity_loss`,
:299 `get_optical_flow_loss`; for the segmentation variant :173 `get_gradient_loss`, `use_alpha=True` of :299 and
:385 `get_optical_flow_alpha_loss`), for callers that keep the reference's function-level structure.

The training script of this repo does NOT go through these (it runs the whole iteration as one fused CUDA
graph, b200.atlas.AtlasTrainer); they are the drop-in boundary: same arguments (CPU video tensors in the
reference layouts, `jif` (3, B, 1) int64 on the CPU, the caller's model objects, `device`), same return
value (0-dim fp32 tensor carrying autograd).  Network evaluations go through the model objects passed in
(the repo's `IMLP` runs them in libb200deflicker.so); the loss arithmetic and its gradient run in the
library's loss-head kernels (csrc/loss_heads.cu, the same per-sample code as the fused step).  Host-side
work is what the reference also does on the host: integer index arithmetic, the gathers from the CPU video
tensors and the upload of the gathered rows.
"""
import ctypes as C

import torch

from b200 import _native as N


def _norm_rows(cols, resx, number_of_frames, shift=(0, 0, 0)):
    """(x + sx) / (resx/2) - 1, (y + sy) / (resx/2) - 1, t / (T/2) - 1 as an fp32 (rows, 3) CPU tensor —
    int64 tensor divided by a Python float, the arithmetic of loss_utils.py:137-146,230-233."""
    x, y, t = cols
    return torch.cat(((x + shift[0]) / (resx / 2) - 1, (y + shift[1]) / (resx / 2) - 1,
                      (t + shift[2]) / (number_of_frames / 2.0) - 1), dim=1)


def _dev(t, device):
    return t.to(device=device, dtype=torch.float32).contiguous()


class _GradientHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, rgb_xp, rgb_yp, dx_gt, dy_gt):
        rgb, rgb_xp, rgb_yp = rgb.contiguous().float(), rgb_xp.contiguous().float(), rgb_yp.contiguous().float()
        loss = torch.empty((), dtype=torch.float32, device=rgb.device)
        grads = [torch.empty_like(rgb) for _ in range(3)]
        N.check(N.lib().b200_gradient_loss_head(N.ptr(rgb), N.ptr(rgb_xp), N.ptr(rgb_yp), N.ptr(dx_gt), N.ptr(dy_gt),
                                                rgb.shape[0], N.ptr(loss), N.ptr(grads[0]), N.ptr(grads[1]),
                                                N.ptr(grads[2]), N.current_stream()), "b200_gradient_loss_head")
        ctx.save_for_backward(*grads)
        return loss

    @staticmethod
    def backward(ctx, g):
        a, b, c = ctx.saved_tensors
        return g * a, g * b, g * c, None, None


class _RigidityHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv, uv_p, resx, uv_mapping_scale, derivative_amount, want_all):
        uv, uv_p = uv.contiguous().float(), uv_p.contiguous().float()
        n = uv.shape[0]
        loss = torch.empty((), dtype=torch.float32, device=uv.device)
        per = torch.empty(n, dtype=torch.float32, device=uv.device) if want_all else None
        d_uv, d_uv_p = torch.empty_like(uv), torch.empty_like(uv_p)
        N.check(N.lib().b200_rigidity_loss_head(N.ptr(uv), N.ptr(uv_p), n, float(resx), float(uv_mapping_scale),
                                                float(derivative_amount), N.ptr(per), N.ptr(loss), N.ptr(d_uv),
                                                N.ptr(d_uv_p), N.current_stream()), "b200_rigidity_loss_head")
        ctx.save_for_backward(d_uv, d_uv_p)
        ctx.want_all, ctx.n = want_all, n
        return per if want_all else loss

    @staticmethod
    def backward(ctx, g):
        d_uv, d_uv_p = ctx.saved_tensors
        if ctx.want_all:
            # per-sample output: the saved gradients are those of the MEAN, i.e. d v_s / d(.) / n per row
            w = g.reshape(-1, 1) * ctx.n
            return w * d_uv, torch.cat((w, w)) * d_uv_p, None, None, None, None
        return g * d_uv, g * d_uv_p, None, None, None, None


class _FlowHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv_rel, uv_match, resx, uv_mapping_scale):
        uv_rel, uv_match = uv_rel.contiguous().float(), uv_match.contiguous().float()
        loss = torch.empty((), dtype=torch.float32, device=uv_match.device)
        d_rel, d_match = torch.zeros_like(uv_rel), torch.zeros_like(uv_match)
        N.check(N.lib().b200_flow_loss_head(N.ptr(uv_rel), N.ptr(uv_match), uv_rel.shape[0], float(resx),
                                            float(uv_mapping_scale), N.ptr(loss), N.ptr(d_rel), N.ptr(d_match),
                                            N.current_stream()), "b200_flow_loss_head")
        ctx.save_for_backward(d_rel, d_match)
        return loss

    @staticmethod
    def backward(ctx, g):
        d_rel, d_match = ctx.saved_tensors
        return g * d_rel, g * d_match, None, None


class _WeightedFlowHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv_rel, uv_match, w, resx, uv_mapping_scale):
        uv_rel, uv_match, w = uv_rel.contiguous().float(), uv_match.contiguous().float(), w.contiguous().float()
        loss = torch.empty((), dtype=torch.float32, device=uv_match.device)
        d_rel, d_match, d_w = torch.zeros_like(uv_rel), torch.zeros_like(uv_match), torch.zeros_like(w)
        N.check(N.lib().b200_flow_loss_head_weighted(N.ptr(uv_rel), N.ptr(uv_match), N.ptr(w), uv_rel.shape[0],
                                                     float(resx), float(uv_mapping_scale), N.ptr(loss), N.ptr(d_rel),
                                                     N.ptr(d_match), N.ptr(d_w), N.current_stream()),
                "b200_flow_loss_head_weighted")
        ctx.save_for_backward(d_rel, d_match, d_w)
        return loss

    @staticmethod
    def backward(ctx, g):
        d_rel, d_match, d_w = ctx.saved_tensors
        return g * d_rel, g * d_match, g * d_w, None, None


def _alpha_of(raw):
    """tanh output of the alpha network -> (0.001, 0.991), the three steps of loss_utils.py:186-192."""
    a = 0.5 * (raw + 1.0)
    a = a * 0.99
    return a + 0.001


def get_gradient_loss(video_frames_dx, video_frames_dy, jif_current, model_F_mapping1, model_F_mapping2, model_F_atlas,
                      rgb_output_foreground, device, resx, number_of_frames, model_alpha):
    """Two-layer form of Eq. 7 (reference :173-224): the shifted reconstructions are alpha composites of the two
    atlas layers.  Each network sees both offsets in one call; the composite of (rows, 3) tensors is elementwise glue,
    the loss and its gradient come from the same head as the single-layer form."""
    cols = (jif_current[0, :], jif_current[1, :], jif_current[2, :])
    n = cols[0].shape[0]
    pts = torch.cat((_norm_rows(cols, resx, number_of_frames, (1, 0, 0)),
                     _norm_rows(cols, resx, number_of_frames, (0, 1, 0)))).to(device)
    a = _alpha_of(model_alpha(pts))
    rgb1 = (model_F_atlas(model_F_mapping1(pts) * 0.5 + 0.5) + 1.0) * 0.5
    rgb2 = (model_F_atlas(model_F_mapping2(pts) * 0.5 - 0.5) + 1.0) * 0.5
    out = rgb1 * a + rgb2 * (1.0 - a)
    dx_gt = _dev(video_frames_dx[cols[1], cols[0], :, cols[2]].squeeze(1), device)
    dy_gt = _dev(video_frames_dy[cols[1], cols[0], :, cols[2]].squeeze(1), device)
    return _GradientHead.apply(rgb_output_foreground, out[:n], out[n:], dx_gt, dy_gt)


def get_gradient_loss_single(video_frames_dx, video_frames_dy, jif_current, model_F_mapping1, model_F_atlas,
                             rgb_output_foreground, device, resx, number_of_frames):
    """Eq. 7: image-gradient consistency for one mapping network (reference :134-170)."""
    cols = (jif_current[0, :], jif_current[1, :], jif_current[2, :])
    n = cols[0].shape[0]
    # rows [0, n): (x+1, y, t); rows [n, 2n): (x, y+1, t) — one pass through each network for both offsets
    pts = torch.cat((_norm_rows(cols, resx, number_of_frames, (1, 0, 0)),
                     _norm_rows(cols, resx, number_of_frames, (0, 1, 0)))).to(device)
    rgb_shift = (model_F_atlas(model_F_mapping1(pts) * 0.5 + 0.5) + 1.0) * 0.5
    dx_gt = _dev(video_frames_dx[cols[1], cols[0], :, cols[2]].squeeze(1), device)
    dy_gt = _dev(video_frames_dy[cols[1], cols[0], :, cols[2]].squeeze(1), device)
    return _GradientHead.apply(rgb_output_foreground, rgb_shift[:n], rgb_shift[n:], dx_gt, dy_gt)


def get_rigidity_loss(jif_foreground, derivative_amount, resx, number_of_frames, model_F_mapping, uv_foreground,
                      device, uv_mapping_scale=1.0, return_all=False):
    """Eq. 9: the mapping's Jacobian should be a rotation (reference :227-278)."""
    x, y, t = jif_foreground[0, :], jif_foreground[1, :], jif_foreground[2, :]
    # (x, y-d, t) for every sample, then (x-d, y, t): the order of the reference's concatenation (:230-233)
    pts = torch.cat((_norm_rows((x, y, t), resx, number_of_frames, (0, -derivative_amount, 0)),
                     _norm_rows((x, y, t), resx, number_of_frames, (-derivative_amount, 0, 0)))).to(device)
    uv_p = model_F_mapping(pts)
    return _RigidityHead.apply(uv_foreground, uv_p, resx, uv_mapping_scale, derivative_amount, bool(return_all))


def _flow_direction(jif, uv, mask, flows, resx, number_of_frames, forward, model_F_mapping, uv_mapping_scale, device,
                    alpha):
    """One direction of Eq. 11 (reference get_corresponding_flow_matches :326-356 + :303-308)."""
    x, y, t = jif[0, :].squeeze(), jif[1, :].squeeze(), jif[2, :].squeeze()
    rows, level = torch.where(mask[y, x, t, :])               # valid (sample, flow level) pairs
    step = 2 ** level
    xs, ys, ts = x[rows], y[rows], t[rows]
    fl = flows[ys, xs, :, ts, level]
    mx, my = xs + fl[:, 0], ys + fl[:, 1]
    mt = ts + step if forward else ts - step
    pts = torch.stack((mx / (resx / 2) - 1, my / (resx / 2) - 1, mt / (number_of_frames / 2) - 1)).T
    rows_d = rows.to(device)
    if rows.numel() == 0:
        return _FlowHead.apply(uv[rows_d], uv.new_zeros((0, 2)), resx, uv_mapping_scale), rows_d
    uv_match = model_F_mapping(pts.to(device=device, dtype=torch.float32))
    if alpha is not None:          # (loss * alpha[rows].squeeze()).mean(), loss_utils.py:316-318
        return _WeightedFlowHead.apply(uv[rows_d], uv_match, alpha[rows_d].reshape(-1), resx, uv_mapping_scale), rows_d
    return _FlowHead.apply(uv[rows_d], uv_match, resx, uv_mapping_scale), rows_d


def get_optical_flow_loss(jif_foreground, uv_foreground, optical_flows_reverse, optical_flows_reverse_mask, resx,
                          number_of_frames, model_F_mapping, optical_flows, optical_flows_mask, uv_mapping_scale,
                          device, use_alpha=False, alpha=1.0):
    """Eq. 11: flow-matched points map to the same atlas point (reference :299-322)."""
    a = alpha if use_alpha else None
    nxt, _ = _flow_direction(jif_foreground, uv_foreground, optical_flows_mask, optical_flows, resx, number_of_frames,
                             True, model_F_mapping, uv_mapping_scale, device, a)
    prv, _ = _flow_direction(jif_foreground, uv_foreground, optical_flows_reverse_mask, optical_flows_reverse, resx,
                             number_of_frames, False, model_F_mapping, uv_mapping_scale, device, a)
    return prv * 0.5 + nxt * 0.5


def _alpha_direction(jif, mask, flows, resx, number_of_frames, forward, model_alpha, device):
    """Alpha at the flow-matched points of one direction (get_corresponding_flow_matches with use_uv=False)."""
    x, y, t = jif[0, :].squeeze(), jif[1, :].squeeze(), jif[2, :].squeeze()
    rows, level = torch.where(mask[y, x, t, :])
    step = 2 ** level
    xs, ys, ts = x[rows], y[rows], t[rows]
    fl = flows[ys, xs, :, ts, level]
    mt = ts + step if forward else ts - step
    pts = torch.stack(((xs + fl[:, 0]) / (resx / 2) - 1, (ys + fl[:, 1]) / (resx / 2) - 1,
                       mt / (number_of_frames / 2) - 1)).T
    return _alpha_of(model_alpha(pts.to(device=device, dtype=torch.float32))), rows.to(device)


def get_optical_flow_alpha_loss(model_alpha, jif_foreground, alpha, optical_flows_reverse, optical_flows_reverse_mask,
                                resx, number_of_frames, optical_flows, optical_flows_mask, device):
    """Eq. 12: alpha of flow-matched points should agree (reference :385-408).  The network evaluations run in the
    library; the L1 mean of two (rows, 1) tensors is elementwise glue."""
    a_f, rows_f = _alpha_direction(jif_foreground, optical_flows_mask, optical_flows, resx, number_of_frames, True,
                                   model_alpha, device)
    nxt = (alpha[rows_f] - a_f).abs().mean()
    a_b, rows_b = _alpha_direction(jif_foreground, optical_flows_reverse_mask, optical_flows_reverse, resx,
                                   number_of_frames, False, model_alpha, device)
    prv = (a_b - alpha[rows_b]).abs().mean()
    return (nxt + prv) * 0.5
