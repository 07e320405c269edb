"""Calibration printout for the parity bounds in tests/test_tc_fullsize_gpu.py / the trajectory tests: error of
the tensor-core path and of the fp32 CPU oracle against a float64 evaluation of the oracle, per tensor, at the
BASELINE configuration (80x432x768, B = 10 000), plus 5-step trajectory statistics (parameters and Adam moments).

    python tests/perf/parity_diag.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "all-in-one-deflicker_b200"))
from b200 import _native as N, atlas as A, synth          # noqa: E402
from oracle import atlas_oracle as O                         # noqa: E402

DEV = "cuda"


def params():
    z = np.load(os.path.join(ROOT, "tests", "golden", "params_seed1234.npz"))
    return ([torch.from_numpy(z[f"map{i}"]) for i in range(12)], [torch.from_numpy(z[f"atl{i}"]) for i in range(16)])


def main():
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    H, W, T, B = 432, 768, 80, 10000
    data = synth.throughput_set(H, W, T, seed=0)
    inds = torch.randint(H * W * T, (B, 1), generator=torch.Generator().manual_seed(1))
    mp, ap = params()
    vid = A.DeviceVideo.from_reference_layout(data, DEV)
    out = {}
    for wg, it in ((True, 0), (False, 6000)):
        tr = A.AtlasTrainer(vid, {"samples_batch": B}, precision=N.PREC_TC, device=DEV)
        tr.load_state(O.state_dict_of(mp), O.state_dict_of(ap))
        tr.indices.copy_(inds.reshape(-1)); tr.loss_grad(wg); torch.cuda.synchronize()
        g_tc = tr.grads.clone(); l_tc = tr.losses.cpu().numpy().copy()
        video = O.Video(**data)
        m32 = [p.clone().requires_grad_(True) for p in mp]; a32 = [p.clone().requires_grad_(True) for p in ap]
        t32 = O.iteration_losses(video, m32, a32, inds, it); t32["total"].backward()
        video64 = O.Video(**{k: v.double() if v.dtype == torch.float32 else v for k, v in data.items()})
        m64 = [p.double().requires_grad_(True) for p in mp]; a64 = [p.double().requires_grad_(True) for p in ap]
        t64 = O.iteration_losses(video64, m64, a64, inds, it); t64["total"].backward()
        rows = []
        i = 0
        for which in ("mapping", "atlas"):
            views = tr._views(g_tc, which)
            for k in views:
                truth = (m64 + a64)[i].grad
                e_tc = (views[k].cpu().double() - truth).abs()
                e_32 = ((m32 + a32)[i].grad.double() - truth).abs()
                i += 1
                rows.append(dict(t=f"{which}.{k}", gmax=float(truth.abs().max()), gfro=float(truth.norm()),
                                 tc_max=float(e_tc.max()), tc_fro=float(e_tc.norm()), tc_q90=float(torch.quantile(e_tc.flatten()[::max(1, e_tc.numel() // 100000)], 0.9)),
                                 o32_max=float(e_32.max()), o32_fro=float(e_32.norm()), o32_q90=float(torch.quantile(e_32.flatten()[::max(1, e_32.numel() // 100000)], 0.9))))
        out["with_global" if wg else "without"] = dict(
            losses_tc=[float(x) for x in l_tc], losses_o32=[float(t32[k]) if k in t32 else 0.0 for k in ("total", "rgb", "gradient", "rigidity", "rigidity_global", "flow")],
            losses_o64=[float(t64[k]) if k in t64 else 0.0 for k in ("total", "rgb", "gradient", "rigidity", "rigidity_global", "flow")], grads=rows)
        for r in rows:
            print(f"{'G' if wg else 'N'} {r['t']:26s} gmax {r['gmax']:.3e} | TC max {r['tc_max']/r['gmax']:.2e} fro {r['tc_fro']/r['gfro']:.2e} q90 {r['tc_q90']/r['gmax']:.2e}"
                  f" | oracle32 max {r['o32_max']/r['gmax']:.2e} fro {r['o32_fro']/r['gfro']:.2e} q90 {r['o32_q90']/r['gmax']:.2e}")
        del tr
    # ---- well-conditioned scenario: quality set, mapping pre-trained for 2 sweeps (J ~ 0.8 I), full size
    dq = synth.quality_set(H, W, T, seed=0); dq.pop("clean")
    vq = A.DeviceVideo.from_reference_layout(dq, DEV)
    trq = A.AtlasTrainer(vq, {"samples_batch": B}, precision=N.PREC_TC, device=DEV)
    trq.load_state(O.state_dict_of(mp), O.state_dict_of(ap))
    torch.manual_seed(11)
    trq.pretrain(T, H, W, 2)
    mq = [v.detach().cpu().clone() for v in trq.param_views("mapping").values()]
    aq = [v.detach().cpu().clone() for v in trq.param_views("atlas").values()]
    for wg, it in ((True, 0), (False, 6000)):
        trq.indices.copy_(inds.reshape(-1)); trq.loss_grad(wg); torch.cuda.synchronize()
        g_tc = trq.grads.clone(); l_tc = trq.losses.cpu().numpy().copy()
        video = O.Video(**dq)
        m32 = [p.clone().requires_grad_(True) for p in mq]; a32 = [p.clone().requires_grad_(True) for p in aq]
        t32 = O.iteration_losses(video, m32, a32, inds, it); t32["total"].backward()
        video64 = O.Video(**{k: v.double() if v.dtype == torch.float32 else v for k, v in dq.items()})
        m64 = [p.double().requires_grad_(True) for p in mq]; a64 = [p.double().requires_grad_(True) for p in aq]
        t64 = O.iteration_losses(video64, m64, a64, inds, it); t64["total"].backward()
        print("Q losses tc", [float(x) for x in l_tc[:6]], "o32", [float(t32[k]) if k in t32 else 0.0 for k in ("total", "rgb", "gradient", "rigidity", "rigidity_global", "flow")])
        i = 0
        for which in ("mapping", "atlas"):
            views = trq._views(g_tc, which)
            for k in views:
                truth = (m64 + a64)[i].grad
                e_tc = (views[k].cpu().double() - truth).abs(); e_32 = ((m32 + a32)[i].grad.double() - truth).abs(); i += 1
                print(f"Q{'G' if wg else 'N'} {which}.{k:18s} gmax {float(truth.abs().max()):.3e} | TC max {float(e_tc.max()/truth.abs().max()):.2e} fro {float(e_tc.norm()/truth.norm()):.2e}"
                      f" | oracle32 max {float(e_32.max()/truth.abs().max()):.2e} fro {float(e_32.norm()/truth.norm()):.2e}")
    # trajectory on the well-conditioned state
    video = O.Video(**dq)
    for prec in (N.PREC_FP32, N.PREC_TC):
        tr = A.AtlasTrainer(vq, {"samples_batch": B}, precision=prec, device=DEV)
        tr.load_state(O.state_dict_of(mq), O.state_dict_of(aq))
        rm = [p.clone().requires_grad_(True) for p in mq]; ra = [p.clone().requires_grad_(True) for p in aq]
        opt = O.make_optimizer(rm, ra)
        gi = torch.Generator().manual_seed(21)
        for it in range(5):
            ii = torch.randint(H * W * T, (B, 1), generator=gi)
            ref = O.train_iteration(video, rm, ra, opt, ii, it)
            got = tr.step_host(ii, it)
            print("Qtraj", "tc" if prec == N.PREC_TC else "fp32", it, "loss rel diff", abs(got[0] - ref["total"]) / abs(ref["total"]))
        for which, ref_p in (("mapping", rm), ("atlas", ra)):
            m_views = tr._views(tr.exp_avg, which)
            for (k, pv), r in zip(tr.param_views(which).items(), ref_p):
                dd = (pv.cpu() - r.detach()).abs(); st = opt.state[r]
                em = (m_views[k].cpu() - st["exp_avg"]).abs()
                print(f"Qtraj {'tc' if prec == N.PREC_TC else 'fp32'} {which}.{k:18s} p_max={float(dd.max()):.2e} p_mean={float(dd.mean()):.2e} p_gt2e5={float((dd > 2e-5).float().mean()):.2e} "
                      f"p_gt1e4={float((dd > 1e-4).float().mean()):.2e} m_rel_fro={float(em.norm() / st['exp_avg'].norm()):.2e} m_rel_max={float(em.max() / st['exp_avg'].abs().max()):.2e}")
        del tr
    del trq, vq
    # ---- trajectory: 5 steps, both precisions, small video (the size of the committed tests) and full size
    for (h, w, t, b) in ((24, 40, 6, 64), (H, W, T, B)):
        d = data if (h, w, t) == (H, W, T) else synth.throughput_set(h, w, t, seed=3)
        v = vid if (h, w, t) == (H, W, T) else A.DeviceVideo.from_reference_layout(d, DEV)
        video = O.Video(**d)
        for prec in (N.PREC_FP32, N.PREC_TC):
            tr = A.AtlasTrainer(v, {"samples_batch": b}, precision=prec, device=DEV)
            tr.load_state(O.state_dict_of(mp), O.state_dict_of(ap))
            rm = [p.clone().requires_grad_(True) for p in mp]; ra = [p.clone().requires_grad_(True) for p in ap]
            opt = O.make_optimizer(rm, ra)
            gi = torch.Generator().manual_seed(21)
            for it in range(5):
                ii = torch.randint(h * w * t, (b, 1), generator=gi)
                O.train_iteration(video, rm, ra, opt, ii, it)
                tr.step_host(ii, it)
            stats = []
            for which, ref_p in (("mapping", rm), ("atlas", ra)):
                m_views = tr._views(tr.exp_avg, which); v_views = tr._views(tr.exp_avg_sq, which)
                for (k, pv), r in zip(tr.param_views(which).items(), ref_p):
                    dd = (pv.cpu() - r.detach()).abs()
                    st = opt.state[r]
                    em = (m_views[k].cpu() - st["exp_avg"]).abs(); ev = (v_views[k].cpu() - st["exp_avg_sq"]).abs()
                    stats.append(dict(t=f"{which}.{k}", p_max=float(dd.max()), p_mean=float(dd.mean()), p_med=float(dd.median()),
                                      p_gt2e5=float((dd > 2e-5).float().mean()), p_gt1e4=float((dd > 1e-4).float().mean()),
                                      m_rel_max=float(em.max() / st["exp_avg"].abs().max()), m_rel_fro=float(em.norm() / st["exp_avg"].norm()),
                                      v_rel_max=float(ev.max() / st["exp_avg_sq"].abs().max())))
            key = f"traj_{h}x{w}x{t}_B{b}_{'tc' if prec == N.PREC_TC else 'fp32'}"
            out[key] = stats
            for s in stats:
                print(key, s["t"], " ".join(f"{k}={v:.2e}" for k, v in s.items() if k != "t"))
            del tr
    print(json.dumps(out))


if __name__ == "__main__":
    main()
