"""Checkpoint + full-video render + PSNR of the reference's `evaluate_model_single`
(src/models/stage_1/evaluate.py:605-793) and of `evaluate_model` (:203-602, segmentation variant).  The dashboards / mp4 dumps / tensorboard images of the
reference are visualisation and out of scope."""
import os

import cv2
import numpy as np
import torch

from b200 import atlas as A


def evaluate_model_single(trainer, resx, resy, number_of_frames, video_frames, results_folder, iteration,
                          vid_name=None, save_checkpoint=True):
    os.makedirs(os.path.join(results_folder, '%06d' % iteration), exist_ok=True)
    os.makedirs(os.path.join(results_folder, "output"), exist_ok=True)
    if save_checkpoint:      # evaluate.py:616-622 — same file name and keys
        torch.save({'F_atlas_state_dict': {k: v.cpu() for k, v in trainer.state_dict("atlas").items()},
                    'iteration': iteration,
                    'model_F_mapping1_state_dict': {k: v.cpu() for k, v in trainer.state_dict("mapping").items()},
                    'optimizer_all_state_dict': trainer.optimizer_state_dict()},
                   '%s/checkpoint' % results_folder)
    psnrs = np.zeros((number_of_frames, 1))
    for f in range(number_of_frames):
        img, u8 = trainer.render_frame(f, int(resy), int(resx), number_of_frames, want_u8=True)
        cv2.imwrite(os.path.join(results_folder, 'output', '%05d.png' % f),
                    cv2.cvtColor(u8.cpu().numpy(), cv2.COLOR_RGB2BGR))       # evaluate.py:732-733
        psnrs[f] = A.psnr(video_frames[:, :, :, f], img.cpu())               # :740-743
    open(os.path.join(results_folder, '%06d' % iteration, "PSNR_%f" % psnrs.mean()), "w").close()   # :782
    print("PSNR: %f" % psnrs.mean())
    return float(psnrs.mean())


def evaluate_model(trainer, resx, resy, number_of_frames, video_frames, results_folder, iteration, mask_frames=None,
                   vid_name=None, save_checkpoint=True):
    """Segmentation variant (evaluate.py:203-602): checkpoint with the reference's keys (:216-233), the composite
    reconstruction (:293-335) written to `output/`, the alpha mattes to `<iteration>/alpha/`, PSNR.  Texture editing,
    atlas dumps and per-pixel loss videos are visualisation (out of scope)."""
    folder = os.path.join(results_folder, '%06d' % iteration)
    os.makedirs(os.path.join(folder, "alpha"), exist_ok=True)
    os.makedirs(os.path.join(results_folder, "output"), exist_ok=True)
    if save_checkpoint:
        cpu = lambda which: {k: v.cpu() for k, v in trainer.state_dict(which).items()}
        ck = {'F_atlas_state_dict': cpu("atlas"), 'iteration': iteration, 'model_F_mapping1_state_dict': cpu("mapping1"),
              'model_F_mapping2_state_dict': cpu("mapping2"), 'model_F_alpha_state_dict': cpu("alpha"),
              'optimizer_all_state_dict': trainer.optimizer_state_dict()}
        torch.save(ck, '%s/checkpoint' % results_folder)
        torch.save(ck, '%s/checkpoint' % folder)
    psnrs = np.zeros((number_of_frames, 1))
    for f in range(number_of_frames):
        img, alpha, u8 = trainer.render_frame(f, int(resy), int(resx), number_of_frames, want_u8=True)
        cv2.imwrite(os.path.join(results_folder, 'output', '%05d.png' % f), cv2.cvtColor(u8.cpu().numpy(), cv2.COLOR_RGB2BGR))
        cv2.imwrite(os.path.join(folder, 'alpha', '%05d.png' % f), (alpha.cpu().numpy() * 255).astype(np.uint8))
        psnrs[f] = A.psnr(video_frames[:, :, :, f], img.cpu())
    open(os.path.join(folder, "PSNR_%f" % psnrs.mean()), "w").close()
    print("PSNR: %f" % psnrs.mean())
    return float(psnrs.mean())
