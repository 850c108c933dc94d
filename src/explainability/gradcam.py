"""Counterpart of reference src/explainability/gradcam.py.

``GradCAMExplainer.apply_gradcam`` computes, per frame, the Grad-CAM map of the predicted class w.r.t. the
last conv layer (reference gradcam.py:68-80) with the device kernel, then the JET overlay
(gradcam.py:81-83) on the host with cv2.  Deviations, documented in DESIGN.md: heat-maps are computed per
frame for any batch size (the reference is only correct for one frame), and an all-zero map yields zeros
instead of NaN.  The Tk file dialog / matplotlib figure (gradcam.py:94-111, visualization.py) are UI.
"""
import os
import sys

import numpy as np
import yaml

sys.path.insert(0, os.getcwd())

from src.models.models import get_model      # noqa: E402
from src.predict import restore_model        # noqa: E402

cfg = yaml.full_load(open(os.getcwd() + "/config.yml", 'r'))


class GradCAMExplainer:

    def __init__(self, model=None):
        self.model = model if model is not None else restore_model(cfg['PATHS']['MODEL_TO_LOAD'])
        self.save_img_dir = cfg['PATHS']['HEATMAPS']
        self.img_dim = tuple(cfg['DATA']['IMG_DIM'])
        self.classes = cfg['DATA']['CLASSES']
        _, preprocessing_fn = get_model(cfg['TRAIN']['MODEL_DEF'])
        self.preprocessing_fn = preprocessing_fn
        layer_name = ''
        for layer in self.model.layers:
            if any('Conv' in l for l in layer._keras_api_names):
                layer_name = layer.name
        self.last_conv_layer = layer_name
        self.hm_intensity = 0.5

    def apply_gradcam(self, frames, batch_size=512, overlay=True):
        '''
        frames: uint8 RGB [N,H,W,3] at IMG_DIM.  Returns (heatmaps fp32 [N,H,W] in [0,1], probs [N,n],
        overlays uint8 BGR [N,H,W,3] or None).
        '''
        frames = np.asarray(frames)
        eng = self.model.engine
        heats, probs = [], []
        for i in range(0, len(frames), batch_size):
            x = self.preprocessing_fn(frames[i:i + batch_size])
            h, p = eng.gradcam_batch(x, out_hw=self.img_dim)
            heats.append(h.cpu().numpy())
            probs.append(p.cpu().numpy())
        heat = np.concatenate(heats, 0)
        prob = np.concatenate(probs, 0)
        overlays = None
        if overlay:
            import cv2
            overlays = np.empty(frames.shape, np.uint8)
            for i in range(len(frames)):
                orig = np.ascontiguousarray(frames[i][..., ::-1])      # cv2.imread gives BGR
                jet = cv2.applyColorMap(np.uint8(255 * heat[i]), cv2.COLORMAP_JET)
                overlays[i] = cv2.addWeighted(jet, self.hm_intensity, orig, 1.0 - self.hm_intensity, 0)
        return heat, prob, overlays


if __name__ == '__main__':
    from ab_line_classifier_b200.synth import synth_frames
    gradcam = GradCAMExplainer()
    frames, _ = synth_frames(8)
    heat, prob, _ = gradcam.apply_gradcam(frames)
    print(heat.shape, prob)
