"""Counterpart of reference src/predict.py for the cutoffvgg16 path.

Keeps ``restore_model``, ``predict_set``, ``compute_metrics``, the clip-aggregation helpers
(``max_contiguous_b_line_preds``, ``predict_with_contiguity_threshold``, ``highest_avg_contiguous_pred_prob``),
``compute_clip_predictions`` / ``compute_frame_predictions`` and ``clock_avg_runtime`` (reference
predict.py:39-122, 125-222, 283-308, 365-376, 425-448).  Frames arrive as uint8/float arrays instead of
JPEG paths (JPEG decode + the frame tables are the reference's dataset plumbing, out of scope).
"""
import os
import sys
import time

import numpy as np
import yaml

sys.path.insert(0, os.getcwd())

from ab_line_classifier_b200.data import ArrayDataset          # noqa: E402
from ab_line_classifier_b200.model import load_model           # noqa: E402
from src.models.models import get_model                        # noqa: E402

cfg = yaml.full_load(open(os.getcwd() + "/config.yml", 'r'))

# reference predict.py:31-32 loads this map from a pickle written by the dataset tooling
CLASS_IDX_MAP = {c: i for i, c in enumerate(cfg['DATA']['CLASSES'])}
IDX_CLASS_MAP = {v: k for k, v in CLASS_IDX_MAP.items()}


def restore_model(model_path):
    '''Restores the model from serialized weights (reference predict.py:39-51; .npz instead of .h5/.onnx).'''
    if os.path.splitext(model_path)[1] == '.onnx':
        raise NotImplementedError("ONNX models are not part of the B200 path")
    return load_model(model_path, compile=False)


def predict_set(model, preprocessing_func, frames, threshold=0.5, batch_size=None):
    '''
    Given frames (uint8/float RGB [N,H,W,3], already resized to IMG_DIM), predict each one.
    :return: (list of predicted classes, array of classwise prediction probabilities)  -- predict.py:54-87
    '''
    x = preprocessing_func(frames) if preprocessing_func is not None else np.asarray(frames, np.float32) / 255.
    ds = ArrayDataset(x, None, batch_size or cfg['TRAIN']['BATCH_SIZE'])
    p = model.predict(ds)
    test_predictions = (p[:, CLASS_IDX_MAP['b_lines']] >= threshold).astype(int)
    pred_classes = [IDX_CLASS_MAP[v] for v in list(test_predictions)]
    test_predictions = [cfg['DATA']['CLASSES'].index(c) for c in pred_classes]
    return test_predictions, p


def compute_metrics(cfg, labels, preds, probs=None):
    '''Common performance metrics (reference predict.py:89-122).'''
    from sklearn.metrics import (accuracy_score, confusion_matrix, f1_score, precision_score, recall_score,
                                 roc_auc_score)
    metrics = {}
    class_names = cfg['DATA']['CLASSES']
    labels = np.asarray(labels)
    recalls = recall_score(labels, preds, average=None)
    metrics['confusion_matrix'] = confusion_matrix(labels, preds).tolist()
    metrics['precision'] = precision_score(labels, preds, average='binary')
    metrics['recall'] = recalls[CLASS_IDX_MAP['b_lines']]
    metrics['specificity'] = recalls[CLASS_IDX_MAP['a_lines']]
    metrics['f1'] = f1_score(labels, preds, average='binary')
    metrics['accuracy'] = accuracy_score(labels, preds)
    if probs is not None:
        metrics['macro_mean_auc'] = roc_auc_score(labels, probs[:, 1], average='macro', multi_class='ovr')
        metrics['weighted_mean_auc'] = roc_auc_score(labels, probs[:, 1], average='weighted', multi_class='ovr')
        for class_name in class_names:
            classwise_labels = (labels == class_names.index(class_name)).astype(int)
            metrics[class_name + '_auc'] = roc_auc_score(classwise_labels, probs[:, class_names.index(class_name)])
    return metrics


def max_contiguous_b_line_preds(pred_series):
    '''Longest run of consecutive positive frame predictions (reference predict.py:283-297).'''
    best = cur = 0
    for p in pred_series:
        cur = cur + 1 if p == 1 else 0
        best = max(best, cur)
    return best


def predict_with_contiguity_threshold(pred_series, contiguity_threshold):
    '''Clip is positive if it has >= contiguity_threshold consecutive positive frames (predict.py:299-308).'''
    return int(max_contiguous_b_line_preds(pred_series) >= contiguity_threshold)


def highest_avg_contiguous_pred_prob(pred_probs, window):
    '''Highest mean B-line probability over any `window` consecutive frames (predict.py:365-376).'''
    pred_probs = np.asarray(pred_probs, np.float64)
    if len(pred_probs) < window:
        return float(np.mean(pred_probs)) if len(pred_probs) else 0.0
    c = np.concatenate([[0.0], np.cumsum(pred_probs)])
    return float(np.max((c[window:] - c[:-window]) / window))


def compute_clip_predictions(model, preprocessing_func, clips, class_thresh=0.5, clip_algorithm='contiguous'):
    '''
    clips: list of (clip_id, frames [n,H,W,3], label or None).  Frame predictions per clip, then the clip-level
    rule selected by CLIP_PREDICTION.ALGORITHM (reference predict.py:125-188).
    '''
    rows = []
    for clip_id, frames, label in clips:
        preds, probs = predict_set(model, preprocessing_func, frames, threshold=class_thresh)
        pb = probs[:, CLASS_IDX_MAP['b_lines']]
        if clip_algorithm == 'contiguous':
            clip_pred = predict_with_contiguity_threshold(preds, cfg['CLIP_PREDICTION']['CONTIGUITY_THRESHOLD'])
            clip_prob = float(max_contiguous_b_line_preds(preds))
        elif clip_algorithm == 'sliding_window':
            clip_prob = highest_avg_contiguous_pred_prob(pb, cfg['CLIP_PREDICTION']['SLIDING_WINDOW'])
            clip_pred = int(clip_prob >= class_thresh)
        else:
            clip_prob = float(np.mean(pb))
            clip_pred = int(clip_prob >= class_thresh)
        rows.append({'clip': clip_id, 'label': label, 'pred': clip_pred, 'score': clip_prob, 'n_frames': len(frames)})
    return rows


def compute_frame_predictions(model, preprocessing_func, frames, labels=None, class_thresh=0.5):
    '''Frame-level predictions (+ metrics when labels are given); reference predict.py:191-222.'''
    preds, probs = predict_set(model, preprocessing_func, frames, threshold=class_thresh)
    metrics = compute_metrics(cfg, labels, preds, probs) if labels is not None else None
    return preds, probs, metrics


def clock_avg_runtime(model, n_gpu_warmup_runs, n_experiment_runs):
    '''Batch-1 latency of model(x) on a random frame, as reference predict.py:425-448.'''
    x = np.random.default_rng(0).normal(size=(1, *cfg['DATA']['IMG_DIM'], 3)).astype(np.float32)
    for _ in range(n_gpu_warmup_runs):
        model(x)
    t = []
    for _ in range(n_experiment_runs):
        t0 = time.time()
        model(x)
        t.append((time.time() - t0) * 1000.0)
    print(f"Average runtime = {np.mean(t):.3f} ms, standard deviation = {np.std(t):.3f} ms")
    return float(np.mean(t)), float(np.std(t))
