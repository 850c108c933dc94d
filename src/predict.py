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


def predict_with_contiguity_threshold(pred_probs, contiguity_threshold, classification_threshold):
    '''[1 - c, c] with c = 1 if the clip has >= contiguity_threshold consecutive frames whose B-line probability
    exceeds classification_threshold (reference predict.py:299-308).'''
    b_preds = (np.asarray(pred_probs)[:, 1] > classification_threshold).astype(int)
    clip_pred = int(max_contiguous_b_line_preds(b_preds) >= contiguity_threshold)
    return np.array([1 - clip_pred, clip_pred])


def highest_avg_contiguous_pred_prob(pred_probs, window_length):
    '''
    Highest mean B-line probability over any `window_length` consecutive frames (reference predict.py:365-376).
    :param pred_probs [N, C]: framewise predictions for each class (column 1 = B-lines, as in the reference); a 1-D
                              vector of B-line probabilities is accepted as an extension
    :return: np.array([1 - m, m]) like the reference; m = 0 when the clip has fewer than `window_length` frames
    '''
    p = np.asarray(pred_probs, np.float64)
    b = p[:, 1] if p.ndim == 2 else p
    m = 0.0
    if window_length > 0 and len(b) >= window_length:
        c = np.concatenate([[0.0], np.cumsum(b)])
        m = float(max(0.0, np.max((c[window_length:] - c[:-window_length]) / window_length)))
    return np.array([1. - m, m])


def compute_clip_predictions(model, preprocessing_func, clips, class_thresh=0.5, clip_algorithm='contiguous'):
    '''
    clips: list of (clip_id, frames [n,H,W,3], label or None).  Frame predictions, then the clip-level rule selected by
    CLIP_PREDICTION.ALGORITHM (reference predict.py:125-188).  The reference loops over clips in Python and rebuilds a
    dataset per clip (predict.py:148-156); here the frames of ALL clips go through the network in full batches and ONE
    kernel (abl_clip_aggregate) evaluates the rule for every clip on the device.
    '''
    import torch
    from ab_line_classifier_b200 import ops
    if clip_algorithm not in ('contiguous', 'sliding_window', 'average'):
        raise Exception('Unknown value for "clip_algorithm" argument.')
    b = CLASS_IDX_MAP['b_lines']
    clips = [c for c in clips]
    if not clips:
        return []
    counts = [len(frames) for _, frames, _ in clips]
    frames_all = np.concatenate([np.asarray(frames) for _, frames, _ in clips], 0)
    x = preprocessing_func(frames_all) if preprocessing_func is not None else np.asarray(frames_all, np.float32) / 255.
    probs = model.predict(ArrayDataset(x, None, max(cfg['TRAIN']['BATCH_SIZE'], 512)), return_device=True)
    offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)).to(probs.device)
    mean, run, win = ops.clip_aggregate(probs, offsets, b, class_thresh, cfg['CLIP_PREDICTION']['SLIDING_WINDOW'])
    mean, run, win = mean.cpu().numpy(), run.cpu().numpy(), win.cpu().numpy()
    rows = []
    for i, (clip_id, frames, label) in enumerate(clips):
        if clip_algorithm == 'contiguous':
            cp = float(run[i] >= cfg['CLIP_PREDICTION']['CONTIGUITY_THRESHOLD'])
            clip_pred_prob = np.array([1. - cp, cp])
        elif clip_algorithm == 'sliding_window':
            clip_pred_prob = np.array([1. - win[i], win[i]])
        else:
            clip_pred_prob = mean[i]
        rows.append({'clip': clip_id, 'label': label, 'pred': int(clip_pred_prob[b] >= class_thresh),
                     'score': float(clip_pred_prob[b]), 'probs': clip_pred_prob, 'max_contiguous': int(run[i]),
                     'n_frames': counts[i]})
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
