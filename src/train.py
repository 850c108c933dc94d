"""Counterpart of reference src/train.py for the cutoffvgg16 path (single_train experiment).

Keeps ``compute_class_weight``, ``compute_output_bias``, ``define_callbacks`` and ``train_classifier``
(reference train.py:41-168) with the same argument meaning; W&B logging, sweeps and artifact download
(train.py:171-371) are the reference's control plane and out of scope.  ``python src/train.py`` trains
on seeded synthetic frames (there is no dataset in this image).
"""
import datetime
import os
import sys

import numpy as np
import yaml

sys.path.insert(0, os.getcwd())

from ab_line_classifier_b200.data import ArrayDataset                      # noqa: E402
from ab_line_classifier_b200.metrics import AUC, Precision, Recall         # noqa: E402
from ab_line_classifier_b200.model import EarlyStopping, ReduceLROnPlateau, save_model   # noqa: E402
from src.models.models import get_model                                    # noqa: E402

cfg = yaml.full_load(open(os.getcwd() + "/config.yml", 'r'))


def compute_class_weight(labels):
    """w_c = (1/C) * N / n_c  (reference train.py:41-54).  ``labels``: int class ids (the 'Class' column)."""
    histogram = np.bincount(np.asarray(labels).astype(int))
    weights = [(1.0 / len(histogram)) * sum(histogram) / histogram[i] for i in range(len(histogram))]
    return {i: weights[i] for i in range(len(histogram))}


def compute_output_bias(labels):
    """log(n_c / (N - n_c)) per class (reference train.py:71-81); returned as an array (the reference
    wraps it in a Keras Constant initializer)."""
    histogram = np.bincount(np.asarray(labels).astype(int))
    return np.log([histogram[i] / (np.sum(histogram) - histogram[i]) for i in range(histogram.shape[0])])


def define_callbacks(log_freq=None):
    """EarlyStopping + ReduceLROnPlateau exactly as configured at reference train.py:155-159 (the
    ClearMemory and W&B logger callbacks have no counterpart here)."""
    early_stopping = EarlyStopping(monitor='val_loss', verbose=1, patience=cfg['TRAIN']['PATIENCE'], mode='min',
                                   restore_best_weights=True)
    reduce_lr = ReduceLROnPlateau(monitor='val_loss', factor=0.5, patience=cfg['TRAIN']['PATIENCE'] // 2, verbose=1,
                                  min_lr=1e-8, min_delta=0.0001)
    return [early_stopping, reduce_lr]


def train_classifier(model_def, train_set, val_set, hparams, output_bias, class_weight, pretrained_path=None,
                     save_weights=False, verbose=True, epochs=None, extra_callbacks=None):
    """Reference train.py:84-146: metrics list, model_def(...), callbacks, model.fit, optional save."""
    n_classes = len(cfg['DATA']['CLASSES'])
    threshold = 1.0 / n_classes
    metrics = ['accuracy', AUC(name='auc')]
    metrics += [Precision(name='precision_' + cfg['DATA']['CLASSES'][c], thresholds=threshold, class_id=c)
                for c in range(n_classes)]
    metrics += [Recall(name='recall_' + cfg['DATA']['CLASSES'][c], thresholds=threshold, class_id=c)
                for c in range(n_classes)]
    input_shape = cfg['DATA']['IMG_DIM'] + [3]
    model = model_def(hparams, input_shape, metrics, cfg['TRAIN']['N_CLASSES'],
                      mixed_precision=cfg['TRAIN']['MIXED_PRECISION'], output_bias=output_bias,
                      weights_path=pretrained_path)
    callbacks = define_callbacks(cfg['TRAIN'].get('LOG_FREQ')) + list(extra_callbacks or [])
    model.fit(train_set, epochs=cfg['TRAIN']['EPOCHS'] if epochs is None else epochs, validation_data=val_set,
              callbacks=callbacks, verbose=verbose, class_weight=class_weight)
    if save_weights:
        model_path = cfg['PATHS']['MODEL_WEIGHTS'] + 'model' + datetime.datetime.now().strftime("%Y%m%d-%H%M%S") + '.npz'
        save_model(model.model, model_path)
    return model


def b200_options(config):
    """(synthetic train frames, synthetic val frames, augment) from the optional B200 section of config.yml.  The section
    is ours; the reference's own config.yml has none and gets these defaults (train set augmented, as train_utils.py:85
    always does)."""
    b200 = config.get('B200') or {}
    return (int(b200.get('SYNTH_TRAIN_FRAMES', 2048)), int(b200.get('SYNTH_VAL_FRAMES', 256)),
            bool(b200.get('AUGMENT', True)))


def synthetic_datasets(n_train, n_val, preprocessing_fn, seed=1234):
    """Seeded synthetic A-/B-line frames through the reference's data path (train_utils.py:84-88):
    ``Preprocessor(preprocessing_fn).prepare(train, shuffle=True, augment=True)`` / ``prepare(val)``."""
    from ab_line_classifier_b200.synth import synth_frames
    from src.data.preprocessor import Preprocessor
    xtr, ltr = synth_frames(n_train, seed=seed)
    xva, lva = synth_frames(n_val, seed=seed + 1)
    preprocessor = Preprocessor(preprocessing_fn)
    train = preprocessor.prepare((xtr, ltr), shuffle=True, augment=b200_options(cfg)[2])
    return train, preprocessor.prepare((xva, lva), shuffle=False, augment=False), ltr


def train_experiment(experiment='single_train', save_weights=False, epochs=None):
    if experiment != 'single_train':
        raise NotImplementedError("only EXPERIMENT_TYPE 'single_train' exists on the B200 path "
                                  "(cross_validation / hparam_search are W&B control plane)")
    model_def, preprocessing_fn = get_model(cfg['TRAIN']['MODEL_DEF'])
    hparams = cfg['HPARAMS'][cfg['TRAIN']['MODEL_DEF'].upper()]
    n_train, n_val, _ = b200_options(cfg)
    train_set, val_set, labels = synthetic_datasets(n_train, n_val, preprocessing_fn)
    model = train_classifier(model_def, train_set, val_set, hparams, compute_output_bias(labels),
                             compute_class_weight(labels), save_weights=save_weights, epochs=epochs)
    print(dict(zip(model.metrics_names, model.evaluate(val_set, verbose=0))))
    return model


if __name__ == '__main__':
    train_experiment(cfg['TRAIN']['EXPERIMENT_TYPE'], save_weights=True)
