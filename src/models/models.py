"""Drop-in for the reference's src/models/models.py on the cutoffvgg16 path.

``get_model(name) -> (model_def, preprocessing_fn)`` keeps the registry contract of
reference src/models/models.py:20-48.  Only 'cutoffvgg16' is accelerated; the other MODEL_DEFs
(ImageNet-pretrained Keras backbones) are out of scope and their model_def raises on use.
"""
import numpy as np

from ab_line_classifier_b200.model import CutoffVGG16  # noqa: F401
from src.models.model_utils import *  # noqa: F401,F403

_VGG_MEAN_BGR = np.array([103.939, 116.779, 123.68], dtype=np.float32)


def vgg16_preprocess(x):
    """keras.applications.vgg16.preprocess_input ('caffe' mode), bound at reference models.py:13:
    RGB -> BGR, subtract the ImageNet BGR means, no scaling."""
    x = np.asarray(x, dtype=np.float32)[..., ::-1]
    return np.ascontiguousarray(x - _VGG_MEAN_BGR)


def _tf_preprocess(x):
    """mobilenet_v2 / xception / resnet_v2 'tf' mode: scale to [-1, 1]."""
    return np.asarray(x, dtype=np.float32) / 127.5 - 1.0


def _not_accelerated(name):
    def model_def(*args, **kwargs):
        raise NotImplementedError(
            f"MODEL_DEF '{name}' is not part of the B200-native hot path; only 'cutoffvgg16' is implemented "
            "(SURVEY.md section 8: the other backbones are out of scope)")
    model_def.__name__ = name
    return model_def


def get_model(model_name):
    '''
    Return the model definition and associated preprocessing function as specified in the config file
    :model_name: String identifier for a model architecture
    :return: (model definition function, preprocessing function)
    '''
    if model_name == 'cutoffvgg16':
        return CutoffVGG16, vgg16_preprocess
    if model_name == 'vgg16':
        return _not_accelerated('vgg16'), vgg16_preprocess
    if model_name == 'efficientnetb7':
        return _not_accelerated('efficientnetb7'), (lambda x: np.asarray(x, dtype=np.float32))
    if model_name in ('mobilenetv2', 'xception', 'custom_resnetv2'):
        return _not_accelerated(model_name), _tf_preprocess
    # the reference silently falls back to cnn0 for unknown names (models.py:46-47)
    return _not_accelerated('cnn0'), _tf_preprocess
