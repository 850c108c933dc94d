"""Counterpart of reference src/models/model_utils.py for the cutoffvgg16 path."""
from ab_line_classifier_b200.model import initialize_with_pretrained_weights  # noqa: F401


def freeze_layers(model, freeze_cutoff, freeze_bn=True):
    '''
    Freeze all layers up to specified index in list of model's layers (reference model_utils.py:26-41)
    '''
    for i in range(len(model.layers)):
        if i <= freeze_cutoff:
            model.layers[i].trainable = False
        elif freeze_bn and (('batch' in model.layers[i].name) or ('bn' in model.layers[i].name)):
            model.layers[i].trainable = False
        else:
            model.layers[i].trainable = True
    return model
