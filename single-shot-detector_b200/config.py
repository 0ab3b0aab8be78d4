"""Loads the reference's JSON configs unchanged (config_mobilenet.json / config_shufflenet.json; the reference
does json.load(open(CONFIG)) at train.py:15-16) and picks out the keys the hot path consumes."""
import json

HOT_PATH_KEYS = ['num_classes', 'score_threshold', 'iou_threshold', 'max_boxes_per_class',
                 'localization_loss_weight', 'classification_loss_weight', 'gamma', 'alpha',
                 'batch_size', 'image_height', 'image_width', 'min_dimension']


def load_config(path):
    with open(path) as f:
        params = json.load(f)
    missing = [k for k in ('num_classes', 'gamma', 'alpha') if k not in params]
    if missing:
        raise ValueError('config %s lacks %s' % (path, missing))
    return params


def postprocess_kwargs(params):
    """Arguments of SSD.get_predictions as model_fn passes them (reference model.py:57-61)."""
    return dict(score_threshold=params['score_threshold'], iou_threshold=params['iou_threshold'],
                max_boxes_per_class=params['max_boxes_per_class'])


def total_loss(losses, params):
    """The caller-side weighting of the two losses (reference model.py:86-87)."""
    return (params['localization_loss_weight'] * losses['localization_loss'] +
            params['classification_loss_weight'] * losses['classification_loss'])
