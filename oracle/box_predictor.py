"""Oracle of the head-output layout (test infrastructure; see oracle/__init__.py).

Follows reference detector/box_predictor.py:67-104 (reshape_and_concatenate): per level, channels_first tensors
[B, n*4, h, w] / [B, n*C, h, w] are transposed to channels-last (:92,:97), reshaped to [B, h, w, n, 4|C] and then to
[B, h*w*n, 4|C] (:93-99), and the levels are concatenated along the anchor axis (:101-102).  Pure data movement: the
GPU results must be bit-identical.  `split_to_levels` is the inverse (not in the reference; used by tests and bench.py
to manufacture tower-shaped inputs from anchor-major synthetic tensors)."""
import numpy as np


def reshape_and_concatenate(encoded_boxes, class_predictions, num_classes, num_anchors_per_location,
                            data_format='channels_first'):
    n = num_anchors_per_location
    batch_size = encoded_boxes[0].shape[0]                                              # :73-75
    boxes_out, classes_out = [], []
    for i in range(len(encoded_boxes)):                                                 # :79
        shape = encoded_boxes[i].shape
        if data_format == 'channels_first':                                             # :83-86
            height_i, width_i = shape[2], shape[3]
        else:
            height_i, width_i = shape[1], shape[2]
        num_anchors_on_feature_map = height_i * width_i * n                             # :89
        y = encoded_boxes[i]
        y = np.transpose(y, [0, 2, 3, 1]) if data_format == 'channels_first' else y     # :92
        y = np.reshape(y, [batch_size, height_i, width_i, n, 4])                        # :93
        boxes_out.append(np.reshape(y, [batch_size, num_anchors_on_feature_map, 4]))    # :94
        y = class_predictions[i]
        y = np.transpose(y, [0, 2, 3, 1]) if data_format == 'channels_first' else y     # :97
        y = np.reshape(y, [batch_size, height_i, width_i, n, num_classes])              # :98
        classes_out.append(np.reshape(y, [batch_size, num_anchors_on_feature_map, num_classes]))  # :99
    return {'encoded_boxes': np.ascontiguousarray(np.concatenate(boxes_out, axis=1)),   # :101
            'class_predictions': np.ascontiguousarray(np.concatenate(classes_out, axis=1))}  # :102


def level_shapes(image_height, image_width, strides):
    """(h_i, w_i) of every feature map: ceil(H / stride), ceil(W / stride) (anchor_generator.py:59-60)."""
    return [(-(-int(image_height) // int(s)), -(-int(image_width) // int(s))) for s in strides]


def split_to_levels(tensor, shapes, num_anchors_per_location, data_format='channels_first'):
    """Inverse of reshape_and_concatenate for one tensor [B, A, D]: a list of [B, n*D, h, w] (or [B, h, w, n*D])."""
    B, A, D = tensor.shape
    n = num_anchors_per_location
    out, off = [], 0
    for h, w in shapes:
        cnt = h * w * n
        y = np.reshape(tensor[:, off:off + cnt], [B, h, w, n * D])
        out.append(np.ascontiguousarray(np.transpose(y, [0, 3, 1, 2]) if data_format == 'channels_first' else y))
        off += cnt
    assert off == A, (off, A)
    return out


def top_fraction_summaries(values, per_level, top_fraction=0.20):
    """Order-independent restatement of _add_scalewise_summaries (reference detector/ssd.py:135-150): per image and level,
    k = int32(ceil(float32(n) * 0.20)) (:146), the k biggest values (tf.nn.top_k, :147); returns (mean of them [B,L],
    the smallest of them [B,L], mean over the batch and k of them [L] = mean of the histogrammed vector, :148-150)."""
    values = np.asarray(values, np.float32)
    B = values.shape[0]
    L = len(per_level)
    mean = np.zeros([B, L], np.float32)
    kth = np.zeros([B, L], np.float32)
    index = 0
    for i, n in enumerate(per_level):
        k = int(np.ceil(np.float32(n) * np.float32(top_fraction)))
        if n > 0 and k > 0:
            big = -np.sort(-values[:, index:index + n], axis=1)[:, :k]                  # top_k
            mean[:, i] = big.astype(np.float64).mean(axis=1)
            kth[:, i] = big[:, -1]
        index += n
    return mean, kth, mean.astype(np.float64).mean(axis=0)


def matches_summaries(matches, per_level):
    """_add_scalewise_matches_summaries (ssd.py:152-163) and total_mean_matches_per_image (ssd.py:129)."""
    w = (np.asarray(matches) >= 0).astype(np.float32)                                   # ssd.py:89
    cols, index = [], 0
    for n in per_level:
        cols.append(w[:, index:index + n].sum(axis=1))                                  # :157
        index += n
    per = np.stack(cols, axis=1)
    return per, per.mean(axis=0), np.float32(w.sum(axis=1).mean())
