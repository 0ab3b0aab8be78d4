"""`reshape_and_concatenate` with the reference's interface (detector/box_predictor.py:67-104) -- without the copy.

The reference's box predictor ends by transposing every per-level tower output from [B, n*C, h, w] to channels-last,
reshaping it to [B, h*w*n, C] and concatenating the levels into `class_predictions` [B,A,C] (and likewise
`encoded_boxes` [B,A,4]): one full read + write of every logit, immediately before SSD.loss / SSD.get_predictions
read them again.  Here the function returns a `HeadPredictions`: a dict-like view that keeps the per-level tensors
where they are.  `SSD` recognises it and runs the head-layout kernels (csrc/head.cu, ssdk_head_*) straight on the tower
outputs; code that really indexes `['class_predictions']` / `['encoded_boxes']` gets the reference's tensors,
materialised on first use by ssdk_head_concat.  The networks (box_net, class_net, RetinaNetBoxPredictor) are the
caller's and out of scope.
"""
import ctypes

import torch

from .. import _lib
from .._tensors import ptr
from .constants import DATA_FORMAT


class HeadPredictions:
    """Per-level tower outputs, presented as the dict the reference's box predictor returns."""

    def __init__(self, encoded_boxes, class_predictions, num_classes, num_anchors_per_location, data_format=None):
        data_format = DATA_FORMAT if data_format is None else data_format
        assert data_format in ('channels_first', 'channels_last')
        assert len(encoded_boxes) == len(class_predictions) and 1 <= len(encoded_boxes) <= _lib.SSDK_MAX_LEVELS
        self.data_format = data_format
        self.num_classes = int(num_classes)
        self.num_anchors_per_location = int(num_anchors_per_location)
        self.encoded_boxes_levels = [self._prepare(t) for t in encoded_boxes]
        self.class_predictions_levels = [self._prepare(t) for t in class_predictions]
        n, C = self.num_anchors_per_location, self.num_classes
        self.batch_size = int(self.class_predictions_levels[0].shape[0])
        self.heights, self.widths = [], []
        for bx, cl in zip(self.encoded_boxes_levels, self.class_predictions_levels):
            if data_format == 'channels_first':                                   # box_predictor.py:83-86
                h, w, cb, cc = int(cl.shape[2]), int(cl.shape[3]), int(bx.shape[1]), int(cl.shape[1])
                ok = tuple(bx.shape) == (self.batch_size, cb, h, w)
            else:
                h, w, cb, cc = int(cl.shape[1]), int(cl.shape[2]), int(bx.shape[3]), int(cl.shape[3])
                ok = tuple(bx.shape) == (self.batch_size, h, w, cb)
            if not ok or cb != n * 4 or cc != n * C or int(cl.shape[0]) != self.batch_size:
                raise ValueError('head tensors do not match num_classes=%d, num_anchors_per_location=%d: boxes %s, classes %s'
                                 % (C, n, tuple(bx.shape), tuple(cl.shape)))
            self.heights.append(h)
            self.widths.append(w)
        self.num_anchors_per_feature_map = [h * w * n for h, w in zip(self.heights, self.widths)]
        self.num_anchors = sum(self.num_anchors_per_feature_map)
        self._cache = {}

    @staticmethod
    def _prepare(t):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t)
        if not t.is_cuda:
            if not torch.cuda.is_available():
                raise _lib.SsdkError('no CUDA device: this package runs only on the GPU (no CPU fallback)')
            t = t.cuda()
        if t.dtype != torch.float32:
            t = t.float()
        if t.requires_grad and torch.is_grad_enabled():
            return t if t.is_contiguous() else t.contiguous()
        return t.contiguous()

    @property
    def device(self):
        return self.class_predictions_levels[0].device

    def descriptor(self):
        """The ssdk_head struct of include/ssdk.h for these tensors."""
        d = _lib.SsdkHead()
        d.num_levels = len(self.heights)
        d.anchors_per_location = self.num_anchors_per_location
        d.data_format = _lib.SSDK_CHANNELS_FIRST if self.data_format == 'channels_first' else _lib.SSDK_CHANNELS_LAST
        for l, (h, w) in enumerate(zip(self.heights, self.widths)):
            d.height[l], d.width[l] = h, w
            d.class_predictions[l] = self.class_predictions_levels[l].data_ptr()
            d.encoded_boxes[l] = self.encoded_boxes_levels[l].data_ptr()
        return d

    def _ctx(self):
        dev = self.device
        h = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
        _lib.check(_lib.load().ssdk_ctx_set_stream(h, torch.cuda.current_stream(dev).cuda_stream))
        return h

    def _materialize_differentiable(self, key):
        """box_predictor.py:83-102 with torch ops, so that gradients flow back to the towers when user code (auxiliary losses,
        regularisers) indexes raw_predictions as in the reference."""
        levels = self.class_predictions_levels if key == 'class_predictions' else self.encoded_boxes_levels
        D = self.num_classes if key == 'class_predictions' else 4
        parts = []
        for t in levels:
            if self.data_format == 'channels_first':
                t = t.permute(0, 2, 3, 1)                                           # :84
            parts.append(t.reshape(self.batch_size, -1, D))                         # :88-97
        return torch.cat(parts, dim=1)                                              # :101-102

    def materialize(self, key):
        """The reference's concatenated tensor for `key` (box_predictor.py:88-102): computed by ssdk_head_concat, or -- when a
        level tensor requires grad -- by differentiable torch ops (the raw kernel's output would be detached from autograd)."""
        if key not in ('class_predictions', 'encoded_boxes'):
            raise KeyError(key)
        levels = self.class_predictions_levels if key == 'class_predictions' else self.encoded_boxes_levels
        if torch.is_grad_enabled() and any(t.requires_grad for t in levels):
            return self._materialize_differentiable(key)                            # not cached: tied to the current autograd graph
        if key not in self._cache:
            B, A = self.batch_size, self.num_anchors
            d = self.descriptor()
            with torch.cuda.device(self.device):
                if key == 'class_predictions':
                    out = torch.empty([B, A, self.num_classes], dtype=torch.float32, device=self.device)
                    args = (None, ptr(out))
                elif key == 'encoded_boxes':
                    out = torch.empty([B, A, 4], dtype=torch.float32, device=self.device)
                    args = (ptr(out), None)
                else:
                    raise KeyError(key)
                _lib.check(_lib.load().ssdk_head_concat(self._ctx(), ctypes.byref(d), B, self.num_classes, *args))
            self._cache[key] = out
        return self._cache[key]

    # ---- the dict the reference returns (box_predictor.py:104)
    def __getitem__(self, key):
        return self.materialize(key)

    def keys(self):
        return ['encoded_boxes', 'class_predictions']

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return 2

    def __contains__(self, key):
        return key in ('encoded_boxes', 'class_predictions')

    def items(self):
        return [(k, self[k]) for k in self.keys()]


def reshape_and_concatenate(encoded_boxes, class_predictions, num_classes, num_anchors_per_location, data_format=None,
                            lazy=True):
    """reference box_predictor.py:67-104.  encoded_boxes / class_predictions: lists with one tensor per FPN level,
    [B, n*4, h_i, w_i] / [B, n*C, h_i, w_i] for data_format 'channels_first' (the reference's constants.py:9; the
    default), [B, h_i, w_i, n*4] / [B, h_i, w_i, n*C] for 'channels_last'.
    Returns {'encoded_boxes': [B,A,4], 'class_predictions': [B,A,C]} -- as a HeadPredictions view (lazy=True) that SSD
    consumes without materialising either tensor, or as plain tensors (lazy=False)."""
    head = HeadPredictions(encoded_boxes, class_predictions, num_classes, num_anchors_per_location, data_format)
    if lazy:
        return head
    return {'encoded_boxes': head['encoded_boxes'], 'class_predictions': head['class_predictions']}
