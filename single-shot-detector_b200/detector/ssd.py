"""SSD head with the reference's interface (detector/ssd.py): anchors + raw predictions in, losses or
detections out.  The network (feature extractor, box predictor) is NOT part of this package: any callables
producing `encoded_boxes` [B,A,4] and `class_predictions` [B,A,C] (layout of box_predictor.py:67-104) plug in."""
import ctypes

import numpy as np
import torch

from .. import _lib
from .._tensors import Call, ptr
from .box_predictor import HeadPredictions
from .constants import MIN_LEVEL, NEGATIVES_THRESHOLD, POSITIVES_THRESHOLD  # noqa: F401
from .training_target_creation import batch_training_targets
from .utils.nms import batch_coco_detections, batch_detections_by_label, batch_multiclass_non_max_suppression


_UPSTREAM_CACHE = {}


def _upstream_tensor(call, upstream, device):
    """(w_loc, w_cls) as a float32 CUDA tensor [2].  Python pairs are uploaded once per (values, device) and cached, so that a
    training step with constant loss weights (model.py:86-87) issues no host-to-device copy and can be captured in a CUDA graph."""
    if isinstance(upstream, torch.Tensor):
        return call.tensor(upstream, torch.float32, (2,))
    key = (float(upstream[0]), float(upstream[1]), str(device))
    t = _UPSTREAM_CACHE.get(key)
    if t is None:
        t = _UPSTREAM_CACHE[key] = torch.tensor(key[:2], dtype=torch.float32, device=device)
    return t


class _ShapeOnly:
    def __init__(self, shape):
        self.shape = shape


class _SSDLossFunction(torch.autograd.Function):
    """SSD.loss as a differentiable torch op: forward = targets + fused loss, backward = ssdk_ssd_loss_backward."""

    @staticmethod
    def forward(ctx, logits, codes, head, groundtruth, params):
        losses = head._loss_forward(groundtruth, params, keep_targets=True)
        ctx.head = head
        ctx.saved = head._saved
        ctx.save_for_backward(logits, codes)          # autograd's version check: an in-place edit before backward() raises
        return torch.stack([losses['localization_loss'], losses['classification_loss']])

    @staticmethod
    def backward(ctx, grad_out):
        head = ctx.head
        _ = ctx.saved_tensors                         # raises if logits / codes were modified in place since forward()
        head._saved = ctx.saved
        grads = head.loss_backward(grad_out.contiguous())
        return grads['class_predictions'], grads['encoded_boxes'], None, None, None


class _SSDHeadLossFunction(torch.autograd.Function):
    """SSD.loss on per-level head tensors as a differentiable torch op (gradients per level, in the head's layout)."""

    @staticmethod
    def forward(ctx, head, groundtruth, params, *levels):
        losses = head._loss_forward(groundtruth, params, keep_targets=True)
        ctx.head = head
        ctx.saved = head._saved
        ctx.save_for_backward(*levels)                # autograd's version check: an in-place edit before backward() raises
        return torch.stack([losses['localization_loss'], losses['classification_loss']])

    @staticmethod
    def backward(ctx, grad_out):
        head = ctx.head
        _ = ctx.saved_tensors                         # raises if a level tensor was modified in place since forward()
        head._saved = ctx.saved
        grads = head.loss_backward(grad_out.contiguous())
        return (None, None, None) + tuple(grads['class_predictions']) + tuple(grads['encoded_boxes'])


class SSD:
    def __init__(self, images, feature_extractor, anchor_generator, box_predictor, num_classes):
        """Same arguments as the reference (ssd.py:10-40).  `images`: [B, H, W, 3] tensor (only its shape is
        used here); `feature_extractor(images)` and `box_predictor(features)` are the caller's network."""
        self.num_classes = num_classes
        image_features = feature_extractor(images)
        image_height, image_width = int(images.shape[1]), int(images.shape[2])           # ssd.py:28-29
        self.raw_predictions = box_predictor(image_features)                               # ssd.py:37
        if isinstance(self.raw_predictions, HeadPredictions):          # per-level tower outputs, consumed as they are
            device = self.raw_predictions.device
        else:
            device = self.raw_predictions['class_predictions'].device \
                if isinstance(self.raw_predictions['class_predictions'], torch.Tensor) else None
        self.anchors = anchor_generator(image_height, image_width,
                                        device=device if device is not None and device.type == 'cuda' else None)  # ssd.py:31
        self.num_anchors_per_feature_map = anchor_generator.num_anchors_per_feature_map   # ssd.py:35
        self.process_group = None      # set to a torch.distributed group (or True for WORLD) to all-reduce the sums
        self.peer_all_reduce = False   # True (after parallel.connect_peers() returned True): the all-reduce of the loss sums
                                       # runs as a peer-memory NVLink kernel fused with the finalisation instead of NCCL
        self._anchors_host = None

    @classmethod
    def from_predictions(cls, image_height, image_width, raw_predictions, anchor_generator, num_classes):
        """Convenience constructor when the head outputs already exist."""
        shape = (raw_predictions['encoded_boxes'].shape[0], int(image_height), int(image_width), 3)
        fake_images = _ShapeOnly(shape)
        return cls(fake_images, lambda x: None, anchor_generator, lambda f: raw_predictions, num_classes)

    @classmethod
    def from_head_outputs(cls, image_height, image_width, encoded_boxes_levels, class_predictions_levels, anchor_generator,
                          num_classes, data_format=None):
        """Head-layout fusion: the per-level tower outputs ([B, n*4, h_i, w_i] / [B, n*C, h_i, w_i] for 'channels_first')
        are consumed directly; reshape_and_concatenate (box_predictor.py:67-104) never runs."""
        head = HeadPredictions(encoded_boxes_levels, class_predictions_levels, num_classes,
                               anchor_generator.num_anchors_per_location, data_format)
        fake_images = _ShapeOnly((head.batch_size, int(image_height), int(image_width), 3))
        return cls(fake_images, lambda x: None, anchor_generator, lambda f: head, num_classes)

    def _head(self):
        """The HeadPredictions view when the box predictor returned one (and the shapes agree with the anchors)."""
        h = self.raw_predictions
        if not isinstance(h, HeadPredictions):
            return None
        if h.num_anchors != int(self.anchors.shape[0]) or h.num_classes != int(self.num_classes):
            raise ValueError('head outputs hold %d anchors x %d classes, the anchor generator made %d anchors, num_classes=%d'
                             % (h.num_anchors, h.num_classes, int(self.anchors.shape[0]), int(self.num_classes)))
        return h

    # ------------------------------------------------------------------ host (NumPy) buffers
    def _host_mode(self):
        if isinstance(self.raw_predictions, HeadPredictions):
            return False
        return isinstance(self.raw_predictions['class_predictions'], np.ndarray)

    def _host_args(self):
        """Head outputs given as NumPy arrays are handed to the *_host C entry points as they are (zero-copy on
        the Python side; pinned memory gives full PCIe speed)."""
        logits = np.ascontiguousarray(self.raw_predictions['class_predictions'], dtype=np.float32)
        codes = np.ascontiguousarray(self.raw_predictions['encoded_boxes'], dtype=np.float32)
        if self._anchors_host is None:
            self._anchors_host = np.ascontiguousarray(self.anchors.cpu().numpy())
        return logits, codes, self._anchors_host

    def _ctx(self):
        dev = self.anchors.device
        h = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
        _lib.check(_lib.load().ssdk_ctx_set_stream(h, torch.cuda.current_stream(dev).cuda_stream))
        return h

    def _get_predictions_host(self, score_threshold, iou_threshold, max_boxes_per_class, out=None):
        logits, codes, anchors = self._host_args()
        B, A, C = logits.shape
        K = int(max_boxes_per_class)
        if out is None:
            out = {'boxes': np.empty([B, C * K, 4], np.float32), 'scores': np.empty([B, C * K], np.float32),
                   'labels': np.empty([B, C * K], np.int32), 'num_boxes': np.empty([B], np.int32)}
        _lib.check(_lib.load().ssdk_postprocess_host(
            self._ctx(), codes.ctypes.data, anchors.ctypes.data, logits.ctypes.data,
            _lib.SSDK_INPUT_LOGITS | _lib.SSDK_BOXES_ENCODED, B, A, C, float(score_threshold), float(iou_threshold), K,
            out['boxes'].ctypes.data, out['scores'].ctypes.data, out['labels'].ctypes.data, out['num_boxes'].ctypes.data))
        return out

    def _loss_host(self, groundtruth, params):
        from . import ssd as this_module
        logits, codes, anchors = self._host_args()
        B, A, C = logits.shape
        gt = np.ascontiguousarray(groundtruth['boxes'], dtype=np.float32)
        labels = np.ascontiguousarray(groundtruth['labels'], dtype=np.int32)
        num = np.ascontiguousarray(groundtruth['num_boxes'], dtype=np.int32)
        sums = np.zeros([3], np.float64)
        losses = np.zeros([2], np.float32)
        _lib.check(_lib.load().ssdk_ssd_targets_and_loss_host(
            self._ctx(), anchors.ctypes.data, logits.ctypes.data, codes.ctypes.data, gt.ctypes.data, labels.ctypes.data,
            num.ctypes.data, B, A, C, gt.shape[1], float(this_module.POSITIVES_THRESHOLD),
            float(this_module.NEGATIVES_THRESHOLD), float(params['gamma']), float(params['alpha']),
            sums.ctypes.data, losses.ctypes.data))
        return sums, losses

    # ------------------------------------------------------------------ inference (ssd.py:42-69)
    def get_predictions(self, score_threshold=0.05, iou_threshold=0.5, max_boxes_per_class=20, out=None,
                        box_scaler=None, final_score_threshold=None, phase=None):
        """Returns {'boxes' [B,N,4], 'labels' [B,N] int, 'scores' [B,N], 'num_boxes' [B]}, N = C * max_boxes_per_class.
        The sigmoid of ssd.py:60 is fused into the score-threshold pass.  Optional extensions fold in the two consumers
        that follow in the reference: `box_scaler` [B,4] (boxes /= box_scaler, model.py:67-68) and
        `final_score_threshold` (keep score > it, order preserved, inference/detector.py:54-58)."""
        if self._host_mode():
            assert box_scaler is None and final_score_threshold is None and phase is None, 'extensions need device tensors' 
            return self._get_predictions_host(score_threshold, iou_threshold, max_boxes_per_class, out)
        head = self._head()
        if head is not None:
            assert phase is None, 'split-phase post-processing: anchor-major tensors only'
            return self._get_predictions_head(head, score_threshold, iou_threshold, max_boxes_per_class, box_scaler,
                                              final_score_threshold)
        # phase='scan' enqueues only the HBM-bound score scan (returns None), phase='finish' with the same arguments the rest:
        # a caller that overlaps sub-paths can put the latency-bound NMS stages next to another sub-path's streaming kernel
        res = batch_multiclass_non_max_suppression(
            self.raw_predictions['encoded_boxes'], self.anchors, self.raw_predictions['class_predictions'],
            score_threshold=score_threshold, iou_threshold=iou_threshold,
            max_boxes_per_class=max_boxes_per_class, scores_are_logits=True,
            box_scaler=box_scaler, final_score_threshold=final_score_threshold, phase=phase)
        if res is None:
            return None
        boxes, scores, classes, num = res
        return {'boxes': boxes, 'labels': classes, 'scores': scores, 'num_boxes': num}

    def _get_predictions_head(self, head, score_threshold, iou_threshold, max_boxes_per_class, box_scaler=None,
                              final_score_threshold=None, return_anchor_indices=False):
        """get_predictions straight from the per-level tower outputs (ssdk_head_detect): the threshold scan streams every
        level's class tensor in place, the NMS stages gather the few box codes they need through the head geometry."""
        call = Call(head.device)
        B, A, C, K = head.batch_size, head.num_anchors, head.num_classes, int(max_boxes_per_class)
        anchors = call.tensor(self.anchors, torch.float32, (A, 4))
        boxes, scores = call.empty([B, C * K, 4], torch.float32), call.empty([B, C * K], torch.float32)
        classes, num = call.empty([B, C * K], torch.int32), call.empty([B], torch.int32)
        aidx = call.empty([B, C * K], torch.int32) if return_anchor_indices else None
        sc = None if box_scaler is None else call.tensor(box_scaler, torch.float32, (B, 4))
        thr2 = float('-inf') if final_score_threshold is None else float(final_score_threshold)
        d = head.descriptor()
        _lib.check(_lib.load().ssdk_head_detect(
            call.ctx(), ctypes.byref(d), ptr(anchors), _lib.SSDK_INPUT_LOGITS, B, A, C, float(score_threshold),
            float(iou_threshold), K, ptr(sc), thr2, ptr(boxes), ptr(scores), ptr(classes), ptr(num), ptr(aidx)))
        self._call_pp = call
        out = {'boxes': boxes, 'labels': classes, 'scores': scores, 'num_boxes': num}
        if return_anchor_indices:
            out['anchor_indices'] = aidx
        return out

    def detect(self, score_threshold=0.1, box_scaler=None, nms_score_threshold=0.05, iou_threshold=0.5, max_boxes_per_class=20):
        """What inference/detector.py:36-60 returns for a batch of one image: (boxes [N,4], labels [N], scores [N]) with
        scores > score_threshold, in the exported graph's order (class-major, score-descending)."""
        p = self.get_predictions(nms_score_threshold, iou_threshold, max_boxes_per_class, box_scaler=box_scaler,
                                 final_score_threshold=score_threshold)
        n = int(p['num_boxes'][0])
        return p['boxes'][0, :n], p['labels'][0, :n], p['scores'][0, :n]

    def _flat_predictions(self):
        """(encoded_boxes [B,A,4], class_predictions [B,A,C]) -- materialised from the per-level tensors if need be."""
        raw = self.raw_predictions
        return raw['encoded_boxes'], raw['class_predictions']

    def detections_by_label(self, score_threshold=0.05, iou_threshold=0.5, max_boxes_per_class=20, box_scaler=None,
                            final_score_threshold=None, image_ids=None):
        """What the reference's evaluator accumulates from get_predictions (metrics.py:113-123, add_detections): for every label the
        list of (image, box, score) records over the images of the batch in order, each image's boxes in descending score.
        Returns {label: {'image': int [n], 'boxes': [n,4], 'scores': [n]}} (labels without detections are absent)."""
        codes, logits = self._flat_predictions()
        boxes, scores, image, counts = batch_detections_by_label(
            codes, self.anchors, logits, score_threshold, iou_threshold, max_boxes_per_class, scores_are_logits=True,
            box_scaler=box_scaler, final_score_threshold=final_score_threshold, image_ids=image_ids)
        cnt = counts.cpu().numpy() if isinstance(counts, torch.Tensor) else counts        # data dependent sizes: one host read
        return {int(c): {'image': image[c, :n], 'boxes': boxes[c, :n], 'scores': scores[c, :n]} for c, n in enumerate(cnt) if n > 0}

    def coco_results(self, image_sizes, image_ids=None, category_ids=None, score_threshold=0.15, nms_score_threshold=0.05,
                     iou_threshold=0.5, max_boxes_per_class=20, box_scaler=None):
        """The `results` list of inference/evaluate_on_COCO.ipynb (cell 10) for a batch: one dict per detection with score >
        score_threshold -- {"image_id", "category_id", "bbox": [x, y, w, h] (ints, pixels), "score"} -- ready for json.dump.
        image_sizes [B,2] = (height, width); category_ids [C] = the notebook's integer_to_coco_id."""
        codes, logits = self._flat_predictions()
        out = batch_coco_detections(codes, self.anchors, logits, image_sizes, nms_score_threshold, iou_threshold, max_boxes_per_class,
                                    scores_are_logits=True, box_scaler=box_scaler, final_score_threshold=score_threshold,
                                    image_ids=image_ids, category_ids=category_ids)
        _, scores, _, num, xywh, cat, img = [t.cpu().numpy() if isinstance(t, torch.Tensor) else t for t in out]
        results = []
        for b in range(len(num)):
            for i in range(int(num[b])):
                results.append({'image_id': int(img[b, i]), 'category_id': int(cat[b, i]), 'bbox': [int(v) for v in xywh[b, i]],
                                'score': float(scores[b, i])})
        return results

    # ------------------------------------------------------------------ training (ssd.py:71-133)
    def loss_sums(self, groundtruth, params, per_anchor=False, keep_targets=False):
        """Un-normalised shard sums: float64 CUDA tensor [3] = (sum loc_losses, sum cls_losses, num_matches).
        With per_anchor=True also returns the tensors the reference's summaries consume (ssd.py:125-129)."""
        from . import ssd as this_module      # thresholds are module constants, as in ssd.py:187-188
        head = self._head()
        if head is not None and not per_anchor:
            return self._loss_sums_head(head, groundtruth, params, keep_targets)
        call = Call()
        logits = call.tensor(self.raw_predictions['class_predictions'], torch.float32)
        B, A, C = logits.shape
        codes = call.tensor(self.raw_predictions['encoded_boxes'], torch.float32, (B, A, 4))
        anchors = call.tensor(self.anchors, torch.float32, (A, 4))
        gt = call.tensor(groundtruth['boxes'], torch.float32)
        G = gt.shape[1]
        labels = call.tensor(groundtruth['labels'], torch.int32, (B, G))
        num = call.tensor(groundtruth['num_boxes'], torch.int32, (B,))
        sums = call.empty([3], torch.float64)
        extra = {}
        if per_anchor or keep_targets:
            extra = {'reg_targets': call.empty([B, A, 4], torch.float32), 'cls_targets': call.empty([B, A], torch.int32),
                     'matches': call.empty([B, A], torch.int32)}
        if per_anchor:
            extra.update({'cls_losses': call.empty([B, A], torch.float32), 'loc_losses': call.empty([B, A], torch.float32)})
        _lib.check(_lib.load().ssdk_ssd_targets_and_loss(
            call.ctx(), ptr(anchors), ptr(logits), ptr(codes), ptr(gt), ptr(labels), ptr(num), B, A, C, G,
            float(this_module.POSITIVES_THRESHOLD), float(this_module.NEGATIVES_THRESHOLD),
            float(params['gamma']), float(params['alpha']), ptr(sums),
            ptr(extra.get('reg_targets')), ptr(extra.get('cls_targets')), ptr(extra.get('matches')),
            ptr(extra.get('cls_losses')), ptr(extra.get('loc_losses'))))
        self._call = call
        if keep_targets:          # what loss_backward needs: the inputs of the loss kernel and the (later all-reduced) sums
            self._saved = dict(logits=logits, codes=codes, sums=sums, gamma=float(params['gamma']), alpha=float(params['alpha']),
                               **{k: extra[k] for k in ('reg_targets', 'cls_targets', 'matches')})
        return (sums, extra) if per_anchor else sums

    def _targets_into(self, call, groundtruth, B, A, count=None):
        """ssd.py:84 for a batch: (reg_targets, cls_targets, matches) as fresh tensors of `call`; `count` (float64 [1]), when
        given, receives the number of matched anchors (ssd.py:121-122), counted inside the matching kernel."""
        from . import ssd as this_module
        anchors = call.tensor(self.anchors, torch.float32, (A, 4))
        gt = call.tensor(groundtruth['boxes'], torch.float32)
        G = gt.shape[1]
        labels = call.tensor(groundtruth['labels'], torch.int32, (B, G))
        num = call.tensor(groundtruth['num_boxes'], torch.int32, (B,))
        reg, cls_t, matches = call.empty([B, A, 4], torch.float32), call.empty([B, A], torch.int32), call.empty([B, A], torch.int32)
        if count is not None:
            _lib.check(_lib.load().ssdk_training_targets_count(
                call.ctx(), ptr(anchors), A, ptr(gt), ptr(labels), ptr(num), B, G, float(this_module.POSITIVES_THRESHOLD),
                float(this_module.NEGATIVES_THRESHOLD), ptr(reg), ptr(cls_t), ptr(matches), ptr(count)))
            return reg, cls_t, matches
        _lib.check(_lib.load().ssdk_training_targets(
            call.ctx(), ptr(anchors), A, ptr(gt), ptr(labels), ptr(num), B, G, float(this_module.POSITIVES_THRESHOLD),
            float(this_module.NEGATIVES_THRESHOLD), ptr(reg), ptr(cls_t), ptr(matches)))
        return reg, cls_t, matches

    def _loss_sums_head(self, head, groundtruth, params, keep_targets):
        """loss_sums on the per-level tower outputs: targets (ssd.py:84), then ssdk_head_ssd_loss."""
        from . import ssd as this_module
        call = Call(head.device)
        B, A, C = head.batch_size, head.num_anchors, head.num_classes
        anchors = call.tensor(self.anchors, torch.float32, (A, 4))
        gt = call.tensor(groundtruth['boxes'], torch.float32)
        G = gt.shape[1]
        labels = call.tensor(groundtruth['labels'], torch.int32, (B, G))
        num = call.tensor(groundtruth['num_boxes'], torch.int32, (B,))
        reg, cls_t, matches = call.empty([B, A, 4], torch.float32), call.empty([B, A], torch.int32), call.empty([B, A], torch.int32)
        sums = call.empty([3], torch.float64)
        d = head.descriptor()
        # targets (ssd.py:84) + losses (ssd.py:89-133): one launch of the fused training-step kernel (csrc/train_step.cu)
        _lib.check(_lib.load().ssdk_head_ssd_targets_and_loss(
            call.ctx(), ctypes.byref(d), ptr(anchors), ptr(gt), ptr(labels), ptr(num), B, A, C, G,
            float(this_module.POSITIVES_THRESHOLD), float(this_module.NEGATIVES_THRESHOLD), float(params['gamma']),
            float(params['alpha']), ptr(sums), ptr(reg), ptr(cls_t), ptr(matches)))
        self._call = call
        if keep_targets:
            self._saved = dict(head=head, sums=sums, gamma=float(params['gamma']), alpha=float(params['alpha']),
                               reg_targets=reg, cls_targets=cls_t, matches=matches)
        return sums

    def _head_grads(self, call, head):
        g_cls = [call.empty(list(t.shape), torch.float32) for t in head.class_predictions_levels]
        g_box = [call.empty(list(t.shape), torch.float32) for t in head.encoded_boxes_levels]
        gd = _lib.SsdkHeadGrads()
        for l in range(len(g_cls)):
            gd.class_predictions[l], gd.encoded_boxes[l] = g_cls[l].data_ptr(), g_box[l].data_ptr()
        return g_cls, g_box, gd

    def loss(self, groundtruth, params):
        """Returns {'localization_loss', 'classification_loss'}: two float32 scalars (0-d CUDA tensors), each
        sum / max(num_matches, 1) (ssd.py:121-133).  When `self.process_group` is set the three sums are
        all-reduced over the image shards first, so every rank returns the global losses."""
        if self._host_mode():
            sums, losses = self._loss_host(groundtruth, params)
            if self.process_group is not None:
                import torch.distributed as dist
                t = torch.from_numpy(sums).to(self.anchors.device)
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=None if self.process_group is True else self.process_group)
                sums = t.cpu().numpy()
                norm = max(sums[2], 1.0)
                losses = np.array([sums[0] / norm, sums[1] / norm], np.float32)
            self.num_matches = sums[2]
            return {'localization_loss': losses[0], 'classification_loss': losses[1]}
        head = self._head()
        if head is not None:
            levels = head.class_predictions_levels + head.encoded_boxes_levels
            if torch.is_grad_enabled() and any(t.requires_grad for t in levels):
                out = _SSDHeadLossFunction.apply(self, groundtruth, params, *levels)          # differentiable (model.py:115-118)
                return {'localization_loss': out[0], 'classification_loss': out[1]}
            return self._loss_forward(groundtruth, params, keep_targets=False)
        logits_in, codes_in = self.raw_predictions['class_predictions'], self.raw_predictions['encoded_boxes']
        if torch.is_grad_enabled() and isinstance(logits_in, torch.Tensor) and (logits_in.requires_grad or codes_in.requires_grad):
            out = _SSDLossFunction.apply(logits_in, codes_in, self, groundtruth, params)      # differentiable (model.py:115-118)
            return {'localization_loss': out[0], 'classification_loss': out[1]}
        return self._loss_forward(groundtruth, params, keep_targets=False)

    def _loss_forward(self, groundtruth, params, keep_targets):
        """ssd.py:71-133 in ONE launch (ssdk_ssd_loss_step: matching, flat pass, matched-anchor corrections, the exchange between
        the image shards over NVLink peer memory, normalisation) whenever no library collective is needed; with NCCL as the
        collective: sums -> all-reduce -> ssdk_loss_finalize."""
        if self.process_group is not None and not self.peer_all_reduce:
            sums = self.loss_sums(groundtruth, params, keep_targets=keep_targets)
            call = self._call
            out = call.empty([2], torch.float32)
            self._reduce_and_finalize(call.ctx(), sums, out)
        else:
            sums, out = self._loss_step(groundtruth, params, keep_targets)
            call = self._call
        self.num_matches = sums[2]
        if call.numpy_mode:
            o = out.cpu().numpy()
            return {'localization_loss': o[0], 'classification_loss': o[1]}
        return {'localization_loss': out[0], 'classification_loss': out[1]}

    def _loss_step(self, groundtruth, params, keep_targets):
        from . import ssd as this_module      # thresholds are module constants, as in ssd.py:187-188
        lib = _lib.load()
        flags = _lib.SSDK_STEP_ALL_REDUCE if (self.process_group is not None and self.peer_all_reduce) else 0
        head = self._head()
        call = Call(head.device) if head is not None else Call()
        if head is not None:
            B, A, C = head.batch_size, head.num_anchors, head.num_classes
        else:
            logits = call.tensor(self.raw_predictions['class_predictions'], torch.float32)
            B, A, C = logits.shape
            codes = call.tensor(self.raw_predictions['encoded_boxes'], torch.float32, (B, A, 4))
        anchors = call.tensor(self.anchors, torch.float32, (A, 4))
        gt = call.tensor(groundtruth['boxes'], torch.float32)
        G = gt.shape[1]
        labels = call.tensor(groundtruth['labels'], torch.int32, (B, G))
        num = call.tensor(groundtruth['num_boxes'], torch.int32, (B,))
        sums, out = call.empty([3], torch.float64), call.empty([2], torch.float32)
        reg = cls_t = matches = None
        if keep_targets:
            reg, cls_t, matches = call.empty([B, A, 4], torch.float32), call.empty([B, A], torch.int32), call.empty([B, A], torch.int32)
        thr = (float(this_module.POSITIVES_THRESHOLD), float(this_module.NEGATIVES_THRESHOLD))
        if head is not None:
            d = head.descriptor()
            _lib.check(lib.ssdk_head_ssd_loss_step(
                call.ctx(), ctypes.byref(d), ptr(anchors), ptr(gt), ptr(labels), ptr(num), B, A, C, G, thr[0], thr[1],
                float(params['gamma']), float(params['alpha']), flags, ptr(sums), ptr(out), ptr(reg), ptr(cls_t), ptr(matches)))
            if keep_targets:
                self._saved = dict(head=head, sums=sums, gamma=float(params['gamma']), alpha=float(params['alpha']),
                                   reg_targets=reg, cls_targets=cls_t, matches=matches)
        else:
            _lib.check(lib.ssdk_ssd_loss_step(
                call.ctx(), ptr(anchors), ptr(logits), ptr(codes), ptr(gt), ptr(labels), ptr(num), B, A, C, G, thr[0], thr[1],
                float(params['gamma']), float(params['alpha']), flags, ptr(sums), ptr(out), ptr(reg), ptr(cls_t), ptr(matches)))
            if keep_targets:
                self._saved = dict(logits=logits, codes=codes, sums=sums, gamma=float(params['gamma']), alpha=float(params['alpha']),
                                   reg_targets=reg, cls_targets=cls_t, matches=matches)
        self._call = call
        return sums, out

    def _reduce_and_finalize(self, ctx, sums, out):
        """ssd.py:121-133 across the image shards: all-reduce(sum) of (sum loc, sum cls, num_matches), then the two ratios."""
        lib = _lib.load()
        if self.process_group is not None and self.peer_all_reduce:
            _lib.check(lib.ssdk_comm_loss_finalize(ctx, ptr(sums), ptr(out)))           # one kernel: NVLink exchange + finalize
            return
        if self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=None if self.process_group is True else self.process_group)
        _lib.check(lib.ssdk_loss_finalize(ctx, ptr(sums), ptr(out)))

    def _reduce_count(self, ctx, count):
        if self.process_group is None:
            return
        if self.peer_all_reduce:
            _lib.check(_lib.load().ssdk_comm_all_reduce_sum(ctx, ptr(count), 1))
            return
        import torch.distributed as dist
        dist.all_reduce(count, op=dist.ReduceOp.SUM, group=None if self.process_group is True else self.process_group)

    def loss_backward(self, upstream=None):
        """Gradients of  upstream[0] * localization_loss + upstream[1] * classification_loss  w.r.t. the head outputs,
        for the last `loss(...)` / `loss_with_gradients(...)` call: {'class_predictions': [B,A,C], 'encoded_boxes': [B,A,4]}.
        This is what the reference obtains from TF autodiff (model.py:115-118); targets and weights are constants
        (ssd.py:197).  `upstream`: None (= 1, 1), a pair of floats (the config's loss weights, model.py:86-87) or a
        float32 CUDA tensor [2].  Uses the global (all-reduced) matched count as normaliser."""
        sv = getattr(self, '_saved', None)
        if sv is None:
            raise RuntimeError('loss_backward() needs a preceding loss_with_gradients() or differentiable loss() call')
        if 'head' in sv:
            head = sv['head']
            call = Call(head.device)
            up = None
            if upstream is not None:
                up = _upstream_tensor(call, upstream, head.device)
            g_cls, g_box, gd = self._head_grads(call, head)
            d = head.descriptor()
            _lib.check(_lib.load().ssdk_head_ssd_loss_forward_backward(
                call.ctx(), ctypes.byref(d), ptr(sv['reg_targets']), ptr(sv['cls_targets']), ptr(sv['matches']),
                head.batch_size, head.num_anchors, head.num_classes, sv['gamma'], sv['alpha'], sv['sums'].data_ptr() + 16,
                ptr(up), None, ctypes.byref(gd)))
            self._call_bw = (call, up)
            return {'class_predictions': g_cls, 'encoded_boxes': g_box}
        logits, codes = sv['logits'], sv['codes']
        B, A, C = logits.shape
        call = Call(logits.device)
        up = None
        if upstream is not None:
            up = _upstream_tensor(call, upstream, logits.device)
        g_logits = call.empty([B, A, C], torch.float32)
        g_codes = call.empty([B, A, 4], torch.float32)
        _lib.check(_lib.load().ssdk_ssd_loss_backward(
            call.ctx(), ptr(logits), ptr(codes), ptr(sv['reg_targets']), ptr(sv['cls_targets']), ptr(sv['matches']), B, A, C,
            sv['gamma'], sv['alpha'], ptr(sv['sums']), ptr(up), ptr(g_logits), ptr(g_codes)))
        self._call_bw = (call, up)
        return {'class_predictions': g_logits, 'encoded_boxes': g_codes}

    @staticmethod
    def assign_targets(anchors, groundtruth, positives_threshold=None, negatives_threshold=None):
        """Target assignment of a batch (ssd.py:84 / :165-199) WITHOUT the head outputs: it needs only the anchors and the
        ground truth, so a training loop can issue it on a side stream while the network's forward pass is still running and
        take it off the critical path.  Returns {'reg_targets' [B,A,4], 'cls_targets' [B,A], 'matches' [B,A], 'count'
        (float64 [1]: matched anchors of this shard)} for `loss_with_gradients(..., targets=...)`."""
        from . import ssd as this_module
        pos = this_module.POSITIVES_THRESHOLD if positives_threshold is None else positives_threshold
        neg = this_module.NEGATIVES_THRESHOLD if negatives_threshold is None else negatives_threshold
        call = Call()
        a = call.tensor(anchors, torch.float32, (-1, 4))
        A = a.shape[0]
        gt = call.tensor(groundtruth['boxes'], torch.float32)
        B, G = gt.shape[0], gt.shape[1]
        labels = call.tensor(groundtruth['labels'], torch.int32, (B, G))
        num = call.tensor(groundtruth['num_boxes'], torch.int32, (B,))
        reg, cls_t, matches = call.empty([B, A, 4], torch.float32), call.empty([B, A], torch.int32), call.empty([B, A], torch.int32)
        count = call.empty([1], torch.float64)
        _lib.check(_lib.load().ssdk_training_targets_count(call.ctx(), ptr(a), A, ptr(gt), ptr(labels), ptr(num), B, G, float(pos),
                                                           float(neg), ptr(reg), ptr(cls_t), ptr(matches), ptr(count)))
        return {'reg_targets': reg, 'cls_targets': cls_t, 'matches': matches, 'count': count}

    def loss_with_gradients(self, groundtruth, params, upstream=None, fused=True, targets=None):
        """One training step of the hot path without autograd: (losses, gradients).
        fused=True (default): targets -> matched count (all-reduced when `process_group` is set) -> ONE pass over the
        logits that produces the loss sums and both gradients (ssdk_ssd_loss_forward_backward) -> sums all-reduced for the
        reported losses.  fused=False: forward pass, then loss_backward (two reads of the logits).
        targets: the dict returned by `SSD.assign_targets` (computed earlier, e.g. during the network's forward pass); the
        step then consists of the streaming pass alone.  `groundtruth` is not read in that case."""
        if not fused:
            losses = self._loss_forward(groundtruth, params, keep_targets=True)
            return losses, self.loss_backward(upstream)
        head = self._head()
        if head is not None:
            return self._loss_with_gradients_head(head, groundtruth, params, upstream, targets)
        from . import ssd as this_module
        lib = _lib.load()
        call = Call()
        logits = call.tensor(self.raw_predictions['class_predictions'], torch.float32)
        B, A, C = logits.shape
        codes = call.tensor(self.raw_predictions['encoded_boxes'], torch.float32, (B, A, 4))
        anchors = call.tensor(self.anchors, torch.float32, (A, 4))
        sums, out = call.empty([3], torch.float64), call.empty([2], torch.float32)
        g_logits, g_codes = call.empty([B, A, C], torch.float32), call.empty([B, A, 4], torch.float32)
        up = None
        if upstream is not None:
            up = _upstream_tensor(call, upstream, logits.device)
        ctx = call.ctx()
        if targets is not None:
            reg, cls_t, matches = self._given_targets(call, targets, B, A)
            count = call.empty([1], torch.float64)
            count.copy_(targets['count'].reshape(1))                 # the all-reduce below works on a private copy
        else:
            gt = call.tensor(groundtruth['boxes'], torch.float32)
            G = gt.shape[1]
            labels = call.tensor(groundtruth['labels'], torch.int32, (B, G))
            num = call.tensor(groundtruth['num_boxes'], torch.int32, (B,))
            reg, cls_t, matches = call.empty([B, A, 4], torch.float32), call.empty([B, A], torch.int32), call.empty([B, A], torch.int32)
            count = call.empty([1], torch.float64)
            _lib.check(lib.ssdk_training_targets_count(ctx, ptr(anchors), A, ptr(gt), ptr(labels), ptr(num), B, G,
                                                       float(this_module.POSITIVES_THRESHOLD), float(this_module.NEGATIVES_THRESHOLD),
                                                       ptr(reg), ptr(cls_t), ptr(matches), ptr(count)))    # ssd.py:84 and :121-122
        self._reduce_count(ctx, count)
        _lib.check(lib.ssdk_ssd_loss_forward_backward(ctx, ptr(logits), ptr(codes), ptr(reg), ptr(cls_t), ptr(matches), B, A, C,
                                                      float(params['gamma']), float(params['alpha']), ptr(count), ptr(up),
                                                      ptr(sums), ptr(g_logits), ptr(g_codes)))
        self._reduce_and_finalize(ctx, sums, out)
        self.num_matches = sums[2]
        self._call = call
        self._saved = dict(logits=logits, codes=codes, sums=sums, gamma=float(params['gamma']), alpha=float(params['alpha']),
                           reg_targets=reg, cls_targets=cls_t, matches=matches)
        return ({'localization_loss': out[0], 'classification_loss': out[1]},
                {'class_predictions': g_logits, 'encoded_boxes': g_codes})

    @staticmethod
    def _given_targets(call, targets, B, A):
        reg = call.tensor(targets['reg_targets'], torch.float32, (B, A, 4))
        cls_t = call.tensor(targets['cls_targets'], torch.int32, (B, A))
        matches = call.tensor(targets['matches'], torch.int32, (B, A))
        return reg, cls_t, matches

    def _loss_with_gradients_head(self, head, groundtruth, params, upstream, targets=None):
        """Fused training step on the per-level tower outputs: targets -> count (all-reduced) -> ONE pass over every level's
        logits that yields the loss sums and the gradients in the head's own layout (ssdk_head_ssd_loss_forward_backward)."""
        lib = _lib.load()
        call = Call(head.device)
        B, A, C = head.batch_size, head.num_anchors, head.num_classes
        count, sums, out = call.empty([1], torch.float64), call.empty([3], torch.float64), call.empty([2], torch.float32)
        if targets is not None:
            reg, cls_t, matches = self._given_targets(call, targets, B, A)
            count.copy_(targets['count'].reshape(1))
        else:
            reg, cls_t, matches = self._targets_into(call, groundtruth, B, A, count=count)         # ssd.py:84 and :121-122
        g_cls, g_box, gd = self._head_grads(call, head)
        up = None
        if upstream is not None:
            up = _upstream_tensor(call, upstream, head.device)
        ctx = call.ctx()
        self._reduce_count(ctx, count)
        d = head.descriptor()
        _lib.check(lib.ssdk_head_ssd_loss_forward_backward(
            ctx, ctypes.byref(d), ptr(reg), ptr(cls_t), ptr(matches), B, A, C, float(params['gamma']), float(params['alpha']),
            ptr(count), ptr(up), ptr(sums), ctypes.byref(gd)))
        self._reduce_and_finalize(ctx, sums, out)
        self.num_matches = sums[2]
        self._call = call
        self._saved = dict(head=head, sums=sums, gamma=float(params['gamma']), alpha=float(params['alpha']),
                           reg_targets=reg, cls_targets=cls_t, matches=matches)
        return ({'localization_loss': out[0], 'classification_loss': out[1]},
                {'class_predictions': g_cls, 'encoded_boxes': g_box})

    def level_summaries(self, cls_losses=None, loc_losses=None, matches=None, top_fraction=0.20):
        """The quantities behind the reference's loss summaries (ssd.py:125-129,135-163), from the per-anchor vectors that
        `loss_sums(..., per_anchor=True)` returns.  For each given loss vector [B,A]: per image and FPN level the mean of the
        ceil(top_fraction * n_level) biggest values ('topk_mean', [B,L]) and the smallest of them ('topk_kth'), plus
        'histogram_mean' [L] = the mean of the vector the reference histograms (tf.reduce_mean(biggest_values, axis=0));
        for `matches`: 'matches' [B,L], 'mean_matches_per_image_on_level' [L] and 'total_mean_matches_per_image'."""
        out = {}
        per_level = np.asarray(self.num_anchors_per_feature_map, np.int32)
        L = len(per_level)
        lib = _lib.load()
        for name, v in (('classification_losses', cls_losses), ('localization_losses', loc_losses)):
            if v is None:
                continue
            call = Call()
            t = call.tensor(v, torch.float32)
            B, A = t.shape
            mean, kth = call.empty([B, L], torch.float32), call.empty([B, L], torch.float32)
            _lib.check(lib.ssdk_level_summaries(call.ctx(), ptr(t), None, B, A, per_level.ctypes.data, L, float(top_fraction),
                                                ptr(mean), ptr(kth), None))
            out[name] = {'topk_mean': mean, 'topk_kth': kth, 'histogram_mean': mean.mean(dim=0)}
        if matches is not None:
            call = Call()
            m = call.tensor(matches, torch.int32)
            B, A = m.shape
            cnt = call.empty([B, L], torch.float32)
            _lib.check(lib.ssdk_level_summaries(call.ctx(), None, ptr(m), B, A, per_level.ctypes.data, L, float(top_fraction),
                                                None, None, ptr(cnt)))
            out['matches'] = cnt
            out['mean_matches_per_image_on_level'] = cnt.mean(dim=0)                                # ssd.py:157-161
            out['total_mean_matches_per_image'] = cnt.sum(dim=1).mean()                              # ssd.py:129
        return out

    def _create_targets(self, groundtruth):
        """reference ssd.py:165-199: reg_targets [B,A,4], cls_targets [B,A], matches [B,A]."""
        from . import ssd as this_module
        return batch_training_targets(self.anchors, groundtruth['boxes'], groundtruth['labels'], groundtruth['num_boxes'],
                                      positives_threshold=this_module.POSITIVES_THRESHOLD,
                                      negatives_threshold=this_module.NEGATIVES_THRESHOLD)

    def matches_per_level(self, matches):
        """Matched-anchor count per image and FPN level, [B, L] (the quantity behind ssd.py:152-163)."""
        w = (matches >= 0).to(torch.float32)
        return torch.stack([c.sum(dim=1) for c in torch.split(w, self.num_anchors_per_feature_map, dim=1)], dim=1)
