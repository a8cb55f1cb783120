"""Criteria of the DefaultSegmentorV2 wrapper (pointcept/models/losses/builder.py:14-51) on the B200 path.

`build_criteria(cfg, loss_type, task_num)` accepts the reference's criteria config (a list of dicts with `type` in
MSELoss / CrossEntropyLoss / LovaszLoss, misc.py:24-129, lovasz.py:210-272) and returns a callable that maps the
reference's `point` dict (n_pred, n_target, optional c_pred / c_target, loss_mode) to the scalar loss -- one C-ABI
call (cdseg_criteria: softmax + CE + per-class radix sort + Lovasz scan + masked MSE + the EW / GLS combination),
the result stays on the device.  Forward values only: the backward pass is the §8(f) training row.
"""
import torch

from . import ops

_SUPPORTED = ("MSELoss", "CrossEntropyLoss", "LovaszLoss")


class Criteria:
    def __init__(self, cfg=None, loss_type="EW", task_num=2):
        self.cfg = list(cfg) if cfg is not None else []
        self.loss_type, self.task_num = loss_type, task_num
        self.has = [False, False, False]
        self.weights = [1.0, 1.0, 1.0]
        ignores = []
        self.mse_use_ignore = False
        for c in self.cfg:
            c = dict(c)
            kind = c.pop("type")
            if kind not in _SUPPORTED:
                raise NotImplementedError(f"criterion {kind!r}: only {_SUPPORTED} are on the CDSegNet path (configs/*/CDSegNet.py)")
            k = _SUPPORTED.index(kind)
            if self.has[k]:
                raise NotImplementedError(f"two {kind} criteria")
            self.has[k] = True
            self.weights[k] = float(c.pop("loss_weight", 1.0))
            ign = c.pop("ignore_index", None if kind != "CrossEntropyLoss" else -1)
            if kind == "MSELoss":
                if c.pop("batch_sample_point", 8192) > 0:
                    raise NotImplementedError("MSELoss.batch_sample_point > 0 (random sub-sampling) is not on the shipped path")
                self.mse_use_ignore = bool(ign)                   # `if(self.ignore_index)` is a truth test (misc.py:77)
                if ign:
                    ignores.append(ign)
                for key, dflt in (("pred", "c_pred"), ("target", "c_target"), ("segment_target", "n_target"), ("reduction", "none")):
                    if c.pop(key, dflt) != dflt:
                        raise NotImplementedError(f"MSELoss.{key}")
            elif kind == "CrossEntropyLoss":
                if not ign:
                    raise NotImplementedError("CrossEntropyLoss without a truthy ignore_index")
                ignores.append(ign)
                for key, dflt in (("pred", "n_pred"), ("target", "n_target"), ("weight", None), ("reduction", "mean"), ("label_smoothing", 0.0)):
                    if c.pop(key, dflt) != dflt:
                        raise NotImplementedError(f"CrossEntropyLoss.{key}")
            else:
                if c.pop("mode") != "multiclass" or c.pop("per_image", False) or c.pop("class_seen", None) is not None:
                    raise NotImplementedError("LovaszLoss: only mode='multiclass', per_image=False, class_seen=None")
                if ign is None:
                    raise NotImplementedError("LovaszLoss without ignore_index")
                ignores.append(ign)
                for key, dflt in (("pred", "n_pred"), ("target", "n_target")):
                    if c.pop(key, dflt) != dflt:
                        raise NotImplementedError(f"LovaszLoss.{key}")
            c.pop("size_average", None); c.pop("reduce", None)
            if c:
                raise NotImplementedError(f"{kind}: unsupported options {sorted(c)}")
        if len(set(ignores)) > 1:
            raise NotImplementedError("criteria with different ignore_index values")
        self.ignore_index = ignores[0] if ignores else -1

    def parts(self, point):
        """fp32 [5] on the device: MSE, CE, Lovasz, EW sum, GLS sqrt(MSE * (CE + Lovasz))"""
        c_pred, c_target = point.get("c_pred"), point.get("c_target")
        tgt = point["n_target"]
        if tgt.dtype is not torch.int64:
            tgt = tgt.long()
        return ops.criteria(point["n_pred"].contiguous(), tgt.contiguous(), self.ignore_index,
                            c_pred.contiguous() if c_pred is not None else None, c_target.contiguous() if c_target is not None else None,
                            self.mse_use_ignore, self.weights, self.has)

    def value_and_grad(self, point):
        """(loss, d loss / d n_pred, d loss / d c_pred or None) of the pass `point["loss_mode"]` selects -- what the backward of the
        network would start from (cdseg_criteria_grad)"""
        gls = point["loss_mode"] == "train" and self.loss_type == "GLS"
        if gls and not (self.task_num == 2 and sum(self.has) == 3):
            raise NotImplementedError("GLS gradients: task_num=2 over [MSELoss, CrossEntropyLoss, LovaszLoss] only")
        c_pred, c_target = point.get("c_pred"), point.get("c_target")
        tgt = point["n_target"]
        tgt = tgt if tgt.dtype is torch.int64 else tgt.long()
        out, gn, gc = ops.criteria_grad(point["n_pred"].contiguous(), tgt.contiguous(), self.ignore_index,
                                        c_pred.contiguous() if c_pred is not None else None,
                                        c_target.contiguous() if c_target is not None else None, self.mse_use_ignore, self.weights, self.has, gls)
        return (out[4] if gls else out[3]), gn, gc

    def __call__(self, point):
        if not self.cfg:
            return point                                            # "loss computation occur in model" (builder.py:25-27)
        out = self.parts(point)
        if point["loss_mode"] == "eval" or self.loss_type == "EW":
            return out[3]
        if point["loss_mode"] == "train" and self.loss_type == "GLS":
            if self.task_num == 2 and sum(self.has) == 3:
                return out[4]
            if self.task_num == 1:                                  # builder.py:41-42: loss[0] + loss[1] of the configured list
                vals = [out[k] for k in range(3) if self.has[k]]
                return vals[0] + vals[1]
            raise NotImplementedError("GLS with task_num=%d over %d criteria" % (self.task_num, sum(self.has)))
        return 0.0


def build_criteria(cfg, loss_type="EW", task_num=2):
    return Criteria(cfg, loss_type=loss_type, task_num=task_num)
