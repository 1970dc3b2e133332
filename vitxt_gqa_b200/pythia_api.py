"""Host-side mirror of the reference's plugin seam for the forward/grounding path.

The reference hangs models on a global registry and a `BaseModel.__call__` that
runs forward -> Losses -> Metrics (reference pythia/common/registry.py:24-338,
pythia/models/base_model.py:52-149, pythia/common/sample.py:60-326,
pythia/modules/losses.py:41-173).  When the real `pythia` package is importable
we bind to ITS registry / BaseModel / SampleList so that
`registry.get_model_class("t2s")` hands back the B200 implementation with no
config change.  When it is not (e.g. the GPU box), the minimal API-compatible
stand-ins below are used: same names, argument meaning and error behaviour, for
exactly the members the model path touches.
"""
import collections
import collections.abc
import os
import warnings
from collections import OrderedDict
from copy import deepcopy

import torch
import yaml
from torch import nn

try:  # bind to the real framework when present (drop-in mode)
    from pythia.common.registry import registry as _real_registry  # type: ignore
    from pythia.models.base_model import BaseModel as _RealBaseModel  # type: ignore
    from pythia.common.sample import SampleList as _RealSampleList  # type: ignore
    HAVE_PYTHIA = True
except Exception:  # pragma: no cover - depends on environment
    _real_registry = _RealBaseModel = _RealSampleList = None
    HAVE_PYTHIA = False


# --------------------------------------------------------------------------- registry
class _Registry:
    """Stand-in for reference `Registry` (registry.py:24-338): class-level maps,
    dotted-path state, decorator registration."""

    mapping = {
        "model_name_mapping": {},
        "loss_name_mapping": {},
        "metric_name_mapping": {},
        "state": {},
    }

    @classmethod
    def register_model(cls, name):
        def wrap(model_cls):
            assert issubclass(model_cls, BaseModel), "All models must inherit BaseModel class"
            cls.mapping["model_name_mapping"][name] = model_cls
            return model_cls
        return wrap

    @classmethod
    def register_loss(cls, name):
        def wrap(loss_cls):
            assert issubclass(loss_cls, nn.Module), "All loss must inherit torch.nn.Module class"
            cls.mapping["loss_name_mapping"][name] = loss_cls
            return loss_cls
        return wrap

    @classmethod
    def register_metric(cls, name):
        def wrap(metric_cls):
            cls.mapping["metric_name_mapping"][name] = metric_cls
            return metric_cls
        return wrap

    @classmethod
    def get_metric_class(cls, name):
        return cls.mapping["metric_name_mapping"].get(name, None)

    @classmethod
    def register(cls, name, obj):
        path = name.split(".")
        cur = cls.mapping["state"]
        for part in path[:-1]:
            cur = cur.setdefault(part, {})
        cur[path[-1]] = obj

    @classmethod
    def get(cls, name, default=None, no_warning=False):
        value = cls.mapping["state"]
        for sub in name.split("."):
            value = value.get(sub, default)
            if value is default:
                break
        return value

    @classmethod
    def unregister(cls, name):
        return cls.mapping["state"].pop(name, None)

    @classmethod
    def get_model_class(cls, name):
        return cls.mapping["model_name_mapping"].get(name, None)

    @classmethod
    def get_loss_class(cls, name):
        return cls.mapping["loss_name_mapping"].get(name, None)


# --------------------------------------------------------------------------- config
class ConfigNode(OrderedDict):
    """Attribute + item access dict (reference pythia/utils/configuration.py
    `ConfigNode`); nested dicts are wrapped, `.get` works as on a dict."""

    def __init__(self, init=None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = self._wrap(v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, collections.abc.Mapping) and not isinstance(v, ConfigNode):
            return ConfigNode(v)
        if isinstance(v, (list, tuple)):
            return [cls._wrap(x) for x in v]
        return v

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self[key] = self._wrap(value)


_CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


def load_yaml_config(path, overrides=None):
    """YAML + `includes:` + dotted overrides, the subset of the reference
    loader the model path needs (reference pythia/utils/configuration.py:96-160).
    A bare file name is looked up among the packaged configs."""
    if not os.path.exists(path):
        cand = os.path.join(_CONFIG_DIR, os.path.basename(path))
        if os.path.exists(cand):
            path = cand
    with open(path, "r", encoding="utf-8") as f:
        data = yaml.safe_load(f) or {}
    merged = {}
    for inc in data.pop("includes", []) or []:
        inc_path = os.path.join(os.path.dirname(path), inc)
        if os.path.exists(inc_path):
            _nested_update(merged, load_yaml_config(inc_path))
    _nested_update(merged, data)
    for key, value in (overrides or {}).items():
        cur = merged
        parts = key.split(".")
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = value
    return ConfigNode(merged)


def _nested_update(dst, src):
    for k, v in src.items():
        if isinstance(v, collections.abc.Mapping) and isinstance(dst.get(k), collections.abc.Mapping):
            _nested_update(dst[k], v)
        else:
            dst[k] = deepcopy(v) if isinstance(v, (dict, list)) else v
    return dst


# --------------------------------------------------------------------------- SampleList
class _SampleList(OrderedDict):
    """Stand-in for reference `SampleList` (sample.py:60-326): an OrderedDict of
    batched fields with attribute read access, `add_field`, `to(device)`.
    Like the reference it defines no `__setattr__`: add fields with
    `add_field()` or item assignment."""

    _TENSOR_FIELD_ = "_tensor_field"

    def __init__(self, samples=None):
        super().__init__()
        if not samples:
            return
        if isinstance(samples, collections.abc.Mapping):
            for k, v in samples.items():
                self.add_field(k, v)
            return
        if isinstance(samples[0], (tuple, list)) and isinstance(samples[0][0], str):
            for k, v in samples:
                self.add_field(k, v)
            return
        for field in samples[0].keys():
            vals = [s[field] for s in samples]
            if isinstance(vals[0], torch.Tensor):
                self.add_field(field, torch.stack(vals, 0))
            else:
                self.add_field(field, vals)

    def __getattr__(self, key):
        if key.startswith("_") or key not in self:
            raise AttributeError(
                "Key {} not found in the SampleList. Valid choices are {}".format(key, list(self.keys())))
        return self[key]

    def fields(self):
        return list(self.keys())

    def get_field(self, field):
        return self[field]

    def get_batch_size(self):
        tf = self.__dict__.get(self._TENSOR_FIELD_)
        assert tf is not None, "There is no tensor yet in SampleList"
        return self[tf].size(0)

    def add_field(self, field, data):
        tf = self.__dict__.get(self._TENSOR_FIELD_)
        if (isinstance(data, torch.Tensor) and data.dim() != 0 and tf is not None
                and data.size(0) != self[tf].size(0)):
            raise AssertionError(
                "A tensor field to be added must have same size as existing tensor fields in SampleList. "
                "Passed size: {}, Required size: {}".format(len(data), len(self[tf])))
        self[field] = data.clone() if isinstance(data, torch.Tensor) else deepcopy(data)
        if isinstance(self[field], torch.Tensor) and tf is None:
            self.__dict__[self._TENSOR_FIELD_] = field

    def copy(self):
        out = type(self)()
        for k in self.keys():
            out.add_field(k, self[k])
        return out

    def to(self, device, non_blocking=True):
        if not isinstance(device, torch.device):
            if not isinstance(device, str):
                raise TypeError("device must be either 'str' or 'torch.device' type, {} found".format(type(device)))
            device = torch.device(device)
        out = type(self)()
        for k in self.keys():
            v = self[k]
            out[k] = v.to(device, non_blocking=non_blocking) if hasattr(v, "to") else v
        if self._TENSOR_FIELD_ in self.__dict__:
            out.__dict__[self._TENSOR_FIELD_] = self.__dict__[self._TENSOR_FIELD_]
        return out


# --------------------------------------------------------------------------- BaseModel / Losses
class _Losses(nn.Module):
    """Stand-in for reference `Losses`/`PythiaLoss` (losses.py:41-173): weighted
    loss dict keyed `<dataset_type>/<dataset_name>/<loss name>`; empty when the
    batch has no `targets`."""

    def __init__(self, loss_list):
        super().__init__()
        self.losses = []
        for params in loss_list:
            if "type" not in params:
                raise ValueError("Parameters to loss must have 'type' field to specify type of loss to instantiate")
            cls = registry.get_loss_class(params["type"])
            if cls is None:
                raise ValueError("No loss named {} is registered to registry".format(params["type"]))
            crit = cls(**(params.get("params", {}) or {}))
            self.losses.append((params["type"], params["weight"], crit))
            self.add_module("loss_%d" % len(self.losses), crit)

    def forward(self, sample_list, model_output, *args, **kwargs):
        output = {}
        if "targets" not in sample_list:
            return output
        for name, weight, crit in self.losses:
            loss = weight * crit(sample_list, model_output, *args, **kwargs)
            if not isinstance(loss, torch.Tensor):
                loss = torch.tensor(loss, dtype=torch.float)
            if loss.dim() == 0:
                loss = loss.view(1)
            key = "{}/{}/{}".format(sample_list["dataset_type"], sample_list["dataset_name"], name)
            output[key] = loss
        registry.register("losses.{}.{}".format(sample_list["dataset_name"], sample_list["dataset_type"]), output)
        return output


class _BaseModel(nn.Module):
    """Stand-in for reference `BaseModel` (base_model.py:52-149): forward, then
    `Losses`, then `Metrics` (vitxt_gqa_b200/metrics.py, the evaluation step on
    the device) built from `config.losses` / `config.metrics`."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.writer = registry.get("writer")

    def build(self):
        raise NotImplementedError("Build method not implemented in the child model class.")

    def init_losses_and_metrics(self):
        losses = self.config.get("losses", [])
        if len(losses) == 0:
            warnings.warn("No losses are defined in model configuration.")
        self.losses = _Losses(losses)
        from .metrics import Metrics      # late: metrics.py imports this module's registry
        self.metrics = Metrics(list(self.config.get("metrics", []) or []))

    def __call__(self, sample_list, *args, **kwargs):
        model_output = super().__call__(sample_list, *args, **kwargs)
        assert isinstance(model_output, collections.abc.Mapping), \
            "A dict must be returned from the forward of the model."
        if "losses" not in model_output:
            model_output["losses"] = self.losses(sample_list, model_output)
        if "metrics" not in model_output:
            model_output["metrics"] = self.metrics(sample_list, model_output)
        return model_output


if HAVE_PYTHIA:  # pragma: no cover - only with the reference on sys.path
    registry = _real_registry
    BaseModel = _RealBaseModel
    SampleList = _RealSampleList
else:
    BaseModel = _BaseModel
    registry = _Registry
    SampleList = _SampleList


class _Writer:
    def write(self, msg, level="info"):
        pass


def register_defaults(vocab_size=5000, ocr_max_num=960, dataset="vtextgqa", bos_idx=1):
    """Registry state the model constructor reads (reference t2s.py:29,138,149;
    base_model.py:66; losses.py:73): `writer`, `config`,
    `<dataset>_num_final_outputs`, `<dataset>_answer_processor`."""
    if registry.get("writer", no_warning=True) is None:
        registry.register("writer", _Writer())
    registry.register("config", ConfigNode({
        "datasets": dataset,
        "training_parameters": {"evalai_inference": False},
    }))
    registry.register(dataset + "_num_final_outputs", vocab_size + ocr_max_num)
    proc = ConfigNode({"BOS_IDX": bos_idx, "EOS_IDX": 2, "PAD_IDX": 0})
    registry.register(dataset + "_answer_processor", proc)
