"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference Python from /root/reference (exists
only in the build container) or from its git-ignored copy `baseline/_ref/` (made by
baseline/install_ref.py; travels to the GPU box) so golden vectors can be generated from the
reference itself and the reference's own consumers / modules can be run on this package's
outputs and operators.

The reference cannot be constructed offline as-is: it needs RoBERTa weights
from the hub, `ipdb`/`termcolor`, a `pointnet2._ext` extension that only has CUDA kernels,
and a cwd containing `data/class_embeddings3d.npy`.  This module supplies:
  * stub modules for `ipdb` / `termcolor`;
  * `pointnet2._ext` := `oracle.point_ops` (CPU restatement, same 9 functions);
  * `RobertaTokenizerFast/RobertaModel.from_pretrained` patched to tiny placeholders — the
    golden vectors start from a synthetic `last_hidden_state`, so RoBERTa is never evaluated;
  * a fake tokenizer / text encoder pair that replays the provided hidden states through the
    reference's own `_run_backbones` (`models/bdetr.py:156-175`).
"""
import contextlib
import os
import sys
import types

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference" if os.path.isdir("/root/reference/models") else os.path.join(_ROOT, "baseline", "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "models"))


class _Tokenized(dict):
    """Mimics transformers.BatchEncoding as far as bdetr.py:164-174 and losses.py need."""

    def to(self, device):
        return self

    @property
    def attention_mask(self):
        return self["attention_mask"]


class _FakeTokenizer:
    def batch_encode_plus(self, text, padding="longest", return_tensors="pt"):
        # `text` is the (hidden_states, attention_mask) pair smuggled through inputs['text']
        hidden, mask = text
        return _Tokenized(attention_mask=mask, _hidden=hidden)


class _FakeTextEncoder(torch.nn.Module):
    class config:
        hidden_size = 768

    def forward(self, attention_mask=None, _hidden=None):
        return types.SimpleNamespace(last_hidden_state=_hidden)


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


_models = None


def import_reference(ext=None):
    """Returns the reference's `models` package (BeaUTyDETR, ...).  `ext`: the module served as
    `pointnet2._ext` — the CPU oracle by default, `butd_detr_b200.pointnet2_ext` to run the
    reference's PointNet++ modules on the CUDA drop-in."""
    global _models
    if _models is not None:
        if ext is not None:
            import pointnet2
            sys.modules["pointnet2._ext"] = ext
            pointnet2._ext = ext
            for name in ("pointnet2_utils", "pointnet2.pointnet2_utils"):
                if name in sys.modules:
                    sys.modules[name]._ext = ext
        return _models
    assert available(), "neither /root/reference nor baseline/_ref is present on this machine"
    if ext is None:
        from . import point_ops
    else:
        point_ops = ext
    for name in ("ipdb", "termcolor"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.set_trace = lambda *a, **k: None
            m.colored = lambda s, *a, **k: s
            sys.modules[name] = m
    for p in (REF, os.path.join(REF, "pointnet2")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pointnet2  # namespace package at /root/reference/pointnet2
    sys.modules["pointnet2._ext"] = point_ops
    pointnet2._ext = point_ops
    import transformers
    transformers.RobertaTokenizerFast.from_pretrained = classmethod(lambda cls, *a, **k: _FakeTokenizer())
    transformers.RobertaModel.from_pretrained = classmethod(lambda cls, *a, **k: _FakeTextEncoder())
    with _cwd(REF):
        import models  # noqa
    _models = models
    return models


def build_reference_model(ext=None, **kwargs):
    """`BeaUTyDETR(**kwargs)` from the reference, eval mode, fake text front-end."""
    models = import_reference(ext)
    with _cwd(REF):
        model = models.BeaUTyDETR(**kwargs)
    return model.eval()


@torch.no_grad()
def run_reference(model, inputs):
    """Run the reference forward on our input schema (text_hidden/text_attention_mask)."""
    ref_in = {
        "point_clouds": inputs["point_clouds"],
        "text": (inputs["text_hidden"], inputs["text_attention_mask"]),
        "det_boxes": inputs["det_boxes"],
        "det_bbox_label_mask": inputs["det_bbox_label_mask"],
        "det_class_ids": inputs["det_class_ids"],
    }
    return model(ref_in)
