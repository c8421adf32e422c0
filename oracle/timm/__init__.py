"""ORACLE (test infrastructure, never shipped): a minimal stand-in for the third-party `timm` package.

The reference calls `timm.create_model(name, pretrained, num_classes=0, global_pool=...)`
(experiments/multimodal_v1/train_mm_joint_dualtask.py:138, ui/gradio_app/infer_mm.py:22,
experiments/vision_v2/train_mil_attention_v1.py:135, ui/gradio_app/infer_mil.py:75) and
`timm.data.create_transform` (train_mm_joint_dualtask.py:75,87; infer_mm.py:13).  timm is an
un-vendored, un-pinned dependency (ui/gradio_app/requirements.txt:6) that is not installed in this
image, so its published `tf_efficientnet_b{0,4}_ns` architecture is restated here in plain PyTorch
 with timm's state-dict key spelling.  With `oracle/` on sys.path the reference
model files import unchanged; `tests/test_oracle_models.py` cross-checks this restatement against
the independent `transformers.EfficientNetModel` implementation (TF 'same' padding, eps 1e-3).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may import this.
"""
from . import data  # noqa: F401
from .efficientnet import EfficientNet, create_model, ARCHS  # noqa: F401

__version__ = "0.0-oracle-shim"
