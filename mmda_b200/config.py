"""Configuration object for the MISA hot path.

The reference builds a ``Config`` from argparse (reference ``src/config.py:71-170``) and the
model / solver read plain attributes off it.  ``MisaConfig`` carries exactly the attributes the
hot path reads under the same names and with the same defaults, so
either object can be handed to :class:`mmda_b200.MISA`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional


# activation name -> (kernel epilogue id, torch class name).  ids are shared with csrc/common.cuh
ACTIVATIONS = {
    "none": 0,
    "leakyrelu": 1,
    "sigmoid": 2,
    "relu": 3,
    "tanh": 4,
}


def activation_name(act: Any) -> str:
    """Map what the reference stores in ``config.activation`` (an ``nn.Module`` *class*,
    reference ``src/config.py:24-27,78-79``) or a plain string onto a kernel epilogue name."""
    if isinstance(act, str):
        name = act.lower()
    else:
        cls = act if isinstance(act, type) else type(act)
        name = cls.__name__.lower()
    if name not in ACTIVATIONS or name == "none":
        raise ValueError(
            f"activation {act!r} has no sm_100a epilogue in this build "
            f"(supported: {sorted(k for k in ACTIVATIONS if k != 'none')})")
    return name


@dataclass
class MisaConfig:
    # --- model (reference src/models.py:20-47, src/config.py:146-154) ---
    embedding_size: int = 300
    visual_size: int = 35
    acoustic_size: int = 74
    num_classes: int = 6
    hidden_size: int = 128
    dropout: float = 0.1
    activation: Any = "leakyrelu"
    extractor: str = "lstm"
    rnncell: str = "lstm"
    use_bert: bool = False
    use_cmd_sim: bool = True
    reverse_grad_weight: float = 1.0
    threshold: float = 0.35
    vocab_size: int = 20000          # len(config.word2id) in the reference
    word2id: Optional[Any] = None
    pretrained_emb: Optional[Any] = None
    # --- solver (reference src/config.py:134-143, src/solver.py:175-186) ---
    data: str = "mosei"
    use_confidNet: bool = False
    diff_weight: float = 0.3
    sim_weight: float = 0.7
    sp_weight: float = 0.0
    recon_weight: float = 0.7
    conf_weight: float = 0.3
    learning_rate: float = 1e-4
    clip: float = 1.0
    batch_size: int = 64
    model: str = "MISA"
    # --- this build only ---
    precision: str = "fp32"          # "fp32" | "bf16" (operands of the hoisted input-projection GEMMs)

    def __post_init__(self):
        if self.word2id is None:
            self.word2id = range(self.vocab_size)
        else:
            self.vocab_size = len(self.word2id)


def mosi_config(**kw) -> MisaConfig:
    """BASELINE.json configs[0]: MOSI-shape (300/47/74), batch 64."""
    d = dict(visual_size=47, acoustic_size=74, batch_size=64, data="mosi")
    d.update(kw)
    return MisaConfig(**d)


def mosei_config(**kw) -> MisaConfig:
    """BASELINE.json configs[1]: MOSEI-shape (300/35/74), batch 256."""
    d = dict(visual_size=35, acoustic_size=74, batch_size=256, data="mosei")
    d.update(kw)
    return MisaConfig(**d)
