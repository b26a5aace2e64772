"""mmda_b200 -- B200-native (sm_100a) implementation of the MISA multimodal-fusion training step
of SoyeonHH/MMDA (reference src/models.py + the loss/step contract of src/solver.py).

Public surface:
  MISA            drop-in model class (same constructor / forward / state_dict keys)
  MisaConfig      the config attributes the hot path reads
  FusedTrainer    level-2 fused step (losses + backward + clip + Adam, optional data parallel)
  synthetic       seeded MOSI/MOSEI-shaped batches
  collate         device-resident dataset + collate; wordpiece: the BERT tokenizer of its host side
All arithmetic runs in libmmda_b200.so (hand-written CUDA behind a C ABI, include/mmda_b200.h).
"""
from .config import MisaConfig, mosei_config, mosi_config  # noqa: F401
from .model import MISA  # noqa: F401


def __getattr__(name):
    if name == "FusedTrainer":
        from .trainer import FusedTrainer
        return FusedTrainer
    if name == "WordPieceTokenizer":
        from .wordpiece import WordPieceTokenizer
        return WordPieceTokenizer
    raise AttributeError(name)
