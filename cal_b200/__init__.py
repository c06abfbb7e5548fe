"""cal_b200 -- B200-native (sm_100a) implementation of the CAL hot path.

CausalGCN / CausalGAT forward + backward over batched mini-graphs as
hand-written CUDA kernels behind a C-ABI library (``include/cal_b200.h``),
wrapped in ``nn.Module``s that keep the reference constructor / forward
signature (model.py:14-22,85 and model.py:316-320,380 upstream).
"""
__version__ = "0.1.0"

from .data import Batch, Data, DataLoader, make_batches, make_dataset  # noqa: F401
from .model import CausalGAT, CausalGCN, CausalGIN, Engine, GATConv, GCNConv, GINConv, flat_offsets  # noqa: F401
from .trainer import (GraphStore, PackedLayout, PeerExchange, Trainer, allreduce_flat_grads, batch_caps,  # noqa: F401
                      cosine_lr, epoch_order)
