"""molclr_b200 -- B200-native MolCLR pre-training hot path.

Drop-in classes with the reference's API: ``GINet``, ``NTXentLoss`` (and ``GCN``), backed by
hand-written sm_100a CUDA kernels behind the C ABI in ``include/molclr_b200.h``.
"""
from .batch import Batch  # noqa: F401
from .ginet import GINet, GINEConv  # noqa: F401
from .gcn import GCN, GCNConv  # noqa: F401
from . import ginet_finetune  # noqa: F401  (ginet_finetune.GINet: models/ginet_finetune.py)
from . import gcn_finetune  # noqa: F401  (gcn_finetune.GCN: models/gcn_finetune.py)
from . import ginet_finetune_mp  # noqa: F401  (ginet_finetune_mp.GINet: models/ginet_finetune_mp.py, motif attention)
from .nt_xent import NTXentLoss  # noqa: F401
from .graph import GraphPlan, get_plan  # noqa: F401
from .functional import normalize, pretrain_loss  # noqa: F401
from .dataset import PackedMolecules, augment_pair  # noqa: F401
