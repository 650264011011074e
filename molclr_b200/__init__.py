"""molclr_b200 -- B200-native MolCLR pre-training hot path."""
from .batch import Batch  # noqa: F401
