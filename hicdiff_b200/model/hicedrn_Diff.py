"""Drop-in for `src/model/hicedrn_Diff.py` (/root/reference/src/model/hicedrn_Diff.py:210-297)."""
from ..nets import hicedrn_Diff

n_feat = 256
kernel_size = 3

__all__ = ["hicedrn_Diff"]
