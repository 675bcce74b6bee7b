"""Drop-in for `src/model/hicedrn_sr3_Diff.py` (/root/reference/src/model/hicedrn_sr3_Diff.py:267-359)."""
from ..nets import hicedrn_sr3_Diff as hicedrn_Diff

n_feat = 256
kernel_size = 3

__all__ = ["hicedrn_Diff"]
