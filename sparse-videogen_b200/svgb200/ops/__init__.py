"""Mirror of svg/kernels/ops (the reference "operator API"): BSR mask builders + sparse attention."""
from .attention_ops import *  # noqa: F401,F403
from .attention_ops_wan import *  # noqa: F401,F403
