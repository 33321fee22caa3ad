from stmask_b200.compat.mmcv_ops import *  # noqa: F401,F403
