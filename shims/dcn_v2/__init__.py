"""`import dcn_v2` -> stmask_b200 (put <repo>/shims on PYTHONPATH)."""
from stmask_b200.compat.dcn_v2 import *  # noqa: F401,F403
from stmask_b200.compat.dcn_v2 import DCN, DCNv2, dcn_v2_conv  # noqa: F401
