from stmask_b200.compat.spatial_correlation_sampler import *  # noqa: F401,F403
