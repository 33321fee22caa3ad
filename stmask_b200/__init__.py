"""stmask_b200 — B200-native (sm_100a) kernels for STMask's feature-calibration and
temporal-fusion hot path, behind the reference's own operator API.

    stmask_b200.compat.dcn_v2                       DCN, DCNv2, dcn_v2_conv
    stmask_b200.compat.mmcv_ops                     DeformConv2d, ModulatedDeformConv2d(+Pack), ...
    stmask_b200.compat.spatial_correlation_sampler  spatial_correlation_sample, SpatialCorrelationSampler
    stmask_b200.feature_align.FeatureAlign          FCB(ada) / FCB(ali), all FPN levels in one launch
    stmask_b200.temporal_fusion                     correlate, correlate_concat
    stmask_b200.backbone_dcn                        DCN placement rule + layer geometry
    stmask_b200.sharding                            clip/frame partition + one-frame halo exchange
    stmask_b200.install_shims()                     make `dcn_v2`, `mmcv.ops`, `spatial_correlation_sampler` resolve here

The arithmetic lives in stmask_b200/lib/libstmask_b200.so (C ABI: include/stmask_b200.h).
There is no CPU path: importing works anywhere, calling an operator needs a B200.
"""
__version__ = "0.2.0"

from .shim_install import install_shims  # noqa: E402,F401
