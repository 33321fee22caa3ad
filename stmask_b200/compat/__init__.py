"""Drop-in replacements for the operator packages the reference imports
(`dcn_v2`, `mmcv.ops`, `spatial_correlation_sampler`); see INTEGRATION.md."""
