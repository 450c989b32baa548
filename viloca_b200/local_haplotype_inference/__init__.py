"""Drop-in replacements for ``viloca.local_haplotype_inference`` (both modes).

``viloca/shotgun.py:57-58`` imports ``use_quality_scores.run_dpm_mfa`` and
``learn_error_params.run_dpm_mfa`` and calls their ``main``; the modules here keep those
names, signatures and on-disk outputs, and run the inference on the CUDA engine.
"""
