"""B200 drop-in for viloca/local_haplotype_inference/learn_error_params."""
