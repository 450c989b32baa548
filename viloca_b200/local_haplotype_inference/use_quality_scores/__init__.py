"""B200 drop-in for viloca/local_haplotype_inference/use_quality_scores."""
