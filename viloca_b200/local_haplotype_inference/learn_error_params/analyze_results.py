"""Finishing steps and writers (learn_error_params/analyze_results.py)."""
from .._common import correct_reads, group_haplotypes as get_unique_haplotypes, haplotypes_to_fasta  # noqa: F401
from .._common import summarize as summarize_results  # noqa: F401
