"""raer_b200: B200-native pileup engines behind raer's .Call boundary."""
from .filter_param import FilterParam, scan_bam_flag, DEFAULT_BULK_BAM_FLAGS, DEFAULT_SC_BAM_FLAGS  # noqa: F401
