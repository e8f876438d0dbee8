"""cgraphs_b200 - B200-native H-bond detection hot path of evabertalan/cgraphs (Bridge HbondAnalysis)."""
__version__ = "0.1.0"
