"""smcdet_b200 -- B200-native implementation of the per-tile SMC hot path of timwhite0/smcdet."""
