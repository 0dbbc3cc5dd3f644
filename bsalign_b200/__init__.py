"""bsalign_b200: B200-native (sm_100a) implementation of bsalign's banded striped DP hot path."""
