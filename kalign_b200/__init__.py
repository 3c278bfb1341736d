"""kalign_b200: B200-native (sm_100a) alignment hot path behind the kalign C surface."""
