"""score_b200: B200-native SCoRe training / scoring hot path behind the reference model interface."""
