"""B200-native GPR hot path of the MQCLE propagator."""
