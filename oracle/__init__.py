"""TEST INFRASTRUCTURE ONLY: CPU oracle of the SalpRobotEnv hot path (see salp_oracle.c)."""
