"""CPU oracle — test infrastructure only (see oracle/mpc_oracle.cpp header)."""
