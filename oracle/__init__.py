"""CPU oracle — test infrastructure only (see oracle/limitador_oracle.h)."""
