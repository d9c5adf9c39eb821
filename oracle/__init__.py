"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/pm4_oracle.c)."""
