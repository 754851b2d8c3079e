"""CPU oracle for the hot path — test infrastructure only (see oracle/bn254.py, oracle/bn254_ref.c)."""
