"""CPU oracle for the hot path: test infrastructure only (see oracle/oracle.py)."""
