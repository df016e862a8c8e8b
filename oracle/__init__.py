"""CPU oracle (test infrastructure only; see reid_oracle.py header)."""
