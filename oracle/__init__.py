"""CPU oracle package (test infrastructure only; see fe_oracle.c)."""
