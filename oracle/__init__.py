"""CPU oracle (test infrastructure only) -- see oracle/fqs_oracle.cpp."""
