"""CPU restatement of the reference algorithm: test infrastructure only (see effq_oracle.py)."""
