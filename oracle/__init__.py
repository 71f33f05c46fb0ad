"""CPU oracle for the AQLM hot path -- TEST INFRASTRUCTURE ONLY (see aqlm_oracle.py header)."""
