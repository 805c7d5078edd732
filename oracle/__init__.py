"""CPU oracle for the retrieval scoring hot path -- TEST INFRASTRUCTURE ONLY (see np_oracle.py)."""
