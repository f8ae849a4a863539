"""CPU oracle of the PGCN hot path — TEST INFRASTRUCTURE ONLY (see pgcn_oracle.py)."""
