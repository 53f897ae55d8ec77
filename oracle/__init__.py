"""CPU oracle of the MOOG Environment.step hot path.  TEST INFRASTRUCTURE ONLY."""
