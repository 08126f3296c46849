"""CPU oracle of the BoatEnv step path: test infrastructure only (see boat_oracle.h)."""
