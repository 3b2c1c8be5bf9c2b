"""CPU oracle for the HPS build+solve hot path.  TEST INFRASTRUCTURE — see hps_oracle.py."""
