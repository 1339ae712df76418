"""oracle/ -- CPU checkers (TEST INFRASTRUCTURE ONLY; see oracle/oracle.py)."""
