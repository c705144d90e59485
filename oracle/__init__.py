"""TEST INFRASTRUCTURE ONLY: CPU checkers for the FLE hot path (see oracle/README.md)."""
