"""CPU oracle for the generator forward path - TEST INFRASTRUCTURE ONLY (see oracle/generators.py)."""
