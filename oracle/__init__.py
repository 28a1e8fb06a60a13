"""CPU oracle package — TEST INFRASTRUCTURE, not product code (see oracle.py)."""
