#!/usr/bin/env python
"""bench.py — placeholder until the C3 pipeline lands (see git log); prints the contract line
for the EWA-Lanczos 2x kernel only."""
import json, sys
print(json.dumps({"metric": "bench not implemented yet"}))
