#!/usr/bin/env python
"""pgsd2vtu.py file.gsd -- one ID-ordered .vtu per frame (see pgsd_sph_b200/vtu.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgsd_sph_b200 import vtu  # noqa: E402

if __name__ == "__main__":
    for name in vtu.convert(sys.argv[1]):
        print(name)
