#!/usr/bin/env python
"""Same name as the reference's installed script (setup.py:15-23); runs the B200 engine."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from protein_gibbs_sampler_b200.cli import pgen_esm as _m
_m.cli()
