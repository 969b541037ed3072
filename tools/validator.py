#!/usr/bin/env python3
"""The reference's validator.py command line on this repo's implementation (no scikit-image needed):
    python tools/validator.py reference_directory own_directory"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

if __name__ == "__main__":
    ge.load_package().validate.main([sys.argv[0]] + sys.argv[1:])
