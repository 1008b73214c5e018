#!/usr/bin/env python3
"""Command-line entry with the reference's flags (see ballermixplus_b200/cli.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ballermixplus_b200.cli import main  # noqa: E402

if __name__ == '__main__':
    main()
