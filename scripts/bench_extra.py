#!/usr/bin/env python
"""Kept for the command lines recorded in profiles/: the secondary workloads now live in bench.py (`--workload cfgN`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    if "--workload" not in sys.argv:
        sys.argv += ["--workload", "cfg2"]
    bench.main()
