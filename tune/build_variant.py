#!/usr/bin/env python3
"""usage: tune/build_variant.py OUT.so [-DFLAG ...]   -- builds a library variant for same-box A/B runs (kernel tuning only)"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fallingsandsurvival_b200 import build as b  # noqa: E402

out, flags = sys.argv[1], sys.argv[2:]
cmd = [b.nvcc()] + b.NVCC_FLAGS + flags + ["-o", out] + b.SOURCES + ["-ldl"]
print(" ".join(cmd))
subprocess.check_call(cmd)
