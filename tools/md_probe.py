#!/usr/bin/env python3
"""bench.py's config-5 section on its own (water2046, velocity Verlet): python tools/md_probe.py [steps].
DFTBP_B200_MD_MIXER=anderson selects the round-1 mixer for comparison."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev)
print(json.dumps(bench.section_md(dev, stream, int(sys.argv[1]) if len(sys.argv) > 1 else 4, potential=os.environ.get("DFTBP_B200_MD_POTENTIAL"))))
