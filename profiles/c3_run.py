#!/usr/bin/env python
"""The C3 section of bench.py alone (allelic matrices at 40 kb + two-step correction), for an ncu launch list:
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
       -k regex:'sym_pass|twostep|vc_|recip|rowstats' --csv --log-file gpurun_out/X.csv python profiles/c3_run.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda:0")
print(json.dumps(bench.c3_section(int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000, dev, 6438.8)))
